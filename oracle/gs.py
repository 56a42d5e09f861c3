"""Groth-Sahai layer of the oracle (TEST INFRASTRUCTURE ONLY -- see bls12_381.py header).

A function-for-function CPU restatement of the reference's hot path, in the
reference's own (deliberately naive) evaluation order:

  src/data_structures.rs:300-541   linear maps, Com scalar_mul, ComT::pairing /
                                   pairing_sum, the four iota_T maps
  src/data_structures.rs:645-742   Mat::left_mul / right_mul on Matrix<Com>
  src/data_structures.rs:768-913   Mat on Matrix<Fr>
  src/generator.rs:81-118          CRS::generate_crs
  src/prover/commit.rs:78-256      batch_commit_G1/G2, batch_commit_scalar_to_B1/B2
  src/prover/prove.rs:92-488       Provable::prove x4 (+ commit_and_prove)
  src/verifier.rs:23-157           Verifiable::verify x4

Randomness is an explicit argument everywhere (the reference draws it from an
``Rng``; the draw ORDER is restated in ``commit_and_prove_*``).

Types: Fr = int mod R; G1 = None | (x, y) ints; G2 = None | (Fp2, Fp2);
Com1 = (G1, G1); Com2 = (G2, G2); ComT = [Fp12]*4 in row-major order
[e(x0,y0), e(x0,y1), e(x1,y0), e(x1,y1)] (data_structures.rs:484-491).
PARITY STATUS: parity unpinned (no arkworks golden vectors exist).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Sequence

from .bls12_381 import (FP12_ONE, Fp12, G1, G2, R, g1_mul, g2_mul,
                        multi_pairing, pairing)

PPE, MSMEG1, MSMEG2, QUAD = 0, 1, 2, 3   # EquType byte, src/statement.rs:68-73


# ---------------------------------------------------------------- Com1 / Com2 / ComT  (data_structures.rs:128-479)
def com1_add(a, b): return (G1.add(a[0], b[0]), G1.add(a[1], b[1]))
def com2_add(a, b): return (G2.add(a[0], b[0]), G2.add(a[1], b[1]))
def com1_neg(a): return (G1.neg(a[0]), G1.neg(a[1]))
def com2_neg(a): return (G2.neg(a[0]), G2.neg(a[1]))
def com1_scalar_mul(a, s): return (g1_mul(a[0], s), g1_mul(a[1], s))     # :336-342
def com2_scalar_mul(a, s): return (g2_mul(a[0], s), g2_mul(a[1], s))     # :381-387
COM1_ZERO = (None, None)
COM2_ZERO = (None, None)


def comt_add(a, b): return [x * y for x, y in zip(a, b)]                 # GT written additively
def comt_neg(a): return [x.conj() for x in a]
def comt_zero(): return [FP12_ONE] * 4
def comt_eq(a, b): return all(x == y for x, y in zip(a, b))


def com1_linear_map(x): return (None, x)                                  # :310-312
def com2_linear_map(y): return (None, y)                                  # :355-357


@dataclass
class CRS:                                                                # generator.rs:36-42
    u: list
    v: list
    g1_gen: tuple
    g2_gen: tuple
    gt_gen: Fp12


def w1(crs): return com1_add(crs.u[1], com1_linear_map(crs.g1_gen))       # :325
def w2(crs): return com2_add(crs.v[1], com2_linear_map(crs.g2_gen))       # :370
def com1_scalar_linear_map(x, crs): return com1_scalar_mul(w1(crs), x)    # :323-326
def com2_scalar_linear_map(y, crs): return com2_scalar_mul(w2(crs), y)    # :368-371


def comt_pairing(x, y):                                                   # :484-491
    return [pairing(x[0], y[0]), pairing(x[0], y[1]), pairing(x[1], y[0]), pairing(x[1], y[1])]


def comt_pairing_sum(xs, ys):                                             # :494-502
    assert len(xs) == len(ys)
    return [multi_pairing([(x[a], y[b]) for x, y in zip(xs, ys)]) for a in (0, 1) for b in (0, 1)]


def comt_linear_map_ppe(t): return [FP12_ONE, FP12_ONE, FP12_ONE, t]      # :509-516
def comt_linear_map_msmeg1(t, crs):                                        # :519-524
    return comt_pairing(com1_linear_map(t), com2_scalar_linear_map(1, crs))
def comt_linear_map_msmeg2(t, crs):                                        # :527-532
    return comt_pairing(com1_scalar_linear_map(1, crs), com2_linear_map(t))
def comt_linear_map_quad(t, crs):                                          # :535-540
    return comt_pairing(com1_scalar_linear_map(1, crs),
                        com2_scalar_mul(com2_scalar_linear_map(1, crs), t))


# ---------------------------------------------------------------- Mat  (data_structures.rs:545-913)
def fr_transpose(m): return [list(r) for r in zip(*m)]
def fr_neg(m): return [[(-x) % R for x in r] for r in m]
def fr_add(a, b):
    assert len(a) == len(b) and len(a[0]) == len(b[0])
    return [[(x + y) % R for x, y in zip(ra, rb)] for ra, rb in zip(a, b)]
def fr_scalar_mul(m, s): return [[x * s % R for x in r] for r in m]


def fr_right_mul(a, rhs):                                                 # :824-868  self * rhs
    if not a or not a[0] or not rhs or not rhs[0]:
        return []
    assert len(a[0]) == len(rhs)
    return [[sum(a[i][k] * rhs[k][j] for k in range(len(rhs))) % R
             for j in range(len(rhs[0]))] for i in range(len(a))]


def fr_left_mul(a, lhs):                                                  # :870-912  lhs * self
    if not a or not a[0] or not lhs or not lhs[0]:
        return []
    assert len(lhs[0]) == len(a)
    return [[sum(a[k][j] * lhs[i][k] for k in range(len(a))) % R
             for j in range(len(a[0]))] for i in range(len(lhs))]


def _com_ops(which):
    return (com1_add, com1_scalar_mul, COM1_ZERO) if which == 1 else (com2_add, com2_scalar_mul, COM2_ZERO)


def com_left_mul(mat, lhs, which):                                        # :696-742: out[i][j] = sum_k mat[k][j] * lhs[i][k]
    """lhs (Fr, r x k)  times  mat (Com, k x c)  ->  r x c matrix of Com."""
    add, smul, zero = _com_ops(which)
    if not lhs or not lhs[0] or not mat or not mat[0]:
        return []
    assert len(lhs[0]) == len(mat)
    out = []
    for i in range(len(lhs)):
        row = []
        for j in range(len(mat[0])):
            acc = zero
            for k in range(len(mat)):
                acc = add(acc, smul(mat[k][j], lhs[i][k]))
            row.append(acc)
        out.append(row)
    return out


def com_right_mul(mat, rhs, which):                                       # :645-694: out[i][j] = sum_k mat[i][k] * rhs[k][j]
    add, smul, zero = _com_ops(which)
    if not rhs or not rhs[0] or not mat or not mat[0]:
        return []
    assert len(mat[0]) == len(rhs)
    out = []
    for i in range(len(mat)):
        row = []
        for j in range(len(rhs[0])):
            acc = zero
            for k in range(len(rhs)):
                acc = add(acc, smul(mat[i][k], rhs[k][j]))
            row.append(acc)
        out.append(row)
    return out


def com_mat_add(a, b, which):                                             # :588-600
    add = _com_ops(which)[0]
    assert len(a) == len(b) and len(a[0]) == len(b[0])
    return [[add(x, y) for x, y in zip(ra, rb)] for ra, rb in zip(a, b)]


def vec_to_col_vec(v): return [[x] for x in v]                            # :153-160
def col_vec_to_vec(m):                                                    # :145-151
    if len(m) == 1:
        return list(m[0])
    return [r[0] for r in m]


# ---------------------------------------------------------------- CRS  (generator.rs:81-118)
def generate_crs(p1, p2, a1, a2, t1, t2) -> CRS:
    """The six values are the reference's draws, in its order (generator.rs:86-93)."""
    q1 = g1_mul(p1, a1)
    q2 = g2_mul(p2, a2)
    u1 = g1_mul(p1, t1)
    u2 = g2_mul(p2, t2)
    v1 = g1_mul(q1, t1)       # prepare_real_binding_key :57-58
    v2 = g2_mul(q2, t2)
    return CRS(u=[(p1, q1), (u1, v1)], v=[(p2, q2), (u2, v2)],
               g1_gen=p1, g2_gen=p2, gt_gen=pairing(p1, p2))


# ---------------------------------------------------------------- commit  (commit.rs)
@dataclass
class Commit:
    coms: list
    rand: list      # Matrix<Fr>


def batch_commit_g1(xvars, crs, rand):                                    # :78-100 ; rand = m x 2
    assert all(len(r) == 2 for r in rand) and len(rand) == len(xvars)
    lin_x = vec_to_col_vec([com1_linear_map(x) for x in xvars])
    coms = com_mat_add(lin_x, com_left_mul(vec_to_col_vec(crs.u), rand, 1), 1)
    return Commit(col_vec_to_vec(coms), [list(r) for r in rand])


def batch_commit_g2(yvars, crs, rand):                                    # :178-200
    assert all(len(r) == 2 for r in rand) and len(rand) == len(yvars)
    lin_y = vec_to_col_vec([com2_linear_map(y) for y in yvars])
    coms = com_mat_add(lin_y, com_left_mul(vec_to_col_vec(crs.v), rand, 2), 2)
    return Commit(col_vec_to_vec(coms), [list(r) for r in rand])


def batch_commit_scalar_to_b1(xs, crs, rand):                             # :125-156 ; rand = m' x 1
    assert all(len(r) == 1 for r in rand) and len(rand) == len(xs)
    coms = [com1_add(com1_scalar_linear_map(x, crs), com1_scalar_mul(crs.u[0], r[0]))
            for x, r in zip(xs, rand)]
    return Commit(coms, [list(r) for r in rand])


def batch_commit_scalar_to_b2(ys, crs, rand):                             # :225-256
    assert all(len(r) == 1 for r in rand) and len(rand) == len(ys)
    coms = [com2_add(com2_scalar_linear_map(y, crs), com2_scalar_mul(crs.v[0], r[0]))
            for y, r in zip(ys, rand)]
    return Commit(coms, [list(r) for r in rand])


# ---------------------------------------------------------------- statements / proofs
@dataclass
class Equation:                                                           # statement.rs:117-185
    equ_type: int
    a_consts: list
    b_consts: list
    gamma: list        # m x n Matrix<Fr>
    target: object     # GT | G1 | G2 | Fr


@dataclass
class EquProof:                                                           # prove.rs:55-61
    pi: list
    theta: list
    equ_type: int
    rand: list


@dataclass
class CProof:                                                             # prove.rs:64-69
    xcoms: Commit
    ycoms: Commit
    equ_proofs: List[EquProof] = field(default_factory=list)


def _map_x(equ_type, xs, crs):
    if equ_type in (PPE, MSMEG1):
        return [com1_linear_map(x) for x in xs]
    return [com1_scalar_linear_map(x, crs) for x in xs]


def _map_y(equ_type, ys, crs):
    if equ_type in (PPE, MSMEG2):
        return [com2_linear_map(y) for y in ys]
    return [com2_scalar_linear_map(y, crs) for y in ys]


def t_shape(equ_type):
    """(rows, cols) of the proof randomness T: prove.rs:123-126, 226-227, 329-332, 440."""
    return {PPE: (2, 2), MSMEG1: (1, 2), MSMEG2: (2, 1), QUAD: (1, 1)}[equ_type]


def prove(equ: Equation, xvars, yvars, xcoms: Commit, ycoms: Commit, crs: CRS, pf_rand) -> EquProof:
    """prove.rs:92-171 (PPE), 195-274 (MSMEG1), 298-379 (MSMEG2), 409-488 (Quad);
    pf_rand is T with shape t_shape(equ_type), row-major in the reference's draw order."""
    ty = equ.equ_type
    cx = 2 if ty in (PPE, MSMEG1) else 1
    cy = 2 if ty in (PPE, MSMEG2) else 1
    assert len(xvars) == len(xcoms.rand) and len(equ.gamma) == len(xcoms.rand)
    assert len(xcoms.rand[0]) == cx
    assert len(yvars) == len(ycoms.rand) and len(equ.gamma[0]) == len(ycoms.rand)
    assert len(ycoms.rand[0]) == cy
    assert (len(pf_rand), len(pf_rand[0])) == t_shape(ty)

    x_rand_trans = fr_transpose(xcoms.rand)
    y_rand_trans = fr_transpose(ycoms.rand)

    x_rand_lin_b = com_left_mul(vec_to_col_vec(_map_y(ty, equ.b_consts, crs)), x_rand_trans, 2)
    x_rand_stmt = fr_right_mul(x_rand_trans, equ.gamma)
    x_rand_stmt_lin_y = com_left_mul(vec_to_col_vec(_map_y(ty, yvars, crs)), x_rand_stmt, 2)
    pf_rand_stmt = fr_add(fr_right_mul(fr_right_mul(x_rand_trans, equ.gamma), ycoms.rand),
                          fr_neg(fr_transpose(pf_rand)))
    vkey = vec_to_col_vec(crs.v) if cy == 2 else [[crs.v[0]]]
    pf_rand_stmt_com2 = com_left_mul(vkey, pf_rand_stmt, 2)
    pi = col_vec_to_vec(com_mat_add(com_mat_add(x_rand_lin_b, x_rand_stmt_lin_y, 2), pf_rand_stmt_com2, 2))
    assert len(pi) == cx

    y_rand_lin_a = com_left_mul(vec_to_col_vec(_map_x(ty, equ.a_consts, crs)), y_rand_trans, 1)
    y_rand_stmt = fr_right_mul(y_rand_trans, fr_transpose(equ.gamma))
    y_rand_stmt_lin_x = com_left_mul(vec_to_col_vec(_map_x(ty, xvars, crs)), y_rand_stmt, 1)
    ukey = vec_to_col_vec(crs.u) if cx == 2 else [[crs.u[0]]]
    pf_rand_com1 = com_left_mul(ukey, pf_rand, 1)
    theta = col_vec_to_vec(com_mat_add(com_mat_add(y_rand_lin_a, y_rand_stmt_lin_x, 1), pf_rand_com1, 1))
    assert len(theta) == cy
    return EquProof(pi, theta, ty, [list(r) for r in pf_rand])


def commit_and_prove(equ: Equation, xvars, yvars, crs: CRS, x_rand, y_rand, pf_rand) -> CProof:
    """prove.rs:72-90 etc.  RNG order in the reference: x_rand rows, then y_rand rows, then T."""
    ty = equ.equ_type
    xcoms = (batch_commit_g1 if ty in (PPE, MSMEG1) else batch_commit_scalar_to_b1)(xvars, crs, x_rand)
    ycoms = (batch_commit_g2 if ty in (PPE, MSMEG2) else batch_commit_scalar_to_b2)(yvars, crs, y_rand)
    return CProof(xcoms, ycoms, [prove(equ, xvars, yvars, xcoms, ycoms, crs, pf_rand)])


def verify(equ: Equation, proof: CProof, crs: CRS) -> bool:
    """verifier.rs:23-157, in the reference's evaluation order (5 pairing_sums / pairings, 20 final exps)."""
    assert len(proof.equ_proofs) == 1
    ep = proof.equ_proofs[0]
    assert ep.equ_type == equ.equ_type
    ty = equ.equ_type
    lin_a_com_y = comt_pairing_sum(_map_x(ty, equ.a_consts, crs), proof.ycoms.coms)
    com_x_lin_b = comt_pairing_sum(proof.xcoms.coms, _map_y(ty, equ.b_consts, crs))
    stmt_com_y = com_left_mul(vec_to_col_vec(proof.ycoms.coms), equ.gamma, 2)
    com_x_stmt_com_y = comt_pairing_sum(proof.xcoms.coms, col_vec_to_vec(stmt_com_y))
    if ty == PPE:
        lin_t = comt_linear_map_ppe(equ.target)
    elif ty == MSMEG1:
        lin_t = comt_linear_map_msmeg1(equ.target, crs)
    elif ty == MSMEG2:
        lin_t = comt_linear_map_msmeg2(equ.target, crs)
    else:
        lin_t = comt_linear_map_quad(equ.target, crs)
    if ty in (PPE, MSMEG1):
        com1_pf2 = comt_pairing_sum(crs.u, ep.pi)
    else:
        com1_pf2 = comt_pairing(crs.u[0], ep.pi[0])
    if ty in (PPE, MSMEG2):
        pf1_com2 = comt_pairing_sum(ep.theta, crs.v)
    else:
        pf1_com2 = comt_pairing(ep.theta[0], crs.v[0])
    lhs = comt_add(comt_add(lin_a_com_y, com_x_lin_b), com_x_stmt_com_y)
    rhs = comt_add(comt_add(lin_t, com1_pf2), pf1_com2)
    return comt_eq(lhs, rhs)
