"""The algebra behind gs_verify_batch_rand, restated with the big-integer oracle and checked on the CPU,
independently of any CUDA code: for weights (sigma_p, tau_p) per proof and beta per call, the four ComT entries of
verifier.rs:23-157's equation  F(iota(A), d) + F(c, iota(B)) + F(c, Gamma d) = iota_T(t) + F(u, pi) + F(theta, v)
fold, by bilinearity, into ONE product of pairings over folded points
        fold1(x) = sigma x.0 + tau x.1   (x in Com1),      fold2(y) = beta y.0 + y.1   (y in Com2),
with entry (a, b) of a ComT raised to w_a * beta_b, (w_0, w_1) = (sigma, tau), (beta_0, beta_1) = (beta, 1).
An honest proof satisfies the folded equation for EVERY choice of weights; a tampered one fails it (for random weights),
and so does a pair of proofs whose errors cancel in an unweighted product."""
import random

import pytest

from gsutil import SeededRng, make_crs, random_instance, draw_rands
from oracle import gs as ogs
from oracle.bls12_381 import G1, G2, R, FP12_ONE, g1_mul, g2_mul, pairing

XA = 0xD201000000010000


def fold1(x, sg, tu):
    return G1.add(g1_mul(x[0], sg), g1_mul(x[1], tu))


def fold2(y, beta):
    return G2.add(g2_mul(y[0], beta), y[1])


def fold_t(m4, sg, tu, beta):
    """ComT [e00, e01, e10, e11] -> e00^(sigma beta) e01^sigma e10^(tau beta) e11^tau"""
    return m4[0].pow(sg * beta % R) * m4[1].pow(sg) * m4[2].pow(tau_beta(tu, beta)) * m4[3].pow(tu)


def tau_beta(tu, beta):
    return tu * beta % R


def folded_sides(equ, proof, crs, sg, tu, beta):
    """(LHS, RHS) of the folded equation of one proof, each ONE GT element."""
    ty = equ.equ_type
    ep = proof.equ_proofs[0]
    pairs_l = list(zip(ogs._map_x(ty, equ.a_consts, crs), proof.ycoms.coms))
    pairs_l += list(zip(proof.xcoms.coms, ogs._map_y(ty, equ.b_consts, crs)))
    gd = ogs.col_vec_to_vec(ogs.com_left_mul(ogs.vec_to_col_vec(proof.ycoms.coms), equ.gamma, 2))
    pairs_l += list(zip(proof.xcoms.coms, gd))
    pairs_r = list(zip(crs.u, ep.pi)) if ty in (ogs.PPE, ogs.MSMEG1) else [(crs.u[0], ep.pi[0])]
    pairs_r += list(zip(ep.theta, crs.v)) if ty in (ogs.PPE, ogs.MSMEG2) else [(ep.theta[0], crs.v[0])]
    lin_t = {ogs.PPE: lambda: ogs.comt_linear_map_ppe(equ.target),
             ogs.MSMEG1: lambda: ogs.comt_linear_map_msmeg1(equ.target, crs),
             ogs.MSMEG2: lambda: ogs.comt_linear_map_msmeg2(equ.target, crs),
             ogs.QUAD: lambda: ogs.comt_linear_map_quad(equ.target, crs)}[ty]()

    def prod(pairs):
        acc = FP12_ONE
        for x, y in pairs:
            acc = acc * pairing(fold1(x, sg, tu), fold2(y, beta))
        return acc

    return prod(pairs_l), fold_t(lin_t, sg, tu, beta) * prod(pairs_r)


@pytest.mark.parametrize("ty", [0, 1, 2, 3])
def test_folded_equation_holds_exactly_for_honest_proofs(ty):
    crs, _ = make_crs(1)
    rng = SeededRng(300 + ty)
    m, n = (2, 1) if ty != 3 else (1, 2)
    equ, xv, yv = random_instance(ty, m, n, crs, rng)
    xr, yr, T = draw_rands(ty, m, n, rng)
    proof = ogs.commit_and_prove(equ, xv, yv, crs, xr, yr, T)
    assert ogs.verify(equ, proof, crs)
    w = random.Random(7 + ty)
    sg, tu = w.getrandbits(63), w.getrandbits(63)
    word = w.getrandbits(64)
    beta = ((word & 0xFFFFFFFF) + (word >> 32) * XA) % R            # the form the G2 fold uses: b0 + b1 |x|
    lhs, rhs = folded_sides(equ, proof, crs, sg, tu, beta)
    assert lhs == rhs
    # degenerate weights: sigma = 0 keeps the (1, .) entries only, beta = 0 the (., 1) entries
    lhs, rhs = folded_sides(equ, proof, crs, 0, 1, 0)
    assert lhs == rhs


def test_folded_equation_rejects_and_weights_matter():
    """PPE: a proof with pi's coordinates swapped fails the folded equation; two proofs whose targets are exchanged pass
    the UNWEIGHTED product of their equations' (1,1) entries, and fail the weighted one."""
    crs, _ = make_crs(1)
    rng = SeededRng(340)
    inst = []
    for _ in range(2):
        equ, xv, yv = random_instance(0, 2, 1, crs, rng)
        xr, yr, T = draw_rands(0, 2, 1, rng)
        inst.append((equ, ogs.commit_and_prove(equ, xv, yv, crs, xr, yr, T)))
    w = random.Random(11)
    (s0, t0), (s1, t1), beta = (w.getrandbits(63), w.getrandbits(63)), (w.getrandbits(63), w.getrandbits(63)), w.getrandbits(64)
    equ, proof = inst[0]
    ep = proof.equ_proofs[0]
    bad = ogs.CProof(proof.xcoms, proof.ycoms, [ogs.EquProof([(ep.pi[0][1], ep.pi[0][0]), ep.pi[1]], ep.theta, ep.equ_type, ep.rand)])
    assert not ogs.verify(equ, bad, crs)
    lhs, rhs = folded_sides(equ, bad, crs, s0, t0, beta)
    assert lhs != rhs
    # exchanged targets
    e0 = ogs.Equation(inst[0][0].equ_type, inst[0][0].a_consts, inst[0][0].b_consts, inst[0][0].gamma, inst[1][0].target)
    e1 = ogs.Equation(inst[1][0].equ_type, inst[1][0].a_consts, inst[1][0].b_consts, inst[1][0].gamma, inst[0][0].target)
    assert not ogs.verify(e0, inst[0][1], crs) and not ogs.verify(e1, inst[1][1], crs)
    l0, r0 = folded_sides(e0, inst[0][1], crs, 1, 1, 1)
    l1, r1 = folded_sides(e1, inst[1][1], crs, 1, 1, 1)
    assert l0 * l1 == r0 * r1                      # equal weights: the two errors cancel
    l0, r0 = folded_sides(e0, inst[0][1], crs, s0, t0, beta)
    l1, r1 = folded_sides(e1, inst[1][1], crs, s1, t1, beta)
    assert l0 * l1 != r0 * r1                      # independent weights per proof: caught
