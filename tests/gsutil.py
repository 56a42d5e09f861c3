"""Shared helpers for the parity tests: seeded inputs in oracle form and their C-ABI bytes."""
import random

from conv import *  # noqa: F401,F403
from oracle import gs as ogs
from oracle.bls12_381 import G1, G2, G1_GEN, G2_GEN_FP2, R, g1_mul, g2_mul, pairing, FP12_ONE


class SeededRng:
    """Deterministic host randomness in ORACLE form (ints / points)."""

    def __init__(self, seed):
        self.r = random.Random(seed)

    def fr(self): return self.r.randrange(R)
    def g1(self): return g1_mul(G1_GEN, self.fr())
    def g2(self): return g2_mul(G2_GEN_FP2, self.fr())


def make_crs(seed=1):
    rng = SeededRng(seed)
    draws = (rng.g1(), rng.g2(), rng.fr(), rng.fr(), rng.fr(), rng.fr())
    return ogs.generate_crs(*draws), draws


def crs_bytes(crs):
    return (com1_b(crs.u[0]) + com1_b(crs.u[1]) + com2_b(crs.v[0]) + com2_b(crs.v[1]) +
            g1_b(crs.g1_gen) + g2_b(crs.g2_gen) + fp12_b(crs.gt_gen))


def enc_A(ty, xs): return b"".join(g1_b(x) for x in xs) if ty in (0, 1) else frs_b(xs)
def enc_B(ty, ys): return b"".join(g2_b(y) for y in ys) if ty in (0, 2) else frs_b(ys)
def enc_T(ty, t): return {0: fp12_b, 1: g1_b, 2: g2_b, 3: fr_b}[ty](t)


def random_instance(ty, m, n, crs, rng, zero_frac=0.0, gamma_small=False):
    """A satisfied equation of type `ty` with m x-variables, n y-variables (oracle form).
    Returns (equation, xvars, yvars).  Some constants / Gamma entries are made trivial on purpose."""
    g1, g2 = crs.g1_gen, crs.g2_gen
    rs = lambda: (0 if rng.r.random() < zero_frac else rng.fr())
    xs = [rng.fr() for _ in range(m)]          # discrete logs / scalar witnesses
    ys = [rng.fr() for _ in range(n)]
    a = [rs() for _ in range(n)]
    b = [rs() for _ in range(m)]
    gamma = [[(rng.r.randrange(8) if gamma_small else rs()) for _ in range(n)] for _ in range(m)]
    # value of the equation "in the exponent"
    val = (sum(a[j] * ys[j] for j in range(n)) + sum(xs[i] * b[i] for i in range(m)) +
           sum(gamma[i][j] * xs[i] * ys[j] for i in range(m) for j in range(n))) % R
    pt1 = lambda k: g1_mul(g1, k) if k else None
    pt2 = lambda k: g2_mul(g2, k) if k else None
    if ty == 0:
        equ = ogs.Equation(0, [pt1(k) for k in a], [pt2(k) for k in b], gamma, crs.gt_gen.pow(val))
        return equ, [pt1(k) for k in xs], [pt2(k) for k in ys]
    if ty == 1:
        equ = ogs.Equation(1, [pt1(k) for k in a], b, gamma, pt1(val))
        return equ, [pt1(k) for k in xs], ys
    if ty == 2:
        equ = ogs.Equation(2, a, [pt2(k) for k in b], gamma, pt2(val))
        return equ, xs, [pt2(k) for k in ys]
    equ = ogs.Equation(3, a, b, gamma, val)
    return equ, xs, ys


def draw_rands(ty, m, n, rng):
    cx = 2 if ty in (0, 1) else 1
    cy = 2 if ty in (0, 2) else 1
    xr = [[rng.fr() for _ in range(cx)] for _ in range(m)]
    yr = [[rng.fr() for _ in range(cy)] for _ in range(n)]
    T = [[rng.fr() for _ in range(cx)] for _ in range(cy)]
    return xr, yr, T


def proof_bytes(ty, equ, proof):
    """The 8 C-ABI arrays of one (equation, CProof), in gs_verify_batch order."""
    ep = proof.equ_proofs[0]
    return [enc_A(ty, equ.a_consts), enc_B(ty, equ.b_consts), frmat_b(equ.gamma), enc_T(ty, equ.target),
            b"".join(com1_b(c) for c in proof.xcoms.coms), b"".join(com2_b(c) for c in proof.ycoms.coms),
            b"".join(com2_b(c) for c in ep.pi), b"".join(com1_b(c) for c in ep.theta)]
