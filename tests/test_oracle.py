"""Pins the oracle: public BLS12-381 anchors, two independent pairing implementations, the
exponent identity, and every algebraic identity / KAT the reference's own tests assert for the
hot path.  arkworks golden bytes do not exist offline => "parity unpinned"."""
import pytest

from gsutil import *  # noqa: F401,F403
from oracle import gs as ogs
from oracle import bls12_381 as o


def test_public_anchors():
    assert o.G1.on_curve(o.G1_GEN) and o.G2.on_curve(o.G2_GEN_FP2)
    assert o.G1.mul(o.G1_GEN, o.R) is None and o.G2.mul(o.G2_GEN_FP2, o.R) is None
    assert o.FINAL_EXP_HARD == 3 * ((o.P ** 4 - o.P ** 2 + 1) // o.R)
    assert o.R_FP == 0x15F65EC3FA80E4935C071A97A256EC6D77CE5853705257455F48985753C758BAEBF4000BC40C0002760900000002FFFD
    assert (-pow(o.P, -1, 1 << 64)) % (1 << 64) == 0x89F3FFFCFFFCFFFD


def test_two_pairing_implementations_agree():
    e1 = o.pairing_textbook(o.G1_GEN, o.G2_GEN_FP2)
    e2 = o.pairing(o.G1_GEN, o.G2_GEN_FP2)
    assert e1 == e2 and not e1.is_one() and e1.pow(o.R).is_one()
    a, b = 0x1234567890ABCDEF, 0xFEDCBA0987654321
    assert o.pairing(g1_mul(o.G1_GEN, a), g2_mul(o.G2_GEN_FP2, b)) == e1.pow(a * b % o.R)


def test_matrix_kats():
    # data_structures.rs:1678-1947
    assert ogs.fr_right_mul([[1, 2, 3]], [[4], [5], [6]]) == [[32]]
    a = [[1, 2, 3], [4, 5, 6]]
    b = [[7, 8, 9, 10], [11, 12, 13, 14], [15, 16, 17, 18]]
    assert ogs.fr_right_mul(a, b) == [[74, 80, 86, 92], [173, 188, 203, 218]]
    assert ogs.fr_left_mul(b, a) == [[74, 80, 86, 92], [173, 188, 203, 218]]
    assert ogs.fr_transpose(a) == [[1, 4], [2, 5], [3, 6]]


def test_comt_structure():
    rng = SeededRng(40)
    x, y = (rng.g1(), rng.g1()), (rng.g2(), rng.g2())
    # identity inputs -> GT identity; (O,X)x(O,Y) -> only entry 3 non-trivial (data_structures.rs:1313-1357)
    assert ogs.comt_pairing((None, None), y) == [o.FP12_ONE] * 4
    t = ogs.comt_pairing(ogs.com1_linear_map(x[1]), ogs.com2_linear_map(y[1]))
    assert t[:3] == [o.FP12_ONE] * 3 and t[3] == o.pairing(x[1], y[1])
    # pairing_sum == sum of pairings (:1381-1407)
    x2, y2 = (rng.g1(), rng.g1()), (rng.g2(), rng.g2())
    assert ogs.comt_eq(ogs.comt_pairing_sum([x, x2], [y, y2]),
                       ogs.comt_add(ogs.comt_pairing(x, y), ogs.comt_pairing(x2, y2)))


def test_crs_and_commutativity():
    crs, (p1, p2, a1, a2, t1, t2) = make_crs(41)
    assert crs.gt_gen == o.pairing(p1, p2)
    assert crs.u[1][1] == g1_mul(g1_mul(p1, a1), t1) and crs.v[1][1] == g2_mul(g2_mul(p2, a2), t2)
    # iota_T(f(x, y)) == F(iota_1(x), iota_2'(y))  for MSMEG1 (tests/commit.rs:22-85)
    rng = SeededRng(42)
    x, s = rng.g1(), rng.fr()
    lhs = ogs.comt_linear_map_msmeg1(g1_mul(x, s), crs)
    rhs = ogs.comt_pairing(ogs.com1_linear_map(x), ogs.com2_scalar_linear_map(s, crs))
    assert ogs.comt_eq(lhs, rhs)


@pytest.mark.parametrize("ty", [0, 3])
def test_completeness_and_soundness_smoke(ty):
    crs, _ = make_crs(43)
    rng = SeededRng(44 + ty)
    equ, xv, yv = random_instance(ty, 2, 1, crs, rng)
    xr, yr, T = draw_rands(ty, 2, 1, rng)
    pf = ogs.commit_and_prove(equ, xv, yv, crs, xr, yr, T)
    assert ogs.verify(equ, pf, crs)
    bad = ogs.Equation(ty, equ.a_consts, equ.b_consts, equ.gamma,
                       equ.target * crs.gt_gen if ty == 0 else (equ.target + 1) % o.R)
    assert not ogs.verify(bad, pf, crs)
