"""Pins the C restatement (checker for big cases + CPU baseline) against the big-int oracle."""
from gsutil import *  # noqa: F401,F403
from oracle import cbaseline as cb
from oracle import gs as ogs
import pytest
from oracle.bls12_381 import multi_pairing, P


def test_c_pairing_and_scalar_mul():
    rng = SeededRng(50)
    p, q = rng.g1(), rng.g2()
    assert fp12_i(cb.pairing(g1_b(p), g2_b(q))) == pairing(p, q)
    assert fp12_i(cb.pairing(g1_b(None), g2_b(q))) == FP12_ONE
    for k in (0, 1, 2, 15, 16, R - 1, rng.fr()):
        assert g1_i(cb.g1_mul(g1_b(p), fr_b(k))) == g1_mul(p, k)
        assert g2_i(cb.g2_mul(g2_b(q), fr_b(k))) == g2_mul(q, k)


def test_c_pairing_sum():
    rng = SeededRng(51)
    xs = [(rng.g1(), rng.g1()), (None, rng.g1())]
    ys = [(rng.g2(), rng.g2()), (rng.g2(), None)]
    got = cb.pairing_sum(b"".join(com1_b(x) for x in xs), b"".join(com2_b(y) for y in ys))
    assert comt_i(got) == ogs.comt_pairing_sum(xs, ys)


def test_c_commit_and_verify():
    crs, _ = make_crs(52)
    rng = SeededRng(53)
    xv = [rng.g1(), None, rng.g1()]
    R2 = [[rng.fr(), rng.fr()], [rng.fr(), 0], [0, 0]]
    got = cb.batch_commit_g1(b"".join(g1_b(x) for x in xv), frmat_b(R2), crs_bytes(crs))
    assert got == b"".join(com1_b(c) for c in ogs.batch_commit_g1(xv, crs, R2).coms)
    yv = [rng.g2(), rng.g2()]
    S2 = [[rng.fr(), rng.fr()], [1, rng.fr()]]
    got = cb.batch_commit_g2(b"".join(g2_b(y) for y in yv), frmat_b(S2), crs_bytes(crs))
    assert got == b"".join(com2_b(c) for c in ogs.batch_commit_g2(yv, crs, S2).coms)
    # verify: honest accepted, tampered rejected, 2 threads
    equ, xs, ys = random_instance(0, 2, 2, crs, rng, zero_frac=0.2)
    xr, yr, T = draw_rands(0, 2, 2, rng)
    pf = ogs.commit_and_prove(equ, xs, ys, crs, xr, yr, T)
    good = proof_bytes(0, equ, pf)
    bad = list(good)
    bad[3] = fp12_b(equ.target * crs.gt_gen)
    arrays = [g + b for g, b in zip(good, bad)]
    assert cb.verify_ppe_batch(2, 2, 2, arrays, crs_bytes(crs), nthreads=2) == b"\x01\x00"


def test_c_inverse_algorithms_agree():
    rng = SeededRng(54)
    for v in (1, 2, P - 1, rng.r.randrange(P), rng.r.randrange(P)):
        assert cb.lib().gsref_selftest_inv(fp_b(v)) == 1




@pytest.mark.parametrize("ty", [0, 1, 2, 3])
def test_c_prove_and_verify_all_types(ty):
    """gsref_prove / gsref_verify / scalar commits against the big-int oracle (3x2, identity constants, zero Gamma
    entries), and the tamper case."""
    crs, _ = make_crs(60 + ty)
    crsb = crs_bytes(crs)
    rng = SeededRng(70 + ty)
    m, n = 3, 2
    equ, xs, ys = random_instance(ty, m, n, crs, rng, zero_frac=0.25)
    xr, yr, T = draw_rands(ty, m, n, rng)
    exp = ogs.commit_and_prove(equ, xs, ys, crs, xr, yr, T)
    xc = cb.commit_x(ty, enc_A(ty, xs), frmat_b(xr), crsb, nthreads=2)
    yc = cb.commit_y(ty, enc_B(ty, ys), frmat_b(yr), crsb, nthreads=2)
    assert xc == b"".join(com1_b(c) for c in exp.xcoms.coms)
    assert yc == b"".join(com2_b(c) for c in exp.ycoms.coms)
    pi, th = cb.prove(ty, m, n, enc_A(ty, equ.a_consts), enc_B(ty, equ.b_consts), frmat_b(equ.gamma), enc_A(ty, xs),
                      enc_B(ty, ys), frmat_b(xr), frmat_b(yr), frmat_b(T), crsb, nthreads=3)
    ep = exp.equ_proofs[0]
    assert pi == b"".join(com2_b(c) for c in ep.pi)
    assert th == b"".join(com1_b(c) for c in ep.theta)
    arrays = proof_bytes(ty, equ, exp)
    assert cb.verify(ty, m, n, arrays, crsb, nthreads=4) is True
    bad = list(arrays)
    bad[7] = bad[7][:96] + bytes(96) if len(bad[7]) == 192 else bad[7][96:192] + bad[7][:96] + bad[7][192:]
    assert cb.verify(ty, m, n, bad, crsb, nthreads=4) is False
    assert cb.verify_batch(ty, 2, m, n, [g + b for g, b in zip(arrays, bad)], crsb, nthreads=2) == b"\x01\x00"


def test_c_prove_batch_shared_vars():
    crs, _ = make_crs(80)
    crsb = crs_bytes(crs)
    rng = SeededRng(81)
    ty, m, n = 1, 2, 2
    equ0, xs, ys = random_instance(ty, m, n, crs, rng)
    equ1, _, _ = random_instance(ty, m, n, crs, rng)
    xr, yr, T0 = draw_rands(ty, m, n, rng)
    _, _, T1 = draw_rands(ty, m, n, rng)
    cat = lambda f, *es: b"".join(f(e) for e in es)
    pi, th = cb.prove_batch(ty, 2, m, n, cat(lambda e: enc_A(ty, e.a_consts), equ0, equ1),
                            cat(lambda e: enc_B(ty, e.b_consts), equ0, equ1), cat(lambda e: frmat_b(e.gamma), equ0, equ1),
                            enc_A(ty, xs), enc_B(ty, ys), frmat_b(xr), frmat_b(yr), frmat_b(T0) + frmat_b(T1), True, crsb, 2)
    for k, (equ, T) in enumerate(((equ0, T0), (equ1, T1))):
        p1, t1 = cb.prove(ty, m, n, enc_A(ty, equ.a_consts), enc_B(ty, equ.b_consts), frmat_b(equ.gamma), enc_A(ty, xs),
                          enc_B(ty, ys), frmat_b(xr), frmat_b(yr), frmat_b(T), crsb)
        assert pi[k * 768:(k + 1) * 768] == p1 and th[k * 192:(k + 1) * 192] == t1
