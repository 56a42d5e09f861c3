"""Pins the C restatement (checker for big cases + CPU baseline) against the big-int oracle."""
from gsutil import *  # noqa: F401,F403
from oracle import cbaseline as cb
from oracle import gs as ogs
from oracle.bls12_381 import multi_pairing


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
