"""gs_verify_batch_rand (opt-in; the reference has no such mode): ONE randomised check of a whole batch --
the four ComT entries of every proof folded into a single pairing product, one final exponentiation per call.
Its verdict must equal "every proof passes the exact verifier" (src/verifier.rs:23-157 through gs_verify_batch, itself
byte-checked against the CPU restatement in test_gpu_bigparity.py): accept honest batches of all four equation types,
reject as soon as one constant, Gamma entry, target, commitment or proof element of one proof is changed -- including
changes that cancel in an UNWEIGHTED product (two targets swapped)."""
import os
import random

import pytest

from bigcase import *  # noqa: F401,F403

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def crs():
    return make_crs(1)[0]


@pytest.fixture(scope="module")
def eng(crs):
    import groth_sahai_rs_b200 as gsb
    e = gsb.Engine(0)
    e.crs_load(crs_bytes(crs))
    return e


def rho_of(count, seed):
    return random.Random(seed).randbytes(8 * (2 * count + 1))


def batch_arrays(cases):
    return [b"".join(c.verify_arrays()[i] for c in cases) for i in range(8)]


def sizes_of(ty, m, n):
    return [n * cb.x_size(ty), m * cb.y_size(ty), m * n * 32, cb.target_size(ty), m * 192, n * 384, cb.cx_of(ty) * 384,
            cb.cy_of(ty) * 192]


def swap(buf, a, b, size):
    """exchange the `size`-byte records at byte offsets a and b"""
    out = bytearray(buf)
    out[a:a + size], out[b:b + size] = buf[b:b + size], buf[a:a + size]
    assert bytes(out) != bytes(buf), "tampering must change something"
    return bytes(out)


def tampered(ty, m, n, arrays, which, p, q):
    """arrays with ONE change in array `which` of proof p (targets: proofs p and q exchanged)."""
    sz = sizes_of(ty, m, n)
    bad = list(arrays)
    o = p * sz[which]
    if which == 0:      # two constants A_0 <-> A_1
        e = cb.x_size(ty)
        bad[0] = swap(arrays[0], o, o + e, e)
    elif which == 1:    # B_0 <-> B_1
        e = cb.y_size(ty)
        bad[1] = swap(arrays[1], o, o + e, e)
    elif which == 2:    # one bit of Gamma[1][1]
        g = bytearray(arrays[2])
        g[o + 32 * (n + 1)] ^= 1
        bad[2] = bytes(g)
    elif which == 3:    # the targets of two proofs exchanged: the unweighted product of all targets does not change
        bad[3] = swap(arrays[3], o, q * sz[3], sz[3])
    elif which == 4:    # commitments c_0 <-> c_1
        bad[4] = swap(arrays[4], o, o + 192, 192)
    elif which == 5:    # d_0 <-> d_1
        bad[5] = swap(arrays[5], o, o + 384, 384)
    elif which == 6:    # the two coordinates of pi_0
        bad[6] = swap(arrays[6], o, o + 192, 192)
    else:               # the two coordinates of theta_0
        bad[7] = swap(arrays[7], o, o + 96, 96)
    return bad


@pytest.mark.parametrize("ty", [0, 1, 2, 3])
def test_rand_batch_4x4(eng, crs, ty):
    count, m, n = 12, 4, 4
    cases = [Case(ty, m, n, crs, seed=2000 + 20 * ty + i, zero_frac=0.25 if i % 4 == 3 else 0.0) for i in range(count)]
    arrays = batch_arrays(cases)
    assert eng.verify_batch(ty, count, m, n, *arrays) == b"\x01" * count
    assert eng.verify_batch_rand(ty, count, m, n, *arrays, rho=rho_of(count, 1)) is True
    assert eng.verify_batch_rand(ty, count, m, n, *arrays) is True                      # OS randomness
    assert eng.verify_batch_rand(ty, count, m, n, *arrays, rho=bytes(8 * (2 * count + 1))) is True   # all-zero weights: vacuous
    for which in range(8):
        bad = tampered(ty, m, n, arrays, which, 5, 9)
        exact = eng.verify_batch(ty, count, m, n, *bad)
        assert exact[5] == 0 and exact.count(0) == (2 if which == 3 else 1), f"array {which}: exact verdicts {exact!r}"
        assert eng.verify_batch_rand(ty, count, m, n, *bad, rho=rho_of(count, 2 + which)) is False, f"array {which} accepted"
    # a single proof, honest and tampered
    one = [a[:s] for a, s in zip(arrays, sizes_of(ty, m, n))]
    assert eng.verify_batch_rand(ty, 1, m, n, *one, rho=rho_of(1, 20)) is True
    assert eng.verify_batch_rand(ty, 1, m, n, *tampered(ty, m, n, one, 7, 0, 0), rho=rho_of(1, 21)) is False


def test_rand_ragged_shapes(eng, crs):
    for ty, m, n in ((0, 1, 1), (0, 7, 3), (1, 1, 5), (2, 6, 1), (3, 5, 9)):
        cases = [Case(ty, m, n, crs, seed=2200 + 10 * ty + m + i) for i in range(3)]
        arrays = batch_arrays(cases)
        assert eng.verify_batch_rand(ty, 3, m, n, *arrays, rho=rho_of(3, 30)) is True
        bad = tampered(ty, m, n, arrays, 6, 2, 0)
        assert eng.verify_batch_rand(ty, 3, m, n, *bad, rho=rho_of(3, 31)) is False


def test_rand_several_passes(crs):
    """More proofs than one pass holds (GS_VERIFY_BATCH_MAX lowered for the test): the passes' products and target
    powers are multiplied before the one final exponentiation; a bad proof in the last pass is caught."""
    import groth_sahai_rs_b200 as gsb
    old = os.environ.get("GS_VERIFY_BATCH_MAX")
    os.environ["GS_VERIFY_BATCH_MAX"] = "40"
    try:
        e = gsb.Engine(0)
    finally:
        os.environ.pop("GS_VERIFY_BATCH_MAX", None) if old is None else os.environ.__setitem__("GS_VERIFY_BATCH_MAX", old)
    e.crs_load(crs_bytes(crs))
    for ty in (0, 3):
        m, n = 4, 4
        cases = [Case(ty, m, n, crs, seed=2400 + 10 * ty + i) for i in range(6)]
        count = 100                                              # passes of 40, 40, 20
        arrays = [b"".join(cases[i % 6].verify_arrays()[k] for i in range(count)) for k in range(8)]
        assert e.verify_batch_rand(ty, count, m, n, *arrays, rho=rho_of(count, 40)) is True
        bad = tampered(ty, m, n, arrays, 7, 97, 0)
        assert e.verify_batch_rand(ty, count, m, n, *bad, rho=rho_of(count, 41)) is False
        assert e.verify_batch(ty, count, m, n, *bad) == b"\x01" * 97 + b"\x00" + b"\x01" * 2
    e.close()


def test_rand_big_statement(eng, crs):
    """One 128 x 128 PPE (the shared-base table path of the statement MSM, several thousand folded pairs)."""
    c = Case(0, 128, 128, crs, seed=2500)
    arrays = c.verify_arrays()
    assert eng.verify_batch_rand(0, 1, 128, 128, *arrays, rho=rho_of(1, 50)) is True
    assert eng.verify_batch_rand(0, 1, 128, 128, *tampered(0, 128, 128, arrays, 2, 0, 0), rho=rho_of(1, 51)) is False


@pytest.mark.parametrize("ty", [0, 1, 2, 3])
def test_rand_shared_commitments(eng, crs, ty):
    """Equations over ONE witness set (the C4 shape): the statement MSM keeps its shared-base tables and both
    coordinates, the slots are folded afterwards (the other order of operations than for independent proofs)."""
    E, m, n = 40, 8, 8
    st = Statement(ty, m, n, E, crs, seed=2600 + ty)
    arrays = st.verify_arrays()
    assert eng.verify_batch(ty, E, m, n, *arrays) == b"\x01" * E
    assert eng.verify_batch_rand(ty, E, m, n, *arrays, rho=rho_of(E, 60)) is True
    for which in (0, 2, 3, 6):
        bad = tampered(ty, m, n, arrays, which, 17, 30)
        assert eng.verify_batch_rand(ty, E, m, n, *bad, rho=rho_of(E, 61 + which)) is False, f"array {which} accepted"


@pytest.mark.parametrize("ty", [0, 1, 2, 3])
def test_rand_big_batch_bucket_sums(crs, ty):
    """>= 4,096 proofs: the pi and theta slots are summed over the proofs with the bucket method (one G2 / G1 MSM with the
    64-bit weights per slot) instead of being paired proof by proof.  A bad pi or theta of ONE proof must still be caught."""
    import groth_sahai_rs_b200 as gsb
    old = os.environ.get("GS_RAND_PIP_MIN")
    os.environ["GS_RAND_PIP_MIN"] = "4096"                              # (the default, pinned here)
    try:
        eng = gsb.Engine(0)
    finally:
        os.environ.pop("GS_RAND_PIP_MIN", None) if old is None else os.environ.__setitem__("GS_RAND_PIP_MIN", old)
    eng.crs_load(crs_bytes(crs))
    m, n, reps = 3, 2, 342
    cases = [Case(ty, m, n, crs, seed=2700 + 20 * ty + i) for i in range(12)]
    count = 12 * reps                                                   # 4,104
    arrays = [b"".join(c.verify_arrays()[k] for c in cases) * reps for k in range(8)]
    assert eng.verify_batch_rand(ty, count, m, n, *arrays, rho=rho_of(count, 70)) is True
    for which, p in ((6, 4000), (7, 17), (3, 2222), (4, 4103)):
        bad = tampered(ty, m, n, arrays, which, p, 5)
        assert eng.verify_batch_rand(ty, count, m, n, *bad, rho=rho_of(count, 71 + which)) is False, f"array {which} accepted"
    eng.close()


def test_rand_two_big_passes(eng, crs):
    """96,000 PPE proofs: more than one pass holds at full size (94,720), so two equal passes of 48,000 run, each with its
    own bucket sums and target powers on the second stream; their products meet in the one final exponentiation.  A bad
    theta in the SECOND pass must be caught."""
    ty, m, n, reps = 0, 3, 2, 8000
    cases = [Case(ty, m, n, crs, seed=2800 + i) for i in range(12)]
    count = 12 * reps
    arrays = [b"".join(c.verify_arrays()[k] for c in cases) * reps for k in range(8)]
    assert eng.verify_batch_rand(ty, count, m, n, *arrays, rho=rho_of(count, 80)) is True
    bad = tampered(ty, m, n, arrays, 7, 95000, 0)
    assert eng.verify_batch_rand(ty, count, m, n, *bad, rho=rho_of(count, 81)) is False
