"""GPU tests at the sizes of BASELINE.json's configs, through size-independent properties (the big-int
oracle cannot follow there): completeness + tamper masks for batch verification (C5), a large dense
statement through the chunked pairing product (C3), linearity of the batched commitments (C2) and the four
equation types side by side (C4).  Everything goes through the C ABI; comparisons are byte equality."""
import os
import sys

import pytest

from gsutil import *  # noqa: F401,F403

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def eng():
    import groth_sahai_rs_b200 as gsb
    e = gsb.Engine(0)
    crs, _ = make_crs(1)
    e.crs_load(crs_bytes(crs))
    e._crs = crs
    return e


from workloads import multiples_g1 as _multiples_g1, multiples_g2 as _multiples_g2, instance as _instance, commit_prove as _commit_prove  # noqa: E402


def test_c5_batch_verify_tamper_mask(eng):
    """C5 shape: 4x4 PPE proofs, distinct instances tiled to 8,192 proofs, 1 % tampered -> exact mask."""
    sys.path.insert(0, ROOT)
    import bench
    arrays, expected, _ = bench.build_workload(eng, distinct=8, proofs=8192, seed=5)
    ok = eng.verify_batch(0, 8192, 4, 4, *[a.tobytes() for a in arrays])
    assert bytes(ok) == expected.tobytes()
    assert expected.sum() == 8192 - len(range(37, 8192, 100))
    crs, _ = make_crs(1)
    eng.crs_load(crs_bytes(crs))           # build_workload loaded its own key


@pytest.mark.parametrize("m,n", [(96, 64)])
def test_c3_large_dense_statement(eng, m, n):
    """One PPE with a dense Gamma (C3 shape, scaled): K = n + m + 4 slots go through the chunked pairing product."""
    rng = SeededRng(3)
    inst = _instance(eng, 0, m, n, rng)
    arrs = _commit_prove(eng, 0, m, n, inst, rng)
    assert eng.verify(0, m, n, *arrs) is True
    bad = list(arrs)
    g = bytearray(bad[2])
    g[32 * (5 * n + 7)] ^= 1                         # one bit of Gamma[5][7]
    bad[2] = bytes(g)
    assert eng.verify(0, m, n, *bad) is False
    bad = list(arrs)
    bad[4] = bad[4][192:384] + bad[4][:192] + bad[4][384:]   # swap two commitments
    assert eng.verify(0, m, n, *bad) is False


def test_pairing_sum_splits(eng):
    """ComT::pairing_sum over 300 pairs == entry-wise product of the sums over the two halves
    (data_structures.rs:1381-1407 at a size that uses several accumulator chunks)."""
    rng = SeededRng(31)
    k = 300
    ks = [rng.fr() for _ in range(4 * k)]
    p = _multiples_g1(eng, ks[:2 * k])
    q = _multiples_g2(eng, ks[2 * k:])
    xs = b"".join(p[2 * i] + p[2 * i + 1] for i in range(k))
    ys = b"".join(q[2 * i] + q[2 * i + 1] for i in range(k))
    whole = eng.comt_pairing_sum(xs, ys)
    h = k // 2
    a = eng.comt_pairing_sum(xs[:h * 192], ys[:h * 384])
    b = eng.comt_pairing_sum(xs[h * 192:], ys[h * 384:])
    from oracle.bls12_381 import Fp12  # noqa: F401  (only the GT product of two 576-byte values is done on the host)
    for e in range(4):
        prod = fp12_i(a[576 * e:576 * (e + 1)]) * fp12_i(b[576 * e:576 * (e + 1)])
        assert fp12_b(prod) == whole[576 * e:576 * (e + 1)]
    # and the discrete-log check of entry (0,0): prod e(x_i g1, y_i g2) = gt^(sum x_i y_i)
    s = sum(ks[2 * i] * ks[2 * k + 2 * i] for i in range(k)) % R
    assert whole[:576] == eng.pairing(_multiples_g1(eng, [s])[0], g2_b(eng._crs.g2_gen))


def test_c2_batch_commit_linearity(eng):
    """C2 shape (scaled to 2^14 variables): commit(X, 0) = iota(X); commit(X, R) - commit(X, 0) does not depend
    on X; and the big-table path (batches >= 8192) equals the small-table path on the same inputs."""
    rng = SeededRng(2)
    n = 1 << 14
    base = _multiples_g1(eng, [rng.fr() for _ in range(64)])
    X = b"".join(base[i % 64] for i in range(n))
    Rr = b"".join(fr_b(rng.fr()) for _ in range(128)) * (n // 64)
    big = eng.batch_commit_g1(X, Rr)
    zero = eng.batch_commit_g1(X, bytes(64 * n))
    assert all(zero[192 * i:192 * i + 96] == bytes(96) and zero[192 * i + 96:192 * (i + 1)] == X[96 * i:96 * (i + 1)]
               for i in range(0, n, 97))
    small = eng.batch_commit_g1(X[:96 * 64], Rr[:64 * 64])          # 64 variables: c = 8 tables
    assert small == big[:192 * 64]
    assert big[192 * 64:192 * 128] == small                          # the inputs repeat with period 64
    q = _multiples_g2(eng, [rng.fr() for _ in range(64)])
    Y = b"".join(q[i % 64] for i in range(n))
    big2 = eng.batch_commit_g2(Y, Rr)
    small2 = eng.batch_commit_g2(Y[:192 * 64], Rr[:64 * 64])
    assert small2 == big2[:384 * 64] == big2[384 * 64:384 * 128]


@pytest.mark.parametrize("ty", [0, 1, 2, 3])
def test_c4_mixed_statement_types(eng, ty):
    """C4 shape (scaled): every equation type at m = n = 16, batch of 6 proofs with one tampered."""
    rng = SeededRng(40 + ty)
    m = n = 16
    rows = [_commit_prove(eng, ty, m, n, _instance(eng, ty, m, n, rng), rng) for _ in range(2)]
    count = 6
    cols = [[] for _ in range(8)]
    for i in range(count):
        r = list(rows[i % 2])
        if i == 4:
            r[5] = r[5][384:768] + r[5][:384] + r[5][768:]           # swap two y-commitments
        for c in range(8):
            cols[c].append(r[c])
    ok = eng.verify_batch(ty, count, m, n, *[b"".join(c) for c in cols])
    assert list(ok) == [1, 1, 1, 1, 0, 1]


@pytest.mark.parametrize("ty,m,n", [(0, 10, 7), (1, 6, 9), (2, 9, 6), (3, 5, 5)])
def test_sharded_statement_equals_single_gpu(eng, ty, m, n):
    """SURVEY.md §8e, one statement split by slot: for every world size the entry-wise product of the ranks'
    Miller partial products, finished with ONE final exponentiation, gives the verdicts of gs_verify_batch
    (honest / tampered), including world sizes that leave ranks without any slot."""
    rng = SeededRng(70 + ty)
    rows = [_commit_prove(eng, ty, m, n, _instance(eng, ty, m, n, rng), rng) for _ in range(2)]
    bad = list(rows[1])
    g = bytearray(bad[2])
    g[32 * (2 * n + 3) + 1] ^= 4                       # one bit of Gamma[2][3] of the second statement
    bad[2] = bytes(g)
    cols = [b"".join(c) for c in zip(rows[0], bad, rows[1])]
    count = 3
    want = eng.verify_batch(ty, count, m, n, *cols)
    assert list(want) == [1, 0, 1]
    K = n + (m if ty in (0, 2) else 1) + (2 if ty in (0, 1) else 1) + (2 if ty in (0, 2) else 1) + (0 if ty == 0 else 1)
    for world in (1, 2, 3, 8, K + 3):
        parts = b"".join(eng.verify_partial(ty, count, m, n, *cols, r, world) for r in range(world))
        assert len(parts) == world * count * 2304
        assert eng.verify_finish(ty, count, parts, cols[3]) == want, f"world={world}"
    # a rank that owns no slot contributes the GT identity
    from oracle.bls12_381 import FP12_ONE
    last = eng.verify_partial(ty, count, m, n, *cols, K + 2, K + 3)
    assert last == fp12_b(FP12_ONE) * (4 * count)
    with pytest.raises(Exception):
        eng.verify_partial(ty, count, m, n, *cols, 2, 2)


@pytest.mark.parametrize("ty", [0, 1, 2, 3])
def test_prove_batch_equals_single_proves(eng, ty):
    """gs_prove_batch (stream pool) == one gs_prove per equation, byte for byte: 9 equations over SHARED
    variables (the C4 shape), and 3 with per-proof variables."""
    from workloads import instance_many
    rng = SeededRng(90 + ty)
    m, n, E = 6, 5, 9
    A, B, G, T, X, Y, xr, yr, Tr = instance_many(eng, ty, m, n, E, rng)
    singles = [eng.prove(ty, m, n, A[e], B[e], G[e], X, Y, xr, yr, Tr[e]) for e in range(E)]
    pi, th = eng.prove_batch(ty, E, m, n, b"".join(A), b"".join(B), b"".join(G), X, Y, xr, yr, b"".join(Tr), shared_vars=True)
    assert pi == b"".join(s[0] for s in singles) and th == b"".join(s[1] for s in singles)
    # per-proof variables: three copies of the same witness arrays must give the first three proofs again
    pi3, th3 = eng.prove_batch(ty, 3, m, n, b"".join(A[:3]), b"".join(B[:3]), b"".join(G[:3]), X * 3, Y * 3, xr * 3, yr * 3,
                               b"".join(Tr[:3]), shared_vars=False)
    assert pi3 == b"".join(s[0] for s in singles[:3]) and th3 == b"".join(s[1] for s in singles[:3])
    # and the proofs verify against commitments of the shared variables
    xc = eng.batch_commit_g1(X, xr) if ty in (0, 1) else eng.batch_commit_scalar_b1(X, xr)
    yc = eng.batch_commit_g2(Y, yr) if ty in (0, 2) else eng.batch_commit_scalar_b2(Y, yr)
    ok = eng.verify_batch(ty, E, m, n, b"".join(A), b"".join(B), b"".join(G), b"".join(T), xc * E, yc * E, pi, th)
    assert ok == b"\x01" * E


@pytest.mark.parametrize("ty,m,n", [(0, 6, 340), (3, 5, 330), (1, 7, 325)])
def test_shared_base_window_tables_single_statement(eng, ty, m, n):
    """Statements with >= 320 MSM outputs per commitment take the shared-base window-table path of verify
    (k_wtab_* / k_vmsm_wsum): honest -> True, any single flipped scalar bit -> False, and the slot-sharded
    evaluation (where each rank is back on the per-problem Straus tables or on smaller table jobs) agrees."""
    rng = SeededRng(120 + ty)
    arrs = _commit_prove(eng, ty, m, n, _instance(eng, ty, m, n, rng), rng)
    assert eng.verify(ty, m, n, *arrs) is True
    for (i, j) in ((0, 0), (m - 1, n - 1), (m // 2, 17)):
        bad = list(arrs)
        g = bytearray(bad[2])
        g[32 * (i * n + j) + 3] ^= 0x10
        bad[2] = bytes(g)
        assert eng.verify(ty, m, n, *bad) is False
    for world in (2, 5):
        parts = b"".join(eng.verify_partial(ty, 1, m, n, *arrs, r, world) for r in range(world))
        assert eng.verify_finish(ty, 1, parts, arrs[3]) == b"\x01"


@pytest.mark.parametrize("ty", [0, 2, 3])
def test_shared_commitments_batch_uses_tables(eng, ty):
    """A batch of equations over ONE set of commitments (C4 shape): gs_verify_batch detects the shared x-commitments
    and builds the window tables once for the whole batch; verdicts equal those of the same equations verified
    one by one (per-problem Straus path)."""
    from workloads import instance_many
    rng = SeededRng(140 + ty)
    m, n, E = 7, 40, 9                                   # 9 x (40 [+1 +1]) outputs per base >= 320
    A, B, G, T, X, Y, xr, yr, Tr = instance_many(eng, ty, m, n, E, rng)
    pi, th = eng.prove_batch(ty, E, m, n, b"".join(A), b"".join(B), b"".join(G), X, Y, xr, yr, b"".join(Tr), shared_vars=True)
    xc = eng.batch_commit_g1(X, xr) if ty in (0, 1) else eng.batch_commit_scalar_b1(X, xr)
    yc = eng.batch_commit_g2(Y, yr) if ty in (0, 2) else eng.batch_commit_scalar_b2(Y, yr)
    Gb = [bytearray(g) for g in G]
    Gb[4][32 * (3 * n + 5)] ^= 2                         # equation 4 tampered
    cols = [b"".join(A), b"".join(B), b"".join(bytes(g) for g in Gb), b"".join(T), xc * E, yc * E, pi, th]
    ok = eng.verify_batch(ty, E, m, n, *cols)
    assert list(ok) == [1, 1, 1, 1, 0, 1, 1, 1, 1]
    cx, cy = (2 if ty in (0, 1) else 1), (2 if ty in (0, 2) else 1)
    for e in (0, 4):
        one = eng.verify(ty, m, n, A[e], B[e], bytes(Gb[e]), T[e], xc, yc, pi[e * cx * 384:(e + 1) * cx * 384],
                         th[e * cy * 192:(e + 1) * cy * 192])
        assert one is bool(ok[e])


def test_c2_chunked_two_stream_commit(eng):
    """Batches above 2^17 variables are committed chunk by chunk on two streams (copy / compute overlap): the
    result must be the concatenation of the results of its halves, each of which takes the single-pass path."""
    rng = SeededRng(222)
    n = (1 << 17) + 777
    base = _multiples_g1(eng, [rng.fr() for _ in range(32)])
    X = b"".join(base[(7 * i) % 32] for i in range(n))
    import numpy as np
    rs = np.random.RandomState(5)
    Rr = rs.randint(0, 2 ** 63 - 1, size=(2 * n, 4), dtype=np.int64).astype(np.uint64)
    Rr[:, 3] &= np.uint64((1 << 62) - 1)
    Rr = Rr.tobytes()
    whole = eng.batch_commit_g1(X, Rr)
    h = n // 2
    assert whole == eng.batch_commit_g1(X[:96 * h], Rr[:64 * h]) + eng.batch_commit_g1(X[96 * h:], Rr[64 * h:])
    q = _multiples_g2(eng, [rng.fr() for _ in range(8)])
    Y = b"".join(q[(3 * i) % 8] for i in range(n))
    whole2 = eng.batch_commit_g2(Y, Rr)
    assert whole2 == eng.batch_commit_g2(Y[:192 * h], Rr[:64 * h]) + eng.batch_commit_g2(Y[192 * h:], Rr[64 * h:])


def test_shared_commitments_wide_window_tables(eng):
    """>= 4,096 MSM outputs per shared commitment (70 equations x 64 columns): the c = 10 window tables (26 windows x
    512 entries) must give the verdicts of the per-equation path; one tampered Gamma entry is the only rejection."""
    from workloads import instance_many
    rng = SeededRng(160)
    ty, m, n, E = 0, 3, 64, 70
    A, B, G, T, X, Y, xr, yr, Tr = instance_many(eng, ty, m, n, E, rng)
    pi, th = eng.prove_batch(ty, E, m, n, b"".join(A), b"".join(B), b"".join(G), X, Y, xr, yr, b"".join(Tr), shared_vars=True)
    xc, yc = eng.batch_commit_g1(X, xr), eng.batch_commit_g2(Y, yr)
    Gb = [bytearray(g) for g in G]
    Gb[33][32 * (2 * n + 63) + 31] ^= 0x20                 # top byte of the last entry of equation 33
    cols = [b"".join(A), b"".join(B), b"".join(bytes(g) for g in Gb), b"".join(T), xc * E, yc * E, pi, th]
    ok = eng.verify_batch(ty, E, m, n, *cols)
    assert list(ok) == [0 if e == 33 else 1 for e in range(E)]
    for e in (0, 33, 69):
        one = eng.verify(ty, m, n, A[e], B[e], bytes(Gb[e]), T[e], xc, yc, pi[e * 768:(e + 1) * 768], th[e * 384:(e + 1) * 384])
        assert one is bool(ok[e])


@pytest.mark.parametrize("ty", [0, 1, 2, 3])
def test_prove_batch_shared_variable_tables(eng, ty):
    """>= 64 proof-element rows over one witness set: the variable terms of gs_prove_batch come from shared-base window
    tables (k_ptab_* / k_msm_var_terms_tab).  Proofs must equal the call-by-call ones byte for byte (identity witness
    included) and verify."""
    from workloads import instance_many
    rng = SeededRng(180 + ty)
    m, n, E = 4, 3, 40
    A, B, G, T, X, Y, xr, yr, Tr = instance_many(eng, ty, m, n, E, rng)
    if ty in (0, 2):                                     # an identity G2 witness: y_0 = O  =>  retarget every equation
        pass
    pi, th = eng.prove_batch(ty, E, m, n, b"".join(A), b"".join(B), b"".join(G), X, Y, xr, yr, b"".join(Tr), shared_vars=True)
    cx, cy = (2 if ty in (0, 1) else 1), (2 if ty in (0, 2) else 1)
    for e in (0, 17, 39):
        one = eng.prove(ty, m, n, A[e], B[e], G[e], X, Y, xr, yr, Tr[e])
        assert pi[e * cx * 384:(e + 1) * cx * 384] == one[0] and th[e * cy * 192:(e + 1) * cy * 192] == one[1], e
    xc = eng.batch_commit_g1(X, xr) if ty in (0, 1) else eng.batch_commit_scalar_b1(X, xr)
    yc = eng.batch_commit_g2(Y, yr) if ty in (0, 2) else eng.batch_commit_scalar_b2(Y, yr)
    ok = eng.verify_batch(ty, E, m, n, b"".join(A), b"".join(B), b"".join(G), b"".join(T), xc * E, yc * E, pi, th)
    assert ok == b"\x01" * E
    # witnesses that include the identity go through the same tables (prove only: byte equality with the single call)
    if ty in (0, 1):
        X2 = bytes(96) + X[96:]
        pi2, th2 = eng.prove_batch(ty, E, m, n, b"".join(A), b"".join(B), b"".join(G), X2, Y, xr, yr, b"".join(Tr), shared_vars=True)
        one = eng.prove(ty, m, n, A[5], B[5], G[5], X2, Y, xr, yr, Tr[5])
        assert pi2[5 * cx * 384:6 * cx * 384] == one[0] and th2[5 * cy * 192:6 * cy * 192] == one[1]


# ---------------------------------------------------------------- gs_verify_sharded: MSM split by base, pairs by slot
def _sharded_on_one_gpu(ty, count, m, n, arrays, crsb, world):
    """`world` ranks emulated by `world` contexts on THIS GPU in `world` host threads: the all-gather callback is an
    in-process exchange (device-to-device copies between the contexts' buffers behind a barrier), everything else is the
    code path the multi-GPU run takes.  Returns every rank's verdict bytes."""
    import threading
    import torch
    import groth_sahai_rs_b200 as gsb
    sh = gsb.shard
    dev = "cuda:0"
    engines = [gsb.Engine(0) for _ in range(world)]
    for e in engines:
        e.crs_load(crsb)
    bar = threading.Barrier(world)
    posted = [None] * world
    out, errs = [None] * world, []

    def run(r):
        def allgather(send_ptr, recv_ptr, nbytes):
            posted[r] = send_ptr
            bar.wait()
            recv = sh.wrap_memory(recv_ptr, nbytes * world, dev)
            for q in range(world):
                recv[q * nbytes:(q + 1) * nbytes].copy_(sh.wrap_memory(posted[q], nbytes, dev))
            torch.cuda.synchronize()
            bar.wait()
        try:
            a, b, gamma, target, xc, yc, pi, th = arrays
            rows = sh.gamma_rows_of(gamma, count, m, n, r, world)
            out[r] = engines[r].verify_sharded(ty, count, m, n, a, b, rows, target, xc, yc, pi, th, r, world, allgather)
        except Exception as ex:  # noqa: BLE001
            errs.append(ex)
            bar.abort()

    ts = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    for e in engines:
        e.close()
    if errs:
        raise errs[0]
    return out


@pytest.mark.parametrize("ty", [0, 1, 2, 3])
def test_verify_sharded_by_base_small(eng, ty):
    """Every equation type, ragged split (m = 5 bases + W1 over 1 / 2 / 3 / 4 ranks, some ranks own no Gamma row or no
    output), honest and tampered: the verdicts equal gs_verify_batch's and the CPU oracle's."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from bigcase import Case, cb, NT
    crs = make_crs(1)[0]
    m, n = 5, 3
    c = Case(ty, m, n, crs, seed=1200 + ty, zero_frac=0.2)
    good = c.verify_arrays()
    bad = list(good)
    g = bytearray(bad[2])
    g[32 * 4] ^= 1
    bad[2] = bytes(g)
    two = [x + y for x, y in zip(good, bad)]                     # count = 2: honest, tampered
    assert cb.verify_batch(ty, 2, m, n, two, c.crsb, NT) == b"\x01\x00"
    assert eng.verify_batch(ty, 2, m, n, *two) == b"\x01\x00"
    for world in (1, 2, 3, 4):
        for ok in _sharded_on_one_gpu(ty, 2, m, n, two, c.crsb, world):
            assert ok == b"\x01\x00", (ty, world, ok)


def test_verify_sharded_by_base_c3_and_c4():
    """The shapes it is meant for: one 256 x 192 PPE (shared-base window tables per rank) over 1 / 2 / 4 ranks, and a C4-style
    statement (48 equations over shared 64-variable commitments, tables shared by the batch) over 3 ranks, with CPU-made
    proofs and one tampered equation."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from bigcase import Case, Statement, cb
    crs = make_crs(1)[0]
    c = Case(0, 256, 192, crs, seed=1300)
    good = c.verify_arrays()
    bad = list(good)
    t = bytearray(bad[4])
    t[0:192], t[192:384] = t[192:384], t[0:192]                  # swap two x-commitments
    bad[4] = bytes(t)
    for world in (1, 2, 4):
        assert _sharded_on_one_gpu(0, 1, 256, 192, good, c.crsb, world) == [b"\x01"] * world
        assert _sharded_on_one_gpu(0, 1, 256, 192, bad, c.crsb, world) == [b"\x00"] * world
    for ty in (0, 3):
        E, m, n = 48, 64, 64
        st = Statement(ty, m, n, E, crs, seed=1310 + ty)
        arrays = st.verify_arrays()
        ts = cb.target_size(ty)
        tg = bytearray(arrays[3])
        tg[7 * ts:8 * ts], tg[8 * ts:9 * ts] = tg[8 * ts:9 * ts], tg[7 * ts:8 * ts]
        arrays[3] = bytes(tg)
        want = bytearray(b"\x01" * E)
        want[7] = want[8] = 0
        assert _sharded_on_one_gpu(ty, E, m, n, arrays, st.crsb, 3) == [bytes(want)] * 3
