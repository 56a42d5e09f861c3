"""Byte parity of the GPU path against the C restatement of the reference (oracle/gs_oracle.c: gsref_prove,
gsref_verify, scalar commits -- reference evaluation order, checked against the big-int oracle in
tests/test_c_oracle.py) AT THE BASELINE.json SHAPES: 4x4 (C1 / C5), 64x64 single and as a shared-variable
statement (C4: the `k_ptab_*` / `k_msm_var_terms_tab` table path), an unshared batch big enough for the unsplit
`k_msm_terms`, 128x128 and one 1024x1024 PPE (C3), and 2^14 commitments (C2: the c = 16 table path).
Inputs, commitments and proofs are built on the CPU only (tests/bigcase.py); the GPU results must equal them."""
from concurrent.futures import ThreadPoolExecutor

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


def gpu_commit_x(eng, ty, X, xr):
    return eng.batch_commit_g1(X, xr) if ty in (0, 1) else eng.batch_commit_scalar_b1(X, xr)


def gpu_commit_y(eng, ty, Y, yr):
    return eng.batch_commit_g2(Y, yr) if ty in (0, 2) else eng.batch_commit_scalar_b2(Y, yr)


def flip_theta(arrays):
    bad = list(arrays)
    th = bytearray(bad[7])
    th[0:96], th[96:192] = th[96:192], th[0:96]          # swap the two coordinates of theta_0
    bad[7] = bytes(th)
    return bad


def check_case(eng, c, oracle_verify=True):
    """commit / prove bytes == reference-order CPU bytes; verify accepts the CPU-made proof, rejects tampered ones."""
    ty, m, n = c.ty, c.m, c.n
    assert gpu_commit_x(eng, ty, c.X, c.xr) == c.xc, "x commitments differ from the reference-order CPU result"
    assert gpu_commit_y(eng, ty, c.Y, c.yr) == c.yc, "y commitments differ"
    pi, th = eng.prove(ty, m, n, *c.prove_args())
    assert pi == c.pi, "pi differs"
    assert th == c.theta, "theta differs"
    arrays = c.verify_arrays()
    assert eng.verify(ty, m, n, *arrays) is True
    bad = flip_theta(arrays)
    assert eng.verify(ty, m, n, *bad) is False
    g = bytearray(arrays[2])
    g[32 * (n + 1 if m > 1 else 0)] ^= 1                  # one bit of one Gamma entry
    bad2 = list(arrays)
    bad2[2] = bytes(g)
    assert eng.verify(ty, m, n, *bad2) is False
    if oracle_verify:                                     # the checker agrees on all three verdicts
        assert cb.verify(ty, m, n, arrays, c.crsb, NT) is True
        assert cb.verify(ty, m, n, bad, c.crsb, NT) is False


@pytest.mark.parametrize("ty", [0, 1, 2, 3])
def test_4x4_all_types(eng, crs, ty):
    """C1 / C5 shape, with identity constants and zero Gamma entries."""
    check_case(eng, Case(ty, 4, 4, crs, seed=100 + ty, zero_frac=0.2))
    check_case(eng, Case(ty, 4, 4, crs, seed=110 + ty))


@pytest.mark.parametrize("ty", [0, 1, 2, 3])
def test_64x64_all_types(eng, crs, ty):
    check_case(eng, Case(ty, 64, 64, crs, seed=200 + ty))


def test_ragged_shapes(eng, crs):
    """m != n, m = 1, n = 1 (col_vec_to_vec's single-row case, data_structures.rs:145-151)."""
    for ty, m, n in ((0, 1, 1), (0, 7, 3), (1, 1, 5), (2, 6, 1), (3, 5, 9)):
        check_case(eng, Case(ty, m, n, crs, seed=300 + 10 * ty + m))


def test_128x128_ppe(eng, crs):
    check_case(eng, Case(0, 128, 128, crs, seed=400))


def test_1024x1024_ppe(eng, crs):
    """C3 at full size: the reference-order CPU proof (~8,200 scalar multiplications per group) equals the GPU's; the
    GPU verifier accepts it and rejects a one-bit change of Gamma.  (The CPU *verifier* would need 2.1 M G2 scalar
    multiplications here, so the verdict check is GPU-only at this size; it is cross-checked with the CPU at 128x128.)"""
    check_case(eng, Case(0, 1024, 1024, crs, seed=500), oracle_verify=False)


@pytest.mark.parametrize("ty,m,n", [(0, 2, 8190), (1, 4100, 3), (2, 3, 4100)])
def test_pippenger_proof_msm(eng, crs, ty, m, n):
    """One statement whose proof MSMs have >= 4,096 terms takes the bucket method (csrc/pippenger.cuh): the proof must
    still equal the reference-order CPU proof byte for byte.  Colliding terms (P + P, P + (-P) in one bucket), identity
    constants, zero / 1 / r-1 scalars included."""
    c = Case(ty, m, n, crs, seed=900 + ty, zero_frac=0.05, collide=True)
    pi, th = eng.prove(ty, m, n, *c.prove_args())
    assert pi == c.pi, "pi differs"
    assert th == c.theta, "theta differs"
    assert eng.verify(ty, m, n, *c.verify_arrays()) is True


def test_pippenger_equals_per_term_path(crs):
    """The same big proof through both MSM paths (GS_PIP_MIN decides at context creation), and at several window widths."""
    import os
    import groth_sahai_rs_b200 as gsb
    c = Case(0, 2, 5000, crs, seed=950, prove=False, collide=True)
    outs = []
    for env in ({"GS_PIP_MIN": "1000000000"}, {"GS_PIP_MIN": "0"}, {"GS_PIP_MIN": "0", "GS_PIP_C": "5"},
                {"GS_PIP_MIN": "0", "GS_PIP_C": "13"}):
        old = {k: os.environ.get(k) for k in ("GS_PIP_MIN", "GS_PIP_C")}
        os.environ.update(env)
        try:
            e = gsb.Engine(0)
        finally:
            for k, v in old.items():
                os.environ.pop(k, None) if v is None else os.environ.__setitem__(k, v)
        e.crs_load(c.crsb)
        outs.append(e.prove(0, 2, 5000, *c.prove_args()))
        e.close()
    assert outs[0] == outs[1] == outs[2] == outs[3]


@pytest.mark.parametrize("ty", [0, 1, 2, 3])
def test_c4_statement_shared_vars(eng, crs, ty):
    """64 equations over one witness set (C4 shape): gs_prove_batch(shared_vars) takes the shared-base window-table
    path; every proof must equal the CPU's, and gs_verify_batch over the shared commitments gives the exact mask."""
    E, m, n = 64, 64, 64
    st = Statement(ty, m, n, E, crs, seed=600 + ty)
    assert gpu_commit_x(eng, ty, st.X, st.xr) == st.xc
    assert gpu_commit_y(eng, ty, st.Y, st.yr) == st.yc
    pi, th = eng.prove_batch(ty, E, m, n, st.A, st.B, st.G, st.X, st.Y, st.xr, st.yr, st.Tr, shared_vars=True)
    assert pi == st.pi and th == st.theta
    arrays = st.verify_arrays()
    tsz = cb.target_size(ty)
    t = bytearray(arrays[3])
    t[5 * tsz:6 * tsz], t[9 * tsz:10 * tsz] = t[9 * tsz:10 * tsz], t[5 * tsz:6 * tsz]   # swap the targets of equations 5 and 9
    arrays[3] = bytes(t)
    expect = bytearray(b"\x01" * E)
    expect[5] = expect[9] = 0
    assert eng.verify_batch(ty, E, m, n, *arrays) == bytes(expect)
    # the CPU verifier on a sample of the same batch (honest 4, 6 and tampered 5)
    sizes = [n * cb.x_size(ty), m * cb.y_size(ty), m * n * 32, tsz, m * 192, n * 384, cb.cx_of(ty) * 384, cb.cy_of(ty) * 192]
    sub = [a[4 * s:7 * s] for a, s in zip(arrays, sizes)]
    assert cb.verify_batch(ty, 3, m, n, sub, st.crsb, NT) == b"\x01\x00\x01"


def test_unshared_batch_unsplit_terms(eng, crs):
    """128 independent 64x64 PPE proofs in one gs_prove_batch (>= 32,768 MSM terms: one thread per whole scalar);
    a sample of them is byte-compared with the CPU."""
    count, m, n = 128, 64, 64
    base = Case(0, m, n, crs, seed=700)
    rng = SeededRng(701)
    Tr = [frs_b(fr_list(rng, 4)) for _ in range(count)]
    xr = [frs_b(fr_list(rng, 2 * m)) for _ in range(count)]
    yr = [frs_b(fr_list(rng, 2 * n)) for _ in range(count)]
    pi, th = eng.prove_batch(0, count, m, n, base.A * count, base.B * count, base.G * count, base.X * count, base.Y * count,
                             b"".join(xr), b"".join(yr), b"".join(Tr), shared_vars=False)
    for k in (0, 1, 63, 127):
        p, t = cb.prove(0, m, n, base.A, base.B, base.G, base.X, base.Y, xr[k], yr[k], Tr[k], base.crsb, NT)
        assert pi[k * 768:(k + 1) * 768] == p and th[k * 384:(k + 1) * 384] == t, f"proof {k} differs"


def test_c2_commit_2p14(eng, crs):
    """2^14 + 2^14 batch commitments (>= 8,192: the c = 16 fixed-base tables) against batch_commit_G1 / G2 in the
    reference's term-by-term order (CPU, chunks spread over host threads)."""
    n = 1 << 14
    rng = SeededRng(800)
    crsb = crs_bytes(crs)
    ks = fr_list(rng, n)
    ks[3] = 0                                            # an identity variable
    X = b"".join(g1_multiples(crs, ks))
    Y = b"".join(g2_multiples(crs, ks))
    rr = fr_list(rng, 2 * n)
    rr[0] = rr[1] = 0                                    # zero randomness
    rr[2], rr[3] = 1, R - 1
    rand = frs_b(rr)
    ch = n // (4 * NT) or 1
    with ThreadPoolExecutor(NT) as ex:
        e1 = b"".join(ex.map(lambda o: cb.batch_commit_g1(X[96 * o:96 * (o + ch)], rand[64 * o:64 * (o + ch)], crsb), range(0, n, ch)))
        e2 = b"".join(ex.map(lambda o: cb.batch_commit_g2(Y[192 * o:192 * (o + ch)], rand[64 * o:64 * (o + ch)], crsb), range(0, n, ch)))
    assert eng.batch_commit_g1(X, rand) == e1
    assert eng.batch_commit_g2(Y, rand) == e2
    # back on the c = 8 tables: small batches give the same bytes
    assert eng.batch_commit_g1(X[:96 * 5], rand[:64 * 5]) == e1[:192 * 5]
