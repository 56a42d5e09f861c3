"""ctypes wrapper of oracle/libgs_oracle.so (the C restatement): checker for larger parity cases
and the CPU baseline of bench.py.  TEST INFRASTRUCTURE ONLY -- never imported by the product."""
import ctypes
import os
import subprocess
import time

HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(HERE, "libgs_oracle.so")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(os.path.join(HERE, "gs_oracle.c")):
            subprocess.check_call(["make", "-s", "-C", HERE])
        _lib = ctypes.CDLL(so)
        _lib.gsref_verify_ppe_batch.argtypes = [ctypes.c_size_t, ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 10 + [ctypes.c_int]
        _lib.gsref_batch_commit_g1.argtypes = [ctypes.c_size_t] + [ctypes.c_void_p] * 4
        _lib.gsref_batch_commit_g2.argtypes = [ctypes.c_size_t] + [ctypes.c_void_p] * 4
        V, Z, I = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
        _lib.gsref_batch_commit_scalar_b1.argtypes = [Z, V, V, V, V, I]
        _lib.gsref_batch_commit_scalar_b2.argtypes = [Z, V, V, V, V, I]
        _lib.gsref_prove.argtypes = [I, Z, Z] + [V] * 11 + [I]
        _lib.gsref_prove_batch.argtypes = [I, Z, Z, Z] + [V] * 8 + [I, V, V, V, I]
        _lib.gsref_verify.argtypes = [I, Z, Z] + [V] * 9 + [I]
        _lib.gsref_verify.restype = I
        _lib.gsref_verify_batch.argtypes = [I, Z, Z, Z] + [V] * 10 + [I]
        _lib.gsref_g1_mul_batch.argtypes = [Z, V, V, V, I]
        _lib.gsref_g2_mul_batch.argtypes = [Z, V, V, V, I]
        _lib.gsref_selftest_inv.argtypes = [V]
    return _lib


def _out(n):
    return ctypes.create_string_buffer(n)


def pairing(p: bytes, q: bytes) -> bytes:
    o = _out(576)
    lib().gsref_pairing(p, q, o)
    return o.raw


def pairing_sum(xs: bytes, ys: bytes) -> bytes:
    k = len(xs) // 192
    o = _out(2304)
    lib().gsref_pairing_sum(k, xs, ys, o)
    return o.raw


def g1_mul(p: bytes, k: bytes) -> bytes:
    o = _out(96)
    lib().gsref_g1_mul(p, k, o)
    return o.raw


def g2_mul(p: bytes, k: bytes) -> bytes:
    o = _out(192)
    lib().gsref_g2_mul(p, k, o)
    return o.raw


def batch_commit_g1(xvars: bytes, rand: bytes, crs: bytes) -> bytes:
    n = len(xvars) // 96
    o = _out(max(1, n * 192))
    lib().gsref_batch_commit_g1(n, xvars, rand, crs, o)
    return o.raw[:n * 192]


def batch_commit_g2(yvars: bytes, rand: bytes, crs: bytes) -> bytes:
    n = len(yvars) // 192
    o = _out(max(1, n * 384))
    lib().gsref_batch_commit_g2(n, yvars, rand, crs, o)
    return o.raw[:n * 384]


def verify_ppe_batch(count, m, n, arrays, crs: bytes, nthreads=1) -> bytes:
    """arrays = the 8 byte strings of gs_verify_batch (a, b, gamma, target, xcoms, ycoms, pi, theta)."""
    ok = _out(max(1, count))
    keep = [bytes(a) for a in arrays]
    lib().gsref_verify_ppe_batch(count, m, n, *keep, crs, ok, nthreads)
    return ok.raw[:count]


def x_is_group(ty): return ty in (0, 1)
def y_is_group(ty): return ty in (0, 2)
def x_size(ty): return 96 if x_is_group(ty) else 32
def y_size(ty): return 192 if y_is_group(ty) else 32
def cx_of(ty): return 2 if x_is_group(ty) else 1
def cy_of(ty): return 2 if y_is_group(ty) else 1
def target_size(ty): return (576, 96, 192, 32)[ty]


def batch_commit_scalar_b1(xs: bytes, rand: bytes, crs: bytes, nthreads=1) -> bytes:
    n = len(xs) // 32
    o = _out(max(1, n * 192))
    lib().gsref_batch_commit_scalar_b1(n, xs, rand, crs, o, nthreads)
    return o.raw[:n * 192]


def batch_commit_scalar_b2(ys: bytes, rand: bytes, crs: bytes, nthreads=1) -> bytes:
    n = len(ys) // 32
    o = _out(max(1, n * 384))
    lib().gsref_batch_commit_scalar_b2(n, ys, rand, crs, o, nthreads)
    return o.raw[:n * 384]


def commit_x(ty, xvars: bytes, rand: bytes, crs: bytes, nthreads=1) -> bytes:
    """batch_commit_G1 or batch_commit_scalar_to_B1 by equation type (what commit_and_prove calls first)."""
    return batch_commit_g1(xvars, rand, crs) if x_is_group(ty) else batch_commit_scalar_b1(xvars, rand, crs, nthreads)


def commit_y(ty, yvars: bytes, rand: bytes, crs: bytes, nthreads=1) -> bytes:
    return batch_commit_g2(yvars, rand, crs) if y_is_group(ty) else batch_commit_scalar_b2(yvars, rand, crs, nthreads)


def prove(ty, m, n, a, b, gamma, xvars, yvars, x_rand, y_rand, pf_rand, crs: bytes, nthreads=1):
    """Provable::prove in the reference's evaluation order (argument layout of gs_prove) -> (pi, theta) bytes."""
    assert len(a) == n * x_size(ty) and len(b) == m * y_size(ty) and len(gamma) == m * n * 32
    assert len(xvars) == m * x_size(ty) and len(yvars) == n * y_size(ty)
    assert len(x_rand) == m * cx_of(ty) * 32 and len(y_rand) == n * cy_of(ty) * 32
    assert len(pf_rand) == cx_of(ty) * cy_of(ty) * 32
    pi, th = _out(cx_of(ty) * 384), _out(cy_of(ty) * 192)
    lib().gsref_prove(ty, m, n, a, b, gamma, xvars, yvars, x_rand, y_rand, pf_rand, crs, pi, th, nthreads)
    return pi.raw, th.raw


def prove_batch(ty, count, m, n, a, b, gamma, xvars, yvars, x_rand, y_rand, pf_rand, shared_vars, crs: bytes, nthreads=1):
    """`count` proofs, array layout of gs_prove_batch."""
    nv = 1 if shared_vars else count
    assert len(a) == count * n * x_size(ty) and len(b) == count * m * y_size(ty) and len(gamma) == count * m * n * 32
    assert len(xvars) == nv * m * x_size(ty) and len(yvars) == nv * n * y_size(ty)
    assert len(x_rand) == nv * m * cx_of(ty) * 32 and len(y_rand) == nv * n * cy_of(ty) * 32
    assert len(pf_rand) == count * cx_of(ty) * cy_of(ty) * 32
    pi, th = _out(count * cx_of(ty) * 384), _out(count * cy_of(ty) * 192)
    lib().gsref_prove_batch(ty, count, m, n, a, b, gamma, xvars, yvars, x_rand, y_rand, pf_rand, int(bool(shared_vars)),
                            crs, pi, th, nthreads)
    return pi.raw, th.raw


def verify(ty, m, n, arrays, crs: bytes, nthreads=1) -> bool:
    """Verifiable::verify for one instance; arrays = the 8 byte strings of gs_verify_batch."""
    a, b, gamma, target, xc, yc, pi, th = [bytes(x) for x in arrays]
    assert len(a) == n * x_size(ty) and len(b) == m * y_size(ty) and len(gamma) == m * n * 32
    assert len(target) == target_size(ty) and len(xc) == m * 192 and len(yc) == n * 384
    assert len(pi) == cx_of(ty) * 384 and len(th) == cy_of(ty) * 192
    return bool(lib().gsref_verify(ty, m, n, a, b, gamma, target, xc, yc, pi, th, crs, nthreads))


def verify_batch(ty, count, m, n, arrays, crs: bytes, nthreads=1) -> bytes:
    keep = [bytes(x) for x in arrays]
    sizes = [n * x_size(ty), m * y_size(ty), m * n * 32, target_size(ty), m * 192, n * 384, cx_of(ty) * 384, cy_of(ty) * 192]
    assert all(len(k) == count * s for k, s in zip(keep, sizes))
    ok = _out(max(1, count))
    lib().gsref_verify_batch(ty, count, m, n, *keep, crs, ok, nthreads)
    return ok.raw[:count]


def g1_mul_batch(base: bytes, ks: bytes, nthreads=1) -> bytes:
    n = len(ks) // 32
    o = _out(max(1, n * 96))
    lib().gsref_g1_mul_batch(n, base, ks, o, nthreads)
    return o.raw[:n * 96]


def g2_mul_batch(base: bytes, ks: bytes, nthreads=1) -> bytes:
    n = len(ks) // 32
    o = _out(max(1, n * 192))
    lib().gsref_g2_mul_batch(n, base, ks, o, nthreads)
    return o.raw[:n * 192]


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def _workload(m, n, distinct, seed=5):
    """`distinct` satisfied PPE instances built with the ORACLE itself (no GPU involved)."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tests"))
    from gsutil import SeededRng, make_crs, crs_bytes, random_instance, draw_rands, proof_bytes
    from . import gs as ogs
    crs, _ = make_crs(seed)
    rng = SeededRng(seed + 1)
    rows = []
    for _ in range(distinct):
        equ, xv, yv = random_instance(0, m, n, crs, rng)
        xr, yr, T = draw_rands(0, m, n, rng)
        rows.append(proof_bytes(0, equ, ogs.commit_and_prove(equ, xv, yv, crs, xr, yr, T)))
    return crs_bytes(crs), rows


def time_ppe_verify(m=4, n=4, sample=0, steps=1, warmup=0):
    """PPE::verify (reference algorithm, C port) on all host cores over a bounded sample of the C5 workload."""
    cores = host_cores()
    crsb, rows = _workload(m, n, distinct=2)
    # calibrate: one verification on one core
    t0 = time.perf_counter()
    ok = verify_ppe_batch(1, m, n, rows[0], crsb, 1)
    t_one = time.perf_counter() - t0
    assert ok == b"\x01", "C oracle rejected an honest proof"
    if sample <= 0:   # ~10-20 s of CPU work per step, a multiple of the core count
        sample = max(cores, int(15.0 / t_one) // cores * cores)
        sample = min(sample, 64 * cores)
    arrays = [b"".join(rows[i % len(rows)][c] for i in range(sample)) for c in range(8)]
    for _ in range(warmup):
        verify_ppe_batch(min(sample, cores), m, n, [a[:len(a) // sample * min(sample, cores)] for a in arrays], crsb, cores)
    t0 = time.perf_counter()
    for _ in range(steps):
        ok = verify_ppe_batch(sample, m, n, arrays, crsb, cores)
    dt = (time.perf_counter() - t0) / steps
    assert ok == b"\x01" * sample
    pairs_ref = 2 * n + 2 * m + 4 * m + 16   # non-trivial Miller loops the reference runs per verify
    return {"verifies_per_sec": sample / dt, "ms_per_step": dt * 1e3, "sample": sample, "cores": cores, "kind": "port",
            "pairings_per_verify": pairs_ref, "single_core_verify_ms": t_one * 1e3,
            "sample_desc": f"{sample} PPE 4x4 verifications per step, reference algorithm (20 final exps, Gamma*d as m*n G2 "
                           f"scalar muls), C restatement (not arkworks), proofs spread over {cores} host threads"}
