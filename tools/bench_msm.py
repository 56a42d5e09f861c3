#!/usr/bin/env python3
"""Proof-element MSM: the bucket method (csrc/pippenger.cuh) against the per-term windowed scalar multiplications, on ONE
statement whose pi / theta MSMs have N terms (PPE with m = 2, n = N - 2).  Same inputs through two contexts
(GS_PIP_MIN = 0 / huge); outputs must be byte-equal; prints one JSON line per N with the per-kernel CUDA-event times.

    python tools/bench_msm.py [--sizes 2050,16384,65536] [--reps 3]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402


def rand_fr(n, seed):
    rs = np.random.RandomState(seed)
    a = rs.randint(0, 2 ** 63 - 1, size=(n, 4), dtype=np.int64).astype(np.uint64)
    a[:, 3] &= np.uint64((1 << 62) - 1)
    return a.tobytes()


def engine(pip_min, c=None):
    import groth_sahai_rs_b200 as gsb
    os.environ["GS_PIP_MIN"] = str(pip_min)
    if c:
        os.environ["GS_PIP_C"] = str(c)
    else:
        os.environ.pop("GS_PIP_C", None)
    return gsb.Engine(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="2050,4096,16384,65536")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    from gsutil import make_crs, crs_bytes, g1_b, g2_b
    crs, _ = make_crs(7)
    crsb = crs_bytes(crs)
    term, pip = engine(10 ** 12), engine(0)
    for e in (term, pip):
        e.crs_load(crsb)
    g1, g2 = g1_b(crs.g1_gen), g2_b(crs.g2_gen)
    for N in [int(x) for x in args.sizes.split(",")]:
        m, n = 2, N - 2

        def pts(eng, k, seed, g2side):
            out = []
            for o in range(0, k, 1 << 16):
                cnt = min(1 << 16, k - o)
                ks = rand_fr(cnt, seed + o)
                if g2side:
                    r = np.frombuffer(eng.com2_matmul(cnt, 1, 1, ks, g2 + g2), dtype=np.uint8).reshape(cnt, 384)[:, :192]
                else:
                    r = np.frombuffer(eng.com1_matmul(cnt, 1, 1, ks, g1 + g1), dtype=np.uint8).reshape(cnt, 192)[:, :96]
                out.append(r)
            return np.ascontiguousarray(np.concatenate(out)).tobytes()

        A, X = pts(term, n, 11, False), pts(term, m, 12, False)
        B, Y = pts(term, m, 13, True), pts(term, n, 14, True)
        G, xr, yr, Tr = rand_fr(m * n, 15), rand_fr(2 * m, 16), rand_fr(2 * n, 17), rand_fr(4, 18)
        argv = (0, m, n, A, B, G, X, Y, xr, yr, Tr)
        res = {}
        outs = {}
        for name, eng in (("per_term", term), ("pippenger", pip)):
            outs[name] = eng.prove(*argv)
            best = None
            for _ in range(args.reps):
                t0 = time.perf_counter()
                eng.prove(*argv)
                dt = time.perf_counter() - t0
                best = dt if best is None or dt < best else best
            eng.profile_enable(True)
            eng.prove(*argv)
            prof = {k: [v[0], round(v[1], 3)] for k, v in eng.profile_read().items()}
            eng.profile_enable(False)
            res[name] = {"prove_ms_e2e": round(best * 1e3, 3), "kernel_ms": round(sum(v[1] for v in prof.values()), 3), "kernels": prof}
        assert outs["per_term"] == outs["pippenger"], f"N={N}: the two MSM paths disagree"
        print(json.dumps({"config": f"proof MSMs of one PPE statement, N = {N} terms per row (m = 2, n = {n}), 2 rows each on G1 and G2",
                          "N": N, "per_term": res["per_term"], "pippenger": res["pippenger"],
                          "speedup_kernels": round(res["per_term"]["kernel_ms"] / res["pippenger"]["kernel_ms"], 2),
                          "parity": "pi / theta byte-equal between the two paths"}), flush=True)


if __name__ == "__main__":
    main()
