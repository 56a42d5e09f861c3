#!/usr/bin/env python3
"""Tiny end-to-end run for compute-sanitizer (memcheck / racecheck): CRS load, commit, prove, verify (single and a
batch of 40 that takes the throughput kernels), pairing_sum, group ops, (de)compression.  Usage:
    compute-sanitizer --tool memcheck  python tools/sanitize_smoke.py
    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g  # noqa: E402

g.smoke()
import groth_sahai_rs_b200 as gsb  # noqa: E402
from gsutil import *  # noqa: E402,F401,F403
from workloads import instance, commit_prove  # noqa: E402

eng = gsb.Engine(0)
crs, _ = make_crs(1)
eng.crs_load(crs_bytes(crs))
eng._crs = crs
rng = SeededRng(5)
for ty in range(4):
    rows = [commit_prove(eng, ty, 3, 2, instance(eng, ty, 3, 2, rng), rng) for _ in range(2)]
    cols = [b"".join(rows[i % 2][c] for i in range(40)) for c in range(8)]
    ok = eng.verify_batch(ty, 40, 3, 2, *cols)
    assert ok == b"\x01" * 40, ok
    parts = b"".join(eng.verify_partial(ty, 2, 3, 2, *[b"".join(r[c] for r in rows) for c in range(8)], r_, 3) for r_ in range(3))
    assert eng.verify_finish(ty, 2, parts, b"".join(r[3] for r in rows)) == b"\x01\x01"
pts = b"".join(g1_b(rng.g1()) for _ in range(5))
w = eng.serialize("g1", pts)
assert eng.deserialize("g1", w) == (pts, b"\x01" * 5)
q = b"".join(g2_b(rng.g2()) for _ in range(3))
assert eng.deserialize("g2", eng.serialize("g2", q)) == (q, b"\x01" * 3)
a = com1_b((rng.g1(), rng.g1()))
assert eng.elementwise("com1", "sub", a, a) == bytes(192)
print("sanitize_smoke ok")
