#!/usr/bin/env python3
"""Summarise the `ncu --page source --csv` dump of one kernel: opcode mix and stall samples by opcode.
usage: tools/ncu_source_summary.py <source.csv.gz> <kernel-substring>"""
import csv, gzip, sys, collections
path, want = sys.argv[1], sys.argv[2]
rows = csv.reader(gzip.open(path, "rt"))
cur = None; hdr = None
ops = collections.Counter(); stall = collections.Counter(); reasons = collections.Counter(); samp_total = 0
for r in rows:
    if not r: continue
    if r[0] == "Kernel Name":
        cur = r[1]; hdr = None; continue
    if r[0] == "Address":
        hdr = r; continue
    if cur is None or want not in cur or hdr is None: continue
    d = dict(zip(hdr, r))
    sass = d["Source"].split()
    if not sass: continue
    op = sass[0] if not sass[0].startswith("@") else sass[1]
    op = ".".join(op.split(".")[:3])
    n = int(d["Instructions Executed"] or 0)
    s = int(d["# Samples"] or 0)
    ops[op] += n; stall[op] += s; samp_total += s
    for k, v in d.items():
        if k.startswith("stall_") and "Not Issued" not in k and v and v != "0":
            reasons[k] += int(v)
tot = sum(ops.values())
print(f"kernel ~ {want}: {tot:.3e} warp instructions, {samp_total} samples")
for op, n in ops.most_common(18):
    print(f"  {op:28s} {n:14d} {100*n/tot:6.2f}%   samples {100*stall[op]/max(1,samp_total):6.2f}%")
print("stall reasons (all samples):")
for k, v in reasons.most_common(10):
    print(f"  {k:24s} {100*v/max(1,samp_total):6.2f}%")
