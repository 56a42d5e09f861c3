#!/usr/bin/env python3
"""C3 (BASELINE.json configs[2]) over N GPUs: ONE PPE with m = n = 1024 and a dense Gamma, verified with the
slots of its pairing-product equation spread round-robin over the ranks (gs_verify_partial), the 4 x 576 B
Miller partial products all-gathered over NCCL, and one final exponentiation + comparison on every rank
(gs_verify_finish).  SURVEY.md §8e.  Strong scaling: the statement is fixed, N grows.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/bench_c3_sharded.py [--size 1024] [--steps 5] [--warmup 2]

Every rank builds the same seeded instance on its own GPU.  Timing: wall clock around the whole sharded call
(host buffers in, verdict out: H2D, kernels, D2H, NCCL all inside), bracketed by barriers, max over ranks.
Rank 0 prints one JSON line; with N = 1 the same code path runs without the collective.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    args = ap.parse_args()
    import torch
    import groth_sahai_rs_b200 as gsb
    from gsutil import SeededRng, make_crs, crs_bytes, fr_b
    from workloads import instance
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = gsb.Engine(local)
    crs, _ = make_crs(3)
    eng.crs_load(crs_bytes(crs))
    eng._crs = crs
    rng = SeededRng(3)
    m = n = args.size
    A, B, G, T, X, Y = instance(eng, 0, m, n, rng)
    xr = b"".join(fr_b(rng.fr()) for _ in range(2 * m))
    yr = b"".join(fr_b(rng.fr()) for _ in range(2 * n))
    Tr = b"".join(fr_b(rng.fr()) for _ in range(4))
    xc, yc = eng.batch_commit_g1(X, xr), eng.batch_commit_g2(Y, yr)
    pi, th = eng.prove(0, m, n, A, B, G, X, Y, xr, yr, Tr)
    arrs = [A, B, G, T, xc, yc, pi, th]
    g = bytearray(G)
    g[32 * (5 * n + 7)] ^= 1
    bad = list(arrs)
    bad[2] = bytes(g)
    dev = torch.device("cuda", local)

    def sharded_verify(a):
        mine = eng.verify_partial(0, 1, m, n, *a, rank, world)
        if world == 1:
            return eng.verify_finish(0, 1, mine, a[3])
        t = torch.frombuffer(bytearray(mine), dtype=torch.uint8).to(dev)
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        return eng.verify_finish(0, 1, torch.cat(parts).cpu().numpy().tobytes(), a[3])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    assert sharded_verify(arrs) == b"\x01", "honest statement rejected"
    assert sharded_verify(bad) == b"\x00", "tampered Gamma accepted"
    assert eng.verify(0, m, n, *arrs) is True            # the single-GPU path agrees
    for _ in range(args.warmup):
        sharded_verify(arrs)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        sharded_verify(arrs)
    barrier()
    dt = (time.perf_counter() - t0) / args.steps
    l0 = eng.launch_count
    eng.profile_enable(True)
    sharded_verify(arrs)
    prof = {k: [v[0], round(v[1], 3)] for k, v in eng.profile_read().items()}
    eng.profile_enable(False)
    launches = eng.launch_count - l0
    t1, _ = (lambda f: (min(f() for _ in range(3)), None))(lambda: _timeit(lambda: eng.verify(0, m, n, *arrs)))
    if world > 1:
        tt = torch.tensor([dt], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
    if rank == 0:
        pairs = 4 * n + 2 * m + 16
        print(json.dumps({
            "metric": "large_ppe_verifies_per_sec", "value": round(1.0 / dt, 3), "unit": "verifies/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt * 1e3, 3), "higher_is_better": True,
            "scaling": "strong", "dtype": "u32x12-montgomery", "data": "synthetic",
            "config": {"workload": f"C3: one PPE, m=n={m}, dense Gamma, verify sharded by slot (BASELINE.json configs[2])",
                       "parallelism": f"slots round-robin x{world}, one all_gather of 2,304 B per rank"},
            "pairings_per_sec": round(pairs / dt, 1), "single_gpu_unsharded_ms": round(t1 * 1e3, 3),
            "rank0_kernels": prof, "gpu_launches": int(launches),
            "parity": "honest -> 1, one flipped Gamma bit -> 0, equal to the unsharded verdict"}))
    if world > 1:
        dist.destroy_process_group()


def _timeit(fn):
    t0 = time.perf_counter()
    fn()
    return time.perf_counter() - t0


if __name__ == "__main__":
    main()
