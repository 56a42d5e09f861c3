#!/usr/bin/env python3
"""bench.py -- headline benchmark of the Groth-Sahai hot path on B200.

Workload (BASELINE.json configs[4], the one `metric` = "pairings/sec and PPE verifies/sec" is
quoted on): batch verification of independent 4x4 PPE proofs (m = n = 4 variables, dense random
Gamma, 1 % tampered).  One step = one pass of Verifiable::verify over `--proofs` proofs PER GPU
(default 65,536 = the whole of C5 on one GPU; weak scaling: every rank verifies its own shard,
verdict bitmaps are all-gathered over NCCL).  Synthetic data: `--distinct` distinct
(equation, proof) instances produced by this engine's own commit/prove path, tiled to the batch.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--proofs P] [--impl reference]

Prints ONE JSON line (see the task contract): value = verifies/s with inputs resident in HBM,
e2e = the same through the C ABI with host buffers (H2D + D2H inside the timed region),
roofline = integer-multiply roofline of the dominant kernel (measured live with CUDA events),
cpu_baseline = the oracle's C restatement of the reference algorithm timed on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

M_VARS = N_VARS = 4
# ---- algorithmic work model (Fp Montgomery products "M"; 1 M = 600 IMAD-equivalents) ----
IMAD_PER_M = 600
M_SQR12, M_LINE, M_014 = 36, 4, 39           # Fp12 squaring, line evaluation at P, sparse product (tower formulas)
M_G2_DBL, M_G2_ADD = 21, 37                  # one doubling / addition step of the projective G2 line walk
M_FE = 8300                                  # final exponentiation (easy ~750 + 5 exp-by-x + products)
M_G1_DBL, M_G1_MADD, M_FP_INV = 7, 11, 490


def work_model(m, n):
    """Algorithmic M per verified 4x4-shaped PPE proof, per kernel: the best sequential (tower / Karatsuba)
    operation counts of SURVEY.md §8d for what each kernel computes -- NOT the instructions it executes
    (the cooperative kernels trade Karatsuba for lazy-reduced schoolbook sums and execute more)."""
    cx = cy = 2
    pairs = 2 * (n + cx + cy) + 2 * (n + m + cx + cy)            # Miller pairs over the 4 ComT entries
    g2_points = 2 * n + m + 2 * cx                                 # G2 coordinates walked per proof (the CRS points
    fixed_pairs = 2 * 2 * cy                                       # v_1, v_2 have their lines stored at key load)
    return {
        "k_miller4": 4 * 62 * M_SQR12 + pairs * 68 * M_014,
        "k_g2_prepare4": g2_points * (63 * M_G2_DBL + 5 * M_G2_ADD) + (pairs - fixed_pairs) * 68 * M_LINE,
        "k_fixed_tiles": fixed_pairs * 68 * M_LINE,
        "k_final_exp3": 4 * M_FE,
        # Straus over the GLV halves, signed 4-bit windows: 32 windows x (4 shared doublings + one addition per
        # half-scalar, 15/16 non-zero)
        "k_vmsm_partial": 2 * n * (128 * M_G1_DBL + 32 * 2 * m * (15 / 16) * M_G1_MADD),
        "k_vmsm_tables": 2 * m * (M_G1_DBL + 6 * M_G1_MADD + 8 * 9),
        "k_vmsm_reduce": 2 * n * (M_G1_MADD + 9),
        "pairs": pairs,
    }


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons}


# ---------------------------------------------------------------- synthetic workload
def build_workload(eng, distinct, proofs, seed=5):
    """`distinct` satisfied 4x4 PPE instances proved by the engine itself, tiled to `proofs`; 1 % tampered.
    Returns (list of 8 numpy uint8 arrays in gs_verify_batch order, expected verdict array)."""
    import numpy as np
    from gsutil import SeededRng, make_crs, crs_bytes, fr_b, g1_b, g2_b, frmat_b
    from oracle.bls12_381 import R
    crs, _ = make_crs(seed)                      # six seeded draws; the CRS itself is recomputed on the GPU below
    eng.crs_load(crs_bytes(crs))
    rng = SeededRng(seed + 1)
    m, n = M_VARS, N_VARS
    g1b, g2b = g1_b(crs.g1_gen), g2_b(crs.g2_gen)
    D = distinct
    # witnesses / constants as multiples of the CRS generators (bench.rs:314 does the same), computed on the GPU
    sc = lambda k: [rng.fr() for _ in range(k)]
    xs, ys, a, b = sc(D * m), sc(D * n), sc(D * n), sc(D * m)
    gam = [rng.fr() for _ in range(D * m * n)]

    def g1_multiples(ks):   # k * g1 via the Mat kernel: (len x 1 Fr) * (1 x 1 Com1)
        out = eng.com1_matmul(len(ks), 1, 1, b"".join(fr_b(k) for k in ks), g1b + g1b)
        return [out[i * 192:i * 192 + 96] for i in range(len(ks))]

    def g2_multiples(ks):
        out = eng.com2_matmul(len(ks), 1, 1, b"".join(fr_b(k) for k in ks), g2b + g2b)
        return [out[i * 384:i * 384 + 192] for i in range(len(ks))]

    X, Y, A, B = g1_multiples(xs), g2_multiples(ys), g1_multiples(a), g2_multiples(b)
    # target = gt^val with val the equation "in the exponent": e(val * g1, g2)
    vals = []
    for d in range(D):
        v = sum(a[d * n + j] * ys[d * n + j] for j in range(n)) + sum(xs[d * m + i] * b[d * m + i] for i in range(m))
        v += sum(gam[(d * m + i) * n + j] * xs[d * m + i] * ys[d * n + j] for i in range(m) for j in range(n))
        vals.append(v % R)
    tg = eng.pairing(b"".join(g1_multiples(vals)), g2b * D)
    cols = [[] for _ in range(8)]
    for d in range(D):
        xr = b"".join(fr_b(rng.fr()) for _ in range(2 * m))
        yr = b"".join(fr_b(rng.fr()) for _ in range(2 * n))
        T = b"".join(fr_b(rng.fr()) for _ in range(4))
        xv, yv = b"".join(X[d * m:(d + 1) * m]), b"".join(Y[d * n:(d + 1) * n])
        av, bv = b"".join(A[d * n:(d + 1) * n]), b"".join(B[d * m:(d + 1) * m])
        gm = b"".join(fr_b(g) for g in gam[d * m * n:(d + 1) * m * n])
        xc = eng.batch_commit_g1(xv, xr)
        yc = eng.batch_commit_g2(yv, yr)
        pi, th = eng.prove(0, m, n, av, bv, gm, xv, yv, xr, yr, T)
        for c, v in enumerate([av, bv, gm, tg[d * 576:(d + 1) * 576], xc, yc, pi, th]):
            cols[c].append(np.frombuffer(v, dtype=np.uint8))
    base = [np.stack(c) for c in cols]                      # [D, bytes]
    idx = (np.arange(proofs, dtype=np.int64) * 7919 + 13) % D
    arrays = [np.ascontiguousarray(bs[idx]) for bs in base]
    expected = np.ones(proofs, dtype=np.uint8)
    bad = np.arange(37, proofs, 100)                        # 1 % tampered: swap pi[0] <-> pi[1]
    pi_arr = arrays[6]
    tmp = pi_arr[bad, :384].copy()
    pi_arr[bad, :384] = pi_arr[bad, 384:]
    pi_arr[bad, 384:] = tmp
    expected[bad] = 0
    return arrays, expected


# ---------------------------------------------------------------- reference arm (CPU)
def run_reference(args):
    """The reference's algorithm on the host cores: oracle C port (oracle/_ref does not exist: no Rust
    toolchain here), all host threads, a bounded sample of the same workload per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cbaseline
    res = cbaseline.time_ppe_verify(m=M_VARS, n=N_VARS, sample=args.ref_sample, steps=args.steps, warmup=args.warmup)
    line = {
        "metric": "ppe_verifies_per_sec", "value": res["verifies_per_sec"], "unit": "verifies/s", "impl": "reference",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64-montgomery", "data": "synthetic",
        "config": {"workload": "C5: independent 4x4 PPE proofs, Verifiable::verify (reference algorithm, CPU)",
                   "m": M_VARS, "n": N_VARS, "proofs_per_step": res["sample"]},
        "pairings_per_sec": res["verifies_per_sec"] * res["pairings_per_verify"],
        "cpu_baseline": {"value": res["verifies_per_sec"], "unit": "verifies/s", "cores": res["cores"], "kind": res["kind"],
                         "sample": res["sample_desc"]},
        "e2e": {"value": res["verifies_per_sec"], "unit": "verifies/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ---------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--proofs", type=int, default=65536, help="proofs per GPU per step")
    ap.add_argument("--distinct", type=int, default=64, help="distinct proved instances tiled to the batch")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ref-sample", type=int, default=0, help="proofs per CPU step (0 = auto, ~10-30 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import groth_sahai_rs_b200 as gsb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    eng = gsb.Engine(local)
    m, n, P = M_VARS, N_VARS, args.proofs
    arrays, expected = build_workload(eng, args.distinct, P, seed=5 + rank)
    h2d_bytes = int(sum(a.nbytes for a in arrays))
    d2h_bytes = P
    stream = torch.cuda.ExternalStream(eng.stream, device=local)
    host = [torch.from_numpy(a).pin_memory() for a in arrays]         # pinned host copies (e2e leg)
    dev = [h.to(f"cuda:{local}") for h in host]                        # resident in HBM (value leg)
    ok_dev = torch.zeros(P, dtype=torch.uint8, device=f"cuda:{local}")
    gathered = [torch.zeros_like(ok_dev) for _ in range(world)] if world > 1 else None
    torch.cuda.synchronize()

    def step_dev():
        eng.verify_batch_dev(0, P, m, n, [t.data_ptr() for t in dev], ok_dev.data_ptr())
        if world > 1:                                                   # verdict bitmaps over NCCL
            torch.cuda.current_stream().wait_stream(stream)
            dist.all_gather(gathered, ok_dev)

    def barrier():
        if world > 1:
            dist.barrier()
        stream.synchronize()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """K steps bracketed by barrier + synchronize; device time from CUDA events on the engine's stream."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        if world > 1:
            stream.wait_stream(torch.cuda.current_stream())
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=f"cuda:{local}")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- warm-up, correctness of the verdicts
    for _ in range(max(1, args.warmup)):
        step_dev()
    barrier()
    got = ok_dev.cpu().numpy()
    if not (got == expected).all():
        raise SystemExit(f"verdict mismatch on rank {rank}: {(got != expected).sum()} wrong of {P}")

    # ---- value: inputs resident in HBM   (working set ~ GBs of line coefficients >> 126 MB L2: no L2 flush needed)
    sampler = ClockSampler(local)
    sampler.start()
    l0 = eng.launch_count
    ms_total = timed(step_dev, args.steps)
    launches = eng.launch_count - l0
    clocks = sampler.stop()
    ms_per_step = ms_total / args.steps
    value = world * P / (ms_per_step * 1e-3)

    # ---- per-kernel device time (CUDA events inside the library, same stream), one extra profiled step
    eng.profile_enable(True)
    step_dev()
    prof = {k.split("<")[0]: v for k, v in eng.profile_read().items()}      # k_g2_prepare4<4> -> k_g2_prepare4
    eng.profile_enable(False)
    wm = work_model(m, n)
    peak_m = eng.fpmul_rate()                       # measured Fp products/s (register-only chain) on this GPU
    peak_imad = peak_m * IMAD_PER_M
    step_ms_prof = sum(v[1] for v in prof.values())
    kern = {}
    for name, (cnt, ms) in prof.items():
        w = wm.get(name)
        kern[name] = {"launches": cnt, "ms": round(ms, 3), "share": round(ms / step_ms_prof, 4)}
        if w:
            kern[name]["frac_of_imad_peak"] = round(w * P / (ms * 1e-3) / peak_m, 4)
    dom = max(prof.items(), key=lambda kv: kv[1][1])[0]
    dom_cnt, dom_ms = prof[dom]
    dom_achieved = wm[dom] * P * IMAD_PER_M / (dom_ms * 1e-3) if dom in wm else None
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")    # per-launch DRAM bytes from the committed ncu capture
    traffic_capture = None
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        traffic_capture = tj.get(dom)
        if traffic_capture:     # the capture ran 16,384 proofs per launch; scale linearly to THIS run's proofs per launch
            traffic = int(traffic_capture * (P / dom_cnt) / tj.get("proofs_per_launch", 16384))
    hbm = None
    try:                                                        # HBM view of the same kernel (north star: report GB/s for the
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))      # point-load phases): ncu DRAM bytes per launch,
        if traffic:                                             # (already scaled to this run's proofs per launch)
            gbs = traffic / (dom_ms / dom_cnt * 1e-3) / 1e9
            hbm = {"achieved_gbs": round(gbs, 1), "peak_gbs": peaks.get("hbm_gbs"), "frac": round(gbs / peaks["hbm_gbs"], 4),
                   "peak_source": "MEASURED_PEAKS.json (driver-measured copy bandwidth)"}
    except Exception:  # noqa: BLE001
        hbm = None
    roofline = {
        "bound": "imad", "kernel": dom, "achieved": dom_achieved and round(dom_achieved / 1e12, 3),
        "peak": round(peak_imad / 1e12, 3), "unit": "TIMAD/s", "frac": dom_achieved and round(dom_achieved / peak_imad, 4),
        "traffic": traffic, "traffic_source": "profiles/traffic.json (ncu --set full at 16,384 proofs per launch: %s B) scaled to this run's "
                                              "proofs per launch" % traffic_capture, "hbm": hbm,
        "peak_source": "measured live: register-only Fp Montgomery chain (gs_diag_fpmul_rate) x 600 IMAD/M; "
                       "MEASURED_PEAKS.json has no integer-pipe figure",
        "whole_step_frac": round(sum(wm[k] for k in wm if k != "pairs") * P / (ms_per_step * 1e-3) / peak_m, 4),
        "kernels": kern,
    }

    # ---- e2e: the public C-ABI call with HOST buffers (pinned), H2D and D2H inside the timed region
    ok_host = torch.zeros(P, dtype=torch.uint8).pin_memory()
    lib, vp = eng.lib, __import__("ctypes").c_void_p

    def step_e2e():
        rc = lib.gs_verify_batch(eng.h, 0, P, m, n, *[vp(h.data_ptr()) for h in host], vp(ok_host.data_ptr()))
        if rc != 0:
            raise SystemExit(lib.gs_last_error(eng.h).decode())

    step_e2e()
    if not (ok_host.numpy() == expected).all():
        raise SystemExit("e2e verdict mismatch")
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()                                  # synchronous: returns after the D2H of the verdicts
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    if world > 1:
        t = torch.tensor([e2e_ms], device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = world * P / (e2e_ms * 1e-3)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            from oracle import cbaseline
            r = cbaseline.time_ppe_verify(m=m, n=n, sample=0, steps=1, warmup=0)
            cpu = {"value": r["verifies_per_sec"], "unit": "verifies/s", "cores": r["cores"], "kind": r["kind"],
                   "sample": r["sample_desc"]}
        except Exception as ex:  # noqa: BLE001
            cpu = {"value": None, "unit": "verifies/s", "cores": 0, "kind": "port", "sample": f"unavailable: {ex}"}

    if rank == 0:
        line = {
            "metric": "ppe_verifies_per_sec", "value": round(value, 1), "unit": "verifies/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32x12-montgomery", "data": "synthetic",
            "config": {"workload": "C5: batch verification of independent 4x4 PPE proofs (BASELINE.json configs[4])",
                       "m": m, "n": n, "proofs_per_gpu": P, "distinct_instances": args.distinct, "tampered": "1%",
                       "l2": "inputs+line scratch (GBs) exceed the 126 MB L2; no flush needed",
                       "parallelism": f"proof-sharded x{world}, verdict all_gather"},
            "pairings_per_sec": round(value * wm["pairs"], 1),
            "miller_pairs_per_verify": wm["pairs"], "final_exps_per_verify": 4,
            "e2e": {"value": round(e2e_value, 1), "unit": "verifies/s", "ms_per_step": round(e2e_ms, 3),
                    "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes},
            "gpu_launches": int(launches), "roofline": roofline, "clocks": clocks,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
