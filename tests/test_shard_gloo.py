"""world_size-2 (and 3) `gloo` test of the multi-GPU host logic: contiguous proof shards,
no data-path collective, verdict bytes all-gathered.  The per-rank "verify" is a stub (the CUDA path needs
a GPU); what is under test is the partition, the ragged all-gather and the ordering of the result."""
import importlib.util
import os
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_shard():
    spec = importlib.util.spec_from_file_location("gs_shard", os.path.join(ROOT, "groth-sahai-rs_b200", "shard.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_shard_range_covers_everything():
    sh = _load_shard()
    for count in (0, 1, 2, 7, 64, 65536, 65537):
        for world in (1, 2, 3, 8):
            got = []
            for r in range(world):
                lo, hi = sh.shard_range(count, r, world)
                assert 0 <= lo <= hi <= count
                got.extend(range(lo, hi))
            assert got == list(range(count))
            sizes = sh.shard_counts(count, world)
            assert max(sizes) - min(sizes) <= 1 and sum(sizes) == count
    with pytest.raises(ValueError):
        sh.shard_range(4, 2, 2)


def _worker(rank, world, port, count, q):
    import torch.distributed as dist
    sh = _load_shard()
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        # every "proof" is 3 bytes in array 0 and 1 byte in array 1; verdict = (tag byte != 0xff)
        a0 = bytes((i * 7 + j) & 0xff for i in range(count) for j in range(3))
        a1 = bytes(0xff if i % 5 == 3 else i & 0x7f for i in range(count))
        seen = []

        def verify_local(arrs, n):
            assert len(arrs[0]) == 3 * n and len(arrs[1]) == n
            seen.append(n)
            return bytes(0 if t == 0xff else 1 for t in arrs[1])

        full = sh.verify_batch_sharded(verify_local, [a0, a1], [3, 1], count, rank, world)
        q.put((rank, full.tolist(), seen))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,count", [(2, 11), (2, 64), (3, 10), (2, 1)])
def test_verdict_all_gather_gloo(world, count):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, count, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [0 if i % 5 == 3 else 1 for i in range(count)]
    sh = _load_shard()
    for rank, full, seen in res:
        assert full == want, (rank, full)
        lo, hi = sh.shard_range(count, rank, world)
        assert seen == ([hi - lo] if hi > lo else [])


def _rand_worker(rank, world, port, count, bad_at, q):
    import torch.distributed as dist
    sh = _load_shard()
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        a0 = bytes(0xff if i == bad_at else i & 0x7f for i in range(count))      # one tag byte per "proof"; 0xff = bad

        def verify_rand_local(arrs, n):                                          # ONE verdict for the rank's block
            assert len(arrs[0]) == n
            return 0xff not in arrs[0]

        q.put((rank, sh.verify_batch_rand_sharded(verify_rand_local, [a0], [1], count, rank, world)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,count,bad_at", [(2, 11, -1), (2, 11, 0), (2, 11, 10), (3, 10, 4), (2, 1, -1), (2, 1, 0)])
def test_rand_verdict_all_reduce_gloo(world, count, bad_at):
    """shard.verify_batch_rand_sharded: per-rank one-byte verdicts AND-ed over the ranks (a bad proof in ANY block, also on
    a rank other than the caller's, turns every rank's answer to False; a rank with an empty block says True)."""
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rand_worker, args=(r, world, port, count, bad_at, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r for r, _ in res) == list(range(world))
    assert all(ok is (bad_at < 0) for _, ok in res), res


# ---------------------------------------------------------------- one statement split by slot (gs_verify_partial)
def test_owned_slots_partition():
    sh = _load_shard()
    for K in (1, 12, 2052):
        for world in (1, 2, 3, 8, 16):
            got = sorted(k for r in range(world) for k in sh.owned_slots(K, r, world))
            assert got == list(range(K))
    with pytest.raises(ValueError):
        sh.owned_slots(4, 3, 3)


def _stmt_worker(rank, world, port, count, q):
    import torch.distributed as dist
    sh = _load_shard()
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        # stub engine: the "partial product" of a rank is 2304 bytes per statement tagged with (rank, statement);
        # finish checks that every rank's block arrived, rank-major, and returns one verdict byte per statement
        def partial_local(r, w):
            assert (r, w) == (rank, world)
            return b"".join(bytes([r, s]) * (sh.PARTIAL_BYTES // 2) for s in range(count))

        def finish(allp):
            assert len(allp) == world * count * sh.PARTIAL_BYTES
            for r in range(world):
                for s in range(count):
                    blk = allp[(r * count + s) * sh.PARTIAL_BYTES:(r * count + s + 1) * sh.PARTIAL_BYTES]
                    assert blk == bytes([r, s]) * (sh.PARTIAL_BYTES // 2)
            return bytes(s % 2 for s in range(count))

        ok = sh.verify_statement_sharded(partial_local, finish, count, rank, world)
        q.put((rank, list(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,count", [(2, 1), (3, 4)])
def test_statement_partials_all_gather_gloo(world, count):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_stmt_worker, args=(r, world, port, count, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok in res:
        assert ok == [s % 2 for s in range(count)]


# ---------------------------------------------------------------- one statement split by EQUATION (C4: shard.*_equations_sharded)
def _eq_worker(rank, world, port, ty, E, q):
    """The per-rank prover / verifier is the C oracle on the CPU (the CUDA path needs a GPU): what is under test is the
    partition by variable (commit) and by equation (prove, verify), the ragged all-gathers and the ordering."""
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, ROOT)
    from bigcase import Statement, cb, make_crs
    sh = _load_shard()
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        m, n = 3, 2
        st = Statement(ty, m, n, E, make_crs(1)[0], seed=900 + ty, prove=False)
        cx, cy, xs, ys, ts = cb.cx_of(ty), cb.cy_of(ty), cb.x_size(ty), cb.y_size(ty), cb.target_size(ty)
        xc = sh.commit_sharded(lambda a, k: cb.commit_x(ty, a[0], a[1], st.crsb), [st.X, st.xr], [xs, cx * 32], m, 192, rank, world)
        yc = sh.commit_sharded(lambda a, k: cb.commit_y(ty, a[0], a[1], st.crsb), [st.Y, st.yr], [ys, cy * 32], n, 384, rank, world)
        pi, th = sh.prove_equations_sharded(
            lambda a, k: cb.prove_batch(ty, k, m, n, a[0], a[1], a[2], st.X, st.Y, st.xr, st.yr, a[3], True, st.crsb),
            [st.A, st.B, st.G, st.Tr], [n * xs, m * ys, m * n * 32, cx * cy * 32], E, cx * 384, cy * 192, rank, world)
        tg = bytearray(st.T)
        tg[ts:2 * ts], tg[2 * ts:3 * ts] = tg[2 * ts:3 * ts], tg[ts:2 * ts]        # swap the targets of equations 1 and 2
        ok = sh.verify_equations_sharded(
            lambda a, k: cb.verify_batch(ty, k, m, n, [a[0], a[1], a[2], a[3], xc * k, yc * k, a[4], a[5]], st.crsb),
            [st.A, st.B, st.G, bytes(tg), pi, th], [n * xs, m * ys, m * n * 32, ts, cx * 384, cy * 192], E, rank, world)
        q.put((rank, xc, yc, pi, th, ok.tolist()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,ty,E", [(2, 0, 5), (3, 3, 4), (2, 1, 1)])
def test_equations_sharded_gloo(world, ty, E):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from bigcase import Statement, make_crs
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_eq_worker, args=(r, world, port, ty, E, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = Statement(ty, 3, 2, E, make_crs(1)[0], seed=900 + ty)      # unsharded: commitments and proofs in one go
    expect_ok = [0 if e in (1, 2) and E > 2 else 1 for e in range(E)]
    for rank, xc, yc, pi, th, ok in res:
        assert xc == want.xc and yc == want.yc, rank
        assert pi == want.pi and th == want.theta, rank
        assert ok == expect_ok, (rank, ok)


# ---------------------------------------------------------------- MSM split by base: host-side pieces of gs_verify_sharded
def test_gamma_rows_of():
    import numpy as np
    sh = _load_shard()
    count, m, n = 2, 5, 3
    g = np.arange(count * m * n * 32, dtype=np.uint32).astype(np.uint8).tobytes()
    full = np.frombuffer(g, dtype=np.uint8).reshape(count, m, n * 32)
    for world in (1, 2, 3, 7):
        seen = []
        for r in range(world):
            rows = np.frombuffer(sh.gamma_rows_of(g, count, m, n, r, world), dtype=np.uint8)
            own = list(range(r, m, world))
            assert rows.size == count * len(own) * n * 32
            rows = rows.reshape(count, len(own), n * 32) if own else rows
            for k, i in enumerate(own):          # row i of Gamma sits at index i // world on rank i % world
                assert k == i // world and (rows[:, k] == full[:, i]).all()
            seen += own
        assert sorted(seen) == list(range(m))
    with pytest.raises(ValueError):
        sh.gamma_rows_of(g[:-1], count, m, n, 0, 1)


def _ag_worker(rank, world, port, q):
    """The all-gather callback on host pointers over gloo: what the C library calls twice per sharded verification."""
    import ctypes
    import torch.distributed as dist
    sh = _load_shard()
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        nbytes = 96 * 5
        send = (ctypes.c_char * nbytes).from_buffer_copy(bytes((rank * 17 + i) & 0xff for i in range(nbytes)))
        recv = (ctypes.c_char * (nbytes * world))()
        sh.make_allgather("cpu")(ctypes.addressof(send), ctypes.addressof(recv), nbytes)
        q.put((rank, bytes(recv)))
    finally:
        dist.destroy_process_group()


def test_allgather_callback_gloo():
    import torch.multiprocessing as mp
    world = 3
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ag_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = b"".join(bytes((r * 17 + i) & 0xff for i in range(96 * 5)) for r in range(world))
    for _, got in res:
        assert got == want


# ---------------------------------------------------------------- one statement: prove split by y-variable block
def _prove_worker(rank, world, port, ty, q):
    """Per-rank prover and the Com sums are the C oracle / big-int oracle on the CPU; under test: the sub-statement
    construction (column block of Gamma, B and T only on rank 0) and the gather + sum."""
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, ROOT)
    from bigcase import Case, cb, make_crs
    from conv import com1_b, com1_i, com2_b, com2_i
    from oracle import gs as ogs
    sh = _load_shard()
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        m, n = 3, 4
        c = Case(ty, m, n, make_crs(1)[0], seed=1400 + ty, prove=False)

        def prove_local(t, mm, nn, a, b, g, x, y, xr, yr, T):
            return cb.prove(t, mm, nn, a, b, g, x, y, xr, yr, T, c.crsb)

        def csum(conv_i, conv_b, add, size):
            def f(blob):
                acc = conv_i(blob[:size])
                for o in range(size, len(blob), size):
                    acc = add(acc, conv_i(blob[o:o + size]))
                return conv_b(acc)
            return f

        pi, th = sh.prove_statement_sharded(prove_local, csum(com2_i, com2_b, ogs.com2_add, 384), csum(com1_i, com1_b, ogs.com1_add, 192),
                                            ty, m, n, c.prove_args(), rank, world)
        q.put((rank, pi, th))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,ty", [(2, 0), (3, 1), (5, 2), (2, 3)])
def test_prove_statement_sharded_gloo(world, ty):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from bigcase import Case, make_crs
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_prove_worker, args=(r, world, port, ty, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = Case(ty, 3, 4, make_crs(1)[0], seed=1400 + ty)     # the unsharded reference-order proof
    for rank, pi, th in res:
        assert pi == want.pi and th == want.theta, rank
