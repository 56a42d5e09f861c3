"""Multi-GPU sharding of the batch workloads: one process per GPU, independent units
(proofs / variables / equations) split into contiguous blocks, NO data-path collective; the only exchange
is the all-gather of the per-unit results (verdict bytes, 1 B per proof).

The reference has no distributed mode (its parallelism is Rayon inside one process,
src/data_structures.rs:657-728); this is the "independent proofs" partition of verifier.rs:23-157.
Pure host logic over torch.distributed: works with NCCL (GPU tensors) and gloo (CPU tensors, tests).
"""
from typing import List, Sequence, Tuple


def shard_range(count: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of `count` units owned by `rank`; sizes differ by at most one."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad rank / world size")
    base, rem = divmod(count, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_counts(count: int, world: int) -> List[int]:
    return [shard_range(count, r, world)[1] - shard_range(count, r, world)[0] for r in range(world)]


def slice_units(arrays: Sequence, unit_sizes: Sequence[int], lo: int, hi: int) -> list:
    """Rows [lo, hi) of every per-unit array (bytes-like or numpy uint8, `unit_sizes[i]` bytes per unit)."""
    return [a[lo * s:hi * s] for a, s in zip(arrays, unit_sizes)]


def gather_verdicts(local_ok, count: int, group=None):
    """All-gather of the per-proof verdict bytes: `local_ok` is this rank's uint8 tensor (its shard, in
    order); returns the full uint8 tensor of `count` verdicts on every rank.  Shards may differ in size by
    one, so they are padded to the largest shard for the collective and trimmed afterwards."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    counts = shard_counts(count, world)
    width = max(counts) if counts else 0
    pad = torch.zeros(width, dtype=torch.uint8, device=local_ok.device)
    pad[:local_ok.numel()] = local_ok
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)]) if count else pad[:0]


def verify_batch_sharded(verify_local, arrays: Sequence, unit_sizes: Sequence[int], count: int, rank: int, world: int,
                         group=None, device="cpu"):
    """Verifiable::verify over `count` independent proofs split across `world` ranks.
    `verify_local(shard_arrays, n) -> sequence of n verdict bytes` is the single-GPU call
    (Engine.verify_batch bound to an equation type and shape)."""
    import torch
    lo, hi = shard_range(count, rank, world)
    ok = verify_local(slice_units(arrays, unit_sizes, lo, hi), hi - lo) if hi > lo else b""
    local = torch.tensor(list(ok), dtype=torch.uint8, device=device)
    if world == 1:
        return local
    return gather_verdicts(local, count, group)


def verify_batch_rand_sharded(verify_rand_local, arrays: Sequence, unit_sizes: Sequence[int], count: int, rank: int, world: int,
                              group=None, device="cpu") -> bool:
    """The opt-in randomised batch check (gs_verify_batch_rand) over `world` ranks: every rank checks its
    contiguous block of proofs with ITS OWN random weights (`verify_rand_local(shard_arrays, n) -> bool`, i.e.
    Engine.verify_batch_rand bound to a type and shape) and the one-byte verdicts are AND-ed (all-reduce MIN).  True iff
    every proof of the batch verifies, up to the per-rank error of 2^-62; no data-path collective."""
    import torch
    import torch.distributed as dist
    lo, hi = shard_range(count, rank, world)
    ok = bool(verify_rand_local(slice_units(arrays, unit_sizes, lo, hi), hi - lo)) if hi > lo else True
    if world == 1:
        return ok
    t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return bool(int(t.item()))


# ---------------------------------------------------------------- one large statement over several GPUs
PARTIAL_BYTES = 4 * 576   # four un-exponentiated GT values per statement (include/gs_b200.h gs_verify_partial)


def owned_slots(num_slots: int, rank: int, world: int) -> range:
    """Slots of the pairing-product equation evaluated by `rank` (round-robin, as the kernels do)."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad rank / world size")
    return range(rank, num_slots, world)


def gather_partials(local_partial, group=None):
    """All-gather of the per-rank Miller partial products (uint8 tensor, count * 2304 bytes, the same size on
    every rank) -> one uint8 tensor, rank-major, ready for gs_verify_finish.  This is the ONLY collective of a
    sharded statement: 2,304 B per rank and statement."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    parts = [torch.empty_like(local_partial) for _ in range(world)]
    dist.all_gather(parts, local_partial, group=group)
    return torch.cat(parts)


def verify_statement_sharded(partial_local, finish, count: int, rank: int, world: int, group=None, device="cpu"):
    """Verifiable::verify of `count` LARGE statements, each split by slot across `world` ranks.
    `partial_local(rank, world) -> bytes` is Engine.verify_partial bound to the statement(s);
    `finish(partials_bytes) -> count verdict bytes` is Engine.verify_finish.  Every rank returns the verdicts."""
    import torch
    mine = partial_local(rank, world)
    if len(mine) != count * PARTIAL_BYTES:
        raise ValueError("verify_partial returned the wrong number of bytes")
    local = torch.frombuffer(bytearray(mine), dtype=torch.uint8).to(device)
    allp = local if world == 1 else gather_partials(local, group)
    return finish(bytes(allp.cpu().numpy().tobytes()))


# ---------------------------------------------------------------- a multi-equation statement split by EQUATION (C4)
def _gather_rows(local_rows, counts, row_bytes: int, group=None):
    """All-gather of per-unit byte rows whose number differs by at most one between ranks: uint8 tensor
    [mine * row_bytes] -> [sum(counts) * row_bytes] in rank order (padded to the widest shard for the collective)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    width = max(counts) * row_bytes
    pad = torch.zeros(width, dtype=torch.uint8, device=local_rows.device)
    pad[:local_rows.numel()] = local_rows
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:c * row_bytes] for p, c in zip(parts, counts)])


def _to_tensor(b, device):
    import torch
    return torch.frombuffer(bytearray(b), dtype=torch.uint8).to(device) if len(b) else torch.zeros(0, dtype=torch.uint8, device=device)


def commit_sharded(commit_local, var_arrays: Sequence, unit_sizes: Sequence[int], nvars: int, out_bytes: int, rank: int,
                   world: int, group=None, device="cpu") -> bytes:
    """batch_commit_* (commit.rs:78-256) with the VARIABLES split across ranks: rank r commits
    its contiguous block (`commit_local(shard_arrays, count) -> count * out_bytes bytes`; var_arrays = the variables
    and their randomness rows) and the commitments (192 / 384 B each) are all-gathered, so every rank ends with the
    full Commit1 / Commit2 -- what every equation's prove and verify then shares."""
    lo, hi = shard_range(nvars, rank, world)
    mine = commit_local(slice_units(var_arrays, unit_sizes, lo, hi), hi - lo) if hi > lo else b""
    if len(mine) != (hi - lo) * out_bytes:
        raise ValueError("commit_local returned the wrong number of bytes")
    if world == 1:
        return bytes(mine)
    full = _gather_rows(_to_tensor(mine, device), shard_counts(nvars, world), out_bytes, group)
    return full.cpu().numpy().tobytes()


def prove_equations_sharded(prove_local, eq_arrays: Sequence, eq_unit_sizes: Sequence[int], num_eqs: int, pi_bytes: int,
                            theta_bytes: int, rank: int, world: int, group=None, device="cpu") -> Tuple[bytes, bytes]:
    """Provable::prove for the `num_eqs` equations of ONE statement (same type and shape, shared witnesses and
    commitment randomness) with the EQUATIONS split contiguously across ranks: the reference proves equation by
    equation (prove.rs:92-488, one EquProof per call), so the shards are independent.
    `prove_local(shard_eq_arrays, count) -> (pi bytes, theta bytes)` is Engine.prove_batch(shared_vars=True) bound
    to the type, shape, witnesses and randomness; eq_arrays = per-equation a_consts, b_consts, gamma, pf_rand.
    The proofs (pi_bytes + theta_bytes per equation) are all-gathered: every rank returns all of them, in order."""
    lo, hi = shard_range(num_eqs, rank, world)
    pi, th = prove_local(slice_units(eq_arrays, eq_unit_sizes, lo, hi), hi - lo) if hi > lo else (b"", b"")
    if len(pi) != (hi - lo) * pi_bytes or len(th) != (hi - lo) * theta_bytes:
        raise ValueError("prove_local returned the wrong number of bytes")
    if world == 1:
        return bytes(pi), bytes(th)
    counts = shard_counts(num_eqs, world)
    # one collective: pi and theta of an equation travel together
    import torch
    row = pi_bytes + theta_bytes
    mine = torch.empty((hi - lo) * row, dtype=torch.uint8, device=device).view(hi - lo, row) if hi > lo else None
    if hi > lo:
        mine[:, :pi_bytes] = _to_tensor(pi, device).view(hi - lo, pi_bytes)
        mine[:, pi_bytes:] = _to_tensor(th, device).view(hi - lo, theta_bytes)
        mine = mine.reshape(-1)
    else:
        mine = torch.zeros(0, dtype=torch.uint8, device=device)
    full = _gather_rows(mine, counts, row, group).view(num_eqs, row).cpu()
    return full[:, :pi_bytes].contiguous().numpy().tobytes(), full[:, pi_bytes:].contiguous().numpy().tobytes()


def verify_equations_sharded(verify_local, eq_arrays: Sequence, eq_unit_sizes: Sequence[int], num_eqs: int, rank: int,
                             world: int, group=None, device="cpu"):
    """Verifiable::verify for the equations of one statement split contiguously across ranks; the shared commitments
    are replicated (bound inside `verify_local(shard_eq_arrays, count) -> count verdict bytes`, which repeats them per
    equation for gs_verify_batch); eq_arrays = per-equation a_consts, b_consts, gamma, target, pi, theta.  One
    verdict byte per equation is all-gathered (verifier.rs:25-26: one EquProof per verify call)."""
    return verify_batch_sharded(verify_local, eq_arrays, eq_unit_sizes, num_eqs, rank, world, group, device)


# ---------------------------------------------------------------- one statement, MSM split by BASE (gs_verify_sharded)
def gamma_rows_of(gamma, count: int, m: int, n: int, rank: int, world: int) -> bytes:
    """Rows i = rank (mod world) of every m x n Gamma in `gamma` ([count][m][n] Fr, 32 B each), compact and in order:
    what gs_verify_sharded uploads on this rank (1 / world of the statement)."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad rank / world size")
    import numpy as np
    g = np.frombuffer(gamma, dtype=np.uint8)
    if g.size != count * m * n * 32:
        raise ValueError("gamma has the wrong size")
    return np.ascontiguousarray(g.reshape(count, m, n * 32)[:, rank::world, :]).tobytes()


class _DevMem:
    """A raw device pointer as an object torch.as_tensor can wrap (CUDA array interface, uint8)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def wrap_memory(ptr: int, nbytes: int, device):
    """uint8 tensor over `nbytes` at `ptr`: device memory (torch device 'cuda:k') or, for the CPU tests, host memory."""
    import torch
    if str(device).startswith("cuda"):
        return torch.as_tensor(_DevMem(ptr, nbytes), device=device)
    import ctypes
    return torch.frombuffer((ctypes.c_char * nbytes).from_address(ptr), dtype=torch.uint8)


def make_allgather(device, group=None):
    """The all-gather callback of gs_verify_sharded over torch.distributed (NCCL on device pointers: GPU to GPU over
    NVLink, no host staging; gloo on host pointers in the CPU tests).  World size 1 degenerates to a copy."""
    import torch
    import torch.distributed as dist

    def allgather(send_ptr, recv_ptr, nbytes):
        world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        send = wrap_memory(send_ptr, nbytes, device)
        recv = wrap_memory(recv_ptr, nbytes * world, device)
        if world == 1:
            recv.copy_(send)
        else:
            dist.all_gather_into_tensor(recv, send, group=group)
        if str(device).startswith("cuda"):
            torch.cuda.current_stream(device).synchronize()

    return allgather


def verify_statement_base_sharded(engine, ty, count, m, n, arrays, rank, world, device, group=None) -> bytes:
    """Verifiable::verify of `count` statements (one big one: C3; or the equations of one statement over shared
    commitments: C4) with the statement MSM split by base and the Miller pairs by slot over `world` ranks.
    arrays = the 8 byte strings of gs_verify_batch with the FULL Gamma; every rank returns the verdict bytes."""
    a, b, gamma, target, xc, yc, pi, th = arrays
    rows = gamma_rows_of(gamma, count, m, n, rank, world)
    return engine.verify_sharded(ty, count, m, n, a, b, rows, target, xc, yc, pi, th, rank, world, make_allgather(device, group))


# ---------------------------------------------------------------- one large statement: prove split by y-variable block
def prove_statement_sharded(prove_local, com2_sum, com1_sum, ty: int, m: int, n: int, arrays: Sequence, rank: int, world: int,
                            group=None, device="cpu") -> Tuple[bytes, bytes]:
    """Provable::prove (prove.rs:92-488) of ONE statement with the y-variables (the columns of Gamma) split contiguously
    across ranks.  Both proof elements are linear in the statement: rank r proves the SUB-statement
        (A[J_r], B', Gamma[:, J_r]) over all x-variables and the y-variables J_r,  B' = B and T' = T on rank 0, zero elsewhere
    with the ordinary single-GPU call, and  pi = sum_r pi^(r),  theta = sum_r theta^(r)  entry-wise in Com2 / Com1
    ("gather <= 4 partial sums per GPU, add, normalise").  A rank uploads 1 / world of Gamma.
    arrays = (a_consts, b_consts, gamma, xvars, yvars, x_rand, y_rand, pf_rand) as gs_prove takes them;
    `prove_local(ty, m, n_r, a, b, gamma, x, y, xr, yr, T) -> (pi, theta)` is Engine.prove; `com2_sum(list_bytes) -> bytes`
    and `com1_sum` add Com elements (Engine.group_sum).  Every rank returns the full proof."""
    import numpy as np
    a, b, gamma, xv, yv, xr, yr, T = arrays
    gx, gy = ty in (0, 1), ty in (0, 2)
    asz, bsz, cx, cy = (96 if gx else 32), (192 if gy else 32), (2 if gx else 1), (2 if gy else 1)
    lo, hi = shard_range(n, rank, world)
    nr = hi - lo
    if nr > 0:
        g = np.frombuffer(gamma, dtype=np.uint8).reshape(m, n * 32)[:, lo * 32:hi * 32]
        zero_b = bytes(len(b)) if rank else b                # identity points / zero scalars: those terms vanish
        zero_t = bytes(len(T)) if rank else T
        pi, th = prove_local(ty, m, nr, a[lo * asz:hi * asz], zero_b, np.ascontiguousarray(g).tobytes(), xv, yv[lo * bsz:hi * bsz], xr,
                             yr[lo * cy * 32:hi * cy * 32], zero_t)
    else:
        pi, th = bytes(cx * 384), bytes(cy * 192)            # a rank without a column contributes the identity
    if world == 1:
        return pi, th
    import torch
    import torch.distributed as dist
    mine = _to_tensor(pi + th, device)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    parts = [bytes(p.cpu().numpy().tobytes()) for p in parts]
    pis = [p[:cx * 384] for p in parts]
    ths = [p[cx * 384:] for p in parts]
    out_pi = b"".join(com2_sum(b"".join(q[i * 384:(i + 1) * 384] for q in pis)) for i in range(cx))
    out_th = b"".join(com1_sum(b"".join(q[i * 192:(i + 1) * 192] for q in ths)) for i in range(cy))
    return out_pi, out_th
