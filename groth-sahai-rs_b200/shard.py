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
