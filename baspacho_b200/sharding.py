"""Batch sharding for identical-structure problems: items of a batch share only the read-only skeleton,
so the split is a contiguous partition with NO data-path collective; torch.distributed is used only to agree on
timings / checksums (NCCL on GPUs, gloo in the CPU tests)."""


def shard_range(batch, rank, world):
    """contiguous split, ceil(batch / world) items per rank (the last ranks may get fewer or none)"""
    per = -(-batch // world)
    lo = min(batch, rank * per)
    return lo, min(batch, lo + per)


def all_shards(batch, world):
    return [shard_range(batch, r, world) for r in range(world)]


def reduce_max(value, device=None):
    """max over ranks of a python float (no-op without an initialised process group)"""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_checksums(values, device=None):
    """all ranks receive the concatenation of every rank's per-item checksums (ordered by rank)"""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return list(map(float, values))
    world = dist.get_world_size()
    counts = [torch.zeros(1, dtype=torch.int64, device=device or "cpu") for _ in range(world)]
    dist.all_gather(counts, torch.tensor([len(values)], dtype=torch.int64, device=device or "cpu"))
    n = max(int(c.item()) for c in counts)
    mine = torch.zeros(max(n, 1), dtype=torch.float64, device=device or "cpu")
    if len(values):
        mine[:len(values)] = torch.tensor(list(values), dtype=torch.float64)
    outs = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(outs, mine)
    res = []
    for c, o in zip(counts, outs):
        res += o[:int(c.item())].cpu().tolist()
    return res
