"""Image sharding across the GPUs of one box.  Images (and the tiles inside them) are independent pictures
(reference: the tile loop of src/heic/decoder.rs:98-119 has no cross-tile state), so every rank decodes its own
own images of the job (image_idx % n_gpus; or a contiguous block) and nothing crosses NVLink on the data path; torch.distributed is used only for the
barrier around the timed region and for reducing the per-rank timings (max) and counters (sum)."""
from __future__ import annotations


def shard_range(n_items: int, rank: int, world: int) -> range:
    """Contiguous block partition: the first n_items % world ranks get one extra item."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def shard_modulo(n_items: int, rank: int, world: int) -> range:
    """Round-robin partition of SURVEY section 8(e): item i belongs to rank i % world."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return range(rank, n_items, world)


def reduce_timing(dist, device, elapsed_ms: float, counters: dict[str, float]):
    """-> (max over ranks of elapsed_ms, {name: sum over ranks}).  dist = torch.distributed or None."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return elapsed_ms, dict(counters)
    import torch

    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    names = sorted(counters)
    c = torch.tensor([float(counters[n]) for n in names], dtype=torch.float64, device=device)
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    return float(t[0]), {n: float(v) for n, v in zip(names, c.tolist())}
