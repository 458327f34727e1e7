"""Frame sharding for multi-GPU runs: frames are independent, so the batch is
block-partitioned across ranks and no data-path collective exists.  torch.distributed is used only
for the barrier / max-over-ranks timing reduction and the optional gather of output frames."""
import torch
import torch.distributed as dist


def partition(total, world_size, rank):
    """Contiguous block [start, stop) of `total` frames owned by `rank`; sizes differ by at most one."""
    if not (0 <= rank < world_size) or total < 0:
        raise ValueError("bad partition arguments")
    base, rem = divmod(total, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def reduce_max(value, device="cpu"):
    """Max of a python float over all ranks (identity when torch.distributed is not initialised)."""
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_sum(value, device="cpu"):
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_frames(local, total, dst=0):
    """Gather per-rank frame blocks [n_r, ...] onto `dst` in frame order (returns None elsewhere)."""
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [partition(total, world, r) for r in range(world)]
    maxn = max(b - a for a, b in sizes)
    pad = torch.zeros((maxn,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst)
    if rank != dst:
        return None
    return torch.cat([bufs[r][: b - a] for r, (a, b) in enumerate(sizes)], dim=0)
