"""Multi-GPU plumbing for the ray-sharded Shoot path (one process per GPU, torch.distributed).

Geometry is replicated on every rank; the ray batch is block-partitioned; the only cross-rank
step is the gather of the per-ray results onto rank 0 (NCCL over NVLink on GPUs, gloo in the CPU
tests).  No collective touches the traversal itself.
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous block of ray indices owned by `rank`: [n*rank/world, n*(rank+1)/world)."""
    return (n * rank) // world, (n * (rank + 1)) // world


def gather_rows(t, dst=0, sizes=None):
    """Gather row-blocks of `t` (same trailing shape on all ranks) to rank `dst`, in rank order.
    `sizes`: rows per rank (defaults to equal blocks).  Returns the concatenated tensor on dst, None elsewhere."""
    world = dist.get_world_size()
    rank = dist.get_rank()
    if sizes is None:
        sizes = [t.shape[0]] * world
    if len(set(sizes)) == 1:
        out = torch.empty((sizes[0] * world,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device) if rank == dst else None
        dist.gather(t.contiguous(), list(out.chunk(world)) if rank == dst else None, dst=dst)
        return out
    # ragged blocks: gather() needs equal sizes, so pad to the largest block and trim on dst
    m = max(sizes)
    pad = torch.zeros((m,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst)
    if rank != dst:
        return None
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)], dim=0)


def max_over_ranks(x, device):
    v = torch.tensor([float(x)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
    return float(v.item())


def sum_over_ranks(x, device):
    v = torch.tensor([int(x)], dtype=torch.int64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(v, op=dist.ReduceOp.SUM)
    return int(v.item())
