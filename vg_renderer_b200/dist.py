"""Multi-GPU plumbing (torch.distributed is used for plumbing only: NCCL on GPUs, gloo in the CPU tests).

Paths are independent, so ranks take contiguous item ranges and run the whole pipeline locally
with no data-path collective.  The ONLY exchange happens when a single merged
vertex/index space is requested: an all-gather of each rank's (vertex, index) totals -- 16 bytes
per rank -- whose exclusive prefix gives the rank's vertex_base (added to every merged uint32
index by the emit kernel) and index_base (where its index block starts in the merged buffer)."""
import torch
import torch.distributed as dist


def shard_range(num_items, rank, world):
    """contiguous, balanced item range of `rank` (i.i.d. items: equal split)"""
    q, r = divmod(num_items, world)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def exclusive_bases(num_vertices, num_indices, device=None, group=None):
    """-> (vertex_base, index_base, total_vertices, total_indices) for the calling rank"""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0, 0, int(num_vertices), int(num_indices)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    mine = torch.tensor([int(num_vertices), int(num_indices)], dtype=torch.int64, device=device or "cpu")
    every = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(every, mine, group=group)
    tot = torch.stack(every).cpu()
    return int(tot[:rank, 0].sum()), int(tot[:rank, 1].sum()), int(tot[:, 0].sum()), int(tot[:, 1].sum())
