"""Multi-GPU plumbing: one process per GPU, images shard across ranks, ONE collective (the weight broadcast).

The reference is single-process / single-GPU (SURVEY.md section 2 "Parallelism"); every image's trajectory is
independent and only scalars are batch-shared, so the path shards with no data-path collective (section 8(e)).
"""
import torch


def shard_slice(total, rank, world):
    """Contiguous image shard [lo, hi) of rank `rank` (remainder spread over the first ranks)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_model(model, src=0, group=None):
    """Broadcast the packed weight arena (kernel-native layouts + tables) from `src` to every rank: rank `src` must have
    loaded a state_dict; the others receive ready-to-run weights without touching the checkpoint."""
    import torch.distributed as dist
    arena = model.weight_arena()
    dist.broadcast(arena, src=src, group=group)
    torch.cuda.synchronize()
    model.mark_weights_loaded()
    return arena.numel()


def gather_results(x_local, total, group=None):
    """Optional end-of-run all_gather of the per-rank result shards (equal shard sizes) back to (total, ...)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    outs = [torch.empty_like(x_local) for _ in range(world)]
    dist.all_gather(outs, x_local.contiguous(), group=group)
    return torch.cat(outs, 0)[:total]
