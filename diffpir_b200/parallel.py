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


class WeightComm:
    """The library's own NCCL communicator for the weight broadcast (`dpir_comm_*`, include/diffpir_b200.h): rank 0 draws the
    128-byte NCCL id, the host program shares it (here: `torch.distributed.broadcast_object_list` over whatever backend the job
    already has - gloo is enough, it is 128 bytes), every rank joins with its GPU current.  `broadcast(model)` then is ONE
    ncclBroadcast of the packed arena issued by the library on the current stream; no torch collective touches the weights."""

    def __init__(self, rank=None, world=None, group=None):
        import ctypes
        import torch.distributed as dist
        from . import _lib
        self._lib = _lib.load()
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        buf = ctypes.create_string_buffer(128)
        if self.rank == 0:
            _lib.check(self._lib.dpir_comm_unique_id(buf, 128), "dpir_comm_unique_id")
        box = [bytes(buf.raw)]
        dist.broadcast_object_list(box, src=0, group=group)
        h = ctypes.c_void_p()
        _lib.check(self._lib.dpir_comm_create(ctypes.c_char_p(box[0]), self.world, self.rank, ctypes.byref(h)), "dpir_comm_create")
        self._h = h

    def broadcast(self, model, src=0):
        from . import _lib
        _lib.check(self._lib.dpir_bcast_weights(model._h, self._h, src, _lib.stream_ptr()), "dpir_bcast_weights")
        return model.weight_arena().numel()

    def close(self):
        if self._h:
            self._lib.dpir_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def gather_results(x_local, total, group=None):
    """Optional end-of-run all_gather of the per-rank result shards (equal shard sizes) back to (total, ...)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    outs = [torch.empty_like(x_local) for _ in range(world)]
    dist.all_gather(outs, x_local.contiguous(), group=group)
    return torch.cat(outs, 0)[:total]
