"""The N>1 host logic on CPU: world_size-2 gloo processes exercise image sharding and the single weight broadcast
(the same code path bench.py / parallel.py use with NCCL on the GPUs)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diffpir_b200 import parallel


def test_shard_slice_partitions_exactly():
    for total in (1, 7, 32, 64, 128, 129):
        for world in (1, 2, 3, 4, 8):
            spans = [parallel.shard_slice(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and b >= a and d >= c
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


class _FakeModel:
    """Stands in for diffpir_b200.UNetModel on CPU: exposes the packed-arena surface used by broadcast_model."""

    def __init__(self, nbytes, fill):
        self.arena = torch.full((nbytes,), fill, dtype=torch.uint8)
        self.loaded = False

    def weight_arena(self):
        return self.arena

    def mark_weights_loaded(self):
        self.loaded = True


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # one broadcast of the packed weights from rank 0
        m = _FakeModel(4096, 7 if rank == 0 else 0)
        orig_sync = torch.cuda.synchronize
        torch.cuda.synchronize = lambda *a, **k: None      # CPU stand-in
        try:
            n = parallel.broadcast_model(m, src=0)
        finally:
            torch.cuda.synchronize = orig_sync
        ok = n == 4096 and m.loaded and bool((m.arena == 7).all())
        # image shards: each rank "restores" its contiguous slice; gather == the full batch in order
        total = 6
        full = torch.arange(total * 4, dtype=torch.float32).reshape(total, 4)
        lo, hi = parallel.shard_slice(total, rank, world)
        local = full[lo:hi] * 2.0
        gathered = parallel.gather_results(local, total)
        ok = ok and torch.equal(gathered, full * 2.0)
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_two_rank_broadcast_and_shard_gather():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = dict(q.get(timeout=10) for _ in range(2))
    assert res == {0: True, 1: True}
