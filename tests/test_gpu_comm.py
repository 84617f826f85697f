"""The path's one collective issued by the library itself (dpir_comm_*, dpir_bcast_weights; include/diffpir_b200.h): rank 0 repacks
the checkpoint, every other rank receives the kernel-native arena with one ncclBroadcast and computes rank 0's answers."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _small_cfg():
    from diffpir_b200 import arch
    return arch.ffhq_config()          # the FFHQ architecture, run on 64 x 64 inputs below (keeps the two forwards short)


def test_library_broadcast_single_rank_world():
    """world = 1: the communicator and the broadcast run on the GPU (ncclCommInitRank, ncclBroadcast arena -> arena) and leave the
    model's answers untouched."""
    import ctypes
    from diffpir_b200 import _lib, synth
    from diffpir_b200.unet import UNetModel
    lib = _lib.load()
    cfg = _small_cfg()
    m = UNetModel(cfg, precision="fp16s")
    m.load_state_dict(synth.synthetic_state_dict(cfg, 0))
    x = torch.randn(2, 3, 64, 64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    t = torch.tensor([900, 30], device="cuda")
    before = m.x0(x, t).clone()
    ident = ctypes.create_string_buffer(128)
    _lib.check(lib.dpir_comm_unique_id(ident, 128), "dpir_comm_unique_id")
    h = ctypes.c_void_p()
    _lib.check(lib.dpir_comm_create(ident, 1, 0, ctypes.byref(h)), "dpir_comm_create")
    _lib.check(lib.dpir_bcast_weights(m._h, h, 0, _lib.stream_ptr()), "dpir_bcast_weights")
    assert lib.dpir_bcast_weights(m._h, h, 1, _lib.stream_ptr()) == -1      # root outside the world
    torch.cuda.synchronize()
    assert torch.equal(m.x0(x, t), before)
    lib.dpir_comm_destroy(h)


def _two_rank_worker(rank, port, out_dir):
    import torch.distributed as dist
    from diffpir_b200 import synth
    from diffpir_b200.parallel import WeightComm
    from diffpir_b200.unet import UNetModel
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=2)   # only carries the 128-byte id
    cfg = _small_cfg()
    m = UNetModel(cfg, precision="fp16s")
    if rank == 0:
        m.load_state_dict(synth.synthetic_state_dict(cfg, 0))
    wc = WeightComm()
    n = wc.broadcast(m, src=0)
    torch.cuda.synchronize()
    wc.close()
    x = torch.randn(2, 3, 64, 64, generator=torch.Generator().manual_seed(7)).cuda()
    out = m.x0(x, torch.tensor([999, 250], device="cuda")).cpu()
    torch.save({"out": out, "arena_bytes": n}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs on one node")
def test_library_broadcast_two_ranks(tmp_path):
    """Rank 1 never sees the checkpoint: after dpir_bcast_weights its x0 is bit-identical to rank 0's."""
    import torch.multiprocessing as mp
    mp.spawn(_two_rank_worker, args=(_free_port(), str(tmp_path)), nprocs=2, join=True)
    r0, r1 = torch.load(tmp_path / "r0.pt"), torch.load(tmp_path / "r1.pt")
    assert r0["arena_bytes"] == r1["arena_bytes"] > 0
    assert torch.isfinite(r0["out"]).all()
    assert torch.equal(r0["out"], r1["out"])
