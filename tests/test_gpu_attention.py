"""The attention core on its own (no isolated attention parity test existed): dpir_attention vs the reference's
QKVAttentionLegacy / QKVAttention math (guided_diffusion/unet.py:337-354, 371-389) evaluated in fp64 with torch."""
import pytest
import torch

from diffpir_b200 import _lib
from dpir_testutil import rel_l2

pytestmark = pytest.mark.gpu


def _reference(qkv_nct, n_heads, new_order):
    """unet.py:337-354 (legacy: [head][q,k,v][ch]) / :371-389 (new: [q,k,v][head][ch]) on (N, 3*H*C, T), fp64."""
    bs, width, length = qkv_nct.shape
    ch = width // (3 * n_heads)
    scale = 1 / (ch ** 0.5) ** 0.5
    if not new_order:
        q, k, v = qkv_nct.reshape(bs * n_heads, ch * 3, length).split(ch, dim=1)
        w = torch.einsum("bct,bcs->bts", q * scale, k * scale)
        w = torch.softmax(w, dim=-1)
        a = torch.einsum("bts,bcs->bct", w, v)
        return a.reshape(bs, -1, length)
    q, k, v = qkv_nct.chunk(3, dim=1)
    w = torch.einsum("bct,bcs->bts", (q * scale).view(bs * n_heads, ch, length), (k * scale).view(bs * n_heads, ch, length))
    w = torch.softmax(w, dim=-1)
    a = torch.einsum("bts,bcs->bct", w, v.reshape(bs * n_heads, ch, length))
    return a.reshape(bs, -1, length)


@pytest.mark.parametrize("out_half", [1, 0])
@pytest.mark.parametrize("B,T,heads,new_order", [
    (2, 64, 8, 0),        # FFHQ middle block (8x8)
    (3, 256, 8, 0),       # FFHQ 16x16
    (1, 1024, 8, 0),      # ImageNet 32x32
    (2, 256, 16, 1),      # new attention order
    (2, 324, 8, 0),       # 18x18: utils_model.test_mode quadrant of a 512^2 image (T % 64 != 0 used to be rejected)
    (1, 1296, 4, 0),      # 36x36
    (2, 100, 2, 0),       # ragged single key block
])
def test_attention_matches_reference_math(B, T, heads, new_order, out_half):
    lib = _lib.load()
    C = heads * 64
    g = torch.Generator().manual_seed(7 + T)
    qkv = torch.randn(B, T, 3 * C, generator=g) * 1.5                      # NHWC: (B, T, 3C)
    ref = _reference(qkv.double().permute(0, 2, 1).contiguous(), heads, new_order).permute(0, 2, 1)      # (B, T, C)
    qd = qkv.cuda().contiguous()
    out = torch.full((B, T, C), float("nan"), device="cuda", dtype=torch.float16 if out_half else torch.float32)
    _lib.check(lib.dpir_attention(_lib.ptr(qd), _lib.ptr(out), out_half, B, T, C, heads, new_order, _lib.stream_ptr()), "dpir_attention")
    torch.cuda.synchronize()
    assert not torch.isnan(out.float()).any()
    err = rel_l2(out.float(), ref)
    print(f"attention T={T} heads={heads} new_order={new_order} {'tcgen05/fp16' if out_half else 'fp32'}: rel-L2 {err:.2e}")
    # fp16 rounding of q, k, v and of the probabilities (tcgen05 path) vs exact fp32 arithmetic
    assert err < (3e-3 if out_half else 2e-6), err
