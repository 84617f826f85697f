"""Per-kernel parity of the two convolution kernels (tcgen05 fp16 and CUDA-core fp32) through the C-ABI."""
import ctypes
import os

import pytest
import torch
import torch.nn.functional as F

from dpir_testutil import rel_l2

pytestmark = pytest.mark.gpu


def _pack_tc(w, cout_pad):
    """(O,I,k,k) fp32 -> [tap][cout_pad][I] fp16 (the layout dpir_load_weight produces for the tcgen05 path)."""
    O, I, kh, kw = w.shape
    p = torch.zeros(kh * kw, cout_pad, I, dtype=torch.float16)
    p[:, :O] = w.permute(2, 3, 0, 1).reshape(kh * kw, O, I).half()
    return p.contiguous()


def _pack_simt(w):
    O, I, kh, kw = w.shape
    ld = (O + 3) // 4 * 4
    p = torch.zeros(kh * kw, I, ld)
    p[:, :, :O] = w.permute(2, 3, 1, 0).reshape(kh * kw, I, O)
    return p.contiguous()


def _case(B, H, W, Cin, Cout, k, seed, residual):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
    b = torch.randn(Cout, generator=g) * 0.1
    r = torch.randn(B, Cout, H, W, generator=g) if residual else None
    return x, w, b, r


SHAPES = [
    # B, H, W, Cin, Cout, k, residual
    (1, 8, 8, 128, 128, 3, False),     # tile spans 2 images (TN=2) with B=1 -> half the M tile is out of range
    (3, 8, 8, 256, 128, 3, True),
    (2, 16, 16, 128, 256, 3, True),
    (2, 32, 32, 384, 128, 3, False),   # 6 K-chunks (skip-concat width)
    (1, 64, 64, 128, 128, 3, True),
    (1, 256, 256, 128, 128, 3, False), # two 128-pixel tiles per row
    (2, 16, 16, 512, 1536, 1, False),  # qkv 1x1
    (2, 16, 16, 256, 128, 1, True),    # skip / proj 1x1 with residual
]


@pytest.mark.parametrize("variant", ["auto", "v1", "v2"])
@pytest.mark.parametrize("B,H,W,Cin,Cout,k,residual", SHAPES + [
    (5, 32, 32, 128, 256, 3, True),     # persistent kernel: 40 tiles, 2 cout tiles, odd batch
    (2, 48, 80, 128, 128, 3, True),     # H, W multiples of 16 but not powers of two
    (1, 24, 40, 192, 128, 3, False),    # ragged: partial 16x16 tiles on both edges
    (7, 16, 16, 256, 512, 3, False),
])
def test_conv_tc_matches_torch(B, H, W, Cin, Cout, k, residual, variant):
    from diffpir_b200 import _lib
    lib = _lib.load()
    lib.dpir_conv_tc_variant({"auto": 0, "v1": 1, "v2": 2}[variant])
    x, w, b, r = _case(B, H, W, Cin, Cout, k, 0, residual)
    # reference: fp32 conv of the fp16-rounded operands (isolates kernel correctness from operand rounding)
    xh, wh = x.half().float(), w.half().float()
    ref = F.conv2d(xh.double(), wh.double(), b.double(), padding=k // 2)
    if r is not None:
        ref = ref + r.double()
    act = x.permute(0, 2, 3, 1).contiguous().half().cuda()
    wp = _pack_tc(w, Cout).cuda()
    bias = b.cuda()
    res = r.permute(0, 2, 3, 1).contiguous().cuda() if r is not None else None
    out = torch.full((B, H, W, Cout), float("nan"), device="cuda")
    _lib.check(lib.dpir_conv_tc(_lib.ptr(act), _lib.ptr(wp), _lib.ptr(bias), _lib.ptr(res), _lib.ptr(out), B, H, W, Cin, Cout,
                                Cout, k, _lib.stream_ptr()), "dpir_conv_tc")
    torch.cuda.synchronize()
    lib.dpir_conv_tc_variant(0)
    got = out.permute(0, 3, 1, 2)
    assert not torch.isnan(got).any()
    assert rel_l2(got, ref) < 1e-5      # fp32 accumulation order over K = 9*Cin
    # and against the true fp32 conv: error is fp16 operand rounding only
    full = F.conv2d(x.double(), w.double(), b.double(), padding=k // 2) + (r.double() if r is not None else 0)
    assert rel_l2(got, full) < 1.5e-3


@pytest.mark.parametrize("B,H,W,Cin,Cout,residual", [
    (2, 32, 32, 128, 8, False),
    (1, 256, 256, 128, 8, False),      # the FFHQ output layer's shape: 19 x 16 tiles of 14 x 16 output pixels, the last tile column ragged
    (3, 24, 40, 256, 4, True),         # ragged on both edges, 4 K chunks, fp32 residual
    (2, 16, 16, 384, 16, False),       # 6 K chunks (the resident-weight limit), all 16 columns of a tap block in use
    (40, 16, 16, 128, 12, False),      # more tiles than SMs at the smallest map the kernel takes
    (2, 8, 8, 128, 8, False),          # below 16 x 16: stays on the one-tile-per-CTA kernel
])
def test_conv_tc_small_cout_pad16(B, H, W, Cin, Cout, residual):
    """3x3 convolutions with <= 16 output channels (the UNet's output layer): k_conv3x3_dxn, horizontal taps in the MMA's N."""
    from diffpir_b200 import _lib
    lib = _lib.load()
    x, w, b, r = _case(B, H, W, Cin, Cout, 3, 1, residual)
    ref = F.conv2d(x.half().double(), w.half().double(), b.double(), padding=1)
    if r is not None:
        ref = ref + r.double()
    act = x.permute(0, 2, 3, 1).contiguous().half().cuda()
    wp = _pack_tc(w, 16).cuda()
    res = r.permute(0, 2, 3, 1).contiguous().cuda() if r is not None else None
    out = torch.full((B, H, W, Cout), float("nan"), device="cuda")
    _lib.check(lib.dpir_conv_tc(_lib.ptr(act), _lib.ptr(wp), _lib.ptr(b.cuda()), _lib.ptr(res), _lib.ptr(out), B, H, W, Cin, Cout, 16,
                                3, _lib.stream_ptr()), "dpir_conv_tc")
    torch.cuda.synchronize()
    assert not torch.isnan(out).any()
    assert rel_l2(out.permute(0, 3, 1, 2), ref) < 2e-6


@pytest.mark.parametrize("B,H,W,Cin,Cout,k,residual", SHAPES[:5] + SHAPES[6:] + [(2, 16, 16, 128, 6, 3, False)])
def test_conv_simt_matches_torch(B, H, W, Cin, Cout, k, residual):
    from diffpir_b200 import _lib
    lib = _lib.load()
    if Cout % 4:
        residual = False
    x, w, b, r = _case(B, H, W, Cin, Cout, k, 2, residual)
    ref = F.conv2d(x.double(), w.double(), b.double(), padding=k // 2)
    if r is not None:
        ref = ref + r.double()
    act = x.permute(0, 2, 3, 1).contiguous().cuda()
    wp = _pack_simt(w).cuda()
    bias = torch.zeros((Cout + 3) // 4 * 4)
    bias[:Cout] = b
    res = r.permute(0, 2, 3, 1).contiguous().cuda() if r is not None else None
    out = torch.full((B, H, W, Cout), float("nan"), device="cuda")
    _lib.check(lib.dpir_conv_simt(_lib.ptr(act), _lib.ptr(wp), _lib.ptr(bias.cuda()), _lib.ptr(res), _lib.ptr(out), B, H, W, Cin,
                                  Cout, k, _lib.stream_ptr()), "dpir_conv_simt")
    torch.cuda.synchronize()
    assert rel_l2(out.permute(0, 3, 1, 2), ref) < 2e-6


@pytest.mark.parametrize("variant", ["auto", "v2", "v1"])
@pytest.mark.parametrize("B,H,W,Cin,Cout,residual", [
    (2, 32, 32, 128, 128, True), (3, 64, 64, 128, 256, False), (1, 24, 40, 192, 128, True),
    (5, 16, 16, 256, 512, True), (2, 48, 80, 128, 128, False), (1, 256, 256, 128, 128, True), (2, 8, 8, 256, 128, True),
])
def test_conv_tc_fp16_stream_io_matches_torch(B, H, W, Cin, Cout, residual, variant):
    """fp16 residual + fp16 output (the fp16-stream pipeline; on >=16x16 maps the swap-AB kernel with TMA-store epilogue)."""
    from diffpir_b200 import _lib
    lib = _lib.load()
    lib.dpir_conv_tc_variant({"auto": 0, "v1": 1, "v2": 2}[variant])
    x, w, b, r = _case(B, H, W, Cin, Cout, 3, 5, residual)
    rh = r.half() if r is not None else None
    ref = F.conv2d(x.half().double(), w.half().double(), b.double(), padding=1)
    if rh is not None:
        ref = ref + rh.double()
    act = x.permute(0, 2, 3, 1).contiguous().half().cuda()
    wp = _pack_tc(w, Cout).cuda()
    res = rh.permute(0, 2, 3, 1).contiguous().cuda() if rh is not None else None
    out = torch.full((B, H, W, Cout), float("nan"), device="cuda", dtype=torch.float16)
    _lib.check(lib.dpir_conv_tc_f16io(_lib.ptr(act), _lib.ptr(wp), _lib.ptr(b.cuda()), _lib.ptr(res), _lib.ptr(out), B, H, W, Cin,
                                      Cout, Cout, 3, _lib.stream_ptr()), "dpir_conv_tc_f16io")
    torch.cuda.synchronize()
    lib.dpir_conv_tc_variant(0)
    got = out.permute(0, 3, 1, 2).float()
    assert not torch.isnan(got).any()
    assert rel_l2(got, ref) < 6e-4      # fp16 rounding of the stored result


def _gn_reference(x, gamma, beta, film, eps=1e-5):
    """GroupNorm32 (nn.py:17-19) + optional FiLM (unet.py:246-249) + SiLU in fp64 on (B,C,H,W)."""
    y = F.group_norm(x.double(), 32, gamma.double(), beta.double(), eps)
    if film is not None:
        C = x.shape[1]
        y = y * (1 + film[:, :C].double()[:, :, None, None]) + film[:, C:].double()[:, :, None, None]
    return F.silu(y)


def _quad_stats(xh):
    """[B][C/4][2] fp64 (sum, sumsq) of an NHWC fp16 tensor - what the producing conv epilogue attaches to it."""
    B, H, W, C = xh.shape
    v = xh.double().reshape(B, H * W, C // 4, 4)
    return torch.stack([v.sum(dim=(1, 3)), (v * v).sum(dim=(1, 3))], dim=-1).contiguous()


# dpir_set_gn_fusion flavours: "1" = activate-once producers (k_conv3x3_v3<3>, the default since round 2), "4" = round-1
# per-tap-box producers with fp32 arithmetic, "2" = per-tap-box with packed-half arithmetic
_FUSE_MODES = ["1", "4", "2"]


@pytest.mark.parametrize("math", _FUSE_MODES)
@pytest.mark.parametrize("B,H,W,C1,C2,Cout,residual,film", [
    (2, 32, 32, 128, 0, 128, True, True),       # norm2 -> conv2 with FiLM and residual
    (3, 64, 64, 128, 0, 256, False, False),     # two cout tiles
    (1, 24, 40, 128, 0, 128, True, False),      # ragged tiles: zero padding must be applied AFTER the activation
    (2, 16, 16, 256, 256, 512, False, False),   # virtual concat (norm1 of an output block), groups straddle nothing
    (2, 32, 32, 512, 256, 256, False, True),    # concat 768: 24 channels per group straddle the two sources
    (1, 256, 256, 128, 128, 128, False, False), # full-resolution concat
    (5, 16, 16, 256, 0, 256, True, True),
])
def test_conv_gn_fused_matches_torch(B, H, W, C1, C2, Cout, residual, film, math):
    """k_conv3x3_v3<FUSE>: out = conv3x3(SiLU(FiLM(GroupNorm(x1|x2)))) + bias (+ residual), the norm evaluated inside the conv."""
    from diffpir_b200 import _lib
    lib = _lib.load()
    _lib.check(lib.dpir_set_gn_fusion(None, int(math)), "dpir_set_gn_fusion")
    C = C1 + C2
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, C, H, W, generator=g) * 1.7 + 0.6 * torch.randn(B, C, 1, 1, generator=g)
    xh = x.half()
    w = torch.randn(Cout, C, 3, 3, generator=g) / (C * 9) ** 0.5
    b = torch.randn(Cout, generator=g) * 0.1
    gamma = 1 + 0.3 * torch.randn(C, generator=g)
    beta = 0.2 * torch.randn(C, generator=g)
    fl = 0.3 * torch.randn(B, 2 * C, generator=g) if film else None
    r = torch.randn(B, Cout, H, W, generator=g).half() if residual else None
    act = _gn_reference(xh.float(), gamma, beta, fl)
    ref = F.conv2d(act, w.half().double(), b.double(), padding=1)
    if r is not None:
        ref = ref + r.double()
    x_nhwc = xh.permute(0, 2, 3, 1).contiguous()
    s1 = x_nhwc[..., :C1].contiguous().cuda()
    s2 = x_nhwc[..., C1:].contiguous().cuda() if C2 else None
    q1 = _quad_stats(s1.cpu()).cuda()
    q2 = _quad_stats(s2.cpu()).cuda() if C2 else None
    coef = torch.empty(B, 2, C, device="cuda")
    flc = fl.cuda().contiguous() if film else None
    gd, bd, biasd = gamma.cuda(), beta.cuda(), b.cuda()      # named: a temporary's block could be recycled before the launch
    _lib.check(lib.dpir_gn_coef(_lib.ptr(q1), _lib.ptr(q2), C1, C2, _lib.ptr(gd), _lib.ptr(bd), _lib.ptr(flc),
                                2 * C, B, H * W, _lib.ptr(coef), _lib.stream_ptr()), "dpir_gn_coef")
    wp = _pack_tc(w, Cout).cuda()
    res = r.permute(0, 2, 3, 1).contiguous().cuda() if r is not None else None
    out = torch.full((B, H, W, Cout), float("nan"), device="cuda", dtype=torch.float16)
    _lib.check(lib.dpir_conv_gnfuse_f16io(_lib.ptr(s1), C1, _lib.ptr(s2), C2, _lib.ptr(coef), _lib.ptr(wp), _lib.ptr(biasd),
                                          _lib.ptr(res), _lib.ptr(out), B, H, W, Cout, Cout, _lib.stream_ptr()),
               "dpir_conv_gnfuse_f16io")
    torch.cuda.synchronize()
    lib.dpir_set_gn_fusion(None, 1)
    got = out.permute(0, 3, 1, 2).float()
    assert not torch.isnan(got).any()
    # fp16 rounding of the activated operand (+ of the stored result); packed-half arithmetic (math 2) rounds twice more
    assert rel_l2(got, ref) < (3e-3 if math == "2" else 1.2e-3)


@pytest.mark.parametrize("B,H,W,C,Cout", [
    (1, 256, 256, 128, 8),      # the FFHQ output head's shape (the NHWC test entry needs Cout % 4 == 0; 6 runs in the UNet tests)
    (3, 24, 40, 256, 4),        # ragged tiles: exact zeros outside the image AFTER the activation; 4 K chunks
    (33, 16, 16, 128, 12),
])
def test_conv_gn_fused_narrow_output(B, H, W, C, Cout):
    """k_conv3x3_dxn<true>: out = conv3x3(SiLU(GroupNorm(x))) + bias for <= 16 output channels (out.0-2, unet.py:617-621): the TMA
    fetches the raw tensor, activator warps rewrite each halo box in place, fp32 NHWC output."""
    from diffpir_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, C, H, W, generator=g) * 1.7 + 0.6 * torch.randn(B, C, 1, 1, generator=g)
    xh = x.half()
    w = torch.randn(Cout, C, 3, 3, generator=g) / (C * 9) ** 0.5
    b = torch.randn(Cout, generator=g) * 0.1
    gamma = 1 + 0.3 * torch.randn(C, generator=g)
    beta = 0.2 * torch.randn(C, generator=g)
    ref = F.conv2d(_gn_reference(xh.float(), gamma, beta, None), w.half().double(), b.double(), padding=1)
    s1 = xh.permute(0, 2, 3, 1).contiguous().cuda()
    q1 = _quad_stats(s1.cpu()).cuda()
    coef = torch.empty(B, 2, C, device="cuda")
    gd, bd, biasd = gamma.cuda(), beta.cuda(), b.cuda()
    _lib.check(lib.dpir_gn_coef(_lib.ptr(q1), None, C, 0, _lib.ptr(gd), _lib.ptr(bd), None, 2 * C, B, H * W, _lib.ptr(coef),
                                _lib.stream_ptr()), "dpir_gn_coef")
    wp = _pack_tc(w, 16).cuda()
    out = torch.full((B, H, W, Cout), float("nan"), device="cuda")
    _lib.check(lib.dpir_conv_gnfuse_f16io(_lib.ptr(s1), C, None, 0, _lib.ptr(coef), _lib.ptr(wp), _lib.ptr(biasd), None, _lib.ptr(out),
                                          B, H, W, Cout, 16, _lib.stream_ptr()), "dpir_conv_gnfuse_f16io")
    torch.cuda.synchronize()
    got = out.permute(0, 3, 1, 2)
    assert not torch.isnan(got).any()
    assert rel_l2(got, ref) < 1.2e-3      # fp16 rounding of the activated operand
