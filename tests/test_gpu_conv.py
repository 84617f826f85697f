"""Per-kernel parity of the two convolution kernels (tcgen05 fp16 and CUDA-core fp32) through the C-ABI."""
import ctypes

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


def test_conv_tc_small_cout_pad16():
    from diffpir_b200 import _lib
    lib = _lib.load()
    B, H, W, Cin, Cout = 2, 32, 32, 128, 8
    x, w, b, _ = _case(B, H, W, Cin, Cout, 3, 1, False)
    ref = F.conv2d(x.half().double(), w.half().double(), b.double(), padding=1)
    act = x.permute(0, 2, 3, 1).contiguous().half().cuda()
    wp = _pack_tc(w, 16).cuda()
    out = torch.full((B, H, W, Cout), float("nan"), device="cuda")
    _lib.check(lib.dpir_conv_tc(_lib.ptr(act), _lib.ptr(wp), _lib.ptr(b.cuda()), None, _lib.ptr(out), B, H, W, Cin, Cout, 16, 3,
                                _lib.stream_ptr()), "dpir_conv_tc")
    torch.cuda.synchronize()
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
