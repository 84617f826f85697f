"""GPU versions of the edges of the sampling loop, with the reference's names (utils/utils_image.py) where they exist:

    tensor2uint_batch(img)                              :238-242   clamp -> uint8 NHWC (returned as numpy, like the reference)
    calculate_psnr_batch(batch1, batch2, max_pixel, eps) :601-610
    calculate_psnr_y_batch(...)   = calculate_psnr_batch(rgb2ycbcr_batch(a, True), rgb2ycbcr_batch(b, True))  (main_ddpir.py:512-515)
    blur_measurement_batch(...)   = the uint8 circular blur + AWGN of CustomDataset.__getitem__ (main_ddpir.py:97-114)
    imresize(img, scale, antialiasing) / imresize_np(...)  :748-881   MATLAB-style bicubic (the SR measurement of
                                                            main_ddpir.py:89 / main_ddpir_sisr.py:202) on the GPU
"""
import math

import numpy as np
import torch

from . import _lib


def tensor2uint_batch(img):
    _lib.require_cuda(img)
    x = img.detach().float().contiguous()
    if x.dim() != 4:
        raise ValueError("expected (B,C,H,W)")
    B, C, H, W = x.shape
    out = torch.empty(B, H, W, C, device=x.device, dtype=torch.uint8)
    _lib.check(_lib.load().dpir_tensor2uint(_lib.ptr(x), _lib.ptr(out), B, C, H, W, _lib.stream_ptr()), "dpir_tensor2uint")
    return out.cpu().numpy()


def _sse(a, b, luma, a_scale=1.0, a_bias=0.0, b_scale=1.0, b_bias=0.0):
    _lib.require_cuda(a, b)
    if a.shape != b.shape:
        raise ValueError('Input images must have the same dimensions.')
    a = a.detach().float().contiguous()
    b = b.detach().float().contiguous()
    B, C = a.shape[:2]
    HW = a[0, 0].numel()
    sse = torch.empty(B, device=a.device, dtype=torch.float64)
    _lib.check(_lib.load().dpir_sqdiff_batch(_lib.ptr(a), _lib.ptr(b), a_scale, a_bias, b_scale, b_bias, int(luma), B, C, HW,
                                             _lib.ptr(sse), _lib.stream_ptr()), "dpir_sqdiff_batch")
    return sse, B, C, HW


def _psnr_from_sse(sse, n, max_pixel, eps):
    mse = (sse / n).float()
    psnr = torch.where(mse == 0, torch.full_like(mse, float('inf')), 20 * torch.log10(max_pixel / torch.sqrt(mse + eps)))
    psnr = torch.where(torch.isnan(psnr), torch.zeros_like(psnr), psnr)
    return torch.mean(psnr).item()


def calculate_psnr_batch(batch1, batch2, max_pixel=2.0, eps=1e-10):
    sse, B, C, HW = _sse(batch1, batch2, luma=False)
    return _psnr_from_sse(sse, C * HW, max_pixel, eps)


def calculate_psnr_y_batch(batch1, batch2, max_pixel=2.0, eps=1e-10):
    """PSNR on the Y channel as main_ddpir.py:512-515 computes it: rgb2ycbcr_batch(only_y=True) leaves channels 1,2 zero,
    so the mean runs over all 3*H*W entries."""
    sse, B, C, HW = _sse(batch1, batch2, luma=True)
    return _psnr_from_sse(sse, C * HW, max_pixel, eps)


def blur_measurement_batch(img_u8, k, noise=None, return_u8=False):
    """img_u8: (B,H,W,3) uint8 (numpy or tensor); k: (B,kh,kw) or (kh,kw) blur kernels; noise: (B,H,W,3) float AWGN in [-1,1]
    units (np.random.normal(0, 2*sigma, ...) drawn by the caller, as the reference does on the host) or None.
    Returns y (B,3,H,W) float32 on the GPU in ~[0,1]."""
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device())
    img = torch.as_tensor(img_u8).to(dev).contiguous()
    if img.dtype != torch.uint8 or img.dim() != 4:
        raise ValueError("img_u8 must be (B,H,W,C) uint8")
    B, H, W, C = img.shape
    kt = torch.as_tensor(np.asarray(k, dtype=np.float64))
    if kt.dim() == 2:
        kt = kt[None].repeat(B, 1, 1)
    kt = kt.to(dev).contiguous()
    nz = None
    if noise is not None:
        nz = torch.as_tensor(np.asarray(noise, dtype=np.float32)).to(dev).contiguous()
    out = torch.empty(B, C, H, W, device=dev, dtype=torch.float32)
    u8 = torch.empty_like(img) if return_u8 else None
    _lib.check(lib.dpir_measure_blur_u8(_lib.ptr(img), _lib.ptr(kt), _lib.ptr(nz), _lib.ptr(u8), _lib.ptr(out), B, H, W, C,
                                        kt.shape[-2], kt.shape[-1], _lib.stream_ptr()), "dpir_measure_blur_u8")
    return (out, u8) if return_u8 else out


# ----------------------------------------------------------------------------------------------- MATLAB-style imresize
def _cubic_taps(d):
    a = np.abs(d)
    a2, a3 = a * a, a * a * a
    return ((1.5 * a3 - 2.5 * a2 + 1) * (a <= 1) + (-0.5 * a3 + 2.5 * a2 - 4 * a + 2) * ((a > 1) & (a <= 2))).astype(np.float32)


def imresize_tables(in_length, out_length, scale, antialiasing=True):
    """Gather form of calculate_weights_indices + the symmetric padding of imresize_np (utils_image.py:680-731, 811-881):
    returns (weights [taps, out] float32, source index [taps, out] int32) such that
    out[o] = sum_t weights[t, o] * in[index[t, o]] - the table layout dpir_resize_separable consumes.
    Host logic (numpy float32, like the reference's float32 torch arithmetic)."""
    f32 = np.float32
    kw = f32(4.0)
    shrink = scale < 1 and antialiasing
    if shrink:
        kw = kw / f32(scale)
    x = np.linspace(1, out_length, out_length, dtype=f32)
    u = x / f32(scale) + f32(0.5 * (1 - 1 / scale))
    left = np.floor(u - kw / f32(2))
    P = int(math.ceil(float(kw))) + 2
    idx = left[:, None] + np.arange(P, dtype=f32)[None, :]                 # 1-based source coordinates
    dist = u[:, None] - idx
    w = (f32(scale) * _cubic_taps(dist * f32(scale))) if shrink else _cubic_taps(dist)
    w = w / w.sum(1, keepdims=True, dtype=f32)
    zeros = (w == 0).sum(0)
    if zeros[0] != 0:                      # :721-723 (any zero in the leading column drops it)
        idx, w = idx[:, 1:P - 1], w[:, 1:P - 1]
    if zeros[-1] != 0:                     # :724-726 (a no-op when the leading column was dropped: kept as in the reference)
        idx, w = idx[:, 0:P - 2], w[:, 0:P - 2]
    taps = w.shape[1]
    # the reference reads a CONTIGUOUS window of the symmetric-padded signal starting at the first tap's position;
    # padded position q -> source index: mirror without repeating the border pixel twice... the pad is img[:s][::-1] in front
    # and img[-e:][::-1] behind, i.e. plain reflection including the edge sample
    sym_s = int(-idx.min() + 1)
    first = (idx[:, 0] + sym_s - 1).astype(np.int64)                       # 0-based start in the padded signal
    q = first[:, None] + np.arange(taps)[None, :]
    src = np.where(q < sym_s, sym_s - 1 - q, np.where(q < sym_s + in_length, q - sym_s, 2 * in_length + sym_s - 1 - q))
    if src.min() < 0 or src.max() >= in_length:
        raise ValueError("imresize: tap window leaves the symmetric padding (unsupported scale / size)")
    return np.ascontiguousarray(w.T.astype(np.float32)), np.ascontiguousarray(src.T.astype(np.int32))


_imresize_cache = {}


def imresize(img, scale, antialiasing=True):
    """img: CUDA float tensor (..., H, W) in any layout with H, W last (the reference takes CHW / HW: utils_image.py:748);
    returns (..., ceil(H*scale), ceil(W*scale)) float32, H pass first, then W, as the reference."""
    _lib.require_cuda(img)
    x = img.detach().float().contiguous()
    Hin, Win = x.shape[-2:]
    Hout, Wout = math.ceil(Hin * scale), math.ceil(Win * scale)
    key = (Hin, Win, float(scale), bool(antialiasing), x.device)
    tabs = _imresize_cache.get(key)
    if tabs is None:
        wh, fh = imresize_tables(Hin, Hout, scale, antialiasing)
        ww, fw = imresize_tables(Win, Wout, scale, antialiasing)
        tabs = tuple(torch.from_numpy(t).to(x.device) for t in (wh, fh, ww, fw))
        _imresize_cache[key] = tabs
    wh, fh, ww, fw = tabs
    N = x.numel() // (Hin * Win)
    tmp = torch.empty(N * Hout * Win, device=x.device, dtype=torch.float32)
    out = torch.empty(*x.shape[:-2], Hout, Wout, device=x.device, dtype=torch.float32)
    _lib.check(_lib.load().dpir_resize_separable(_lib.ptr(x), N, Hin, Win, Hout, Wout, _lib.ptr(wh), _lib.ptr(fh), wh.shape[0],
                                                 _lib.ptr(ww), _lib.ptr(fw), ww.shape[0], _lib.ptr(tmp), _lib.ptr(out),
                                                 _lib.stream_ptr()), "dpir_resize_separable")
    return out


def imresize_np(img, scale, antialiasing=True):
    """The reference's numpy entry point (HWC or HW float in [0,1] -> numpy, utils_image.py:811): computed on the GPU."""
    a = np.asarray(img, dtype=np.float32)
    t = torch.from_numpy(a if a.ndim == 2 else np.ascontiguousarray(a.transpose(2, 0, 1))).cuda()
    o = imresize(t, scale, antialiasing).cpu().numpy()
    return o if a.ndim == 2 else np.ascontiguousarray(o.transpose(1, 2, 0))
