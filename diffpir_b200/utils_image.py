"""GPU versions of the edges of the sampling loop, with the reference's names (utils/utils_image.py) where they exist:

    tensor2uint_batch(img)                              :238-242   clamp -> uint8 NHWC (returned as numpy, like the reference)
    calculate_psnr_batch(batch1, batch2, max_pixel, eps) :601-610
    calculate_psnr_y_batch(...)   = calculate_psnr_batch(rgb2ycbcr_batch(a, True), rgb2ycbcr_batch(b, True))  (main_ddpir.py:512-515)
    blur_measurement_batch(...)   = the uint8 circular blur + AWGN of CustomDataset.__getitem__ (main_ddpir.py:97-114)
"""
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
