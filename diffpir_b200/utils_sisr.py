"""GPU stand-ins for the FFT data-fidelity solver of the reference, same names and call signatures:

    reference utils/utils_sisr.py:78-95  pre_calculate(x, k, sf) -> (FB, FBC, F2B, FBFy)
    reference utils/utils_sisr.py:65-75  data_solution(x, FB, FBC, F2B, FBFy, alpha, sf) -> Xest

The 4-tuple is opaque to the drivers (main_ddpir.py:318-320, 396-398), so it carries Hermitian
half-spectra (H, W/2+1) instead of the reference's full complex spectra; both functions agree on it.
`data_solution_fused` additionally folds the loop's x0/2+.5, *2-1 and guidance blend
(main_ddpir.py:396-400) into the same three launches (cuFFT R2C -> one frequency kernel -> cuFFT C2R).
"""
import ctypes

import torch

from . import _lib

_plans = {}


def _plan(batch, H, W, device):
    key = (batch, H, W, device.index)
    p = _plans.get(key)
    if p is None:
        h = ctypes.c_void_p()
        _lib.check(_lib.load().dpir_fft_plan_create(batch, H, W, ctypes.byref(h)), "dpir_fft_plan_create")
        _plans[key] = p = h
    return p


def pre_calculate(x, k, sf):
    """x: (B,C,H/sf,W/sf) LR measurement in [0,1]; k: (Bk,1,kh,kw), Bk in {1,B}."""
    _lib.require_cuda(x, k)
    lib = _lib.load()
    x = x.float().contiguous()
    k = k.float().contiguous()
    B, C, h, w = x.shape
    H, W = h * sf, w * sf
    Wh = W // 2 + 1
    Bk = k.shape[0]
    dev = x.device
    FB = torch.empty(Bk, 1, H, Wh, 2, device=dev, dtype=torch.float32)
    F2B = torch.empty(Bk, 1, H, Wh, device=dev, dtype=torch.float32)
    FBFy = torch.empty(B, C, H, Wh, 2, device=dev, dtype=torch.float32)
    scratch = torch.empty(max(Bk, B * C) * H * W, device=dev, dtype=torch.float32)
    _lib.check(lib.dpir_fft_precalc(_plan(Bk, H, W, dev), _plan(B * C, H, W, dev), _lib.ptr(x), _lib.ptr(k), B, C, Bk,
                                    k.shape[-2], k.shape[-1], sf, _lib.ptr(FB), _lib.ptr(F2B), _lib.ptr(FBFy),
                                    _lib.ptr(scratch), _lib.stream_ptr()), "dpir_fft_precalc")
    FBc = torch.view_as_complex(FB)
    return FBc, torch.conj(FBc), F2B, torch.view_as_complex(FBFy)


def _as_float(alpha):
    return float(alpha.reshape(-1)[0].item()) if torch.is_tensor(alpha) else float(alpha)


def data_solution_fused(x, FB, F2B, FBFy, alpha, sf, in_scale=1.0, in_bias=0.0, out_scale=1.0, out_bias=0.0, blend=1.0,
                        out=None, spec=None):
    _lib.require_cuda(x, FB, F2B, FBFy)
    lib = _lib.load()
    x = x.float().contiguous()
    B, C, H, W = x.shape
    Wh = W // 2 + 1
    Bk = FB.shape[0]
    if out is None:
        out = torch.empty_like(x)
    need = B * C * H * Wh * 2 * (2 if sf > 1 else 1)
    if spec is None or spec.numel() < need:
        spec = torch.empty(need, device=x.device, dtype=torch.float32)
    FBr = torch.view_as_real(FB) if FB.is_complex() else FB
    FYr = torch.view_as_real(FBFy) if FBFy.is_complex() else FBFy
    alpha_dev = _lib.device_scalar(alpha, x.device)
    if alpha_dev is not None:      # alpha lives on the GPU (main_ddpir.py:389): read it in the kernel, no host sync
        _lib.check(lib.dpir_fft_solve_dev(_plan(B * C, H, W, x.device), _lib.ptr(x), _lib.ptr(FBr), _lib.ptr(F2B), _lib.ptr(FYr),
                                          _lib.ptr(alpha_dev), B, C, Bk, sf, in_scale, in_bias, out_scale, out_bias, blend,
                                          _lib.ptr(spec), _lib.ptr(out), _lib.stream_ptr()), "dpir_fft_solve_dev")
    else:
        _lib.check(lib.dpir_fft_solve(_plan(B * C, H, W, x.device), _lib.ptr(x), _lib.ptr(FBr), _lib.ptr(F2B), _lib.ptr(FYr),
                                      _as_float(alpha), B, C, Bk, sf, in_scale, in_bias, out_scale, out_bias, blend,
                                      _lib.ptr(spec), _lib.ptr(out), _lib.stream_ptr()), "dpir_fft_solve")
    return out


def data_solution(x, FB, FBC, F2B, FBFy, alpha, sf):
    """Reference signature and semantics: x -> Xest = argmin ||y - SHx||^2 + alpha ||x - x_in||^2."""
    return data_solution_fused(x, FB, F2B, FBFy, alpha, sf)
