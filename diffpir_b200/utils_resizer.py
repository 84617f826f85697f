"""GPU stand-in for the reference's separable antialiased resizer (utils/utils_resizer.py:8-178).

Same constructor and call surface: Resizer(in_shape, scale_factor=None, output_shape=None, kernel=None,
antialiasing=True), `.to(device)`, `.forward(x)` / `resizer(x)` on NCHW tensors.  The weight / field-of-view
tables are built on the host at init exactly like Resizer.contributions (:104-167); forward is two CUDA
gather passes (dpir_resize_separable).  Only the last two dims may be resized (what the drivers do:
main_ddpir.py:294, main_ddpir_sisr.py:209).
"""
from math import pi

import numpy as np
import torch

from . import _lib


def _cubic(x):
    a = np.abs(x)
    a2, a3 = a ** 2, a ** 3
    return ((1.5 * a3 - 2.5 * a2 + 1) * (a <= 1) + (-0.5 * a3 + 2.5 * a2 - 4 * a + 2) * ((1 < a) & (a <= 2)))


def _lanczos2(x):
    return (((np.sin(pi * x) * np.sin(pi * x / 2) + np.finfo(np.float32).eps) /
             ((pi ** 2 * x ** 2 / 2) + np.finfo(np.float32).eps)) * (abs(x) < 2))


def _lanczos3(x):
    return (((np.sin(pi * x) * np.sin(pi * x / 3) + np.finfo(np.float32).eps) /
             ((pi ** 2 * x ** 2 / 3) + np.finfo(np.float32).eps)) * (abs(x) < 3))


def _box(x):
    return ((-0.5 <= x) & (x < 0.5)) * 1.0


def _linear(x):
    return (x + 1) * ((-1 <= x) & (x < 0)) + (1 - x) * ((0 <= x) & (x <= 1))


_METHODS = {"cubic": (_cubic, 4.0), "lanczos2": (_lanczos2, 4.0), "lanczos3": (_lanczos3, 6.0), "box": (_box, 1.0),
            "linear": (_linear, 2.0), None: (_cubic, 4.0)}


def contributions(in_length, out_length, scale, kernel, kernel_width, antialiasing):
    """Per-dimension taps: returns (weights [out, taps], field_of_view [out, taps]); utils_resizer.py:104-167."""
    fixed_kernel = (lambda arg: scale * kernel(scale * arg)) if antialiasing else kernel
    kernel_width *= 1.0 / scale if antialiasing else 1.0
    out_coordinates = np.arange(1, out_length + 1)
    shifted = out_coordinates - (out_length - in_length * scale) / 2
    match = shifted / scale + 0.5 * (1 - 1 / scale)
    left = np.floor(match - kernel_width / 2)
    expanded = np.ceil(kernel_width) + 2
    fov = np.squeeze(np.int16(np.expand_dims(left, axis=1) + np.arange(expanded) - 1))
    weights = fixed_kernel(1.0 * np.expand_dims(match, axis=1) - fov - 1)
    sw = np.sum(weights, axis=1)
    sw[sw == 0] = 1.0
    weights = 1.0 * weights / np.expand_dims(sw, axis=1)
    mirror = np.uint(np.concatenate((np.arange(in_length), np.arange(in_length - 1, -1, step=-1))))
    fov = mirror[np.mod(fov, mirror.shape[0])]
    nz = np.nonzero(np.any(weights, axis=0))
    weights = np.squeeze(weights[:, nz])
    fov = np.squeeze(fov[:, nz])
    return weights, fov


class Resizer:
    def __init__(self, in_shape, scale_factor=None, output_shape=None, kernel=None, antialiasing=True):
        in_shape = tuple(int(s) for s in in_shape)
        nd = len(in_shape)
        if scale_factor is not None:
            if np.isscalar(scale_factor) and nd > 1:
                scale_factor = [scale_factor, scale_factor]
            scale_factor = list(scale_factor)
            scale_factor = [1] * (nd - len(scale_factor)) + scale_factor
        if output_shape is not None:
            output_shape = list(in_shape[len(output_shape):]) + list(np.uint(np.array(output_shape)))
        if scale_factor is None:
            scale_factor = 1.0 * np.array(output_shape) / np.array(in_shape)
        if output_shape is None:
            output_shape = np.uint(np.ceil(np.array(in_shape) * np.array(scale_factor)))
        if any(scale_factor[d] != 1 for d in range(nd - 2)):
            raise NotImplementedError("only the last two dims can be resized")
        method, kernel_width = _METHODS[kernel]
        antialiasing = antialiasing * np.any(np.array(scale_factor) < 1)
        self.in_shape = in_shape
        self.output_shape = tuple(int(s) for s in output_shape)
        self._tabs = {}
        for d in (nd - 2, nd - 1):
            if scale_factor[d] == 1:
                n = in_shape[d]
                w = np.ones((1, n), np.float32)
                f = np.arange(n, dtype=np.int32)[None]
            else:
                wts, fov = contributions(in_shape[d], int(output_shape[d]), scale_factor[d], method, kernel_width, antialiasing)
                w = np.ascontiguousarray(wts.T.astype(np.float32))          # [taps, out], float32 as utils_resizer.py:46
                f = np.ascontiguousarray(fov.T.astype(np.int32))
            self._tabs[d - (nd - 2)] = (torch.from_numpy(w.copy()), torch.from_numpy(f.copy()))
        self.device = torch.device("cpu")

    def to(self, device):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("diffpir_b200.Resizer runs on the GPU only (no CPU fallback)")
        self._tabs = {k: (w.to(device).contiguous(), f.to(device).contiguous()) for k, (w, f) in self._tabs.items()}
        self.device = device
        return self

    def forward(self, in_tensor):
        _lib.require_cuda(in_tensor)
        if self.device.type != "cuda":
            self.to(in_tensor.device)
        x = in_tensor.float().contiguous()
        Hin, Win = x.shape[-2:]
        (wh, fh), (ww, fw) = self._tabs[0], self._tabs[1]
        Hout, Wout = wh.shape[1], ww.shape[1]
        N = x.numel() // (Hin * Win)
        tmp = torch.empty(N * Hout * Win, device=x.device, dtype=torch.float32)
        out = torch.empty(*x.shape[:-2], Hout, Wout, device=x.device, dtype=torch.float32)
        _lib.check(_lib.load().dpir_resize_separable(_lib.ptr(x), N, Hin, Win, Hout, Wout, _lib.ptr(wh), _lib.ptr(fh),
                                                     wh.shape[0], _lib.ptr(ww), _lib.ptr(fw), ww.shape[0], _lib.ptr(tmp),
                                                     _lib.ptr(out), _lib.stream_ptr()), "dpir_resize_separable")
        return out

    __call__ = forward


def ibp_update(x0, y, down, gamma, rho, sf, out=None):
    """One IBP inner iteration, main_ddpir.py:403-406 (nearest up-sampling of the low-res residual)."""
    _lib.require_cuda(x0, y, down)
    x0 = x0.float().contiguous(); y = y.float().contiguous(); down = down.float().contiguous()
    H, W = x0.shape[-2:]
    N = x0.numel() // (H * W)
    out = torch.empty_like(x0) if out is None else out
    rho_dev = _lib.device_scalar(rho, x0.device)
    if rho_dev is not None:        # rho lives on the GPU (main_ddpir.py:389): read it in the kernel, no host sync
        _lib.check(_lib.load().dpir_ibp_update_dev(_lib.ptr(x0), _lib.ptr(y), _lib.ptr(down), float(gamma), _lib.ptr(rho_dev), N, H,
                                                   W, sf, _lib.ptr(out), _lib.stream_ptr()), "dpir_ibp_update_dev")
    else:
        _lib.check(_lib.load().dpir_ibp_update(_lib.ptr(x0), _lib.ptr(y), _lib.ptr(down), float(gamma), float(rho), N, H, W, sf,
                                               _lib.ptr(out), _lib.stream_ptr()), "dpir_ibp_update")
    return out
