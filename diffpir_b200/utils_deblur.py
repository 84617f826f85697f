"""Blur operator objects with the reference's H / H^T convention (utils/utils_deblur.py:641-727):

    op = GaussialBlurOperator(kernel_size, intensity, device)   # sic: the reference's spelling
    op.forward(data)      # H: ReflectionPad2d(k//2) + depthwise k x k correlation           (:653-655, :697-699)
    op.transpose(data)    # the reference returns `data` unchanged (NOT a true adjoint)        (:700-701, :723-724)
    op.get_kernel()       # (1,1,k,k) float32                                                  (:703-706, :726-727)

The Gaussian kernel is built exactly like Blurkernel.weights_init (:658-664).  The reference's motion kernels
come from the un-vendored `motionblur` package (:7, :668); MotionBlurOperator here uses our own seeded
random-walk generator (diffpir_b200.synth.motion_kernel) - same interface, different random kernels.
"""
import numpy as np
import torch

from . import _lib, synth


class _BlurBase:
    def __init__(self, kernel_size, device):
        self.kernel_size = kernel_size
        self.device = torch.device(device)
        self.kernel = None

    def forward(self, data, **kwargs):
        _lib.require_cuda(data)
        x = data.float().contiguous()
        k = self.kernel.to(device=x.device, dtype=torch.float32).contiguous()
        H, W = x.shape[-2:]
        out = torch.empty_like(x)
        _lib.check(_lib.load().dpir_blur_reflect(_lib.ptr(x), _lib.ptr(k), x.numel() // (H * W), H, W, k.shape[-2], k.shape[-1],
                                                 _lib.ptr(out), _lib.stream_ptr()), "dpir_blur_reflect")
        return out

    def transpose(self, data, **kwargs):
        return data

    def get_kernel(self):
        return self.kernel.view(1, 1, self.kernel_size, self.kernel_size)


class GaussialBlurOperator(_BlurBase):
    def __init__(self, kernel_size, intensity, device):
        super().__init__(kernel_size, device)
        # float64, as Blurkernel.weights_init / get_kernel return it (the drivers cast: main_ddpir_deblur.py:168)
        self.kernel = torch.from_numpy(synth.gaussian_kernel(kernel_size, intensity, dtype=np.float64))


GaussianBlurOperator = GaussialBlurOperator


class MotionBlurOperator(_BlurBase):
    def __init__(self, kernel_size, intensity, device, seed=None):
        super().__init__(kernel_size, device)
        seed = int(np.random.randint(0, 2 ** 31 - 1)) if seed is None else seed
        self.kernel = torch.tensor(synth.motion_kernel(kernel_size, intensity, seed), dtype=torch.float32)
