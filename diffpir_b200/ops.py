"""Elementwise steps that are inline in the reference drivers, as single fused CUDA kernels.

    inpaint_prox  - main_ddpir.py:393-394, main_ddpir_inpainting.py:267-268
    renoise       - main_ddpir.py:451-456 (same formulas in main_ddpir_deblur.py:339-347, _sisr.py:364-372,
                    _inpainting.py:285-293)
    eps_to_x0     - guided_diffusion/gaussian_diffusion.py:328-333 (+ clamp :305-307)
"""
import torch

from . import _lib


def _f32c(*ts):
    out = []
    for t in ts:
        _lib.require_cuda(t)
        out.append(t if (t is None or (t.dtype == torch.float32 and t.is_contiguous())) else t.float().contiguous())
    return out


def affine(x, a, b, out=None):
    (x,) = _f32c(x)
    out = torch.empty_like(x) if out is None else out
    _lib.check(_lib.load().dpir_affine(_lib.ptr(x), float(a), float(b), _lib.ptr(out), x.numel(), _lib.stream_ptr()), "dpir_affine")
    return out


def inpaint_prox(x0, y, mask, tau, guidance_scale=1.0, y_scale=2.0, y_bias=-1.0, out=None):
    """x0_p = (mask*(y_scale*y+y_bias) + tau*x0)/(mask+tau); x0 + guidance*(x0_p-x0).
    Defaults follow main_ddpir.py:393 (y in [0,1] -> 2y-1); main_ddpir_inpainting.py:267 passes y_scale=1, y_bias=0."""
    x0, y, mask = _f32c(x0, y, mask)
    if mask.shape != x0.shape:
        mask = mask.expand_as(x0).contiguous()
    out = torch.empty_like(x0) if out is None else out
    tau_dev = _lib.device_scalar(tau, x0.device)
    if tau_dev is not None:        # tau lives on the GPU (main_ddpir.py:389): read it in the kernel, no host sync
        _lib.check(_lib.load().dpir_inpaint_prox_dev(_lib.ptr(x0), _lib.ptr(y), _lib.ptr(mask), _lib.ptr(tau_dev), float(y_scale),
                                                     float(y_bias), float(guidance_scale), _lib.ptr(out), x0.numel(),
                                                     _lib.stream_ptr()), "dpir_inpaint_prox_dev")
    else:
        tau = float(tau.reshape(-1)[0]) if torch.is_tensor(tau) else float(tau)
        _lib.check(_lib.load().dpir_inpaint_prox(_lib.ptr(x0), _lib.ptr(y), _lib.ptr(mask), tau, float(y_scale), float(y_bias),
                                                 float(guidance_scale), _lib.ptr(out), x0.numel(), _lib.stream_ptr()),
                   "dpir_inpaint_prox")
    return out


def renoise(x, x0, noise_zeta, sa_t, s1_t, sa_p, c_eps, c_zeta, noise_eta=None, c_eta=0.0, out=None):
    """eps = (x - sa_t*x0)/s1_t; out = sa_p*x0 + c_eps*eps + c_eta*noise_eta + c_zeta*noise_zeta."""
    x, x0, noise_zeta, noise_eta = _f32c(x, x0, noise_zeta, noise_eta)
    out = torch.empty_like(x) if out is None else out
    _lib.check(_lib.load().dpir_renoise(_lib.ptr(x), _lib.ptr(x0), _lib.ptr(noise_eta), _lib.ptr(noise_zeta), float(sa_t),
                                        float(s1_t), float(sa_p), float(c_eps), float(c_eta), float(c_zeta), _lib.ptr(out),
                                        x.numel(), _lib.stream_ptr()), "dpir_renoise")
    return out


def eps_to_x0(model_out, x, coef, clip=True, out=None):
    """model_out (B,Cout>=3,H,W); coef (B,2) = (sqrt_recip_alphas_cumprod[t], sqrt_recipm1_alphas_cumprod[t])."""
    model_out, x, coef = _f32c(model_out, x, coef)
    B, Cout, H, W = model_out.shape
    out = torch.empty_like(x) if out is None else out
    _lib.check(_lib.load().dpir_eps_to_x0(_lib.ptr(model_out), _lib.ptr(x), _lib.ptr(coef), _lib.ptr(out), B, Cout, H * W,
                                          int(clip), _lib.stream_ptr()), "dpir_eps_to_x0")
    return out
