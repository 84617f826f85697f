"""Drop-in stand-in for the reference denoiser object.

Mirrors the surface the DiffPIR drivers use on `guided_diffusion.unet.UNetModel`
(reference main_ddpir.py:231-240; guided_diffusion/respace.py:123-128 calls it as model(x, ts)):
    model = UNetModel(cfg); model.load_state_dict(sd); model.eval(); model.named_parameters();
    model.to(device); out = model(x, timesteps)            # (B,6,H,W) fp32
plus `x0(x, timesteps)` = `diffusion.p_sample(model, x, t)['pred_xstart']` with the eps->x0 step fused
(utils/utils_model.py:221-236, guided_diffusion/gaussian_diffusion.py:258-311).
All compute happens in libdiffpir_b200.so; there is no CPU path.
"""
import ctypes
import math

import numpy as np
import torch

from . import _lib
from .arch import UNetConfig, param_specs


def _coef_table(T):
    """(sqrt_recip_alphas_cumprod, sqrt_recipm1_alphas_cumprod) exactly as gaussian_diffusion.py:133-151
    + get_named_beta_schedule('linear') computes them (float64), cast to float like _extract_into_tensor :905."""
    scale = 1000.0 / T
    betas = np.linspace(scale * 0.0001, scale * 0.02, T, dtype=np.float64)
    ac = np.cumprod(1.0 - betas, axis=0)
    return np.stack([np.sqrt(1.0 / ac), np.sqrt(1.0 / ac - 1)], axis=1).astype(np.float32)


class _CudaMem:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


class UNetModel:
    def __init__(self, cfg: UNetConfig = None, precision="fp16", diffusion_steps=1000, graph=False, **kwargs):
        if cfg is None:
            cfg = UNetConfig(**kwargs)
        self.cfg = cfg
        self.precision = precision
        self.diffusion_steps = diffusion_steps
        _lib.require_cuda()
        lib = _lib.load()
        c = _lib.UNetConfigC()
        c.image_size, c.in_channels, c.model_channels = cfg.image_size, cfg.in_channels, cfg.model_channels
        c.out_channels, c.num_res_blocks = cfg.out_channels, cfg.num_res_blocks
        # fractional multipliers (the 512x512 model's 0.5, script_util.py:149-150) travel as numerators over a common denominator
        den = 1
        while any(abs(m * den - round(m * den)) > 1e-9 for m in cfg.channel_mult) and den < 64:
            den *= 2
        cm = [int(round(m * den)) for m in cfg.channel_mult]
        if any(abs(c_ / den - m) > 1e-9 for c_, m in zip(cm, cfg.channel_mult)):
            raise ValueError(f"channel_mult {cfg.channel_mult} is not representable")
        c.channel_mult_den = den
        c.n_channel_mult = len(cm)
        for i, m in enumerate(cm):
            c.channel_mult[i] = m
        ads = list(cfg.attention_ds)
        c.n_attention_ds = len(ads)
        for i, a in enumerate(ads):
            c.attention_ds[i] = a
        c.num_heads, c.num_head_channels, c.num_heads_upsample = cfg.num_heads, cfg.num_head_channels, cfg.num_heads_upsample
        c.use_new_attention_order = int(cfg.use_new_attention_order)
        c.precision = {"fp32": _lib.PRECISION_FP32, "fp16": _lib.PRECISION_FP16, "fp16s": _lib.PRECISION_FP16S,
                       "fp16x3": _lib.PRECISION_FP16X3}[precision]
        c.diffusion_steps = diffusion_steps
        h = ctypes.c_void_p()
        _lib.check(lib.dpir_create(ctypes.byref(c), ctypes.byref(h)), "dpir_create")
        self._h = h
        self._lib = lib
        self.device = torch.device("cuda", torch.cuda.current_device())
        half = cfg.model_channels // 2
        freqs = torch.exp(-math.log(10000) * torch.arange(start=0, end=half, dtype=torch.float32) / half).contiguous()
        coef = np.ascontiguousarray(_coef_table(diffusion_steps))
        _lib.check(lib.dpir_set_tables(h, ctypes.c_void_p(freqs.data_ptr()), half,
                                       ctypes.c_void_p(coef.ctypes.data), diffusion_steps), "dpir_set_tables")
        self._sd = {}
        self._film_table = True
        if graph:
            self.set_graph_mode(True)

    # ---- nn.Module-like surface used by the drivers
    def load_state_dict(self, sd, strict=True):
        specs = dict(param_specs(self.cfg))
        missing = [k for k in specs if k not in sd]
        unexpected = [k for k in sd if k not in specs]
        if strict and (missing or unexpected):
            raise RuntimeError(f"load_state_dict: missing={missing[:5]} unexpected={unexpected[:5]}")
        for name, shape in specs.items():
            if name not in sd:
                continue
            t = sd[name].detach().to("cpu", torch.float32).contiguous()
            if tuple(t.shape) != tuple(shape) and t.numel() != int(np.prod(shape)):
                raise RuntimeError(f"size mismatch for {name}: {tuple(t.shape)} vs {shape}")
            shp = (ctypes.c_int64 * t.dim())(*t.shape)
            _lib.check(self._lib.dpir_load_weight(self._h, name.encode(), ctypes.c_void_p(t.data_ptr()), shp, t.dim()),
                       f"dpir_load_weight({name})")
            self._sd[name] = t
        return self

    def state_dict(self):
        return dict(self._sd)

    def named_parameters(self):
        for k, v in self._sd.items():
            yield k, v

    def parameters(self):
        for _, v in self.named_parameters():
            yield v

    def eval(self):
        return self

    def train(self, mode=False):
        if mode:
            raise RuntimeError("diffpir_b200.UNetModel is inference-only")
        return self

    def to(self, device=None, *a, **k):
        if device is not None and torch.device(device).type != "cuda":
            raise RuntimeError("diffpir_b200.UNetModel lives on the GPU only (no CPU fallback)")
        return self

    def requires_grad_(self, flag=False):
        return self

    def set_graph_mode(self, enable=True):
        _lib.check(self._lib.dpir_set_graph_mode(self._h, int(enable)), "dpir_set_graph_mode")

    def set_film_table(self, enable=True):
        """True (default): per-timestep FiLM table (integer timesteps in [0, diffusion_steps)); False: evaluate the embedding
        MLPs per call (arbitrary timesteps)."""
        _lib.check(self._lib.dpir_set_film_table(self._h, int(enable)), "dpir_set_film_table")
        self._film_table = bool(enable)

    def _prep(self, x, timesteps):
        _lib.require_cuda(x)
        if x.dtype != torch.float32:
            x = x.float()
        x = x.contiguous()
        B = x.shape[0]
        if not torch.is_tensor(timesteps):
            timesteps = torch.tensor(timesteps)
        if not timesteps.is_cuda and timesteps.numel() and self._film_table:
            # host-side timesteps can be validated for free (the reference raises IndexError); device-side ones are clamped
            # inside the engine before any table lookup
            lo, hi = int(timesteps.min()), int(timesteps.max())
            if lo < 0 or hi >= self.diffusion_steps:
                raise ValueError(f"timesteps must lie in [0, {self.diffusion_steps}) (got {lo}..{hi})")
        t = timesteps.to(device=x.device, dtype=torch.int64).reshape(-1).contiguous()
        if t.numel() == 1 and B > 1:
            t = t.expand(B).contiguous()
        if t.numel() != B:
            raise ValueError("timesteps must have one entry per sample")
        return x, t

    def __call__(self, x, timesteps, y=None, **kwargs):
        """UNetModel.forward (unet.py:634-663); class-conditioning (`y`) is not part of DiffPIR's models."""
        if y is not None:
            raise NotImplementedError("class-conditional models are out of scope")
        x, t = self._prep(x, timesteps)
        B, C, H, W = x.shape
        out = torch.empty(B, self.cfg.out_channels, H, W, device=x.device, dtype=torch.float32)
        _lib.check(self._lib.dpir_unet_forward(self._h, _lib.ptr(x), _lib.ptr(t), _lib.ptr(out), B, H, W, _lib.stream_ptr()),
                   "dpir_unet_forward")
        return out

    forward = __call__

    def x0(self, x, timesteps, out=None):
        x, t = self._prep(x, timesteps)
        B, C, H, W = x.shape
        if out is None:
            out = torch.empty(B, 3, H, W, device=x.device, dtype=torch.float32)
        _lib.check(self._lib.dpir_unet_x0(self._h, _lib.ptr(x), _lib.ptr(t), _lib.ptr(out), B, H, W, _lib.stream_ptr()),
                   "dpir_unet_x0")
        return out

    # ---- multi-GPU plumbing: the packed weight arena as a torch byte tensor (one NCCL broadcast moves the model)
    def weight_arena(self):
        p = ctypes.c_void_p()
        n = ctypes.c_size_t()
        _lib.check(self._lib.dpir_weight_arena(self._h, ctypes.byref(p), ctypes.byref(n)), "dpir_weight_arena")
        return torch.as_tensor(_CudaMem(p.value, n.value), device=self.device)

    def mark_weights_loaded(self):
        _lib.check(self._lib.dpir_mark_weights_loaded(self._h), "dpir_mark_weights_loaded")

    def missing_weights(self):
        return int(self._lib.dpir_missing_weights(self._h))

    def workspace_bytes(self):
        return int(self._lib.dpir_workspace_bytes(self._h))

    # ---- per-op device timing (CUDA events inside the engine)
    OP_KINDS = ("memset", "time_embed", "film", "conv_in", "gn_stats", "gn_apply", "conv", "attention", "gn_coef", "split_cast", "pack_stem")

    def set_gn_fusion(self, enable):
        """A/B switch (this model only): 1/True ResBlock GroupNorm+SiLU inside the convolutions (default: producers that activate
        each element once per chunk), 4 round-1 per-tap-box producers, 2 per-tap-box with packed-half arithmetic, 0/False
        separate passes; drops this model's cached plans."""
        _lib.check(self._lib.dpir_set_gn_fusion(self._h, int(enable)), "dpir_set_gn_fusion")

    def profile_begin(self):
        _lib.check(self._lib.dpir_profile_begin(self._h), "dpir_profile_begin")

    def profile_end(self):
        """-> list of dicts (kind, dims, ms) per tape op, summed over the forwards since profile_begin."""
        cap = 4096
        kind = (ctypes.c_int32 * cap)()
        dims = (ctypes.c_int32 * (6 * cap))()
        ms = (ctypes.c_float * cap)()
        n = self._lib.dpir_profile_end(self._h, kind, dims, ms, cap)
        if n < 0:
            _lib.check(n, "dpir_profile_end")
        rows = []
        for i in range(min(n, cap)):
            d = list(dims[6 * i:6 * i + 6])
            fused = d[4] >= 10          # conv with GroupNorm+SiLU applied inside the kernel (ksize reported + 10)
            d[4] %= 10
            rows.append(dict(kind=self.OP_KINDS[kind[i]], dims=tuple(d), ms=float(ms[i]), gn_fused=bool(fused)))
        return rows

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._lib.dpir_destroy(self._h)
                self._h = None
        except Exception:
            pass
