"""Native stand-in for the part of the reference `diffusion` object that utils_model.model_fn reaches:

    GaussianDiffusion.p_sample      guided_diffusion/gaussian_diffusion.py:395-439  (via p_mean_variance :232-326)
    GaussianDiffusion.ddim_sample   guided_diffusion/gaussian_diffusion.py:537-585
    SpacedDiffusion / _WrappedModel guided_diffusion/respace.py:63-128 (timestep_map, rescale_timesteps)

for the configuration DiffPIR builds (utils/utils_model.py:354-383: learn_sigma=True -> LEARNED_RANGE variance,
predict_xstart=False -> EPSILON mean, linear schedule).  One UNet forward (tape of CUDA kernels) + ONE elementwise kernel
(dpir_sample_step) replace the ~25 ATen launches of the reference; tables are the reference's float64 numpy tables cast to
float exactly like _extract_into_tensor (:895-908).  Noise is drawn with torch.randn_like on the device like the reference
does, or injected (`noise=`) for bit-reproducible parity runs.
"""
import numpy as np
import torch

from . import _lib


def linear_betas(T=1000):
    """get_named_beta_schedule('linear', T), gaussian_diffusion.py:28-38."""
    scale = 1000.0 / T
    return np.linspace(scale * 0.0001, scale * 0.02, T, dtype=np.float64)


class GaussianDiffusion:
    """`p_sample(model, x, t, ...)` / `ddim_sample(model, x, t, ..., eta=0.0)` -> {'sample', 'pred_xstart'}."""

    def __init__(self, betas=None, timestep_map=None, rescale_timesteps=False, original_num_steps=None):
        betas = linear_betas() if betas is None else np.array(betas, dtype=np.float64)
        self.betas = betas
        self.num_timesteps = T = int(betas.shape[0])
        self.timestep_map = None if timestep_map is None else [int(v) for v in timestep_map]
        self.rescale_timesteps = bool(rescale_timesteps)
        self.original_num_steps = original_num_steps or (max(self.timestep_map) + 1 if self.timestep_map else T)
        alphas = 1.0 - betas
        ac = np.cumprod(alphas, axis=0)                                       # :142-144
        ac_prev = np.append(1.0, ac[:-1])
        self.alphas_cumprod, self.alphas_cumprod_prev = ac, ac_prev
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / ac)                     # :150-151
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / ac - 1)
        post_var = betas * (1.0 - ac_prev) / (1.0 - ac)                        # :154-167
        self.posterior_log_variance_clipped = np.log(np.append(post_var[1], post_var[1:]))
        self.posterior_mean_coef1 = betas * np.sqrt(ac_prev) / (1.0 - ac)
        self.posterior_mean_coef2 = (1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac)
        self._tabs = {}

    @classmethod
    def from_reference(cls, diffusion):
        """Mirror a reference GaussianDiffusion / SpacedDiffusion object (same betas, same timestep map); raises
        NotImplementedError for model types DiffPIR never builds."""
        mv, mm = str(getattr(diffusion, "model_var_type", "")), str(getattr(diffusion, "model_mean_type", ""))
        if "LEARNED_RANGE" not in mv or "EPSILON" not in mm:
            raise NotImplementedError(f"native diffusion covers LEARNED_RANGE/EPSILON models only (got {mv}, {mm})")
        return cls(betas=diffusion.betas, timestep_map=getattr(diffusion, "timestep_map", None),
                   rescale_timesteps=getattr(diffusion, "rescale_timesteps", False),
                   original_num_steps=getattr(diffusion, "original_num_steps", None))

    # ---- tables
    def _table(self, device, eta):
        key = (str(device), float(eta))
        tab = self._tabs.get(key)
        if tab is None:
            T = self.num_timesteps
            f32 = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float64)).float()     # noqa: E731  (_extract_into_tensor :905)
            ab, abp = f32(self.alphas_cumprod), f32(self.alphas_cumprod_prev)
            sigma = eta * torch.sqrt((1 - abp) / (1 - ab)) * torch.sqrt(1 - ab / abp)     # :569-573, float32 like the reference
            cols = [f32(self.sqrt_recip_alphas_cumprod), f32(self.sqrt_recipm1_alphas_cumprod), f32(self.posterior_mean_coef1),
                    f32(self.posterior_mean_coef2), f32(self.posterior_log_variance_clipped), f32(np.log(self.betas)),
                    torch.sqrt(abp), torch.sqrt(1 - abp - sigma ** 2), sigma, torch.zeros(T)]
            tab = torch.stack(cols, dim=1).contiguous().to(device)
            self._tabs[key] = tab
        return tab

    def model_timesteps(self, t):
        """What the wrapped model is called with: map_tensor[ts] (+ rescale), respace.py:123-128."""
        if self.timestep_map is not None:
            mt = torch.tensor(self.timestep_map, device=t.device, dtype=t.dtype)
            t = mt[t]
        if self.rescale_timesteps:
            t = (t.float() * (1000.0 / self.original_num_steps)).long()
        return t

    def _step(self, mode, model, x, t, clip_denoised, noise, eta, want_sample=True):
        _lib.require_cuda(x)
        x = x.float().contiguous()
        t = t.to(device=x.device, dtype=torch.int64).reshape(-1).contiguous()
        B, C, H, W = x.shape
        if C != 3 or t.numel() != B:
            raise ValueError("p_sample/ddim_sample: x must be (B,3,H,W) with one timestep per sample")
        mo = model(x, self.model_timesteps(t))                                # (B,6,H,W): eps | learned-range variance values
        if noise is None and want_sample:
            noise = torch.randn_like(x)                                       # :430 / :575
        x0 = torch.empty_like(x)
        sample = torch.empty_like(x) if want_sample else None
        tab = self._table(x.device, eta)
        _lib.check(_lib.load().dpir_sample_step(_lib.ptr(mo), _lib.ptr(x), _lib.ptr(noise), _lib.ptr(t), _lib.ptr(tab),
                                                self.num_timesteps, _lib.ptr(sample), _lib.ptr(x0), B, mo.shape[1], H * W, mode,
                                                int(bool(clip_denoised)), _lib.stream_ptr()), "dpir_sample_step")
        return {"sample": sample, "pred_xstart": x0}

    def p_sample(self, model, x, t, clip_denoised=True, denoised_fn=None, cond_fn=None, model_kwargs=None, noise=None):
        if denoised_fn is not None or cond_fn is not None or model_kwargs:
            raise NotImplementedError("denoised_fn / cond_fn / model_kwargs are outside the DiffPIR path")
        return self._step(0, model, x, t, clip_denoised, noise, 0.0)

    def ddim_sample(self, model, x, t, clip_denoised=True, denoised_fn=None, cond_fn=None, model_kwargs=None, eta=0.0, noise=None):
        if denoised_fn is not None or cond_fn is not None or model_kwargs:
            raise NotImplementedError("denoised_fn / cond_fn / model_kwargs are outside the DiffPIR path")
        return self._step(1, model, x, t, clip_denoised, noise, eta)
