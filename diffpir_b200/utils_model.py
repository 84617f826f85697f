"""Denoiser wrapper with the reference's names and signatures (utils/utils_model.py):

    find_nearest(array, value)                                   :202-205
    model_fn(x, noise_level, model_diffusion, vec_t=None, model_out_type='pred_xstart',
             diffusion=None, ddim_sample=False, alphas_cumprod=None, **model_kwargs)   :207-258
    test_mode(model_fn, model_diffusion, L, mode, refield, min_size, sf, modulo, noise_level, vec_t,
              model_out_type, diffusion, ddim_sample, alphas_cumprod)                   :16-195

When `model_diffusion` is a diffpir_b200.UNetModel every output type runs natively:
  'pred_xstart' (p_sample)                 one fused UNet + eps->x0 launch sequence (dpir_unet_x0), none of p_sample's dead
                                           variance / posterior work
  'pred_x_prev', 'pred_x_prev_and_start',  UNet forward + ONE elementwise kernel (diffpir_b200.diffusion.GaussianDiffusion =
  ddim_sample=True                         gaussian_diffusion.py:395-439 / 537-585 with the reference's float64 tables)
  'epsilon', 'score'                       derived from pred_xstart exactly as :247-256
The `diffusion` argument may be the reference's own SpacedDiffusion (its betas / timestep map are mirrored), a native
GaussianDiffusion, or None (DiffPIR's linear 1000-step schedule).  Any other model object is handed to the caller's
`diffusion` unchanged (reference behaviour).
"""
import numpy as np
import torch

# The reference draws torch.randn_like(x) inside p_sample / ddim_sample on EVERY call, also when only pred_xstart is used
# (gaussian_diffusion.py:430, 575).  True keeps the torch generator in lockstep with the reference (same seed -> same
# trajectory under the unmodified drivers); False skips the dead draw (one launch less per step).
RNG_COMPAT = True


def find_nearest(array, value):
    if torch.is_tensor(array):
        array = array.detach().cpu().numpy()
    array = np.asarray(array)
    return int((np.abs(array - value)).argmin())


_sigma_cache = {}
_native_cache = {}


def _reduced_alpha_cumprod(alphas_cumprod):
    """sigma_bar table (utils_model.py:210-212) on the host, cached by CONTENT (a freed and re-allocated tensor with another
    schedule must not hit a stale entry: the key holds size, first/last/mid values, not the data pointer)."""
    ac = alphas_cumprod.detach().to("cpu")
    n = ac.numel()
    key = (n, float(ac[0]), float(ac[n // 2]), float(ac[-1]), str(ac.dtype))
    tab = _sigma_cache.get(key)
    if tab is None:
        tab = torch.div(torch.sqrt(1.0 - ac), torch.sqrt(ac)).numpy()
        _sigma_cache[key] = tab
    return tab


def _native_diffusion(diffusion):
    from .diffusion import GaussianDiffusion
    if isinstance(diffusion, GaussianDiffusion):
        return diffusion
    key = id(diffusion) if diffusion is not None else None
    ent = _native_cache.get(key)
    if ent is None or ent[0] is not diffusion:
        nat = GaussianDiffusion() if diffusion is None else GaussianDiffusion.from_reference(diffusion)
        nat.is_default = (nat.timestep_map is None or list(nat.timestep_map) == list(range(nat.num_timesteps))) and \
            not nat.rescale_timesteps
        ent = (diffusion, nat)
        _native_cache[key] = ent
    return ent[1]


def _engine_matches(model, nat):
    """dpir_unet_x0 converts eps->x0 with the engine's built-in linear-schedule table: usable when `nat` is that schedule."""
    if not getattr(nat, "is_default", False) or nat.num_timesteps != getattr(model, "diffusion_steps", -1):
        return False
    if not hasattr(nat, "_linear"):
        from .diffusion import linear_betas
        nat._linear = bool(np.array_equal(nat.betas, linear_betas(nat.num_timesteps)))
    return nat._linear


def model_fn(x, noise_level, model_diffusion, vec_t=None, model_out_type='pred_xstart',
             diffusion=None, ddim_sample=False, alphas_cumprod=None, **model_kwargs):
    t_step = None
    if not torch.is_tensor(vec_t):
        t_step = find_nearest(_reduced_alpha_cumprod(alphas_cumprod), float(noise_level) / 255.)
        vec_t = torch.full((x.shape[0],), t_step, device=x.device, dtype=torch.int64)
    native = hasattr(model_diffusion, "x0") and not model_kwargs
    if native:
        try:
            nat = _native_diffusion(diffusion)
        except NotImplementedError:
            native = False
    if native:
        need_sample = model_out_type in ('pred_x_prev', 'pred_x_prev_and_start')
        if not need_sample and not ddim_sample and _engine_matches(model_diffusion, nat):
            if RNG_COMPAT:
                torch.randn_like(x)                    # p_sample's draw (gaussian_diffusion.py:430), result unused
            out = {"pred_xstart": model_diffusion.x0(x, vec_t)}
        elif ddim_sample:
            out = nat.ddim_sample(model_diffusion, x, vec_t, clip_denoised=True, eta=0)
        else:
            out = nat.p_sample(model_diffusion, x, vec_t, clip_denoised=True)
    else:
        if diffusion is None:
            raise RuntimeError("model_fn: a model that is not a diffpir_b200.UNetModel needs the `diffusion` object")
        sampler = diffusion.ddim_sample if ddim_sample else diffusion.p_sample
        kw = dict(clip_denoised=True, denoised_fn=None, cond_fn=None, model_kwargs=model_kwargs)
        if ddim_sample:
            kw["eta"] = 0
        out = sampler(model_diffusion, x, vec_t, **kw)
    if model_out_type == 'pred_x_prev_and_start':
        return out["sample"], out["pred_xstart"]
    if model_out_type == 'pred_x_prev':
        return out["sample"]
    if model_out_type == 'pred_xstart':
        return out["pred_xstart"]
    if model_out_type not in ('epsilon', 'score'):
        raise ValueError(f"unknown model_out_type {model_out_type!r}")
    # :247-256.  The reference indexes alphas_cumprod[int(t_step)] (t_step only exists when vec_t was not given); with an
    # explicit vec_t the first entry is used, indexed ON THE DEVICE (no .item(), no host sync)
    if t_step is not None:
        alpha_prod_t = alphas_cumprod[int(t_step)]
    else:
        alpha_prod_t = torch.index_select(alphas_cumprod.to(x.device), 0, vec_t.reshape(-1)[:1]).reshape(())
    beta_prod_t = 1 - alpha_prod_t
    e = (x - alpha_prod_t ** (0.5) * out["pred_xstart"]) / beta_prod_t ** (0.5)
    return e if model_out_type == 'epsilon' else - e / beta_prod_t ** (0.5)


# ----------------------------------------------------------------------------- test-time wrappers
def _augment(img, mode):
    """The 8 dihedral transforms of utils_image.augment_img_tensor (utils/utils_image.py:399-413) on (B,C,H,W)."""
    if mode == 0:
        return img
    if mode == 1:
        return img.rot90(1, [2, 3]).flip([2])
    if mode == 2:
        return img.flip([2])
    if mode == 3:
        return img.rot90(3, [2, 3])
    if mode == 4:
        return img.rot90(2, [2, 3]).flip([2])
    if mode == 5:
        return img.rot90(1, [2, 3])
    if mode == 6:
        return img.rot90(2, [2, 3])
    return img.rot90(3, [2, 3]).flip([2])


def _pad_to(L, modulo):
    h, w = L.shape[-2:]
    pb = int(np.ceil(h / modulo) * modulo - h)
    pr = int(np.ceil(w / modulo) * modulo - w)
    if pb or pr:
        L = torch.nn.functional.pad(L, (0, pr, 0, pb), mode="replicate")
    return L, h, w


def _run(model, L, noise_level, vec_t, out_type):
    return model(L, noise_level, vec_t=vec_t, model_out_type=out_type)


def _padded(model, L, modulo, noise_level, vec_t, out_type):
    Lp, h, w = _pad_to(L, modulo)
    return _run(model, Lp, noise_level, vec_t, out_type)[..., :h, :w]


def _quadrants(L, refield):
    h, w = L.shape[-2:]
    hh, ww = (h // 2 // refield + 1) * refield, (w // 2 // refield + 1) * refield
    rows, cols = (slice(0, hh), slice(h - hh, h)), (slice(0, ww), slice(w - ww, w))
    return [L[..., r, c] for r in rows for c in cols]


def _stitch(Es, L, h, w, sf):
    b, c = Es[0].shape[:2]
    E = torch.zeros(b, c, sf * h, sf * w, dtype=L.dtype, device=L.device)
    hs, ws = h // 2 * sf, w // 2 * sf
    E[..., :hs, :ws] = Es[0][..., :hs, :ws]
    E[..., :hs, ws:w * sf] = Es[1][..., :hs, (-w + w // 2) * sf:]
    E[..., hs:h * sf, :ws] = Es[2][..., (-h + h // 2) * sf:, :ws]
    E[..., hs:h * sf, ws:w * sf] = Es[3][..., (-h + h // 2) * sf:, (-w + w // 2) * sf:]
    return E


def _split(model, L, refield, min_size, sf, modulo, noise_level, vec_t, out_type, once=False):
    h, w = L.shape[-2:]
    if not once and h * w <= min_size ** 2:
        Lp, _, _ = _pad_to(L, modulo)
        return _run(model, Lp, noise_level, vec_t, out_type)[..., :h * sf, :w * sf]
    Ls = _quadrants(L, refield)
    if once or h * w <= 4 * (min_size ** 2):
        Es = [_run(model, q, noise_level, vec_t, out_type) for q in Ls]
    else:
        Es = [_split(model, q, refield, min_size, sf, modulo, noise_level, vec_t, out_type) for q in Ls]
    return _stitch(Es, L, h, w, sf)


def _x8(fn, L):
    outs = []
    for i in range(8):
        E = fn(_augment(L, i))
        outs.append(_augment(E, 8 - i if i in (3, 5) else i))
    return torch.stack(outs, dim=0).mean(dim=0)


def test_mode(model_fn, model_diffusion, L, mode=0, refield=32, min_size=256, sf=1, modulo=1, noise_level=0, vec_t=None,
              model_out_type='pred_xstart', diffusion=None, ddim_sample=False, alphas_cumprod=None):
    """(0) normal (1) pad (2) split (3) x8 self-ensemble (4) split + x8 (5) split once.  As in the reference,
    `ddim_sample` is accepted but the wrapped call always uses ddim_sample=False (utils_model.py:32)."""
    def model(L_, nl, vec_t=None, model_out_type='pred_xstart'):
        return model_fn(L_, nl, model_diffusion=model_diffusion, vec_t=vec_t, model_out_type=model_out_type,
                        diffusion=diffusion, ddim_sample=False, alphas_cumprod=alphas_cumprod)
    a = (noise_level, vec_t, model_out_type)
    if mode == 0:
        return _run(model, L, *a)
    if mode == 1:
        return _padded(model, L, modulo, *a)
    if mode == 2:
        return _split(model, L, refield, min_size, sf, modulo, *a)
    if mode == 3:
        return _x8(lambda z: _padded(model, z, modulo, *a), L)
    if mode == 4:
        return _x8(lambda z: _split(model, z, refield, min_size, sf, modulo, *a), L)
    if mode == 5:
        return _split(model, L, refield, min_size, sf, modulo, *a, once=True)
    raise ValueError(f"unknown test mode {mode}")
