"""Denoiser wrapper with the reference's names and signatures (utils/utils_model.py):

    find_nearest(array, value)                                   :202-205
    model_fn(x, noise_level, model_diffusion, vec_t=None, model_out_type='pred_xstart',
             diffusion=None, ddim_sample=False, alphas_cumprod=None, **model_kwargs)   :207-258
    test_mode(model_fn, model_diffusion, L, mode, refield, min_size, sf, modulo, noise_level, vec_t,
              model_out_type, diffusion, ddim_sample, alphas_cumprod)                   :16-195

`model_fn` takes the fast path (one fused UNet + eps->x0 launch sequence, no dead variance / posterior /
randn work of p_sample) when the output type is 'pred_xstart' and `model_diffusion` is a
diffpir_b200.UNetModel; 'epsilon' / 'score' derive from it as the reference does (:241-256).  Output types
that need p_sample's random `sample` ('pred_x_prev', 'pred_x_prev_and_start') or DDIM are delegated to the
`diffusion` object passed by the caller (the reference's own SpacedDiffusion), with our model plugged in.
"""
import numpy as np
import torch


def find_nearest(array, value):
    if torch.is_tensor(array):
        array = array.detach().cpu().numpy()
    array = np.asarray(array)
    return int((np.abs(array - value)).argmin())


_sigma_cache = {}


def _reduced_alpha_cumprod(alphas_cumprod):
    """sigma_bar table (utils_model.py:210-212), cached on the host per alphas_cumprod tensor."""
    key = (alphas_cumprod.data_ptr(), alphas_cumprod.numel(), str(alphas_cumprod.device))
    tab = _sigma_cache.get(key)
    if tab is None:
        ac = alphas_cumprod.detach().to("cpu")
        tab = torch.div(torch.sqrt(1.0 - ac), torch.sqrt(ac)).numpy()
        _sigma_cache[key] = tab
    return tab


def model_fn(x, noise_level, model_diffusion, vec_t=None, model_out_type='pred_xstart',
             diffusion=None, ddim_sample=False, alphas_cumprod=None, **model_kwargs):
    t_step = None
    if not torch.is_tensor(vec_t):
        t_step = find_nearest(_reduced_alpha_cumprod(alphas_cumprod), float(noise_level) / 255.)
        vec_t = torch.full((x.shape[0],), t_step, device=x.device, dtype=torch.int64)
    fast = hasattr(model_diffusion, "x0") and not ddim_sample and not model_kwargs
    if fast and model_out_type in ('pred_xstart', 'epsilon', 'score'):
        x0 = model_diffusion.x0(x, vec_t)
        if model_out_type == 'pred_xstart':
            return x0
        if t_step is None:
            t_step = int(vec_t.reshape(-1)[0].item())
        alpha_prod_t = alphas_cumprod[int(t_step)]
        beta_prod_t = 1 - alpha_prod_t
        out = (x - alpha_prod_t ** (0.5) * x0) / beta_prod_t ** (0.5)
        if model_out_type == 'score':
            out = - out / beta_prod_t ** (0.5)
        return out
    if diffusion is None:
        raise RuntimeError(f"model_fn: output type '{model_out_type}' / ddim needs the `diffusion` object")
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
    if t_step is None:
        t_step = int(vec_t.reshape(-1)[0].item())
    alpha_prod_t = alphas_cumprod[int(t_step)]
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
