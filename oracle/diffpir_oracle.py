"""CPU restatement of the DiffPIR per-step math and sampling loop (oracle; test infra only).

Everything here is plain torch/numpy on CPU and cites the reference lines it follows
(paths relative to /root/reference).  Works in fp32 (as the reference) or fp64.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import unet_oracle


# ----------------------------------------------------------------------------- schedules
def driver_schedule(beta_start=0.0001, beta_end=0.02, T=1000):
    """float32 driver tables, main_ddpir.py:184-190 (identical in the single-image drivers)."""
    betas = torch.from_numpy(np.linspace(beta_start, beta_end, T, dtype=np.float32))
    alphas = 1.0 - betas
    alphas_cumprod = torch.from_numpy(np.cumprod(alphas.numpy(), axis=0))
    sqrt_ac = torch.sqrt(alphas_cumprod)
    sqrt_1mac = torch.sqrt(1.0 - alphas_cumprod)
    reduced = torch.div(sqrt_1mac, sqrt_ac)
    return dict(betas=betas, alphas=alphas, alphas_cumprod=alphas_cumprod, sqrt_alphas_cumprod=sqrt_ac,
                sqrt_1m_alphas_cumprod=sqrt_1mac, reduced_alpha_cumprod=reduced)


def diffusion_tables(T=1000):
    """float64 tables of GaussianDiffusion.__init__, gaussian_diffusion.py:133-169, for the
    'linear' schedule of get_named_beta_schedule (scale = 1000/T, linspace(1e-4*scale, 2e-2*scale))."""
    scale = 1000.0 / T
    betas = np.linspace(scale * 0.0001, scale * 0.02, T, dtype=np.float64)
    ac = np.cumprod(1.0 - betas, axis=0)
    return dict(betas=betas, alphas_cumprod=ac, sqrt_recip_alphas_cumprod=np.sqrt(1.0 / ac),
                sqrt_recipm1_alphas_cumprod=np.sqrt(1.0 / ac - 1))


def find_nearest(array, value):
    """utils/utils_model.py:202-205."""
    array = np.asarray(array)
    return int((np.abs(array - value)).argmin())


def make_seq(T, iter_num, skip_type="quad"):
    """main_ddpir.py:327-335."""
    skip = T // iter_num
    if skip_type == "uniform":
        seq = [i * skip for i in range(iter_num)]
        if skip > 1:
            seq.append(T - 1)
    else:
        seq = np.sqrt(np.linspace(0, T ** 2, iter_num))
        seq = [int(s) for s in list(seq)]
        seq[-1] = seq[-1] - 1
    return seq


def rhos_sigmas(sched, lambda_, sigma, T=1000):
    """main_ddpir.py:274-286 (pred_xstart + DiffPIR branch): float32 torch scalars."""
    red = sched["reduced_alpha_cumprod"]
    sigmas = [red[T - 1 - i] for i in range(T)]
    sigma_ks = [sched["sqrt_1m_alphas_cumprod"][i] / sched["sqrt_alphas_cumprod"][i] for i in range(T)]
    rhos = [lambda_ * (sigma ** 2) / (sigma_ks[i] ** 2) for i in range(T)]
    return torch.tensor(rhos), torch.tensor(sigmas)


# ----------------------------------------------------------------------------- denoiser wrapper
def pred_xstart(eps_model_out, x, t, tables=None):
    """p_mean_variance with EPSILON mean / LEARNED_RANGE var, gaussian_diffusion.py:258-311, 328-333:
    split 6->(eps, var), x0 = clamp(c1[t] x - c2[t] eps, -1, 1); tables are f64 -> .float()
    (_extract_into_tensor :895-908).  t: (B,) long."""
    tables = tables or diffusion_tables()
    C = x.shape[1]
    eps = eps_model_out[:, :C]
    c1 = torch.from_numpy(tables["sqrt_recip_alphas_cumprod"])[t].float().to(x.dtype)[:, None, None, None]
    c2 = torch.from_numpy(tables["sqrt_recipm1_alphas_cumprod"])[t].float().to(x.dtype)[:, None, None, None]
    return (c1 * x - c2 * eps).clamp(-1, 1)


def model_fn(sd, cfg, x, noise_level, sched, tables=None):
    """utils/utils_model.py:207-258 with model_out_type='pred_xstart', ddim_sample=False:
    t = find_nearest(sigma_bar, noise_level/255); x0 = p_sample(...)['pred_xstart']."""
    t_step = find_nearest(sched["reduced_alpha_cumprod"].numpy(), noise_level / 255.0)
    vec_t = torch.tensor([t_step] * x.shape[0])
    out = unet_oracle.unet_forward(sd, cfg, x, vec_t)
    return pred_xstart(out, x, vec_t, tables)


# ----------------------------------------------------------------------------- FFT data solver
def splits(a, sf):
    """utils/utils_sisr.py:9-19."""
    b = torch.stack(torch.chunk(a, sf, dim=2), dim=4)
    return torch.cat(torch.chunk(b, sf, dim=3), dim=4)


def p2o(psf, shape):
    """utils/utils_sisr.py:22-41: zero-pad the PSF to `shape`, circularly shift its centre to (0,0), FFT."""
    otf = torch.zeros(psf.shape[:-2] + tuple(shape), dtype=psf.dtype)
    otf[..., :psf.shape[2], :psf.shape[3]] = psf
    for axis, axis_size in enumerate(psf.shape[2:]):
        otf = torch.roll(otf, -int(axis_size / 2), dims=axis + 2)
    return torch.fft.fftn(otf, dim=(-2, -1))


def pre_calculate(y, k, sf):
    """utils/utils_sisr.py:78-95."""
    w, h = y.shape[-2:]
    FB = p2o(k, (w * sf, h * sf))
    FBC = torch.conj(FB)
    F2B = torch.abs(FB) ** 2
    STy = torch.zeros(y.shape[0], y.shape[1], w * sf, h * sf, dtype=y.dtype)
    STy[..., 0::sf, 0::sf] = y
    FBFy = FBC * torch.fft.fftn(STy, dim=(-2, -1))
    return FB, FBC, F2B, FBFy


def data_solution(x, FB, FBC, F2B, FBFy, alpha, sf):
    """utils/utils_sisr.py:65-75 (reference-literal full-spectrum formula)."""
    FR = FBFy + torch.fft.fftn(alpha * x, dim=(-2, -1))
    x1 = FB * FR
    FBR = torch.mean(splits(x1, sf), dim=-1)
    invW = torch.mean(splits(F2B, sf), dim=-1)
    invWBR = FBR / (invW + alpha)
    FCBinvWBR = FBC * invWBR.repeat(1, 1, sf, sf)
    FX = (FR - FCBinvWBR) / alpha
    return torch.real(torch.fft.ifftn(FX, dim=(-2, -1)))


# ----------------------------------------------------------------------------- inpaint prox, re-noise
def inpaint_prox(x0, y_pm1, mask, tau, guidance_scale=1.0):
    """main_ddpir.py:393-394 (y_pm1 = 2y-1 there; main_ddpir_inpainting.py:267-268 passes y directly)."""
    x0_p = (mask * y_pm1 + tau * x0) / (mask + tau)
    return x0 + guidance_scale * (x0_p - x0)


def renoise(x, x0, t_i, t_im1, sched, eta, zeta, noise_eta, noise_zeta):
    """main_ddpir.py:451-456.  Scalar tensors are float32; np.sqrt(1-zeta), np.sqrt(zeta) are Python
    float64 scalars multiplying float32 tensors (stay float32).  Noise tensors are passed in."""
    sa, s1 = sched["sqrt_alphas_cumprod"], sched["sqrt_1m_alphas_cumprod"]
    dt = x.dtype
    eps = (x - sa[t_i].to(dt) * x0) / s1[t_i].to(dt)
    eta_sigma = eta * s1[t_im1] / s1[t_i] * torch.sqrt(sched["betas"][t_i])
    return (sa[t_im1].to(dt) * x0
            + np.sqrt(1 - zeta) * (torch.sqrt(s1[t_im1] ** 2 - eta_sigma ** 2).to(dt) * eps
                                   + eta_sigma.to(dt) * noise_eta)
            + np.sqrt(zeta) * s1[t_im1].to(dt) * noise_zeta)


# ----------------------------------------------------------------------------- bicubic resizer / IBP
def _cubic(x):
    """utils/utils_resizer.py:173-178."""
    a = np.abs(x)
    a2, a3 = a ** 2, a ** 3
    return ((1.5 * a3 - 2.5 * a2 + 1) * (a <= 1) + (-0.5 * a3 + 2.5 * a2 - 4 * a + 2) * ((1 < a) & (a <= 2)))


def resizer_contributions(in_length, out_length, scale, antialiasing=True):
    """utils/utils_resizer.py:104-167 for the cubic kernel (kernel_width 4).
    Returns (weights [out, taps] float64, field_of_view [out, taps] int64)."""
    kernel_width = 4.0
    fixed = (lambda a: scale * _cubic(scale * a)) if antialiasing else _cubic
    kernel_width *= 1.0 / scale if antialiasing else 1.0
    out_c = np.arange(1, out_length + 1)
    shifted = out_c - (out_length - in_length * scale) / 2
    match = shifted / scale + 0.5 * (1 - 1 / scale)
    left = np.floor(match - kernel_width / 2)
    expanded = np.ceil(kernel_width) + 2
    fov = np.squeeze(np.int16(np.expand_dims(left, axis=1) + np.arange(expanded) - 1))
    weights = fixed(1.0 * np.expand_dims(match, axis=1) - fov - 1)
    sw = np.sum(weights, axis=1)
    sw[sw == 0] = 1.0
    weights = 1.0 * weights / np.expand_dims(sw, axis=1)
    mirror = np.uint(np.concatenate((np.arange(in_length), np.arange(in_length - 1, -1, step=-1))))
    fov = mirror[np.mod(fov, mirror.shape[0])]
    nz = np.nonzero(np.any(weights, axis=0))
    weights = np.squeeze(weights[:, nz])
    fov = np.squeeze(fov[:, nz])
    return weights, fov.astype(np.int64)


def resizer_forward(x, scale):
    """Resizer(in_shape, scale).forward, utils/utils_resizer.py:8-74, NCHW, same scale on H and W
    (sorted_dims = [2, 3] for equal scales since argsort is stable): gather + weighted sum per dim,
    weights cast to float32 as the reference does (:46)."""
    for dim in (2, 3):
        n_in = x.shape[dim]
        n_out = int(np.ceil(n_in * scale))
        w, fov = resizer_contributions(n_in, n_out, scale, antialiasing=scale < 1)
        w = torch.tensor(w.T, dtype=torch.float32).to(x.dtype)      # [taps, out]
        fov = torch.tensor(fov.T, dtype=torch.long)                  # [taps, out]
        xt = torch.transpose(x, dim, 0)
        w_b = w.reshape(list(w.shape) + [1] * (x.dim() - 1))
        xt = torch.sum(xt[fov] * w_b, dim=0)
        x = torch.transpose(xt, dim, 0)
    return x


def _matlab_cubic(x):
    """utils/utils_image.py:672-677."""
    ax = torch.abs(x)
    ax2, ax3 = ax ** 2, ax ** 3
    return (1.5 * ax3 - 2.5 * ax2 + 1) * (ax <= 1).type_as(ax) + \
        (-0.5 * ax3 + 2.5 * ax2 - 4 * ax + 2) * ((ax > 1) * (ax <= 2)).type_as(ax)


def imresize_weights_indices(in_length, out_length, scale, kernel_width=4, antialiasing=True):
    """MATLAB-style bicubic taps, utils/utils_image.py:680-731 (calculate_weights_indices): float32 torch arithmetic as there.
    Returns (weights [out, taps], first index into the symmetric-padded signal [out], sym_len_start, sym_len_end).
    Note :721-726: the leading column is dropped when it contains ANY zero weight, and after that the trailing column is
    only dropped if the leading one was not (the second narrow is then a no-op) - restated as is."""
    if scale < 1 and antialiasing:
        kernel_width = kernel_width / scale
    x = torch.linspace(1, out_length, out_length)
    u = x / scale + 0.5 * (1 - 1 / scale)
    left = torch.floor(u - kernel_width / 2)
    P = math.ceil(kernel_width) + 2
    indices = left.view(out_length, 1).expand(out_length, P) + torch.linspace(0, P - 1, P).view(1, P).expand(out_length, P)
    dist = u.view(out_length, 1).expand(out_length, P) - indices
    weights = scale * _matlab_cubic(dist * scale) if (scale < 1 and antialiasing) else _matlab_cubic(dist)
    weights = weights / torch.sum(weights, 1).view(out_length, 1).expand(out_length, P)
    zeros = torch.sum((weights == 0), 0)
    if int(zeros[0]) != 0:
        indices = indices.narrow(1, 1, P - 2)
        weights = weights.narrow(1, 1, P - 2)
    if int(zeros[-1]) != 0:
        indices = indices.narrow(1, 0, P - 2)
        weights = weights.narrow(1, 0, P - 2)
    sym_s = int(-indices.min() + 1)
    sym_e = int(indices.max() - in_length)
    first = (indices[:, 0] + sym_s - 1).long()
    return weights.contiguous(), first, sym_s, sym_e


def imresize_np(img, scale, antialiasing=True):
    """utils/utils_image.py:811-881 on an (H,W,C) or (H,W) float array: symmetric padding, H pass, then W pass."""
    x = torch.from_numpy(np.asarray(img, dtype=np.float32))
    squeeze = x.dim() == 2
    if squeeze:
        x = x[:, :, None]
    H, W, C = x.shape
    oH, oW = math.ceil(H * scale), math.ceil(W * scale)

    def one_axis(t, n_in, n_out):            # resize along dim 0 of (n_in, M)
        w, first, ss, se = imresize_weights_indices(n_in, n_out, scale, 4, antialiasing)
        aug = torch.cat([t[:ss].flip(0), t, t[n_in - se:].flip(0) if se > 0 else t[:0]], 0)
        kw = w.shape[1]
        return torch.stack([(aug[int(first[i]):int(first[i]) + kw] * w[i].view(kw, 1)).sum(0) for i in range(n_out)], 0)

    y = one_axis(x.reshape(H, W * C), H, oH).reshape(oH, W, C)
    y = one_axis(y.permute(1, 0, 2).reshape(W, oH * C), W, oW).reshape(oW, oH, C).permute(1, 0, 2)
    y = y.contiguous().numpy()
    return y[:, :, 0] if squeeze else y


def ibp_step(x0, y, sf, gamma, rho, in_iter=1):
    """main_ddpir.py:401-406: x0 in [-1,1]; y in [0,1] at low resolution;
    up_sample = F.interpolate(scale_factor=sf) (default mode = nearest, :297)."""
    for _ in range(in_iter):
        x0 = x0 / 2 + 0.5
        x0 = x0 + gamma * F.interpolate(y - resizer_forward(x0, 1.0 / sf), scale_factor=sf) / (1 + rho)
        x0 = x0 * 2 - 1
    return x0


# ----------------------------------------------------------------------------- full loop
def diffpir_loop(sd, cfg, task, y, *, k=None, mask=None, sf=1, sigma=0.05, lambda_=7.0, zeta=0.3,
                 eta=0.0, guidance_scale=1.0, iter_num=100, skip_type="quad", T=1000,
                 x_init=None, init_noise=None, zeta_noise=None, eta_noise=None, sr_mode="blur",
                 gamma=1.0 / 100, in_iter=1, unet=None, trace=None, max_steps=None, iter_num_U=1, u_noise=None):
    """The batched DiffPIR loop of main_ddpir.py:259-470 (generate_mode='DiffPIR', sub_1_analytic,
    model_output_type='pred_xstart', iter_num_U=1, noise_init_img='max', skip_noise_model_t=False)
    with all noise pre-drawn:
        init_noise (B,3,H,W): the randn of :315;  zeta_noise [steps-1] x (B,3,H,W): 2nd randn of :456;
        eta_noise: 1st randn of :456 (only matters when eta != 0).
    task: 'deblur' | 'sr' | 'inpaint'.  y in [0,1] (for inpaint y is the un-masked image, mask applied
    as in :310-312).  `x_init` is the pre-noise x of :295-312 for sr (bicubic upsample); defaults to y / y*mask.
    `unet(x, vec_t)` may override the denoiser (e.g. the reference nn.Module).  Returns x_0 = x/2+0.5 (:470).
    iter_num_U > 1 (main_ddpir.py:349, 460-467): every non-final step is repeated U times - re-noise to t_{i-1}, then back to
    t_i with fresh noise u_noise[.]; zeta_noise then holds one tensor per re-noise.  (On the final sequence entry the reference's
    repeats would index seq[i+1] - IndexError at :451 - so it is evaluated once.)
    """
    sched = driver_schedule(T=T)
    tables = diffusion_tables(T)
    dt = y.dtype
    rhos, sigmas = rhos_sigmas(sched, lambda_, max(0.001, sigma), T)
    t_start = T - 1
    if task == "inpaint":
        x = y * mask
    elif x_init is not None:
        x = x_init
    else:
        x = y
    x = sched["sqrt_alphas_cumprod"][t_start].to(dt) * (2 * x - 1) + sched["sqrt_1m_alphas_cumprod"][t_start].to(dt) * init_noise
    if task in ("deblur", "sr") and sr_mode == "blur":
        FB, FBC, F2B, FBFy = pre_calculate(y, k, sf)
    seq = make_seq(T, iter_num, skip_type)
    red = sched["reduced_alpha_cumprod"].numpy()
    n_steps = len(seq) if max_steps is None else min(max_steps, len(seq))
    nz = nu = 0
    for i in range(n_steps):
        curr_sigma = sigmas[seq[i]].numpy()
        t_i = find_nearest(red, curr_sigma)
        vec_t = torch.tensor([t_i] * x.shape[0])
        last = seq[i] == seq[-1]
        for u in range(1 if last else iter_num_U):
            out = unet(x, vec_t) if unet is not None else unet_oracle.unet_forward(sd, cfg, x, vec_t)
            x0 = pred_xstart(out, x, vec_t, tables)
            if trace is not None:
                trace.append(("x0_unet", i, x0.clone()))
            if not last:
                tau = rhos[t_i].float().reshape(1, 1, 1, 1).to(dt)
                if task == "inpaint":
                    x0 = inpaint_prox(x0, 2 * y - 1, mask, tau, guidance_scale)
                elif sr_mode == "blur":
                    x0_p = data_solution(x0 / 2 + 0.5, FB, FBC, F2B, FBFy, tau, sf) * 2 - 1
                    x0 = x0 + guidance_scale * (x0_p - x0)
                else:
                    x0 = ibp_step(x0, y, sf, gamma, rhos[t_i].to(dt), in_iter)
                if trace is not None:
                    trace.append(("x0_prox", i, x0.clone()))
                t_im1 = find_nearest(red, sigmas[seq[i + 1]].numpy())
                n_eta = eta_noise[nz] if eta_noise is not None else torch.zeros_like(x)
                x = renoise(x, x0, t_i, t_im1, sched, eta, zeta, n_eta, zeta_noise[nz])
                nz += 1
                if u < iter_num_U - 1:                   # :460-467 set back to x_t from x_{t-1}
                    sa, s1 = sched["sqrt_alphas_cumprod"], sched["sqrt_1m_alphas_cumprod"]
                    a_eff = (sa[t_i] / sa[t_im1]).to(dt)
                    x = a_eff * x + torch.sqrt(s1[t_i].to(dt) ** 2 - a_eff ** 2 * s1[t_im1].to(dt) ** 2) * u_noise[nu]
                    nu += 1
                if trace is not None:
                    trace.append(("x", i, x.clone()))
    return x / 2 + 0.5
