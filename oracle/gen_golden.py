"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference).

Build-container only (the reference does not travel to the GPU box).  The reference ships no tests or
golden vectors, so these fixtures - outputs of the reference's own modules on
seeded synthetic inputs - are what pins the oracle.  Usage:

    python -m oracle.gen_golden            # (re)write the fixtures
    python -m oracle.gen_golden --check    # recompute and compare with the committed fixtures

Every entry is produced by reference code only: guided_diffusion.unet.UNetModel (via
script_util.create_model_and_diffusion), SpacedDiffusion.p_sample, utils.utils_model.model_fn,
utils.utils_sisr.pre_calculate / data_solution, utils.utils_resizer.Resizer, utils.utils_inpaint.mask_generator,
utils.utils_deblur.GaussialBlurOperator and the loop arithmetic of main_ddpir.py:341-470 re-typed around them.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from diffpir_b200 import arch, synth  # noqa: E402
from oracle import ref_shim  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def small_cfg():
    return arch.UNetConfig(image_size=64, model_channels=128, num_res_blocks=1, attention_resolutions="16,8",
                           num_head_channels=64)


def gen_unet_small():
    cfg = small_cfg()
    model, diffusion = ref_shim.create_reference_model(ref_shim.ref_model_config(cfg))
    model.load_state_dict(synth.synthetic_state_dict(cfg, 0))
    g = torch.Generator().manual_seed(11)
    x = torch.randn(2, 3, 64, 64, generator=g)
    t = torch.tensor([999, 37])
    with torch.no_grad():
        out = model(x, t)
        ps = diffusion.p_sample(model, x, t, clip_denoised=True)["pred_xstart"]
    return dict(x=x.numpy(), t=t.numpy(), out=out.numpy(), pred_xstart=ps.numpy())


def gen_unet_ffhq():
    cfg = arch.ffhq_config()
    model, _ = ref_shim.create_reference_model(ref_shim.ref_model_config(cfg))
    model.load_state_dict(synth.synthetic_state_dict(cfg, 0))
    g = torch.Generator().manual_seed(12)
    x = torch.randn(1, 3, 256, 256, generator=g)
    t = torch.tensor([321])
    with torch.no_grad():
        out = model(x, t)
    return dict(seed=np.array(12), t=t.numpy(), out_sub=out[:, :, ::4, ::4].contiguous().numpy(),
                out_mean=np.array(out.double().mean().item()), out_std=np.array(out.double().std().item()))


def sisr_cases():
    """Seeded inputs of the FFT-solver fixtures (regenerated identically by the tests)."""
    g = torch.Generator().manual_seed(13)
    for sf in (1, 4):
        y = torch.rand(1, 3, 64 // sf, 64 // sf, generator=g)
        x = torch.rand(1, 3, 64, 64, generator=g)
        k = torch.rand(1, 1, 25, 25, generator=g)
        k = k / k.sum(dim=(-2, -1), keepdim=True)
        for alpha in (1e-3, 0.3):
            yield f"sf{sf}_a{alpha:g}", sf, alpha, y, x, k


def gen_sisr():
    ref_shim.install()
    from utils import utils_sisr as sr
    res = {}
    for tag, sf, alpha, y, x, k in sisr_cases():
        a = torch.tensor(alpha).view(1, 1, 1, 1)
        FB, FBC, F2B, FBFy = sr.pre_calculate(y, k, sf)
        res[tag + "_out32"] = sr.data_solution(x, FB, FBC, F2B, FBFy, a, sf).numpy()
        if sf == 4 and alpha == 0.3:
            res["FB_sf4"] = torch.view_as_real(FB).numpy()
            res["F2B_sf4"] = F2B.numpy()
            res["FBFy_sf4"] = torch.view_as_real(FBFy).numpy()
        FB, FBC, F2B, FBFy = sr.pre_calculate(y.double(), k.double(), sf)
        res[tag + "_out64"] = sr.data_solution(x.double(), FB, FBC, F2B, FBFy, a.double(), sf).numpy()
    return res


def gen_misc():
    ref_shim.install()
    from utils.utils_resizer import Resizer
    from utils.utils_inpaint import mask_generator
    from utils.utils_deblur import GaussialBlurOperator
    res = {}
    g = torch.Generator().manual_seed(14)
    x = torch.rand(2, 3, 64, 64, generator=g)
    res["resize_x"] = x.numpy()
    res["resize_down4"] = Resizer(x.shape, 0.25)(x).numpy()
    res["resize_up2"] = Resizer(x.shape, 2.0)(x).numpy()
    r = Resizer((1, 3, 256, 256), 0.25)
    res["fov256_h"] = r.field_of_view[0].numpy().astype(np.int32)
    res["w256_h"] = r.weights[0].numpy().reshape(r.weights[0].shape[0], -1)
    np.random.seed(123)
    mg = mask_generator("random", mask_prob_range=(0.5, 0.5), image_size=64)
    res["mask_random"] = mg(torch.zeros(1, 3, 64, 64)).numpy().astype(np.uint8)
    np.random.seed(123)
    mg = mask_generator("box", mask_len_range=(16, 17), image_size=64, margin=(4, 4))
    res["mask_box"] = mg(torch.zeros(1, 3, 64, 64)).numpy().astype(np.uint8)
    op = GaussialBlurOperator(kernel_size=21, intensity=2.5, device="cpu")
    res["gauss_k"] = op.get_kernel().numpy()
    xb = torch.rand(1, 3, 40, 48, generator=g)
    res["blur_x"] = xb.numpy()
    with torch.no_grad():
        res["blur_out"] = op.forward(xb).numpy()
    return res


def reference_loop(model, diffusion, task, y, k, mask, sf, hp, iters, init_noise, zeta_noise, x_init, iter_num_U=1, u_noise=None):
    """main_ddpir.py:259-470 (DiffPIR / sub_1_analytic / pred_xstart branch), with the reference's own model_fn,
    pre_calculate and data_solution, and the torch.randn_like draws replaced by the pre-drawn tensors."""
    from utils import utils_model, utils_sisr as sr
    T = 1000
    betas = torch.from_numpy(np.linspace(0.0001, 0.02, T, dtype=np.float32))
    alphas = 1.0 - betas
    alphas_cumprod = np.cumprod(alphas.cpu(), axis=0)
    sqrt_alphas_cumprod = torch.sqrt(alphas_cumprod)
    sqrt_1m_alphas_cumprod = torch.sqrt(1. - alphas_cumprod)
    reduced_alpha_cumprod = torch.div(sqrt_1m_alphas_cumprod, sqrt_alphas_cumprod)
    sigma = max(0.001, hp["sigma"])
    sigmas, rhos = [], []
    for i in range(T):
        sigmas.append(reduced_alpha_cumprod[T - 1 - i])
        rhos.append(hp["lambda_"] * (sigma ** 2) / ((sqrt_1m_alphas_cumprod[i] / sqrt_alphas_cumprod[i]) ** 2))
    rhos, sigmas = torch.tensor(rhos), torch.tensor(sigmas)
    t_start = T - 1
    x = y * mask if task == "inpaint" else (x_init if x_init is not None else y)
    x = sqrt_alphas_cumprod[t_start] * (2 * x - 1) + sqrt_1m_alphas_cumprod[t_start] * init_noise
    if task in ("sr", "deblur"):
        FB, FBC, F2B, FBFy = sr.pre_calculate(y, k, sf)
    seq = np.sqrt(np.linspace(0, T ** 2, iters))
    seq = [int(s) for s in list(seq)]
    seq[-1] = seq[-1] - 1
    zeta, eta = hp["zeta"], 0.0
    nz = nu = 0
    for i in range(len(seq)):
        curr_sigma = sigmas[seq[i]].cpu().numpy()
        t_i = utils_model.find_nearest(reduced_alpha_cumprod, curr_sigma)
        # main_ddpir.py:349 `for u in range(config.iter_num_U)`.  On the final sequence entry the reference's re-noise would
        # index seq[i+1] for u < iter_num_U-1 (IndexError), so the repeats only exist before it.
        for u in range(iter_num_U if seq[i] != seq[-1] else 1):
            x0 = utils_model.model_fn(x, noise_level=curr_sigma * 255, model_out_type="pred_xstart", model_diffusion=model,
                                      diffusion=diffusion, ddim_sample=False, alphas_cumprod=alphas_cumprod)
            if seq[i] != seq[-1]:
                tau = rhos[t_i].float().repeat(1, 1, 1, 1)
                if task == "inpaint":
                    x0_p = (mask * (2 * y - 1) + tau * x0).div(mask + tau)
                    x0 = x0 + 1.0 * (x0_p - x0)
                else:
                    x0_p = x0 / 2 + 0.5
                    x0_p = sr.data_solution(x0_p.float(), FB, FBC, F2B, FBFy, tau, sf)
                    x0_p = x0_p * 2 - 1
                    x0 = x0 + 1.0 * (x0_p - x0)
                t_im1 = utils_model.find_nearest(reduced_alpha_cumprod, sigmas[seq[i + 1]].cpu().numpy())
                eps = (x - sqrt_alphas_cumprod[t_i] * x0) / sqrt_1m_alphas_cumprod[t_i]
                eta_sigma = eta * sqrt_1m_alphas_cumprod[t_im1] / sqrt_1m_alphas_cumprod[t_i] * torch.sqrt(betas[t_i])
                x = sqrt_alphas_cumprod[t_im1] * x0 + np.sqrt(1 - zeta) * (
                    torch.sqrt(sqrt_1m_alphas_cumprod[t_im1] ** 2 - eta_sigma ** 2) * eps
                    + eta_sigma * torch.zeros_like(x)) + np.sqrt(zeta) * sqrt_1m_alphas_cumprod[t_im1] * zeta_noise[nz]
                nz += 1
                if u < iter_num_U - 1:                   # :460-467 set back to x_t from x_{t-1}
                    sqrt_alpha_effective = sqrt_alphas_cumprod[t_i] / sqrt_alphas_cumprod[t_im1]
                    x = sqrt_alpha_effective * x + torch.sqrt(sqrt_1m_alphas_cumprod[t_i] ** 2 -
                                                              sqrt_alpha_effective ** 2 * sqrt_1m_alphas_cumprod[t_im1] ** 2) * u_noise[nu]
                    nu += 1
    return x / 2 + 0.5


def loop_inputs(task, B=2, size=64, sf=1):
    imgs = synth.synthetic_images(B, size, seed=100)
    gt = torch.from_numpy(imgs).permute(0, 3, 1, 2).float() / 255.0
    k = mask = None
    if task == "deblur":
        kk = synth.gaussian_kernel(21, 2.0)
        y = np.stack([synth.blur_measurement(imgs[b], kk, 0.05, seed=b) for b in range(B)])
        y = torch.from_numpy(y).permute(0, 3, 1, 2).contiguous()
        k = torch.from_numpy(kk)[None, None].repeat(B, 1, 1, 1)
    elif task == "sr":
        kk = synth.gaussian_kernel(25, 1.6)
        k = torch.from_numpy(kk)[None, None].repeat(B, 1, 1, 1)
        y = torch.nn.functional.avg_pool2d(gt, sf)
    else:
        y = gt.clone()
        mask = synth.random_mask(B, size, 0.5, seed=3)
    return gt, y, k, mask


LOOP_CASES = [("deblur", 1, dict(lambda_=7.0, zeta=0.3, sigma=0.05)),
              ("sr", 4, dict(lambda_=6.0, zeta=0.25, sigma=0.05)),
              ("inpaint", 1, dict(lambda_=1.0, zeta=1.0, sigma=0.0))]


def gen_loops(iters=6):
    cfg = small_cfg()
    model, diffusion = ref_shim.create_reference_model(ref_shim.ref_model_config(cfg))
    model.load_state_dict(synth.synthetic_state_dict(cfg, 0))
    res = {"iters": np.array(iters)}
    for task, sf, hp in LOOP_CASES:
        gt, y, k, mask = loop_inputs(task, sf=sf)
        init, zs = synth.predrawn_noise((2, 3, 64, 64), iters, seed=7)
        x_init = torch.nn.functional.interpolate(y, scale_factor=sf, mode="bicubic", align_corners=False) if task == "sr" else None
        with torch.no_grad():
            out = reference_loop(model, diffusion, task, y, k, mask, sf, hp, iters, init, zs, x_init)
        res[task + "_x0"] = out.numpy()
    return res


def gen_loop_u2(iters=4):
    """The same loop with iter_num_U = 2 (main_ddpir.py:349, 460-467), deblur."""
    cfg = small_cfg()
    model, diffusion = ref_shim.create_reference_model(ref_shim.ref_model_config(cfg))
    model.load_state_dict(synth.synthetic_state_dict(cfg, 0))
    task, sf, hp = LOOP_CASES[0]
    gt, y, k, mask = loop_inputs(task, sf=sf)
    init, zs = synth.predrawn_noise((2, 3, 64, 64), 2 * iters, seed=8)
    _, us = synth.predrawn_noise((2, 3, 64, 64), iters, seed=9)
    with torch.no_grad():
        out = reference_loop(model, diffusion, task, y, k, mask, sf, hp, iters, init, zs, None, iter_num_U=2, u_noise=us)
    return {"iters": np.array(iters), "deblur_U2_x0": out.numpy()}


MODEL_FN_SEED = 4321


def model_fn_inputs():
    g = torch.Generator().manual_seed(21)
    return torch.randn(2, 3, 64, 64, generator=g), 37.0          # x, noise_level (255 * sigma)


def gen_model_fn():
    """utils_model.model_fn (utils/utils_model.py:207-258) for every output type, p_sample and ddim_sample, small UNet.  The
    reference draws its noise with th.randn_like(x) from the global generator: seeded with MODEL_FN_SEED right before each call,
    so a test can re-draw the identical tensor on the CPU."""
    from utils import utils_model
    cfg = small_cfg()
    model, diffusion = ref_shim.create_reference_model(ref_shim.ref_model_config(cfg))
    model.load_state_dict(synth.synthetic_state_dict(cfg, 0))
    x, nl = model_fn_inputs()
    betas = torch.from_numpy(np.linspace(0.0001, 0.02, 1000, dtype=np.float32))
    alphas_cumprod = np.cumprod((1.0 - betas).cpu(), axis=0)
    res = {}
    with torch.no_grad():
        for ddim in (False, True):
            tag = "ddim" if ddim else "p"
            for typ in ("pred_xstart", "pred_x_prev", "epsilon", "score"):
                torch.manual_seed(MODEL_FN_SEED)
                res[f"{tag}_{typ}"] = utils_model.model_fn(x, noise_level=nl, model_out_type=typ, model_diffusion=model,
                                                           diffusion=diffusion, ddim_sample=ddim, alphas_cumprod=alphas_cumprod).numpy()
            torch.manual_seed(MODEL_FN_SEED)
            a, b = utils_model.model_fn(x, noise_level=nl, model_out_type="pred_x_prev_and_start", model_diffusion=model,
                                        diffusion=diffusion, ddim_sample=ddim, alphas_cumprod=alphas_cumprod)
            res[f"{tag}_both_prev"], res[f"{tag}_both_start"] = a.numpy(), b.numpy()
        # explicit vec_t including t = 0 (no noise added there, gaussian_diffusion.py:431-433)
        vt = torch.tensor([0, 500])
        torch.manual_seed(MODEL_FN_SEED)
        res["p_vec_t_prev"] = utils_model.model_fn(x, noise_level=0, vec_t=vt, model_out_type="pred_x_prev", model_diffusion=model,
                                                   diffusion=diffusion, alphas_cumprod=alphas_cumprod).numpy()
    return res


def fake_model_fn(x, noise_level, model_diffusion=None, vec_t=None, model_out_type="pred_xstart", diffusion=None,
                  ddim_sample=False, alphas_cumprod=None, **kw):
    """Deterministic, position-sensitive stand-in for utils_model.model_fn (an asymmetric 3x3 depthwise filter), used to
    pin the test-time wrappers of utils_model.test_mode without a network."""
    k = torch.tensor([[0.1, 0.3, -0.2], [0.5, 1.0, 0.05], [-0.4, 0.2, 0.15]]).view(1, 1, 3, 3).repeat(3, 1, 1, 1)
    return torch.nn.functional.conv2d(x, k.to(x), padding=1, groups=3) * (1.0 + float(noise_level) / 255.0)


TEST_MODE_CASES = [dict(mode=0), dict(mode=1, modulo=32), dict(mode=2, refield=16, min_size=48, modulo=16),
                   dict(mode=3, modulo=16), dict(mode=4, refield=16, min_size=48, modulo=16),
                   dict(mode=5, refield=16, min_size=48)]


def test_mode_input():
    return torch.rand(2, 3, 96, 80, generator=torch.Generator().manual_seed(15))


def gen_test_mode():
    ref_shim.install()
    from utils import utils_model
    x = test_mode_input()
    res = {}
    for i, kw in enumerate(TEST_MODE_CASES):
        res[f"case{i}"] = utils_model.test_mode(fake_model_fn, None, x, noise_level=12.0, **kw).numpy()
    return res


def gen_imresize():
    """utils_image.imresize_np (the SR measurement of main_ddpir.py:89 / main_ddpir_sisr.py:202) on small seeded images."""
    ref_shim.install()
    from utils import utils_image as util
    rng = np.random.RandomState(5)
    res = {}
    a = rng.rand(48, 40, 3).astype(np.float32)
    b = rng.rand(37, 52).astype(np.float32)          # odd sizes, single channel
    res["a"], res["b"] = a, b
    for sf in (2, 3, 4):
        res[f"a_down{sf}"] = util.imresize_np(a.copy(), 1 / sf)
    res["a_up2"] = util.imresize_np(a.copy(), 2)
    res["b_down2"] = util.imresize_np(b.copy(), 1 / 2)
    res["a_up3_noaa"] = util.imresize_np(a.copy(), 3, antialiasing=False)
    # (a zero-length symmetric pad - e.g. b at 1/4 without antialiasing - raises inside the reference itself, :870)
    return res


FILES = {"imresize.npz": gen_imresize, "test_mode.npz": gen_test_mode, "unet_small.npz": gen_unet_small, "unet_ffhq.npz": gen_unet_ffhq, "sisr.npz": gen_sisr,
         "misc.npz": gen_misc, "loops.npz": gen_loops, "loop_u2.npz": gen_loop_u2, "model_fn.npz": gen_model_fn}


def main(check=False):
    os.makedirs(GOLD, exist_ok=True)
    ok = True
    for fn, gen in FILES.items():
        torch.manual_seed(0)
        data = gen()
        path = os.path.join(GOLD, fn)
        if check:
            old = np.load(path)
            for key, v in data.items():
                same = np.allclose(old[key], v, rtol=1e-5, atol=1e-6)
                ok &= same
                if not same:
                    print(f"MISMATCH {fn}:{key} max|d|={np.abs(old[key] - v).max()}")
        else:
            np.savez_compressed(path, **data)
            print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")
    return ok


if __name__ == "__main__":
    good = main(check="--check" in sys.argv)
    sys.exit(0 if good else 1)
