"""Pin the CPU oracle against the committed golden fixtures (outputs of the UNMODIFIED reference, produced by
oracle/gen_golden.py in the build container).  Runs anywhere - no GPU, no /root/reference."""
import os

import numpy as np
import pytest
import torch

from diffpir_b200 import arch, synth
from oracle import diffpir_oracle as orc
from oracle import gen_golden, unet_oracle
from dpir_testutil import rel_l2

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    return np.load(os.path.join(GOLD, name))


def test_unet_small_and_pred_xstart():
    g = _load("unet_small.npz")
    cfg = gen_golden.small_cfg()
    sd = synth.synthetic_state_dict(cfg, 0)
    x, t = torch.from_numpy(g["x"]), torch.from_numpy(g["t"])
    out = unet_oracle.unet_forward(sd, cfg.as_dict(), x, t)
    assert rel_l2(out, torch.from_numpy(g["out"])) < 5e-6
    x0 = orc.pred_xstart(out, x, t)
    assert rel_l2(x0, torch.from_numpy(g["pred_xstart"])) < 2e-5
    # fp64 evaluation of the oracle stays within fp32 round-off of the fp32 reference
    out64 = unet_oracle.unet_forward(sd, cfg.as_dict(), x.double(), t)
    assert rel_l2(out64, torch.from_numpy(g["out"])) < 5e-6


def test_unet_ffhq_architecture():
    g = _load("unet_ffhq.npz")
    cfg = arch.ffhq_config()
    sd = synth.synthetic_state_dict(cfg, 0)
    x = torch.randn(1, 3, 256, 256, generator=torch.Generator().manual_seed(int(g["seed"])))
    out = unet_oracle.unet_forward(sd, cfg.as_dict(), x, torch.from_numpy(g["t"]))
    assert rel_l2(out[:, :, ::4, ::4], torch.from_numpy(g["out_sub"])) < 5e-6
    assert abs(float(out.double().std()) - float(g["out_std"])) < 1e-5


def test_fft_solver():
    g = _load("sisr.npz")
    for tag, sf, alpha, y, x, k in gen_golden.sisr_cases():
        a = torch.tensor(alpha).view(1, 1, 1, 1)
        FB, FBC, F2B, FBFy = orc.pre_calculate(y, k, sf)
        o32 = orc.data_solution(x, FB, FBC, F2B, FBFy, a, sf)
        assert rel_l2(o32, torch.from_numpy(g[tag + "_out32"])) < 1e-5, tag
        t64 = orc.pre_calculate(y.double(), k.double(), sf)
        o64 = orc.data_solution(x.double(), *t64, a.double(), sf)
        assert rel_l2(o64, torch.from_numpy(g[tag + "_out64"])) < 1e-12, tag
        if sf == 4 and alpha == 0.3:
            assert rel_l2(torch.view_as_real(FB), torch.from_numpy(g["FB_sf4"])) < 1e-6
            assert rel_l2(F2B, torch.from_numpy(g["F2B_sf4"])) < 1e-6
            assert rel_l2(torch.view_as_real(FBFy), torch.from_numpy(g["FBFy_sf4"])) < 1e-6


def test_resizer_masks_blur():
    g = _load("misc.npz")
    x = torch.from_numpy(g["resize_x"])
    assert rel_l2(orc.resizer_forward(x, 0.25), torch.from_numpy(g["resize_down4"])) < 1e-6
    assert rel_l2(orc.resizer_forward(x, 2.0), torch.from_numpy(g["resize_up2"])) < 1e-6
    w, fov = orc.resizer_contributions(256, 64, 0.25)
    assert np.array_equal(fov.T.astype(np.int32), g["fov256_h"])           # index table: bit-exact
    assert np.array_equal(w.T.astype(np.float32), g["w256_h"])
    # masks: bit-exact against utils_inpaint.mask_generator under the same numpy seed
    np.random.seed(123)
    prob = np.random.uniform(0.5, 0.5)
    samples = np.random.choice(64 * 64, int(64 * 64 * prob), replace=False)
    m = np.ones(64 * 64, np.uint8)
    m[samples] = 0
    assert np.array_equal(np.broadcast_to(m.reshape(1, 1, 64, 64), (1, 3, 64, 64)), g["mask_random"])
    assert int((g["mask_box"][0, 0] == 0).sum()) == 16 * 16
    assert np.array_equal(synth.gaussian_kernel(21, 2.5, dtype=np.float64), g["gauss_k"][0, 0])


@pytest.mark.parametrize("task,sf,hp", gen_golden.LOOP_CASES)
def test_full_loop(task, sf, hp):
    g = _load("loops.npz")
    iters = int(g["iters"])
    cfg = gen_golden.small_cfg()
    sd = synth.synthetic_state_dict(cfg, 0)
    gt, y, k, mask = gen_golden.loop_inputs(task, sf=sf)
    init, zs = synth.predrawn_noise((2, 3, 64, 64), iters, seed=7)
    x_init = torch.nn.functional.interpolate(y, scale_factor=sf, mode="bicubic", align_corners=False) if task == "sr" else None
    out = orc.diffpir_loop(sd, cfg.as_dict(), task, y, k=k, mask=mask, sf=sf, iter_num=iters, init_noise=init,
                           zeta_noise=zs, x_init=x_init, **hp)
    assert rel_l2(out, torch.from_numpy(g[task + "_x0"])) < 2e-5


def test_full_loop_iter_num_U():
    """The oracle's iter_num_U = 2 branch (main_ddpir.py:349, 460-467) against the golden of the reference loop arithmetic."""
    g = _load("loop_u2.npz")
    iters = int(g["iters"])
    cfg = gen_golden.small_cfg()
    sd = synth.synthetic_state_dict(cfg, 0)
    task, sf, hp = gen_golden.LOOP_CASES[0]
    gt, y, k, mask = gen_golden.loop_inputs(task, sf=sf)
    init, zs = synth.predrawn_noise((2, 3, 64, 64), 2 * iters, seed=8)
    _, us = synth.predrawn_noise((2, 3, 64, 64), iters, seed=9)
    out = orc.diffpir_loop(sd, cfg.as_dict(), task, y, k=k, sf=sf, iter_num=iters, init_noise=init, zeta_noise=zs,
                           iter_num_U=2, u_noise=us, **hp)
    assert rel_l2(out, torch.from_numpy(g["deblur_U2_x0"])) < 2e-5


def test_model_fn_goldens_present():
    """model_fn.npz (utils_model.model_fn output types from the reference's p_sample / ddim_sample) feeds the GPU drop-in tests;
    here: the identities that tie its entries together (utils/utils_model.py:241-256)."""
    g = _load("model_fn.npz")
    betas = np.linspace(0.0001, 0.02, 1000, dtype=np.float32)
    ac = np.cumprod(1.0 - betas, axis=0)
    red = np.sqrt(1.0 - ac) / np.sqrt(ac)
    t_step = int(np.abs(red - 37.0 / 255.).argmin())
    a, b = ac[t_step], 1 - ac[t_step]
    for tag in ("p", "ddim"):
        assert np.allclose(g[f"{tag}_score"], -g[f"{tag}_epsilon"] / b ** 0.5, rtol=1e-5, atol=1e-6)
        assert np.array_equal(g[f"{tag}_both_start"], g[f"{tag}_pred_xstart"])
        assert np.array_equal(g[f"{tag}_both_prev"], g[f"{tag}_pred_x_prev"])
    assert np.array_equal(g["p_pred_xstart"], g["ddim_pred_xstart"])
    assert float(np.abs(g["p_pred_xstart"]).max()) <= 1.0
