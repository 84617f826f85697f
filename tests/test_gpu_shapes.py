"""Shapes beyond the 256x256 benchmark: non-square inputs whose attention maps have token counts
that are not multiples of 64, utils_model.test_mode's quadrant splits through the REAL engine, and the 512x512 model family with
its fractional channel_mult (script_util.py:149-150)."""
import numpy as np
import pytest
import torch

from diffpir_b200 import arch
from oracle import diffpir_oracle as orc
from oracle import unet_oracle
from dpir_testutil import make_model, rel_l2, small_config, weights

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("fp16", 4e-3)])
def test_nonsquare_input_ragged_attention(precision, tol):
    """64-size config on a 96 x 80 image: attention at 24x20 = 480 and 12x10 = 120 tokens (neither a multiple of 64)."""
    cfg = small_config()
    sd = weights(cfg)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 3, 96, 80, generator=g)
    t = torch.tensor([700, 20])
    ref = unet_oracle.unet_forward(sd, cfg.as_dict(), x.double(), t)
    m = make_model(cfg, precision)
    out = m(x.cuda(), t.cuda())
    err = rel_l2(out, ref)
    print(f"96x80 input, {precision}: rel-L2 {err:.3e}")
    assert err < tol, err


@pytest.mark.parametrize("mode", [2, 5, 3])
def test_test_mode_through_the_real_engine(mode):
    """utils_model.test_mode (utils/utils_model.py:16-195) split / split-once / x8 modes with the real UNet behind model_fn, vs
    the same wrapper around the CPU oracle (the wrapper itself is golden-pinned against the reference in the CPU suite)."""
    from diffpir_b200 import utils_model
    cfg = small_config()
    sd = weights(cfg)
    tabs = orc.diffusion_tables()
    betas = torch.from_numpy(np.linspace(0.0001, 0.02, 1000, dtype=np.float32))
    alphas_cumprod = np.cumprod((1.0 - betas).cpu(), axis=0)
    red = torch.div(torch.sqrt(1.0 - alphas_cumprod), torch.sqrt(alphas_cumprod)).numpy()

    def oracle_model_fn(x, noise_level, model_diffusion=None, vec_t=None, model_out_type="pred_xstart", **kw):
        t_step = utils_model.find_nearest(red, float(noise_level) / 255.)
        vt = torch.tensor([t_step] * x.shape[0])
        return orc.pred_xstart(unet_oracle.unet_forward(sd, cfg.as_dict(), x, vt), x, vt, tabs)

    g = torch.Generator().manual_seed(15)
    L = torch.randn(1, 3, 96, 80, generator=g)
    kw = dict(mode=mode, refield=16, min_size=48, modulo=16, noise_level=40.0, alphas_cumprod=alphas_cumprod)
    ref = utils_model.test_mode(oracle_model_fn, None, L, **kw)
    m = make_model(cfg, "fp32")
    old = utils_model.RNG_COMPAT
    utils_model.RNG_COMPAT = False
    try:
        out = utils_model.test_mode(utils_model.model_fn, m, L.cuda(), **kw)
    finally:
        utils_model.RNG_COMPAT = old
    err = rel_l2(out, ref)
    print(f"test_mode {mode} through the engine: rel-L2 {err:.3e}")
    assert tuple(out.shape) == tuple(ref.shape) and err < 2e-4, err


def test_512_model_family_fractional_channel_mult():
    """image_size=512 -> channel_mult (0.5, 1, 1, 2, 2, 4, 4) (script_util.py:149-150): 7 levels, first level int(0.5 * 256) = 128
    channels.  The UNet is fully convolutional, so the check runs on a 128x128 crop (H, W % 64 == 0) to keep the CPU oracle cheap."""
    cfg = arch.UNetConfig(image_size=512, model_channels=256, num_res_blocks=1, attention_resolutions="32,16")
    assert cfg.channel_mult == (0.5, 1, 1, 2, 2, 4, 4)
    sd = weights(cfg)
    g = torch.Generator().manual_seed(9)
    x = torch.randn(1, 3, 128, 128, generator=g)
    t = torch.tensor([345])
    ref = unet_oracle.unet_forward(sd, cfg.as_dict(), x, t)
    for precision, tol in (("fp16", 4e-3), ("fp16s", 6e-3)):
        m = make_model(cfg, precision)
        out = m(x.cuda(), t.cuda())
        err = rel_l2(out, ref)
        print(f"512-family UNet {precision}: rel-L2 {err:.3e}")
        assert tuple(out.shape) == (1, 6, 128, 128) and err < tol, err
        del m
