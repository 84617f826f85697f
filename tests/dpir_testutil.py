"""Shared helpers for the test-suite."""
import numpy as np
import torch

from diffpir_b200 import arch, synth


def rel_l2(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def psnr(a, b, peak=1.0):
    mse = float(((a.detach().double().cpu() - b.detach().double().cpu()) ** 2).mean())
    return 10 * np.log10(peak ** 2 / max(mse, 1e-30))


def small_config():
    """A 64x64 UNet with the reference's structure (4 levels, attention at 16x16 and 8x8, head dim 64)."""
    return arch.UNetConfig(image_size=64, model_channels=128, num_res_blocks=1, attention_resolutions="16,8",
                           num_head_channels=64)


_sd_cache = {}


def weights(cfg, seed=0):
    key = (cfg.model_channels, cfg.num_res_blocks, cfg.attention_resolutions, cfg.image_size, seed)
    if key not in _sd_cache:
        _sd_cache[key] = synth.synthetic_state_dict(cfg, seed)
    return _sd_cache[key]


def make_model(cfg, precision, seed=0):
    from diffpir_b200.unet import UNetModel
    m = UNetModel(cfg, precision=precision)
    m.load_state_dict(weights(cfg, seed))
    return m
