"""Seeded synthetic weights, images, blur kernels, masks and noise.

The reference checkpoints (model_zoo/*.pt) and validation images are not available offline
(SURVEY.md section 8(c)), so every test and benchmark runs on deterministic synthetic data of the
reference's shapes.  All generation happens on the CPU with explicit torch/numpy generators so the
same bytes are produced in the build container and on the GPU box.
"""
import math

import numpy as np
import torch

from .arch import UNetConfig, param_specs


def synthetic_state_dict(cfg: UNetConfig, seed: int = 0, dtype=torch.float32):
    """A checkpoint-shaped state_dict with non-trivial values everywhere (the reference zero-inits
    out_layers.3 / proj_out / out.2 via zero_module, unet.py:210,294,615 - those are overwritten too,
    otherwise the network would output zeros)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in param_specs(cfg):
        if name.endswith("bias"):
            if ".in_layers.0." in name or ".out_layers.0." in name or ".norm." in name or name == "out.0.bias":
                t = 0.1 * torch.randn(shape, generator=g)
            else:
                t = 0.02 * torch.randn(shape, generator=g)
        elif len(shape) == 1:            # norm weight
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:
            fan_in = int(np.prod(shape[1:]))
            gain = 0.7
            if "emb_layers" in name:
                gain = 0.3
            t = (gain / math.sqrt(fan_in)) * torch.randn(shape, generator=g)
        sd[name] = t.to(dtype)
    return sd


def synthetic_images(batch, size=256, seed=1234):
    """uint8 HWC 'FFHQ-shape' images: smoothed uniform noise, so blurs/PSNR are meaningful."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(batch, 3, size // 8, size // 8, generator=g)
    x = torch.nn.functional.interpolate(x, size=(size, size), mode="bicubic", align_corners=False)
    x = x + 0.05 * torch.randn(x.shape, generator=g)
    x = (x.clamp(0, 1) * 255.0).round().to(torch.uint8)
    return x.permute(0, 2, 3, 1).contiguous().numpy()      # (B,H,W,3) uint8


def gaussian_kernel(size=61, std=3.0, dtype=np.float32):
    """Same construction as Blurkernel.weights_init, utils/utils_deblur.py:658-664 (scipy gaussian_filter
    of a centred delta)."""
    import scipy.ndimage
    n = np.zeros((size, size))
    n[size // 2, size // 2] = 1
    return scipy.ndimage.gaussian_filter(n, sigma=std).astype(dtype)


def motion_kernel(size=61, intensity=0.5, seed=0):
    """Random-walk motion-blur PSF.  The reference delegates to the un-vendored `motionblur` package
    (utils/utils_deblur.py:7,668); parity only needs the same k on both sides, so this is our own
    seeded generator: a smoothed random walk rasterised into a size x size grid, normalised to sum 1."""
    rng = np.random.RandomState(seed)
    steps = 64
    ang = rng.uniform(0, 2 * np.pi)
    pos = np.zeros(2)
    pts = [pos.copy()]
    for _ in range(steps):
        ang += intensity * rng.randn() * 1.5
        pos = pos + np.array([np.cos(ang), np.sin(ang)]) * (size / 3.0 / steps) * (1 + rng.rand())
        pts.append(pos.copy())
    pts = np.array(pts)
    pts -= pts.mean(0)
    k = np.zeros((size, size), np.float64)
    c = size // 2
    for p in pts:
        fx, fy = p[0] + c, p[1] + c
        x0, y0 = int(np.floor(fx)), int(np.floor(fy))
        for dy in (0, 1):
            for dx in (0, 1):
                xx, yy = x0 + dx, y0 + dy
                if 0 <= xx < size and 0 <= yy < size:
                    k[yy, xx] += (1 - abs(fx - xx)) * (1 - abs(fy - yy))
    k /= k.sum()
    return k.astype(np.float32)


def levin_like_kernels(seed=9):
    """Stand-ins for kernels/Levin09.mat (MATLAB v7.3/HDF5, unreadable offline - SURVEY.md section 7 #7):
    eight seeded motion PSFs with the Levin sizes 19,17,15,27,13,21,23,23."""
    return [motion_kernel(s, 0.6, seed + i) for i, s in enumerate((19, 17, 15, 27, 13, 21, 23, 23))]


def blur_measurement(img_u8, k, noise_level_img, seed=0):
    """Measurement synthesis of main_ddpir.py:97-114: circular blur on uint8 (scipy returns uint8,
    truncating), /255, AWGN of std 2*sigma in [-1,1] space.  Returns float32 HWC in ~[0,1]."""
    from scipy import ndimage
    img_L = ndimage.convolve(img_u8, np.expand_dims(k, axis=2), mode="wrap")
    img_L = np.float32(img_L / 255.0)
    rng = np.random.RandomState(seed)
    img_L = img_L * 2 - 1
    img_L = img_L + rng.normal(0, noise_level_img * 2, img_L.shape)
    img_L = img_L / 2 + 0.5
    return img_L.astype(np.float32)


def random_mask(batch, size=256, prob=0.5, seed=0):
    """mask_generator._retrieve_random, utils/utils_inpaint.py:113-125: one (H,W) 0/1 mask per image,
    repeated over the 3 channels; exactly int(H*W*prob) pixels are masked."""
    rng = np.random.RandomState(seed)
    masks = np.ones((batch, 1, size * size), np.float32)
    for b in range(batch):
        samples = rng.choice(size * size, int(size * size * prob), replace=False)
        masks[b, 0, samples] = 0
    return torch.from_numpy(masks.reshape(batch, 1, size, size)).repeat(1, 3, 1, 1)


def box_mask(batch, size=256, box=128, margin=16, seed=0):
    """random_sq_bbox, utils/utils_inpaint.py:67-84."""
    rng = np.random.RandomState(seed)
    m = torch.ones(batch, 3, size, size)
    for b in range(batch):
        t = rng.randint(margin, size - margin - box)
        l = rng.randint(margin, size - margin - box)
        m[b, :, t:t + box, l:l + box] = 0
    return m


def predrawn_noise(shape, n_steps, seed=42):
    """x_T noise + one zeta-noise tensor per step (RNG parity is handled by feeding the same pre-drawn
    noise to both sides)."""
    g = torch.Generator().manual_seed(seed)
    init = torch.randn(shape, generator=g)
    steps = [torch.randn(shape, generator=g) for _ in range(n_steps)]
    return init, steps
