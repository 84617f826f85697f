"""GPU edges of the loop added late in round 1 (kept in a file that sorts last): MATLAB-style bicubic imresize (the SR
measurement synthesis of main_ddpir.py:89 / main_ddpir_sisr.py:202) through dpir_resize_separable, against fixtures produced by
the unmodified reference (tests/golden/imresize.npz) and the CPU oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import diffpir_oracle as orc

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "imresize.npz")


@pytest.mark.parametrize("key,scale,aa", [("a_down2", 1 / 2, True), ("a_down3", 1 / 3, True), ("a_down4", 1 / 4, True),
                                          ("a_up2", 2, True), ("b_down2", 1 / 2, True), ("a_up3_noaa", 3, False)])
def test_imresize_np_matches_reference_golden(key, scale, aa):
    from diffpir_b200 import utils_image as ui
    g = np.load(GOLD)
    out = ui.imresize_np(g[key[0]], scale, aa)
    assert out.shape == g[key].shape and out.dtype == np.float32
    assert np.abs(out - g[key]).max() < 2e-6          # fp32 summation order only


def test_imresize_batch_256_sf4_matches_oracle():
    """The benchmark-size case: (B,3,256,256) -> (B,3,64,64), each image equal to the oracle's imresize_np of it."""
    from diffpir_b200 import utils_image as ui
    x = torch.rand(3, 3, 256, 256, generator=torch.Generator().manual_seed(21))
    out = ui.imresize(x.cuda(), 1 / 4)
    assert tuple(out.shape) == (3, 3, 64, 64)
    for b in range(3):
        ref = orc.imresize_np(x[b].permute(1, 2, 0).numpy(), 1 / 4)
        assert np.abs(out[b].permute(1, 2, 0).cpu().numpy() - ref).max() < 2e-6
