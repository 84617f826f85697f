"""Inpainting masks with the reference's name and call convention (utils/utils_inpaint.py:67-137):

    mask_generator(mask_type, mask_len_range=None, mask_prob_range=None, image_size=256, margin=(16, 16))(img) -> mask

Host logic, as in the reference: the draws come from the GLOBAL numpy stream in the reference's order (so a driver that
calls np.random.seed(...) - main_ddpir.py:150 - gets bit-identical masks), the mask is a 0/1 float tensor shaped like `img`
on `img.device`.  The masked least-squares prox that consumes it is the CUDA kernel behind ops.inpaint_prox.
"""
import numpy as np
import torch


def random_sq_bbox(img, mask_shape, image_size=256, margin=(16, 16)):
    """One box of zeros per call, the same box for every image / channel of the batch (utils_inpaint.py:67-84).
    Returns (mask, top, bottom, left, right)."""
    B, C, H, W = img.shape
    h, w = mask_shape
    top = np.random.randint(margin[0], image_size - margin[0] - h)
    left = np.random.randint(margin[1], image_size - margin[1] - w)
    mask = torch.ones(B, C, H, W, device=img.device)
    mask[..., top:top + h, left:left + w] = 0
    return mask, top, top + h, left, left + w


class mask_generator:
    TYPES = ('box', 'random', 'both', 'extreme')

    def __init__(self, mask_type, mask_len_range=None, mask_prob_range=None, image_size=256, margin=(16, 16)):
        assert mask_type in self.TYPES
        self.mask_type = mask_type
        self.mask_len_range = mask_len_range
        self.mask_prob_range = mask_prob_range
        self.image_size = image_size
        self.margin = margin

    def _retrieve_box(self, img):
        lo, hi = int(self.mask_len_range[0]), int(self.mask_len_range[1])
        mask_h = np.random.randint(lo, hi)
        mask_w = np.random.randint(lo, hi)
        return random_sq_bbox(img, (mask_h, mask_w), self.image_size, self.margin)

    def _retrieve_random(self, img):
        n = self.image_size * self.image_size
        prob = np.random.uniform(*self.mask_prob_range)
        dropped = np.random.choice(n, int(n * prob), replace=False)      # exactly int(n * prob) masked pixels
        flat = torch.ones(n)
        flat[torch.from_numpy(dropped)] = 0
        plane = flat.view(1, 1, self.image_size, self.image_size).to(img.device)
        return plane.expand(img.shape[0], img.shape[1], -1, -1).clone() if img.dim() == 4 else plane[0].expand_as(img).clone()

    def __call__(self, img):
        if self.mask_type == 'random':
            return self._retrieve_random(img)
        if self.mask_type == 'box':
            return self._retrieve_box(img)[0]
        if self.mask_type == 'extreme':
            return 1. - self._retrieve_box(img)[0]
        # 'both' is accepted by the constructor and returns None in the reference as well (utils_inpaint.py:127-137)
        return None
