#!/usr/bin/env python
"""Design check (pure numpy, no GPU) of the narrow-output convolution k_conv3x3_dxn (conv_tc.cu): replay its tiling and index
arithmetic on the CPU and compare with a direct 3x3 convolution.

  * tiles of 14 x 16 output pixels; per 64-channel chunk ONE {64 ch, 16 px, 18 rows} halo box at (w0 - 1, h0 - 1), zero outside the image;
  * the weights [tap = 3*dy + dx][16][Cin] viewed as [dy][dx*16 + cout][Cin]: 48 rows per vertical tap (the {64, 48, 1} TMA box);
  * per (chunk, dy, half-tile mb): D[p][dx*16 + o] += sum_c box[row(p) + dy + 8*mb][col(p)][c] * W3[dy][dx*16 + o][c]
    with M row p = 16 * yy + xb (yy 0..7, xb 0..15) - the tcgen05 MMA of the kernel;
  * epilogue: out(y = h0 + 8*mb + yy, x = w0 + xb - 1) = D[p - 1][o] + D[p][16 + o] + D[p + 1][32 + o] for xb in 1..14 (the two
    neighbours are lane - 1 / lane + 1 of the same warp: 16 consecutive lanes hold one image row);
  * fused flavour: the box is activated in place (here: any elementwise function with f(0) != 0) and set to exact zero outside the
    image, because the convolution pads the ACTIVATED tensor.

    python tools/check_dxn_formulation.py
"""
import numpy as np


def direct_conv(x, w, act):
    """x (H, W, C), w (O, C, 3, 3) -> (H, W, O): conv3x3(act(x)) with zero padding of the activated tensor."""
    H, W, C = x.shape
    a = np.zeros((H + 2, W + 2, C))
    a[1:-1, 1:-1] = act(x)
    out = np.zeros((H, W, w.shape[0]))
    for dy in range(3):
        for dx in range(3):
            out += a[dy:dy + H, dx:dx + W] @ w[:, :, dy, dx].T
    return out


def dxn_conv(x, w, act):
    H, W, C = x.shape
    O = w.shape[0]
    assert O <= 16 and C % 64 == 0
    packed = np.zeros((9, 16, C))                                   # [tap][16][Cin], rows O..15 zero
    packed[:, :O] = w.transpose(2, 3, 0, 1).reshape(9, O, C)
    w3 = packed.reshape(3, 48, C)                                   # [dy][dx*16 + o][Cin]: the same bytes, a different view
    out = np.full((H, W, O), np.nan)
    tiles_w, tiles_h = -(-W // 14), -(-H // 16)
    for th in range(tiles_h):
        for tw in range(tiles_w):
            h0, w0 = th * 16, tw * 14
            D = np.zeros((2, 128, 48))
            for ch in range(C // 64):
                box = np.zeros((18, 16, 64))                        # TMA zero fill
                valid = np.zeros((18, 16), bool)
                for r in range(18):
                    for c in range(16):
                        y, xx = h0 - 1 + r, w0 - 1 + c
                        if 0 <= y < H and 0 <= xx < W:
                            box[r, c] = x[y, xx, ch * 64:(ch + 1) * 64]
                            valid[r, c] = True
                box = np.where(valid[..., None], act(box), 0.0)    # activators: in place, exact zeros outside the image
                for dy in range(3):
                    for mb in range(2):
                        A = box[dy + 8 * mb: dy + 8 * mb + 8].reshape(128, 64)      # rows p = 16*yy + xb: one contiguous view
                        D[mb] += A @ w3[dy][:, ch * 64:(ch + 1) * 64].T
            for mb in range(2):
                for p in range(128):
                    yy, xb = divmod(p, 16)
                    y, xo = h0 + 8 * mb + yy, w0 + xb - 1
                    if 1 <= xb <= 14 and xo < W and y < H:
                        out[y, xo] = D[mb, p - 1, 0:O] + D[mb, p, 16:16 + O] + D[mb, p + 1, 32:32 + O]
    return out


def check(H, W, C, O, seed, fused):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((H, W, C))
    w = rng.standard_normal((O, C, 3, 3)) / np.sqrt(9 * C)
    act = (lambda v: v / (1 + np.exp(-(1.3 * v + 0.4))) + 0.25) if fused else (lambda v: v)    # f(0) != 0 on purpose
    ref, got = direct_conv(x, w, act), dxn_conv(x, w, act)
    assert not np.isnan(got).any(), "a pixel was never written"
    err = np.abs(got - ref).max()
    assert err < 1e-10, (H, W, C, O, fused, err)
    return err


def main():
    for (H, W, C, O) in [(16, 16, 64, 6), (32, 32, 128, 6), (24, 40, 128, 3), (48, 30, 64, 16), (17, 45, 64, 1)]:
        for fused in (False, True):
            print(f"H {H} W {W} C {C} O {O} fused {fused}: max |diff| {check(H, W, C, O, 5, fused):.2e}")
    print("k_conv3x3_dxn formulation: ok")


if __name__ == "__main__":
    main()
