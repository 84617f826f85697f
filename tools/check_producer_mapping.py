#!/usr/bin/env python
"""Design check (pure numpy, no GPU): replay the thread -> (global element, shared-memory address) mapping of the operand
producers of k_conv3x3_v3<1> (shipped) and k_conv3x3_v3<3> (activate-once variant, not yet run on hardware) for one tile and
verify that every stage's halo box is written completely, exactly once, at the SWIZZLE_128B address the tensor core reads,
with zeros outside the image.  This checks the index arithmetic of conv_tc.cu, not the CUDA code itself.

    python tools/check_producer_mapping.py
"""
import numpy as np

BOX_ROWS, BOX_COLS, OCTETS = 18, 16, 8          # halo box: 18 x 16 pixels x 8 channel octets (64 fp16) = 2304 16-byte units


def expected_box(raw, h0, w0, dx):
    """Ground truth: unit (yy, xx, j) = id of raw[h0-1+yy, w0+dx-1+xx, octet j], or -1 (zero padding) outside the image."""
    H, W, _ = raw.shape
    box = -np.ones((BOX_ROWS, BOX_COLS, OCTETS), np.int64)
    for yy in range(BOX_ROWS):
        for xx in range(BOX_COLS):
            y, x = h0 - 1 + yy, w0 + dx - 1 + xx
            if 0 <= y < H and 0 <= x < W:
                box[yy, xx] = raw[y, x]
    return box


def swz(yy, xx, j):
    r = yy * 16 + xx
    return r * 128 + ((j ^ (r & 7)) << 4)


def read_back(smem):
    """What the MMA sees: unit (yy, xx, j) lives at swz(yy, xx, j)."""
    box = np.empty((BOX_ROWS, BOX_COLS, OCTETS), np.int64)
    for yy in range(BOX_ROWS):
        for xx in range(BOX_COLS):
            for j in range(OCTETS):
                box[yy, xx, j] = smem[swz(yy, xx, j) // 16]
    return box


def v4_stage(raw, h0, w0, dx):
    """k_conv3x3_v3<1>: 256 threads, j = p & 7, xx = (p >> 3) & 15, yh = p >> 7, rows yy = 2i + yh."""
    H, W, _ = raw.shape
    smem = np.full(2304, -7, np.int64)
    writes = np.zeros(2304, np.int64)
    for p in range(256):
        j, xx, yh = p & 7, (p >> 3) & 15, p >> 7
        dst_off = yh * 2048 + xx * 128 + ((j ^ (xx & 7)) << 4)
        x = w0 + dx - 1 + xx
        for i in range(9):
            y = h0 - 1 + yh + 2 * i
            ok = 0 <= x < W and 0 <= y < H
            a = (dst_off + i * 4096) // 16
            smem[a] = raw[y, x, j] if ok else -1
            writes[a] += 1
    return smem, writes


def v5_stage(raw, h0, w0, dx):
    """k_conv3x3_v3<3>: 288 threads, j = t & 7, Xl = t >> 3, X = Xl % 18, rg = Xl // 18, rows yy = 2i + rg; box column X - dx."""
    H, W, _ = raw.shape
    smem = np.full(2304, -7, np.int64)
    writes = np.zeros(2304, np.int64)
    for t in range(288):
        j, Xl = t & 7, t >> 3
        X, rg = Xl % 18, Xl // 18
        xx = X - dx
        if not 0 <= xx < 16:
            continue
        x = w0 - 1 + X
        for i in range(9):
            y = h0 - 1 + rg + 2 * i
            ok = 0 <= x < W and 0 <= y < H
            a = (rg * 2048 + xx * 128 + ((j ^ (xx & 7)) << 4) + i * 4096) // 16
            smem[a] = raw[y, x, j] if ok else -1
            writes[a] += 1
    return smem, writes


def main():
    H, W = 40, 24                                   # ragged: 24 is not a multiple of 16 -> partial tiles on the right edge
    raw = np.arange(H * W * OCTETS, dtype=np.int64).reshape(H, W, OCTETS)      # element ids
    for name, stage in (("v3<1>", v4_stage), ("v3<3>", v5_stage)):
        for h0 in (0, 16, 32):
            for w0 in (0, 16):
                for dx in range(3):
                    smem, writes = stage(raw, h0, w0, dx)
                    assert (writes == 1).all(), (name, h0, w0, dx, "every unit written exactly once")
                    assert (read_back(smem) == expected_box(raw, h0, w0, dx)).all(), (name, h0, w0, dx)
        print(f"{name}: 18 stages (3 x 2 tiles x 3 taps) - boxes complete, single-writer, swizzle-correct, zero padded")


if __name__ == "__main__":
    main()
