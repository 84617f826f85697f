#!/usr/bin/env python
"""Per-op device time of one UNet forward (CUDA events inside the engine) as a table: dims, ms, TFLOP/s or GB/s.
Usage: python tools/profile_ops.py [--batch 32] [--precision fp16] [--reps 3] [--model ffhq|imagenet] [--size 256]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from diffpir_b200 import arch, synth  # noqa: E402
from diffpir_b200.unet import UNetModel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--precision", default="fp16")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--model", default="ffhq")
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--json", default="")
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--variant", type=int, default=0, help="dpir_conv_tc_variant: 0 auto (v3), 1 v1, 2 v2")
    ap.add_argument("--gn-fusion", type=int, default=1, help="0 separate GroupNorm passes, 1 fused into the convs (activate-once producers), 4 round-1 per-tap producers, 2 per-tap half math")
    a = ap.parse_args()
    from diffpir_b200 import _lib
    _lib.check(_lib.load().dpir_conv_tc_variant(a.variant), "dpir_conv_tc_variant")
    cfg = arch.ffhq_config() if a.model == "ffhq" else arch.imagenet_config()
    m = UNetModel(cfg, precision=a.precision)
    m.load_state_dict(synth.synthetic_state_dict(cfg, 0))
    m.set_gn_fusion(a.gn_fusion)
    B = a.batch
    x = torch.randn(B, 3, a.size, a.size, device="cuda")
    t = torch.full((B,), 500, device="cuda", dtype=torch.int64)
    for _ in range(a.warmup):
        m.x0(x, t)
    torch.cuda.synchronize()
    if a.reps <= 0:
        torch.cuda.synchronize()
        print("warm-up forwards only (no per-op timing requested)")
        return
    m.profile_begin()
    for _ in range(a.reps):
        m.x0(x, t)
    rows = m.profile_end()
    tot = sum(r["ms"] for r in rows) / a.reps
    agg = {}
    print(f"{'kind':10s} {'dims':28s} {'ms':>8s} {'rate':>14s}")
    by_kind = {}
    for r in rows:
        ms = r["ms"] / a.reps
        d = r["dims"]
        rate = ""
        if r["kind"] == "conv":
            H, W, Cin, Cout, ks, Cin2 = d
            fl = 2.0 * B * H * W * (ks * ks * Cin + Cin2) * Cout
            rate = f"{fl / ms / 1e9:8.1f} TF/s" + (" +gn" if r.get("gn_fused") else "")
            key = ("conv+gn" if r.get("gn_fused") else "conv", H, Cin, Cout, ks, Cin2)
        elif r["kind"] == "gn_apply":
            H, W, C, rs = d[:4]
            by = B * H * W * C * ((2 if a.precision == "fp16s" else 4) + 2)      # one read + one write of the tensor
            rate = f"{by / ms / 1e6:8.1f} GB/s~"
            key = ("gn_apply", H, C, rs, 0, 0)
        elif r["kind"] == "gn_stats":
            HW, _, C = d[:3]
            by = B * HW * C * 4
            rate = f"{by / ms / 1e6:8.1f} GB/s"
            key = ("gn_stats", HW, C, 0, 0, 0)
        else:
            key = (r["kind"], 0, 0, 0, 0, 0)
        by_kind[r["kind"]] = by_kind.get(r["kind"], 0.0) + ms
        agg.setdefault(key, [0.0, 0])
        agg[key][0] += ms
        agg[key][1] += 1
        print(f"{r['kind']:10s} {str(d):28s} {ms:8.4f} {rate:>14s}")
    print(f"\ntotal serialised {tot:.3f} ms per forward (B={B}); by kind: " + ", ".join(f"{k} {v:.3f}" for k, v in sorted(by_kind.items())))
    print("\naggregated by shape (ms per forward, count):")
    for k, (ms, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
        print(f"  {str(k):40s} {ms:8.3f} ms  x{n}")
    if a.json:
        json.dump(rows, open(a.json, "w"))


if __name__ == "__main__":
    main()
