#!/usr/bin/env python
"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`) of the bench command: share of every kernel in the
steady-state launch time, one-time set-up kernels listed apart.
Usage: python tools/launch_summary.py gpurun_out/r02_launches_bench.csv profiles/r02_launches_bench_summary.md "<command line>" """
import csv
import re
import sys
from collections import defaultdict

SETUP = ("k_linear_rows", "k_measure_blur_u8", "k_time_embed0", "k_mul_conj", "k_abs2", "k_psf_place", "k_cast_w", "k_split_w")
CONV = ("k_conv3x3_v3", "k_conv3x3_v2", "k_conv_tc", "k_conv3x3_dxn")


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"^dpir::", "", name)
    m = re.match(r"([\w:<>, ]+?)\(", name)
    name = m.group(1) if m else name
    return name[:100]


def main():
    src, dst, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr, data = rows[0], rows[1:]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    steady, setup = defaultdict(lambda: [0.0, 0]), defaultdict(lambda: [0.0, 0])
    for r in data:
        us = float(r[iv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[iu], 1e-3)
        k = short(r[ik])
        d = setup if any(s in k for s in SETUP) else steady
        d[k][0] += us
        d[k][1] += 1
    tot = sum(v[0] for v in steady.values())
    n = sum(v[1] for v in steady.values())
    conv = sum(v[0] for k, v in steady.items() if any(c in k for c in CONV))
    with open(dst, "w") as f:
        f.write("# r02 - ncu launch list of the bench command\n\n")
        f.write(f"`{cmd}` on one B200 (raw: `{dst.rsplit('/', 1)[-1].replace('_summary.md', '.csv.gz')}`). Per-launch times under ncu are "
                "cold-cache and serialised: compare SHARES, not absolutes. The capture window covers the model set-up (listed "
                "separately) and the warm-up steps (graph capture replays the same kernels).\n\n")
        f.write(f"Steady-state kernels ({n} launches, {tot / 1e3:.2f} ms under ncu):\n\n| share | time, us | launches | kernel |\n|---|---|---|---|\n")
        for k, v in sorted(steady.items(), key=lambda kv: -kv[1][0]):
            f.write(f"| {100 * v[0] / tot:.2f} % | {v[0]:.1f} | {v[1]} | `{k}` |\n")
        f.write(f"\ntcgen05 convolution kernels ({', '.join('`' + c + '`' for c in CONV)}): **{100 * conv / tot:.1f} %** of the steady-state "
                "launch time.\n\nOne-time set-up kernels in the capture window:\n\n| time, us | launches | kernel |\n|---|---|---|\n")
        for k, v in sorted(setup.items(), key=lambda kv: -kv[1][0]):
            f.write(f"| {v[0]:.1f} | {v[1]} | `{k}` |\n")
    print(f"steady {n} launches {tot / 1e3:.2f} ms, conv share {100 * conv / tot:.1f} %")


if __name__ == "__main__":
    main()
