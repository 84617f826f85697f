#!/usr/bin/env python
"""Full 100-NFE parity of the headline configuration against the fp64 evaluation of the reference formulae (evidence run, not
collected by pytest: the fp64 CPU oracle needs minutes; it lives under tests/ because it executes oracle/, which is test infrastructure).

BASELINE config-2 shape - FFHQ architecture, 256x256, a Levin-size kernel, sigma 12.75/255, lambda 7, zeta 0.3, the full 100-entry
quad schedule, one image, identical pre-drawn noise on every side:
    ref64 = CPU oracle in fp64          (the reference's formulae without its fp32 cancellation)
    ref32 = CPU oracle in fp32          (pinned to the unmodified reference's goldens; `dev` = its own distance to ref64)
    GPU   = fp16s (throughput mode), fp16x3 (split-precision tensor-core mode), fp32 (CUDA cores), through the C-ABI
Prints rel-L2 of the iterate x at steps 1 / 25 / 50 / 75 / 99 and of the final image, PSNR differences vs ground truth.
Usage: python tests/parity_100nfe.py > profiles/rNN_parity_100nfe.txt
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from diffpir_b200 import arch, synth  # noqa: E402
from diffpir_b200.sampler import DiffPIRSampler  # noqa: E402
from diffpir_b200.unet import UNetModel  # noqa: E402
from oracle import diffpir_oracle as orc  # noqa: E402


def rel_l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm())


def psnr(a, b):
    mse = float(((a.detach().double().cpu() - b.detach().double().cpu()) ** 2).mean())
    return 10 * np.log10(1.0 / max(mse, 1e-30))


def main():
    torch.set_num_threads(min(32, len(os.sched_getaffinity(0))))
    cfg = arch.ffhq_config()
    sd = synth.synthetic_state_dict(cfg, 0)
    iters = 100
    imgs = synth.synthetic_images(1, 256, seed=33)
    gt = torch.from_numpy(imgs).permute(0, 3, 1, 2).float() / 255.0
    k = synth.levin_like_kernels()[3]
    y = torch.from_numpy(np.stack([synth.blur_measurement(imgs[0], k, 12.75 / 255, seed=0)])).permute(0, 3, 1, 2).contiguous()
    kt = torch.from_numpy(k)[None, None]
    init, zs = synth.predrawn_noise((1, 3, 256, 256), iters, seed=17)
    hp = dict(lambda_=7.0, zeta=0.3, sigma=12.75 / 255)
    t0 = time.time()
    tr64 = []
    ref64 = orc.diffpir_loop(sd, cfg.as_dict(), "deblur", y.double(), k=kt.double(), sf=1, iter_num=iters, init_noise=init.double(),
                             zeta_noise=[z.double() for z in zs], trace=tr64, **hp)
    t64 = time.time() - t0
    t0 = time.time()
    tr32 = []
    ref32 = orc.diffpir_loop(sd, cfg.as_dict(), "deblur", y, k=kt, sf=1, iter_num=iters, init_noise=init, zeta_noise=zs, trace=tr32, **hp)
    t32 = time.time() - t0
    x64 = {i: v for (n, i, v) in tr64 if n == "x"}
    x32 = {i: v for (n, i, v) in tr32 if n == "x"}
    marks = (0, 24, 49, 74, 98)
    print(f"# 100-NFE FFHQ deblur, one image, identical pre-drawn noise; CPU oracle fp64 {t64:.0f} s, fp32 {t32:.0f} s")
    print(f"# rel-L2 of the iterate x after step 1/25/50/75/99 and of the final x_0, |dPSNR| vs ground truth, all against the fp64 oracle")

    def line(name, xs, out):
        r = " / ".join(f"{rel_l2(xs[i], x64[i]):.2e}" for i in marks)
        print(f"{name:28s} x: {r}   final: {rel_l2(out, ref64):.2e}   |dPSNR| {abs(psnr(out, gt) - psnr(ref64, gt)):.5f} dB   "
              f"PSNR(out, ref64) {psnr(out, ref64):.1f} dB")

    line("CPU oracle fp32 (= reference)", x32, ref32)
    for prec in ("fp16s", "fp16", "fp16x3", "fp32"):
        m = UNetModel(cfg, precision=prec)
        m.load_state_dict(sd)
        m.set_graph_mode(True)
        tr = []
        out = DiffPIRSampler(m, "deblur", iter_num=iters, **hp).run(y.cuda(), k=kt.cuda(), init_noise=init.cuda(),
                                                                     zeta_noise=[z.cuda() for z in zs], trace=tr)
        torch.cuda.synchronize()
        line(f"GPU {prec}", {i: v for (n, i, v) in tr if n == "x"}, out)
        del m
    dev = rel_l2(ref32, ref64)
    print(f"# bar for the strict modes (fp16x3, fp32): max(1e-4, 1.5 x dev) with dev = rel-L2(ref32, ref64) = {dev:.2e} -> {max(1e-4, 1.5 * dev):.2e}")


if __name__ == "__main__":
    main()
