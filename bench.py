#!/usr/bin/env python
"""bench.py - images/sec @ 100 NFE on 256x256 synthetic batches (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config 2|3|4|5] [--impl reference] [--batch B]
                    [--precision fp16s|fp16|fp16x3|fp32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" is one iteration of the DiffPIR loop over one batch: UNet denoise (eps->x0 fused) -> closed-form data-fidelity prox
-> re-noise, i.e. one NFE for every image of the batch; the timed K steps walk the real 100-entry quad schedule (t_i, rho_i,
re-noise coefficients) cyclically, the zeta-noise is drawn on the device inside the timed loop (the reference draws it with
randn_like every step, main_ddpir.py:455-456).  images/sec @ 100 NFE = (images per step over all ranks) / (100 x s per step).

Workloads = BASELINE.json `configs` (config.workload names the one that ran; --config selects):
  2 (default, the headline): Gaussian/motion deblur with Levin-size kernels, FFHQ architecture, 32 images PER GPU (weak scaling)
  3: SISR x4 (FFT prox with the sf x sf alias fold), ImageNet architecture, 32 images IN TOTAL over the N GPUs (strong scaling)
  4: random 50 % inpainting, FFHQ architecture, 64 images in total (strong scaling)
  5: motion deblur 61x61, ImageNet architecture, 128 images in total (strong scaling)
Image shards are independent: the only collective is ONE NCCL broadcast of the packed weight arena at
init; after it every rank proves it computes rank 0's answers (`rank_consistency`).

JSON keys follow the driver contract.  `roofline` is for the dominant kernel family (the tcgen05 implicit-GEMM convolutions):
achieved = algorithmic conv FLOPs per step / conv-kernel time per step IN the timed region = (conv share of a per-launch CUDA-event
profile) x (the CUDA-graph forward as the loop runs it, events around the whole forward); the per-launch event intervals themselves
(which include the launch gap an event pair opens) are reported under `serialised`.  peak = MEASURED_PEAKS.json
bf16_tflops_sustained (fp16 and bf16 share the tensor-pipe rate).  `strict_mode` (N = 1,
config 2) is the same step in the split-precision tensor-core mode fp16x3 that meets north_star's <= 1e-4 rel-L2 bar (three MMA
passes per convolution, FLOPs counted once), with its own images/s, roofline and parity sample.
"""
import argparse
import glob
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from diffpir_b200 import arch, synth  # noqa: E402

NFE, SIZE = 100, 256
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}

# BASELINE.json configs[1..4]; hyper-parameters from the reference's own sweeps (main_ddpir.py:548-580 on configs/*.yaml):
# deblur lambda 1*7, zeta 0.1*3; sr lambda 1*{2..12} (6 used), zeta 0.25; inpaint lambda 1, zeta 1, sigma max(0.001, 0)
WORKLOADS = {
    2: dict(task="deblur", model="ffhq", per_gpu=32, total=None, sf=1, sigma=12.75 / 255, lam=7.0, zeta=0.3, kernels="levin",
            text="Gaussian/motion deblur (Levin-size kernels), {b} synthetic FFHQ-shape 256x256 images per GPU, 100 NFE quad schedule, "
                 "diffusion_ffhq_10m architecture (random-init), sigma_n=12.75/255, lambda=7, zeta=0.3"),
    3: dict(task="sr", model="imagenet", per_gpu=None, total=32, sf=4, sigma=12.75 / 255, lam=6.0, zeta=0.25, kernels="lowpass25",
            text="4x SISR (closed-form FFT prox, 25x25 low-pass kernel), {t} synthetic ImageNet-shape 256x256 images in total ({b} per "
                 "GPU), 100 NFE, 256x256_diffusion_uncond architecture (random-init), sigma_n=12.75/255, lambda=6, zeta=0.25"),
    4: dict(task="inpaint", model="ffhq", per_gpu=None, total=64, sf=1, sigma=0.0, lam=1.0, zeta=1.0, kernels=None,
            text="random 50% inpainting, {t} synthetic FFHQ-shape 256x256 images in total ({b} per GPU), 100 NFE, diffusion_ffhq_10m "
                 "architecture (random-init), sigma_n=0 (sigma=0.001), lambda=1, zeta=1"),
    5: dict(task="deblur", model="imagenet", per_gpu=None, total=128, sf=1, sigma=12.75 / 255, lam=7.0, zeta=0.3, kernels="motion61",
            text="motion deblur (61x61 seeded random-walk kernels, intensity 0.5), {t} synthetic ImageNet-shape 256x256 images in total "
                 "({b} per GPU), 100 NFE, 256x256_diffusion_uncond architecture (random-init), sigma_n=12.75/255, lambda=7, zeta=0.3"),
}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        d["_source"] = "measured"
        return d
    d = dict(FALLBACK_PEAKS)
    d["_source"] = "fallback"
    return d


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_cores():
    """Usable host threads: scheduler affinity, capped by the cgroup CPU quota (os.cpu_count() over-reports in containers)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        q, p = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = max(1, min(n, int(float(q) / float(p))))
    except Exception:
        pass
    return min(n, 32)      # the oracle's torch-CPU convs stop scaling (and thrash) far below the box's thread count


def metric_name(wl):
    return "images/sec @100-NFE 256x256 " + {"deblur": "deblur", "sr": "sisr x4", "inpaint": "inpaint"}[wl["task"]]


def model_cfg(wl):
    return arch.ffhq_config() if wl["model"] == "ffhq" else arch.imagenet_config()


def kernels_for(wl, B, seed):
    """(B,1,kh,kw) host tensor of per-image blur kernels (zero-padded to the largest), or None."""
    if wl["kernels"] is None:
        return None
    if wl["kernels"] == "levin":
        ks = synth.levin_like_kernels()
        ks = [ks[b % len(ks)] for b in range(B)]
    elif wl["kernels"] == "motion61":
        ks = [synth.motion_kernel(61, 0.5, seed=1000 * seed + b) for b in range(B)]
    else:                                   # the 25x25 SR kernel (kernels_bicubicx234.mat is not shippable: a seeded low-pass stand-in)
        ks = [synth.gaussian_kernel(25, 1.6)] * B
    kmax = max(k.shape[0] for k in ks)
    kt = torch.zeros(B, 1, kmax, kmax)
    for b, k in enumerate(ks):
        o = (kmax - k.shape[0]) // 2
        kt[b, 0, o:o + k.shape[0], o:o + k.shape[1]] = torch.from_numpy(k)
    return kt


def make_inputs(wl, B, seed, dev=None):
    """Per-rank synthetic measurement batch: (y in [0,1], k or None, mask or None).  With `dev` the measurement is synthesised
    on the GPU with the library's own data-edge kernels (uint8 wrap blur, MATLAB bicubic); without, on the host (CPU baseline)."""
    imgs = synth.synthetic_images(B, SIZE, seed=1234 + seed)
    kt = kernels_for(wl, B, seed)
    mask = None
    if wl["task"] == "deblur":
        if dev is not None:
            from diffpir_b200 import utils_image
            noise = np.random.RandomState(77 + seed).normal(0, 2 * wl["sigma"], imgs.shape).astype(np.float32)
            y = utils_image.blur_measurement_batch(imgs, kt[:, 0].numpy().astype(np.float64), noise=noise).cpu()
        else:
            ks = kt[:, 0].numpy()
            y = np.stack([synth.blur_measurement(imgs[b], ks[b], wl["sigma"], seed=b) for b in range(B)])
            y = torch.from_numpy(y).permute(0, 3, 1, 2).contiguous()
    elif wl["task"] == "sr":
        gt = torch.from_numpy(imgs).permute(0, 3, 1, 2).float() / 255.0
        y = torch.nn.functional.avg_pool2d(gt, wl["sf"])                       # any LR measurement of the right shape
        y = (y + torch.from_numpy(np.random.RandomState(78 + seed).normal(0, wl["sigma"], tuple(y.shape)).astype(np.float32))).contiguous()
    else:
        y = (torch.from_numpy(imgs).permute(0, 3, 1, 2).float() / 255.0).contiguous()
        mask = synth.random_mask(B, SIZE, 0.5, seed=3 + seed)
    return y, kt, mask


def workload_config(cfg_id, wl, B, world, precision):
    total = B * world
    return {"workload": f"config {cfg_id}: " + wl["text"].format(b=B, t=total), "baseline_config": cfg_id, "task": wl["task"],
            "per_gpu_batch": B, "global_batch": total, "nfe": NFE, "precision": precision,
            "parallelism": f"image-shard over {world} GPU(s), one NCCL weight broadcast, no collective in the loop",
            "l2": "per-step activations (GBs) exceed the 126 MB L2; no explicit flush between steps"}


# ----------------------------------------------------------------------------------------------- CPU oracle legs
class CpuLoop:
    """The reference's CPU path for one loop step of the selected workload: the oracle port of the PyTorch code (the reference
    itself is Python and cannot travel to the GPU box), B = 1 image."""

    def __init__(self, wl, cores):
        from oracle import diffpir_oracle as orc
        from oracle import unet_oracle
        torch.set_num_threads(cores)
        self.orc, self.unet, self.wl = orc, unet_oracle, wl
        self.cfg = model_cfg(wl)
        self.sd = synth.synthetic_state_dict(self.cfg, 0)
        self.y, self.k, self.mask = make_inputs(wl, 1, 0)
        self.sched = orc.driver_schedule()
        self.tabs = orc.diffusion_tables()
        self.FBt = orc.pre_calculate(self.y, self.k, wl["sf"]) if wl["task"] != "inpaint" else None

    def step(self, x, t_i, t_im1, rho, noise):
        orc, wl = self.orc, self.wl
        vec_t = torch.tensor([t_i] * x.shape[0])
        x0 = orc.pred_xstart(self.unet.unet_forward(self.sd, self.cfg.as_dict(), x, vec_t), x, vec_t, self.tabs)
        tau = torch.tensor(rho, dtype=torch.float32).view(1, 1, 1, 1)
        if wl["task"] == "inpaint":
            x0 = orc.inpaint_prox(x0, 2 * self.y - 1, self.mask, tau, 1.0)
        else:
            x0 = orc.data_solution(x0 / 2 + 0.5, *self.FBt, tau, wl["sf"]) * 2 - 1
        x = orc.renoise(x, x0, t_i, t_im1, self.sched, 0.0, wl["zeta"], torch.zeros_like(x), noise)
        return x, x0


def run_reference(args, wl):
    """--impl reference: the CPU path timed on the box's host cores, all threads, bounded sample (B = 1 image per step)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from diffpir_b200.sampler import StepTable
    cores = host_cores()
    cpu = CpuLoop(wl, cores)
    tb = StepTable(sigma=wl["sigma"], lambda_=wl["lam"], zeta=wl["zeta"], iter_num=NFE)
    steps = [st for st in tb.steps if not st["last"]]
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 3, SIZE, SIZE, generator=g)
    with torch.no_grad():
        def one(j, x):
            st = steps[j % len(steps)]
            return cpu.step(x, st["t_i"], st["t_im1"], st["rho"], torch.randn(x.shape, generator=g))[0]
        for w in range(args.warmup):
            x = one(w, x)
        t0 = time.perf_counter()
        for s in range(args.steps):
            x = one(args.warmup + s, x)
        dt = time.perf_counter() - t0
    ms = dt / max(args.steps, 1) * 1e3
    value = 1 / (NFE * ms / 1e3)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    B = args.batch or wl["per_gpu"] or max(1, wl["total"] // world)
    line = {"impl": "reference", "metric": metric_name(wl), "value": value, "unit": "images/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak" if wl["per_gpu"] else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(args.config, wl, B, world, "fp32-cpu"), sample_batch=1),
            "cpu_baseline": {"value": value, "unit": "images/s", "cores": cores, "kind": "port",
                             "sample": f"{args.steps} loop steps (UNet + prox + re-noise) of B=1 image on {cores} host threads, scaled to 100 NFE"},
            "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def cpu_baseline_sample(wl, keep=None):
    """~10-30 s of the oracle's CPU evaluation of the same step, rank 0 / N=1 only.  With `keep` (a dict) the inputs, the
    pre-drawn noise and the result of the short trajectory are recorded so that parity_sample() can replay it on the GPU."""
    cores = host_cores()
    cpu = CpuLoop(wl, cores)
    nsteps = 3 if wl["model"] == "ffhq" else 2
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 3, SIZE, SIZE, generator=g)
    noises = [torch.randn(1, 3, SIZE, SIZE, generator=g) for _ in range(nsteps)]
    x_init = x.clone()
    t_i, t_im1, rho = 500, 480, 0.01
    with torch.no_grad():
        cpu.unet.unet_forward(cpu.sd, cpu.cfg.as_dict(), x, torch.tensor([t_i]))      # warm-up
        t0 = time.perf_counter()
        for i in range(nsteps):
            x, x0 = cpu.step(x, t_i, t_im1, rho, noises[i])
        dt = (time.perf_counter() - t0) / nsteps
    if keep is not None:
        keep.update(y=cpu.y, k=cpu.k, mask=cpu.mask, x_init=x_init, noises=noises, x_final=x, x0_last=x0, t_i=t_i, t_im1=t_im1, rho=rho,
                    nsteps=nsteps)
    return {"value": 1 / (NFE * dt), "unit": "images/s", "cores": cores, "kind": "port",
            "sample": f"{nsteps} loop steps of B=1 image ({wl['model']} UNet + {wl['task']} prox + re-noise) with the CPU oracle on {cores} "
                      f"threads, scaled to 100 NFE ({dt:.2f} s/step)"}


class GpuLoop:
    """The product path for one workload: model.x0 -> prox -> renoise, all through the C-ABI."""

    def __init__(self, wl, model, y, k, mask, dev):
        from diffpir_b200 import ops, utils_sisr
        self.ops, self.sr, self.wl, self.model = ops, utils_sisr, wl, model
        B = y.shape[0]
        self.y, self.mask = y.to(dev), (mask.to(dev) if mask is not None else None)
        self.fft = wl["task"] != "inpaint"
        if self.fft:
            self.FB, _, self.F2B, self.FBFy = utils_sisr.pre_calculate(self.y, k.to(dev), wl["sf"])
            self.spec = torch.empty(B * 3 * SIZE * (SIZE // 2 + 1) * 2 * (2 if wl["sf"] > 1 else 1), device=dev, dtype=torch.float32)
        self.x0 = torch.empty(B, 3, SIZE, SIZE, device=dev)

    def prox(self, rho):
        if self.fft:
            self.sr.data_solution_fused(self.x0, self.FB, self.F2B, self.FBFy, rho, self.wl["sf"], 0.5, 0.5, 2.0, -1.0, 1.0,
                                        out=self.x0, spec=self.spec)
        else:
            self.ops.inpaint_prox(self.x0, self.y, self.mask, rho, 1.0, out=self.x0)

    def step(self, x, t_vec, st, noise):
        self.model.x0(x, t_vec, out=self.x0)
        self.prox(st["rho"])
        self.ops.renoise(x, self.x0, noise, st["sa_t"], st["s1_t"], st["sa_p"], st["c_eps"], st["c_zeta"], out=x)


def parity_sample(wl, model, keep, dev):
    """The trajectory cpu_baseline_sample() just computed on the CPU (fp32 oracle), replayed through the product path on the GPU
    with the same weights, measurement, start and pre-drawn noise: the "PSNR vs ref" half of the metric on a bounded sample.
    Never allowed to take the bench line down: any failure is reported in the result."""
    try:
        from diffpir_b200.sampler import driver_schedule
        s = driver_schedule()
        sa, s1 = s["sqrt_alphas_cumprod"], s["sqrt_1m_alphas_cumprod"]
        t_i, t_im1 = keep["t_i"], keep["t_im1"]
        st = dict(rho=keep["rho"], sa_t=float(sa[t_i]), s1_t=float(s1[t_i]), sa_p=float(sa[t_im1]),
                  c_eps=float(np.sqrt(1 - wl["zeta"]) * s1[t_im1]), c_zeta=float(np.sqrt(wl["zeta"]) * s1[t_im1]))
        loop = GpuLoop(wl, model, keep["y"], keep["k"], keep["mask"], dev)
        x = keep["x_init"].to(dev).clone()
        t = torch.full((1,), t_i, device=dev, dtype=torch.int64)
        for i in range(keep["nsteps"]):
            loop.step(x, t, st, keep["noises"][i].to(dev))
        torch.cuda.synchronize()
        xg, x0g = x.cpu().double(), loop.x0.cpu().double()
        xr, x0r = keep["x_final"].double(), keep["x0_last"].double()
        mse = float((((x0g - x0r) / 2) ** 2).mean())               # images in [0,1]
        return {"what": f"{keep['nsteps']}-step trajectory (UNet -> {wl['task']} prox -> re-noise at t={t_i}) of one 256x256 image, "
                        f"GPU product path ({model.precision}) vs the fp32 CPU oracle, identical weights / measurement / noise",
                "rel_l2_x": float((xg - xr).norm() / xr.norm()), "rel_l2_x0": float((x0g - x0r).norm() / x0r.norm()),
                "psnr_db_gpu_vs_oracle_x0": float(10 * np.log10(1.0 / max(mse, 1e-30)))}
    except Exception as exc:        # noqa: BLE001 - reported, not raised
        return {"error": f"{type(exc).__name__}: {exc}"}


def conv_traffic():
    """DRAM bytes of the dominant kernel from the newest committed `ncu --set full` capture (profiles/r*_ncu_conv*.json written
    by tools/ncu_summary.py; it cannot be measured live): launch 0 = the first ResBlock convolution of the forward (256x256,
    128->128, B=32, fp16 in/out), beside its algorithmic bytes (read x once + write h once)."""
    best = None
    for p in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_conv*.json"))):
        try:
            d = json.load(open(p))
        except Exception:
            continue
        if not isinstance(d, dict):
            continue
        pl = d.get("per_launch")
        if not (isinstance(pl, list) and pl):
            continue              # round-1 summaries hold totals over launches of different shapes: not a per-launch figure
        # the first fused-operand launch of the capture = in[1].conv1 (captures that start at the forward's first launch also hold the
        # stem, a k_conv3x3_v3<0> launch with a 3-channel input, before it)
        fused = [l for l in pl if "k_conv3x3_v3<3>" in l.get("kernel", "")]
        if not fused:
            continue
        rd, wr, kern, n = fused[0]["dram__bytes_read.sum"], fused[0]["dram__bytes_write.sum"], fused[0].get("kernel", ""), 1
        best = {"dram_bytes_per_launch": (float(rd) + float(wr)) / n, "dram_read_bytes_per_launch": float(rd) / n, "kernel": kern,
                "algorithmic_bytes_per_launch": 2 * 32 * 256 * 256 * 128 * 2,
                "launch": "first ResBlock convolution of the forward: 256x256, 128->128, B=32, fp16 in/out (read x once + write h once)",
                "source": os.path.relpath(p, ROOT) + " (dram__bytes_read.sum + dram__bytes_write.sum)"}
    return best


def conv_roofline(model, x, x0, ts, nsteps, B, peaks, profile_steps, graph_reps=10):
    """Roofline of the convolution kernels of one forward.

    Two live CUDA-event measurements on the launch stream:
      (1) the forward exactly as the timed loop issues it (CUDA-graph replay, no events inside): `unet_ms_per_step` - what the
          kernels cost IN the timed region;
      (2) `profile_steps` forwards launched eagerly with an event pair around EVERY launch (dpir_profile_*): per-kernel-kind times.
          These intervals also contain the launch gap an event pair opens (~5-10 us per launch, 145 launches: the serialised sum
          exceeds (1) by 5-13 %), so they are used for what they measure robustly - the SHARE of each kernel kind (cross-check:
          the ncu launch list of the same command, profiles/r*_launches_bench_summary.md).
    conv time in the timed region = share_conv x (1); small kernels carry relatively more gap than the convolutions, so the
    convolutions' share - and with it their time - is if anything over-estimated.  `achieved` = algorithmic conv FLOPs / that time;
    the serialised figures are reported beside it (`serialised`)."""
    for j in range(2):                        # re-plan / re-capture after a mode switch, warm-up
        model.x0(x, ts[(7 * j) % nsteps], out=x0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for j in range(graph_reps):
        model.x0(x, ts[(7 * j) % nsteps], out=x0)
    e1.record()
    torch.cuda.synchronize()
    unet_ms_region = e0.elapsed_time(e1) / graph_reps
    model.profile_begin()
    for j in range(profile_steps):
        model.x0(x, ts[(7 * j) % nsteps], out=x0)
    rows = model.profile_end()
    by_kind = {}
    conv_flops = 0.0
    for r in rows:
        by_kind[r["kind"]] = by_kind.get(r["kind"], 0.0) + r["ms"] / profile_steps
        if r["kind"] == "conv":
            H, W, Cin, Cout, ks, Cin2 = r["dims"]
            conv_flops += 2.0 * B * H * W * (ks * ks * Cin + Cin2) * Cout
    conv_ser = by_kind.get("conv", 0.0)
    unet_ser = sum(by_kind.values())
    share = conv_ser / unet_ser if unet_ser else 0.0
    conv_ms = share * unet_ms_region
    peak = peaks.get("bf16_tflops_sustained", FALLBACK_PEAKS["bf16_tflops_sustained"])
    achieved = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    ach_ser = conv_flops / (conv_ser * 1e-3) / 1e12 if conv_ser > 0 else 0.0
    return {"achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
            "conv_ms_per_step": conv_ms, "unet_ms_per_step": unet_ms_region, "share_of_unet": share,
            "method": "conv time = (conv share of the per-launch event profile) x (CUDA-graph forward as the timed loop runs it, "
                      "events around the whole forward)",
            "serialised": {"note": "eager launches with an event pair around every launch (includes ~5-10 us of launch gap per launch)",
                           "conv_ms_per_step": conv_ser, "unet_ms_per_step": unet_ser, "achieved": ach_ser,
                           "frac": ach_ser / peak if peak else None},
            "conv_ms_per_step_serialised": conv_ser, "unet_ms_per_step_serialised": unet_ser,
            "ms_by_kernel_kind": {k2: round(v, 4) for k2, v in sorted(by_kind.items())}, "algorithmic_gflop_per_step": conv_flops / 1e9}


# ----------------------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--config", type=int, default=2, choices=sorted(WORKLOADS), help="BASELINE.json config (2 = headline)")
    ap.add_argument("--batch", type=int, default=0, help="images per GPU (default: the workload's)")
    ap.add_argument("--precision", default="fp16s", choices=["fp16", "fp16s", "fp16x3", "fp32"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--weight-bcast", default="torch", choices=["torch", "library"],
                    help="N > 1: who issues the one collective - torch.distributed.broadcast of the arena view (default), or the "
                         "library's own NCCL communicator (dpir_bcast_weights)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strict-leg", action="store_true", help="skip the fp16x3 (<= 1e-4 parity mode) leg of config 2 at N = 1")
    ap.add_argument("--profile-steps", type=int, default=3)
    ap.add_argument("--gn-fusion", type=int, default=1, choices=[0, 1, 2, 4],
                    help="ResBlock GroupNorm+SiLU inside the convolutions: 1 (default, activate-once producers), 4 (round-1 per-tap "
                         "producers), 2 (per-tap packed half), 0 separate passes")
    args = ap.parse_args()
    wl = WORKLOADS[args.config]
    if args.impl == "reference":
        run_reference(args, wl)
        return
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (the engine has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    # stdout carries exactly ONE JSON line.  Libraries may write there from C (NCCL prints its version banner when NCCL_DEBUG=VERSION
    # is set in the environment or in /etc/nccl.conf), so file descriptor 1 points at stderr for the whole run and the line is
    # written to the saved descriptor at the end.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        import torch.distributed as dist
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)

    from diffpir_b200 import _lib, ops
    from diffpir_b200.sampler import StepTable
    from diffpir_b200.unet import UNetModel

    if args.batch:
        B, scaling = args.batch, "weak"
    elif wl["per_gpu"]:
        B, scaling = wl["per_gpu"], "weak"
    else:
        if wl["total"] % world:
            raise RuntimeError(f"config {args.config}: {wl['total']} images do not split over {world} GPUs")
        B, scaling = wl["total"] // world, "strong"
    cfg = model_cfg(wl)
    model = UNetModel(cfg, precision=args.precision)
    t_bcast_ms = 0.0
    sd = None
    if rank == 0:
        sd = synth.synthetic_state_dict(cfg, 0)
        model.load_state_dict(sd)
    if world > 1:
        arena = model.weight_arena()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if args.weight_bcast == "library":
            from diffpir_b200.parallel import WeightComm
            wcomm = WeightComm()              # the library's own communicator (id shared through the job's process group)
            torch.cuda.synchronize()
            dist.barrier()
            e0.record()
            wcomm.broadcast(model, src=0)     # ONE ncclBroadcast issued by libdiffpir_b200 on the current stream
            e1.record()
            torch.cuda.synchronize()
            wcomm.close()
        else:
            e0.record()
            dist.broadcast(arena, src=0)          # the ONE collective: packed weights over NVLink
            e1.record()
            torch.cuda.synchronize()
            model.mark_weights_loaded()
        t_bcast_ms = e0.elapsed_time(e1)
    model.set_gn_fusion(args.gn_fusion)
    if not args.no_graph:
        model.set_graph_mode(True)

    # every rank proves that the weights it RECEIVED compute rank 0's answers: same seeded input, max |diff| of x0 vs rank 0
    rank_consistency = None
    if world > 1:
        gchk = torch.Generator().manual_seed(4242)
        xc = torch.randn(2, 3, SIZE, SIZE, generator=gchk).to(dev)
        oc = model.x0(xc, torch.tensor([999, 250], device=dev))
        gathered = [torch.empty_like(oc) for _ in range(world)]
        dist.all_gather(gathered, oc)
        if rank == 0:
            diffs = [float((g_ - gathered[0]).abs().max()) for g_ in gathered]
            rank_consistency = {"what": "x0 of a seeded 2-image batch on every rank after the weight broadcast vs rank 0",
                                "max_abs_diff_per_rank": diffs, "bit_identical": all(d == 0.0 for d in diffs)}

    y_h, k_h, m_h = make_inputs(wl, B, rank, dev)
    loop = GpuLoop(wl, model, y_h, k_h, m_h, dev)
    tb = StepTable(sigma=wl["sigma"], lambda_=wl["lam"], zeta=wl["zeta"], iter_num=NFE)
    steps_tab = tb.steps
    nst = len(steps_tab)
    ts = torch.tensor([[st["t_i"]] * B for st in steps_tab], dtype=torch.int64, device=dev)
    gen = torch.Generator(device=dev).manual_seed(1000 + rank)
    x = torch.randn(B, 3, SIZE, SIZE, device=dev, generator=gen)
    noise = torch.empty_like(x)
    x0 = loop.x0

    def one_step(j, x, noise, draw=True):
        st = steps_tab[j % nst]
        if draw:
            noise.normal_(generator=gen)             # the step's zeta-noise (main_ddpir.py:456), drawn inside the timed region
        if st["last"]:
            # trajectory finished (its x_0 = x/2+0.5 would be read out here): the final NFE is executed, its output unused
            # (SURVEY 3.1); the next trajectory starts from fresh noise
            model.x0(x, ts[j % nst], out=x0)
            x.copy_(noise)
            return
        loop.step(x, ts[j % nst], st, noise)

    def sync_all():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, n, finish=None):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        e0.record()
        for j in range(n):
            fn(j)
        if finish is not None:
            finish()                                   # e.g. join side streams so that their work is inside the timed region
        e1.record()
        sync_all()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            tt = torch.tensor([ms], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        return ms / n

    # ------------------------------------------------------------------ device-resident timing ("value")
    for j in range(args.warmup):
        one_step(j, x, noise)
    sync_all()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    _lib.launch_count_reset()
    ms_step = timed(lambda j: one_step(args.warmup + j, x, noise), args.steps)
    launches = _lib.launch_count()
    clocks = sampler.stop() if rank == 0 else None
    value = world * B / (NFE * ms_step / 1e3)

    # ------------------------------------------------------------------ end-to-end timing through host buffers ("e2e")
    # Every step's state + noise come from pinned host memory and its result goes back to the host, all inside the timed region.
    # The copies are pipelined like a real host-fed deployment would: step j+1's inputs are uploaded on a copy stream while step j
    # computes, step j's result is downloaded on a second copy stream; two device buffer sets, events order the reuse.
    xh = torch.randn(B, 3, SIZE, SIZE).pin_memory()
    nh = torch.randn(B, 3, SIZE, SIZE).pin_memory()
    oh = [torch.empty(B, 3, SIZE, SIZE).pin_memory() for _ in range(2)]
    xd = [torch.empty_like(x) for _ in range(2)]
    nd = [torch.empty_like(x) for _ in range(2)]
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_done = [torch.cuda.Event() for _ in range(2)]
    ev_out = [torch.cuda.Event() for _ in range(2)]
    cur = torch.cuda.current_stream()

    def upload(b):
        with torch.cuda.stream(s_in):
            s_in.wait_event(ev_out[b])                 # the previous result held in this buffer set has left the device
            xd[b].copy_(xh, non_blocking=True)
            nd[b].copy_(nh, non_blocking=True)
            ev_in[b].record(s_in)

    def e2e_step(j):
        b = j & 1
        cur.wait_event(ev_in[b])                       # this step's inputs (uploaded during the previous step)
        upload(b ^ 1)                                  # next step's inputs, overlapped with this step's compute
        one_step(j, xd[b], nd[b], draw=False)
        ev_done[b].record(cur)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_done[b])
            oh[b].copy_(xd[b], non_blocking=True)      # the step's result back to the host
            ev_out[b].record(s_out)
    for b_ in range(2):
        ev_out[b_].record(cur)
    upload(0)
    k_e2e = max(10, min(args.steps, 30))
    for j in range(3):
        e2e_step(j)
    ms_e2e = timed(lambda j: e2e_step(3 + j), k_e2e, finish=lambda: [cur.wait_event(ev) for ev in ev_out])
    e2e_value = world * B / (NFE * ms_e2e / 1e3)
    img_bytes = B * 3 * SIZE * SIZE * 4

    peaks = load_peaks()
    # ------------------------------------------------------------------ roofline of the dominant kernel family (conv)
    roofline = None
    if rank == 0:
        rf = conv_roofline(model, x, x0, ts, nst, B, peaks, args.profile_steps)
        ab = None
        if args.gn_fusion and args.precision == "fp16s":
            # A/B in the same process: the same forward with every GroupNorm as a separate pass (convolution-only kernels)
            model.set_gn_fusion(0)
            model.x0(x, ts[0], out=x0)                  # re-plan + warm-up
            rab = conv_roofline(model, x, x0, ts, nst, B, peaks, args.profile_steps)
            model.set_gn_fusion(args.gn_fusion)
            model.x0(x, ts[0], out=x0)
            ab = {"note": "same forward with GroupNorm+SiLU as separate k_gn_apply passes (gn_fusion=0): convolution-only kernels",
                  "achieved": rab["achieved"], "frac": rab["frac"], "conv_ms_per_step": rab["conv_ms_per_step"],
                  "unet_ms_per_step": rab["unet_ms_per_step"], "serialised": rab["serialised"],
                  "unet_ms_per_step_serialised": rab["unet_ms_per_step_serialised"], "ms_by_kernel_kind": rab["ms_by_kernel_kind"]}
        tr = conv_traffic()
        roofline = dict(rf, bound="tensor",
                        kernel="k_conv3x3_v3<FUSE> / k_conv_tc / k_conv3x3_dxn (tcgen05 implicit-GEMM conv: every 3x3 / 1x1 layer of one forward, "
                               "stem and output head included; with gn_fusion the ResBlock and head convs also evaluate GroupNorm+FiLM+SiLU "
                               "on their operand)",
                        separate_norm_passes_ab=ab, peak_source=f"MEASURED_PEAKS.json bf16_tflops_sustained ({peaks['_source']})",
                        traffic=(tr or {}).get("dram_bytes_per_launch"), traffic_detail=tr,
                        whole_step_tflops=world * B * arch.flops_per_image(cfg) / (ms_step * 1e-3) / 1e12)

    # ------------------------------------------------------------------ HBM roofline of the prox / re-noise kernels
    prox = None
    if rank == 0:
        hbm = peaks.get("hbm_gbs", FALLBACK_PEAKS["hbm_gbs"])
        st = steps_tab[10]
        reps = 20

        def timed_local(fn):
            for i in range(3):
                fn(i)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(reps):
                fn(i)
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps

        ms_prox = timed_local(lambda i: loop.prox(st["rho"]))
        # re-noise with the L2 defeated: rotate over buffer sets whose total footprint (8 x 4 x B images) exceeds the 126 MB L2
        nset = max(2, int(np.ceil(160e6 / (4 * B * 3 * SIZE * SIZE * 4))) + 1)
        sets = [[torch.randn(B, 3, SIZE, SIZE, device=dev) for _ in range(4)] for _ in range(nset)]
        ms_rn = timed_local(lambda i: ops.renoise(sets[i % nset][0], sets[i % nset][1], sets[i % nset][2], st["sa_t"], st["s1_t"],
                                                  st["sa_p"], st["c_eps"], st["c_zeta"], out=sets[i % nset][3]))
        del sets
        by_prox = {"deblur": 2497536, "sr": 2761728, "inpaint": 3145728}[wl["task"]] * B     # SURVEY.md section 8(d), per image per step
        by_rn = 3145728 * B
        prox = {"bound": "hbm", "peak": hbm, "unit": "GB/s", "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peaks['_source']})",
                "data_fidelity_prox": {"task": wl["task"], "ms": ms_prox, "algorithmic_bytes": by_prox, "achieved": by_prox / ms_prox / 1e6,
                                       "frac": by_prox / ms_prox / 1e6 / hbm,
                                       "note": "cuFFT R2C + frequency-domain solve + cuFFT C2R (3 launches), compulsory traffic only"
                                       if wl["task"] != "inpaint" else "k_inpaint_prox (1 launch)"},
                "renoise": {"ms": ms_rn, "algorithmic_bytes": by_rn, "achieved": by_rn / ms_rn / 1e6, "frac": by_rn / ms_rn / 1e6 / hbm,
                            "note": f"out-of-place over {nset} rotating buffer sets ({nset * 4 * B * 3 * SIZE * SIZE * 4 / 1e6:.0f} MB > L2)"},
                "share_of_step": (ms_prox + ms_rn) / ms_step}

    # ------------------------------------------------------------------ CPU baseline + parity sample + strict-mode leg
    cpu = parity = strict = None
    if rank == 0 and world == 1:
        keep = {}
        if not args.no_cpu_baseline:
            cpu = cpu_baseline_sample(wl, keep)
            parity = parity_sample(wl, model, keep, dev)
        if args.config == 2 and not args.no_strict_leg and args.precision != "fp16x3":
            try:
                del loop
                mx = UNetModel(cfg, precision="fp16x3")
                mx.load_state_dict(sd)
                mx.set_graph_mode(not args.no_graph)
                lx = GpuLoop(wl, mx, y_h, k_h, m_h, dev)

                def strict_step(j):
                    st = steps_tab[j % nst]
                    noise.normal_(generator=gen)
                    if st["last"]:
                        mx.x0(x, ts[j % nst], out=lx.x0)
                        x.copy_(noise)
                    else:
                        lx.step(x, ts[j % nst], st, noise)
                for j in range(3):
                    strict_step(j)
                ksx = max(5, min(args.steps, 12))
                ms_x = timed(lambda j: strict_step(3 + j), ksx)
                rfx = conv_roofline(mx, x, lx.x0, ts, nst, B, peaks, 2)
                strict = {"precision": "fp16x3", "what": "the same step in the split-precision tensor-core mode (fp16 hi/lo operand pairs, "
                          "three tcgen05 passes per convolution, fp32 residual stream and attention): the <= 1e-4 rel-L2 configuration",
                          "value": B / (NFE * ms_x / 1e3), "unit": "images/s", "ms_per_step": ms_x, "steps": ksx,
                          "roofline": dict(rfx, bound="tensor", note="algorithmic conv FLOPs counted ONCE (the tensor core executes 3x)"),
                          "parity_sample": parity_sample(wl, mx, keep, dev) if keep else None}
                del mx, lx
            except Exception as exc:        # noqa: BLE001 - reported, never takes the bench line down
                strict = {"error": f"{type(exc).__name__}: {exc}"}

    if rank == 0:
        line = {"metric": metric_name(wl), "value": value, "unit": "images/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": scaling, "vs_baseline": None, "dtype": {"fp32": "f32", "fp16x3": "f16x3"}.get(args.precision, "f16"),
                "data": "synthetic", "config": workload_config(args.config, wl, B, world, args.precision),
                "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": 2 * img_bytes,
                        "d2h_bytes_per_step": img_bytes, "ms_per_step": ms_e2e},
                "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "roofline_prox": prox, "cpu_baseline": cpu,
                "parity_sample": parity, "strict_mode": strict, "rank_consistency": rank_consistency,
                "weight_broadcast_ms": t_bcast_ms, "weight_broadcast_by": args.weight_bcast if world > 1 else None, "graph": not args.no_graph, "gn_fusion": args.gn_fusion}
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
