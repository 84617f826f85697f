#!/usr/bin/env python
"""bench.py - images/sec @ 100 NFE, 256x256 Gaussian deblur (BASELINE.json configs[1]) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch B] [--precision fp16|fp32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" is one iteration of the DiffPIR loop over one batch: UNet denoise (eps->x0 fused) -> FFT-domain
data-fidelity prox -> re-noise, i.e. one NFE for every image of the batch; the timed K steps walk the real
100-entry quad schedule (t_i, rho_i, re-noise coefficients) cyclically.  images/sec @ 100 NFE =
(images per step over all ranks) / (100 x seconds per step).

Workload (config.workload): per GPU B=32 synthetic FFHQ-shape 256x256 images, diffusion_ffhq_10m architecture with
seeded random weights (no checkpoints offline), per-image Levin-size blur kernels, sigma_n = 12.75/255, lambda = 7,
zeta = 0.3, 100-NFE quad schedule.  Weak scaling: every rank processes its own B images (image shards are
independent); the only collective is one NCCL broadcast of the packed weight arena at init.

JSON keys follow the driver contract; `roofline` is for the dominant kernel (the tcgen05 implicit-GEMM convolution):
achieved = algorithmic conv FLOPs per step / summed conv-kernel time per step (CUDA events inside the engine, on
the launch stream), peak = MEASURED_PEAKS.json bf16_tflops_sustained (fp16 and bf16 share the tensor pipe rate).
"""
import argparse
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

SIGMA = 12.75 / 255.0
LAMBDA, ZETA, NFE, SIZE = 7.0, 0.3, 100, 256
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


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


def make_inputs(B, seed):
    """Per-rank synthetic measurement batch (host tensors): y in [0,1] (B,3,256,256), k (B,1,27,27) zero-padded Levin-size PSFs."""
    imgs = synth.synthetic_images(B, SIZE, seed=1234 + seed)
    ks = synth.levin_like_kernels()
    kmax = max(k.shape[0] for k in ks)
    kt = torch.zeros(B, 1, kmax, kmax)
    ys = []
    for b in range(B):
        k = ks[b % len(ks)]
        ys.append(synth.blur_measurement(imgs[b], k, SIGMA, seed=b))
        o = (kmax - k.shape[0]) // 2
        kt[b, 0, o:o + k.shape[0], o:o + k.shape[1]] = torch.from_numpy(k)
    y = torch.from_numpy(np.stack(ys)).permute(0, 3, 1, 2).contiguous()
    return y, kt


# ----------------------------------------------------------------------------------------------- reference arm
def run_reference(args):
    """The reference's CPU path for the same step (oracle port of the PyTorch code; the reference itself is Python and
    cannot travel to the GPU box), all host threads, bounded sample: B=1 image per step."""
    from oracle import diffpir_oracle as orc
    from oracle import unet_oracle
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    torch.set_num_threads(cores)
    cfg = arch.ffhq_config()
    sd = synth.synthetic_state_dict(cfg, 0)
    B = 1
    y, k = make_inputs(B, 0)
    sched = orc.driver_schedule()
    tabs = orc.diffusion_tables()
    rhos, sigmas = orc.rhos_sigmas(sched, LAMBDA, SIGMA)
    seq = orc.make_seq(1000, NFE)
    red = sched["reduced_alpha_cumprod"].numpy()
    FBt = orc.pre_calculate(y, k, 1)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, 3, SIZE, SIZE, generator=g)

    def step(i, x):
        t_i = orc.find_nearest(red, sigmas[seq[i]].numpy())
        vec_t = torch.tensor([t_i] * B)
        x0 = orc.pred_xstart(unet_oracle.unet_forward(sd, cfg.as_dict(), x, vec_t), x, vec_t, tabs)
        if seq[i] != seq[-1]:
            tau = rhos[t_i].float().reshape(1, 1, 1, 1)
            x0 = orc.data_solution(x0 / 2 + 0.5, *FBt, tau, 1) * 2 - 1
            t_im1 = orc.find_nearest(red, sigmas[seq[i + 1]].numpy())
            x = orc.renoise(x, x0, t_i, t_im1, sched, 0.0, ZETA, torch.zeros_like(x), torch.randn(x.shape, generator=g))
        return x

    with torch.no_grad():
        for w in range(args.warmup):
            x = step(w % (NFE - 1), x)
        t0 = time.perf_counter()
        for s in range(args.steps):
            x = step(s % (NFE - 1), x)
        dt = time.perf_counter() - t0
    ms = dt / max(args.steps, 1) * 1e3
    value = B / (NFE * ms / 1e3)
    line = {"impl": "reference", "metric": "images/sec @100-NFE 256x256 deblur", "value": value, "unit": "images/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(args.batch, "fp32-cpu"), sample_batch=B),
            "cpu_baseline": {"value": value, "unit": "images/s", "cores": cores, "kind": "port",
                             "sample": f"{args.steps} loop steps (UNet+FFT prox+re-noise) of B={B} image on {cores} host threads, "
                                       f"scaled to 100 NFE"},
            "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def conv_traffic():
    """DRAM bytes of the dominant kernel from the committed `ncu --set full` capture (cannot be measured live): the first
    ResBlock convolution of the forward (256x256, 128->128, B=32, GroupNorm+SiLU fused), beside its algorithmic bytes."""
    p = os.path.join(ROOT, "profiles", "r01_ncu_conv_v4_fused_full.md")
    if not os.path.exists(p):
        return None
    return {"dram_bytes_per_launch": 1074328000 + 509015552, "algorithmic_bytes_per_launch": 2 * 32 * 256 * 256 * 128 * 2,
            "launch": "k_conv3x3_v3<1> 256x256 128->128 B=32 fp16 in/out (read x once + write h once)",
            "source": "profiles/r01_ncu_conv_v4_fused_full.md (dram__bytes_read.sum + dram__bytes_write.sum, launch 0)"}


def workload_config(B, precision):
    return {"workload": f"Gaussian/motion deblur (Levin-size kernels), batch {B} synthetic FFHQ-shape 256x256 per GPU, 100 NFE "
                        f"quad schedule, diffusion_ffhq_10m architecture (random-init), sigma_n=12.75/255, lambda=7, zeta=0.3",
            "per_gpu_batch": B, "nfe": NFE, "precision": precision, "parallelism": "image-shard (weak scaling), one NCCL weight broadcast",
            "l2": "per-step activations (GBs) exceed the 126 MB L2; no explicit flush"}


# ----------------------------------------------------------------------------------------------- cpu baseline leg
def cpu_baseline_sample(keep=None):
    """~10-30 s of the oracle's CPU evaluation of the same step, rank 0 / N=1 only.  With `keep` (a dict) the inputs, the
    pre-drawn noise and the result of the short trajectory are recorded so that parity_sample() can replay it on the GPU."""
    from oracle import diffpir_oracle as orc
    from oracle import unet_oracle
    cores = host_cores()
    torch.set_num_threads(cores)
    cfg = arch.ffhq_config()
    sd = synth.synthetic_state_dict(cfg, 0)
    B, nsteps = 1, 3
    y, k = make_inputs(B, 0)
    sched = orc.driver_schedule()
    tabs = orc.diffusion_tables()
    FBt = orc.pre_calculate(y, k, 1)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, 3, SIZE, SIZE, generator=g)
    noises = [torch.randn(B, 3, SIZE, SIZE, generator=g) for _ in range(nsteps)]
    x_init = x.clone()
    vec_t = torch.tensor([500] * B)
    with torch.no_grad():
        unet_oracle.unet_forward(sd, cfg.as_dict(), x, vec_t)      # warm-up
        t0 = time.perf_counter()
        for i in range(nsteps):
            x0 = orc.pred_xstart(unet_oracle.unet_forward(sd, cfg.as_dict(), x, vec_t), x, vec_t, tabs)
            x0 = orc.data_solution(x0 / 2 + 0.5, *FBt, torch.tensor(0.01).view(1, 1, 1, 1), 1) * 2 - 1
            x = orc.renoise(x, x0, 500, 480, sched, 0.0, ZETA, torch.zeros_like(x), noises[i])
        dt = (time.perf_counter() - t0) / nsteps
    if keep is not None:
        keep.update(y=y, k=k, x_init=x_init, noises=noises, x_final=x, x0_last=x0, t_i=500, t_im1=480, rho=0.01, nsteps=nsteps)
    return {"value": B / (NFE * dt), "unit": "images/s", "cores": cores, "kind": "port",
            "sample": f"{nsteps} loop steps of B={B} image (FFHQ UNet + FFT prox + re-noise) with the CPU oracle on {cores} threads, "
                      f"scaled to 100 NFE ({dt:.2f} s/step)"}


def parity_sample(model, keep, dev):
    """The trajectory cpu_baseline_sample() just computed on the CPU (fp32 oracle), replayed through the product path on the
    GPU with the same weights, measurement, start and pre-drawn noise: the "PSNR vs ref" half of the metric on a bounded
    sample.  Never allowed to take the bench line down: any failure is reported in the result."""
    try:
        from diffpir_b200 import ops, utils_sisr
        from diffpir_b200.sampler import driver_schedule
        s = driver_schedule()
        sa, s1 = s["sqrt_alphas_cumprod"], s["sqrt_1m_alphas_cumprod"]
        t_i, t_im1 = keep["t_i"], keep["t_im1"]
        c_eps = float(np.sqrt(1 - ZETA) * s1[t_im1])            # eta = 0
        c_zeta = float(np.sqrt(ZETA) * s1[t_im1])
        y, k = keep["y"].to(dev), keep["k"].to(dev)
        FB, _, F2B, FBFy = utils_sisr.pre_calculate(y, k, 1)
        x = keep["x_init"].to(dev).clone()
        x0 = torch.empty_like(x)
        spec = torch.empty(3 * SIZE * (SIZE // 2 + 1) * 2, device=dev, dtype=torch.float32)
        t = torch.full((1,), t_i, device=dev, dtype=torch.int64)
        for i in range(keep["nsteps"]):
            model.x0(x, t, out=x0)
            utils_sisr.data_solution_fused(x0, FB, F2B, FBFy, keep["rho"], 1, 0.5, 0.5, 2.0, -1.0, 1.0, out=x0, spec=spec)
            ops.renoise(x, x0, keep["noises"][i].to(dev), float(sa[t_i]), float(s1[t_i]), float(sa[t_im1]), c_eps, c_zeta, out=x)
        torch.cuda.synchronize()
        xg, x0g = x.cpu().double(), x0.cpu().double()
        xr, x0r = keep["x_final"].double(), keep["x0_last"].double()
        mse = float((((x0g - x0r) / 2) ** 2).mean())               # images in [0,1]
        return {"what": f"{keep['nsteps']}-step trajectory (UNet -> FFT prox -> re-noise at t={t_i}) of one 256x256 image, "
                        f"GPU product path vs the fp32 CPU oracle, identical weights / measurement / noise",
                "rel_l2_x": float((xg - xr).norm() / xr.norm()), "rel_l2_x0": float((x0g - x0r).norm() / x0r.norm()),
                "psnr_db_gpu_vs_oracle_x0": float(10 * np.log10(1.0 / max(mse, 1e-30)))}
    except Exception as exc:        # noqa: BLE001 - reported, not raised
        return {"error": f"{type(exc).__name__}: {exc}"}


# ----------------------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--precision", default="fp16s", choices=["fp16", "fp16s", "fp32"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-steps", type=int, default=3)
    ap.add_argument("--gn-fusion", type=int, default=1, choices=[0, 1, 2, 4],
                    help="ResBlock GroupNorm+SiLU inside the convolutions: 1 (default, activate-once producers), 4 (round-1 per-tap "
                         "producers), 2 (per-tap packed half), 0 separate passes")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
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
    if world > 1:
        import torch.distributed as dist
        # stdout carries exactly one JSON line: NCCL_DEBUG=VERSION (set on the GPU boxes) makes NCCL print its banner there
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)

    from diffpir_b200 import _lib, ops, utils_sisr
    from diffpir_b200.sampler import StepTable
    from diffpir_b200.unet import UNetModel

    B = args.batch
    cfg = arch.ffhq_config()
    model = UNetModel(cfg, precision=args.precision)
    t_bcast_ms = 0.0
    if rank == 0:
        model.load_state_dict(synth.synthetic_state_dict(cfg, 0))
    if world > 1:
        arena = model.weight_arena()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dist.broadcast(arena, src=0)          # the ONE collective: packed weights over NVLink
        e1.record()
        torch.cuda.synchronize()
        t_bcast_ms = e0.elapsed_time(e1)
        model.mark_weights_loaded()
    model.set_gn_fusion(args.gn_fusion)
    if not args.no_graph:
        model.set_graph_mode(True)

    y_h, k_h = make_inputs(B, rank)
    y, k = y_h.to(dev), k_h.to(dev)
    tb = StepTable(sigma=SIGMA, lambda_=LAMBDA, zeta=ZETA, iter_num=NFE)
    steps_tab = tb.steps
    ts = torch.tensor([[st["t_i"]] * B for st in steps_tab], dtype=torch.int64, device=dev)
    FB, _, F2B, FBFy = utils_sisr.pre_calculate(y, k, 1)
    spec = torch.empty(B * 3 * SIZE * (SIZE // 2 + 1) * 2, device=dev, dtype=torch.float32)
    gen = torch.Generator(device=dev).manual_seed(1000 + rank)
    x = torch.randn(B, 3, SIZE, SIZE, device=dev, generator=gen)
    x0 = torch.empty_like(x)
    noise = torch.randn(B, 3, SIZE, SIZE, device=dev, generator=gen)

    def one_step(j, x, noise):
        st = steps_tab[j % len(steps_tab)]
        model.x0(x, ts[j % len(steps_tab)], out=x0)
        if st["last"]:
            # trajectory finished (its x_0 = x/2+0.5 would be read out here); start the next one from fresh noise
            x.copy_(noise)
            return
        utils_sisr.data_solution_fused(x0, FB, F2B, FBFy, st["rho"], 1, 0.5, 0.5, 2.0, -1.0, 1.0, out=x0, spec=spec)
        ops.renoise(x, x0, noise, st["sa_t"], st["s1_t"], st["sa_p"], st["c_eps"], st["c_zeta"], out=x)

    def sync_all():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    # ------------------------------------------------------------------ device-resident timing ("value")
    for j in range(args.warmup):
        one_step(j, x, noise)
    sync_all()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    _lib.launch_count_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for j in range(args.steps):
        one_step(args.warmup + j, x, noise)
    e1.record()
    sync_all()
    launches = _lib.launch_count()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = e0.elapsed_time(e1)
    if dist is not None:
        tt = torch.tensor([ms_total], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total = float(tt.item())
    ms_step = ms_total / args.steps
    value = world * B / (NFE * ms_step / 1e3)

    # ------------------------------------------------------------------ end-to-end timing through host buffers ("e2e")
    xh = torch.randn(B, 3, SIZE, SIZE).pin_memory()
    nh = torch.randn(B, 3, SIZE, SIZE).pin_memory()
    oh = torch.empty(B, 3, SIZE, SIZE).pin_memory()
    xd, nd = torch.empty_like(x), torch.empty_like(x)

    def e2e_step(j):
        xd.copy_(xh, non_blocking=True)           # this step's state and noise from pinned host memory
        nd.copy_(nh, non_blocking=True)
        one_step(j, xd, nd)
        oh.copy_(xd, non_blocking=True)           # the step's result back to the host
    k_e2e = max(10, min(args.steps, 30))
    for j in range(3):
        e2e_step(j)
    sync_all()
    e0.record()
    for j in range(k_e2e):
        e2e_step(3 + j)
    e1.record()
    sync_all()
    ms_e2e = e0.elapsed_time(e1)
    if dist is not None:
        tt = torch.tensor([ms_e2e], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_e2e = float(tt.item())
    ms_e2e /= k_e2e
    e2e_value = world * B / (NFE * ms_e2e / 1e3)
    img_bytes = B * 3 * SIZE * SIZE * 4

    # ------------------------------------------------------------------ roofline of the dominant kernel (conv)
    roofline = None
    if rank == 0:
        peaks = load_peaks()
        model.profile_begin()
        for j in range(args.profile_steps):
            model.x0(x, ts[(7 * j) % len(steps_tab)], out=x0)
        rows = model.profile_end()
        by_kind = {}
        for r in rows:
            by_kind[r["kind"]] = by_kind.get(r["kind"], 0.0) + r["ms"] / args.profile_steps
        conv_ms = by_kind.get("conv", 0.0)
        # algorithmic FLOPs of the tensor-core convs in one forward (everything but the 3->C stem, linears, attention core)
        ted = cfg.time_embed_dim
        conv_flops = 0.0
        for r in rows:
            if r["kind"] == "conv":
                H, W, Cin, Cout, ks, Cin2 = r["dims"]
                conv_flops += 2.0 * B * H * W * (ks * ks * Cin + Cin2) * Cout
        peak = peaks.get("bf16_tflops_sustained", FALLBACK_PEAKS["bf16_tflops_sustained"])
        achieved = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
        unet_ms = sum(by_kind.values())
        # A/B in the same process: the same forward with every GroupNorm as a separate pass (conv kernels do convolution only)
        ab = None
        if args.gn_fusion:
            model.set_gn_fusion(0)
            model.x0(x, ts[0], out=x0)                  # re-plan + warm-up
            model.profile_begin()
            for j in range(args.profile_steps):
                model.x0(x, ts[(7 * j) % len(steps_tab)], out=x0)
            rows_ab = model.profile_end()
            model.set_gn_fusion(args.gn_fusion)
            kind_ab = {}
            for r in rows_ab:
                kind_ab[r["kind"]] = kind_ab.get(r["kind"], 0.0) + r["ms"] / args.profile_steps
            ach_ab = conv_flops / (kind_ab.get("conv", 0.0) * 1e-3) / 1e12 if kind_ab.get("conv") else 0.0
            ab = {"note": "same forward with GroupNorm+SiLU as separate k_gn_apply passes (gn_fusion=0): convolution-only kernels",
                  "achieved": ach_ab, "frac": ach_ab / peak if peak else None, "conv_ms_per_step": kind_ab.get("conv", 0.0),
                  "unet_ms_per_step_serialised": sum(kind_ab.values()),
                  "ms_by_kernel_kind": {k2: round(v, 4) for k2, v in sorted(kind_ab.items())}}
        roofline = {"bound": "tensor", "kernel": "k_conv3x3_v3<FUSE> / k_conv_tc (tcgen05 implicit-GEMM conv, all 3x3/1x1 layers of one forward; "
                                                 "with gn_fusion the ResBlock convs also evaluate GroupNorm+FiLM+SiLU on their operand)",
                    "separate_norm_passes_ab": ab,
                    "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                    "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({peaks['_source']})",
                    "traffic": (conv_traffic() or {}).get("dram_bytes_per_launch"), "traffic_detail": conv_traffic(),
                    "conv_ms_per_step": conv_ms, "unet_ms_per_step_serialised": unet_ms,
                    "share_of_unet": conv_ms / unet_ms if unet_ms else None,
                    "ms_by_kernel_kind": {k2: round(v, 4) for k2, v in sorted(by_kind.items())},
                    "algorithmic_gflop_per_step": conv_flops / 1e9,
                    "whole_step_tflops": world * B * arch.flops_per_image(cfg) / (ms_step * 1e-3) / 1e12}

    # ------------------------------------------------------------------ HBM roofline of the prox / re-noise kernels
    prox = None
    if rank == 0:
        peaks = load_peaks()
        hbm = peaks.get("hbm_gbs", FALLBACK_PEAKS["hbm_gbs"])
        st = steps_tab[10]
        reps = 20

        def timed(fn):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps

        ms_fft = timed(lambda: utils_sisr.data_solution_fused(x0, FB, F2B, FBFy, st["rho"], 1, 0.5, 0.5, 2.0, -1.0, 1.0, out=x0, spec=spec))
        ms_rn = timed(lambda: ops.renoise(x, x0, noise, st["sa_t"], st["s1_t"], st["sa_p"], st["c_eps"], st["c_zeta"], out=x))
        by_fft, by_rn = 2497536 * B, 3145728 * B          # algorithmic bytes per image per step, SURVEY.md section 8(d)
        prox = {"bound": "hbm", "peak": hbm, "unit": "GB/s", "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peaks['_source']})",
                "fft_data_solution": {"ms": ms_fft, "algorithmic_bytes": by_fft, "achieved": by_fft / ms_fft / 1e6,
                                      "frac": by_fft / ms_fft / 1e6 / hbm,
                                      "note": "cuFFT R2C + k_solve_sf1 + cuFFT C2R (3 launches); compulsory traffic only"},
                "renoise": {"ms": ms_rn, "algorithmic_bytes": by_rn, "achieved": by_rn / ms_rn / 1e6, "frac": by_rn / ms_rn / 1e6 / hbm},
                "share_of_step": (ms_fft + ms_rn) / ms_step}

    if rank == 0:
        keep = {}
        cpu = None if (args.no_cpu_baseline or world > 1) else cpu_baseline_sample(keep)
        parity = parity_sample(model, keep, dev) if keep else None
        line = {"metric": "images/sec @100-NFE 256x256 deblur", "value": value, "unit": "images/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32" if args.precision == "fp32" else "f16",
                "data": "synthetic", "config": workload_config(B, args.precision),
                "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": 2 * img_bytes,
                        "d2h_bytes_per_step": img_bytes, "ms_per_step": ms_e2e},
                "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "roofline_prox": prox, "cpu_baseline": cpu, "parity_sample": parity,
                "weight_broadcast_ms": t_bcast_ms, "graph": not args.no_graph, "gn_fusion": args.gn_fusion}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
