"""DiffPIR sampling loop engine: UNet denoise -> closed-form data-fidelity prox -> re-noise, <= 100 NFE.

The reference has no loop function - the loop is inline in its drivers (main_ddpir.py:259-470 batched;
main_ddpir_deblur.py:257-360, main_ddpir_sisr.py:280-387, main_ddpir_inpainting.py:231-303 single image).
This class restates that loop for generate_mode='DiffPIR', sub_1_analytic=True,
model_output_type='pred_xstart', iter_num_U=1 on top of the C-ABI operators:

    per step:  x0 = model.x0(x, t_i)                         1 tape of UNet kernels (eps->x0 fused)
               x0 = prox(x0; y, H, rho_t)                    cuFFT R2C + 1 kernel + cuFFT C2R | 1 kernel (inpaint)
               x  = renoise(x, x0, noise)                    1 kernel

Every per-step scalar (t_i, t_{i-1}, rho_t, schedule coefficients) is batch-shared and is precomputed on
the host once per run with the reference's float32 arithmetic (main_ddpir.py:184-190, 274-286, 327-344,
451-456), so nothing synchronises inside the loop.
"""
import numpy as np
import torch

from . import ops, utils_sisr
from .utils_model import find_nearest
from .utils_resizer import Resizer, ibp_update


def driver_schedule(beta_start=0.0001, beta_end=0.02, T=1000):
    """float32 tables exactly as main_ddpir.py:184-190."""
    betas = torch.from_numpy(np.linspace(beta_start, beta_end, T, dtype=np.float32))
    alphas = 1.0 - betas
    alphas_cumprod = torch.from_numpy(np.cumprod(alphas.numpy(), axis=0))
    sa = torch.sqrt(alphas_cumprod)
    s1 = torch.sqrt(1.0 - alphas_cumprod)
    return dict(betas=betas, alphas=alphas, alphas_cumprod=alphas_cumprod, sqrt_alphas_cumprod=sa,
                sqrt_1m_alphas_cumprod=s1, reduced_alpha_cumprod=torch.div(s1, sa))


def make_seq(T, iter_num, skip_type="quad"):
    """main_ddpir.py:327-335."""
    skip = T // iter_num
    if skip_type == "uniform":
        seq = [i * skip for i in range(iter_num)]
        if skip > 1:
            seq.append(T - 1)
    elif skip_type == "quad":
        seq = np.sqrt(np.linspace(0, T ** 2, iter_num))
        seq = [int(s) for s in list(seq)]
        seq[-1] = seq[-1] - 1
    else:
        raise ValueError(skip_type)
    return seq


class StepTable:
    """All batch-shared scalars of one run, host side."""

    def __init__(self, *, sigma, lambda_, zeta, eta=0.0, iter_num=100, skip_type="quad", T=1000,
                 beta_start=0.0001, beta_end=0.02, t_start=None):
        s = driver_schedule(beta_start, beta_end, T)
        self.sched = s
        self.T = T
        red = s["reduced_alpha_cumprod"]
        red_np = red.numpy()
        sigma = max(0.001, sigma)                                             # main_ddpir.py:141
        sigmas = [red[T - 1 - i] for i in range(T)]                            # :280
        self.t_start = T - 1 if t_start is None else t_start
        self.seq = make_seq(T, iter_num, skip_type)
        sa, s1, betas = s["sqrt_alphas_cumprod"], s["sqrt_1m_alphas_cumprod"], s["betas"]
        self.steps = []
        for i in range(len(self.seq)):
            t_i = find_nearest(red_np, sigmas[self.seq[i]].numpy())           # :342-344
            if t_i > self.t_start:                                            # :346-347
                continue
            last = self.seq[i] == self.seq[-1]
            st = dict(i=i, t_i=t_i, last=last)
            sigma_k = s1[t_i] / sa[t_i]                                       # :282
            st["rho"] = float((lambda_ * (sigma ** 2) / (sigma_k ** 2)).float())   # :286, :389
            if not last:
                t_im1 = find_nearest(red_np, sigmas[self.seq[i + 1]].numpy())      # :451
                eta_sigma = eta * s1[t_im1] / s1[t_i] * torch.sqrt(betas[t_i])     # :454
                c_eps = torch.tensor(np.sqrt(1 - zeta), dtype=torch.float32) * torch.sqrt(s1[t_im1] ** 2 - eta_sigma ** 2)
                c_eta = torch.tensor(np.sqrt(1 - zeta), dtype=torch.float32) * eta_sigma
                c_zeta = torch.tensor(np.sqrt(zeta), dtype=torch.float32) * s1[t_im1]
                u_a = sa[t_i] / sa[t_im1]                                          # :465 sqrt_alpha_effective
                u_b = torch.sqrt(s1[t_i] ** 2 - u_a ** 2 * s1[t_im1] ** 2)         # :466-467
                st.update(t_im1=t_im1, sa_t=float(sa[t_i]), s1_t=float(s1[t_i]), sa_p=float(sa[t_im1]),
                          c_eps=float(c_eps), c_eta=float(c_eta), c_zeta=float(c_zeta), u_a=float(u_a), u_b=float(u_b))
            self.steps.append(st)
        self.sa_start = float(sa[self.t_start])
        self.s1_start = float(s1[self.t_start])

    @property
    def n_renoise(self):
        return sum(0 if st["last"] else 1 for st in self.steps)


class DiffPIRSampler:
    """task: 'deblur' | 'sr' | 'inpaint'.  For 'sr', sr_mode 'blur' uses the FFT solver with the (bicubic) kernel
    k and scale sf; sr_mode 'cubic' uses iterative back-projection through the Resizer (main_ddpir.py:401-406)."""

    def __init__(self, model, task, *, sigma, lambda_, zeta, eta=0.0, guidance_scale=1.0, sf=1, iter_num=100,
                 skip_type="quad", T=1000, sr_mode="blur", gamma=1.0 / 100, in_iter=1, iter_num_U=1):
        if task not in ("deblur", "sr", "inpaint"):
            raise ValueError(task)
        self.model, self.task, self.sf = model, task, (sf if task == "sr" else 1)
        self.guidance_scale, self.sr_mode, self.gamma, self.in_iter = guidance_scale, sr_mode, gamma, in_iter
        self.eta = eta
        self.iter_num_U = int(iter_num_U)        # main_ddpir.py:349 inner repeats ("repeat for semantic consistence: from repaint")
        if self.iter_num_U < 1:
            raise ValueError("iter_num_U must be >= 1")
        self.table = StepTable(sigma=sigma, lambda_=lambda_, zeta=zeta, eta=eta, iter_num=iter_num, skip_type=skip_type, T=T)
        self._t_rows = None

    def _t_vectors(self, B, device):
        key = (B, str(device))
        if self._t_rows is None or self._t_rows[0] != key:
            ts = torch.tensor([[st["t_i"]] * B for st in self.table.steps], dtype=torch.int64, device=device)
            self._t_rows = (key, ts)
        return self._t_rows[1]

    @torch.no_grad()
    def run(self, y, *, k=None, mask=None, x_init=None, init_noise=None, zeta_noise=None, eta_noise=None, u_noise=None,
            generator=None, max_steps=None, trace=None):
        """y: measurement in [0,1] on the GPU ((B,3,H/sf,W/sf) for sr); k: (Bk,1,kh,kw); mask: (B,3,H,W) of 0/1.
        Noise: pass `init_noise` and `zeta_noise` (sequence of per-step tensors) for bit-reproducible parity runs,
        otherwise it is drawn on the device from `generator`.  With iter_num_U > 1 every non-final step is repeated U times
        (main_ddpir.py:349): re-noise to t_{i-1}, then back to t_i with fresh noise `u_noise` (:460-467); `zeta_noise` then
        holds one tensor per re-noise (U per step).  On the final sequence entry the reference's repeats would index past the
        sequence (IndexError at :451), so it is evaluated once, as for U = 1.  Returns x_0 = x/2+0.5 (main_ddpir.py:470)."""
        tb = self.table
        dev = y.device
        y = y.float().contiguous()
        if self.task == "inpaint":
            mask = mask.float().contiguous()
            x = (y * mask) if x_init is None else x_init
        elif x_init is not None:
            x = x_init.float().contiguous()
        elif self.task == "sr":
            # main_ddpir.py:295: the SR loop starts from the bicubic upsample of the measurement, never from y itself
            x = torch.nn.functional.interpolate(y, scale_factor=self.sf, mode="bicubic", align_corners=False).contiguous()
        else:
            x = y
        B, C, H, W = x.shape
        if self.task == "sr" and tuple(x.shape[-2:]) != (y.shape[-2] * self.sf, y.shape[-1] * self.sf):
            raise ValueError(f"sr: x_init must be {self.sf}x the measurement size, got {tuple(x.shape[-2:])} for y {tuple(y.shape[-2:])}")

        def draw():
            return torch.randn(B, C, H, W, device=dev, dtype=torch.float32, generator=generator)

        # x_T = sqrt(abar_T)(2x-1) + sqrt(1-abar_T) n          main_ddpir.py:315
        n0 = init_noise if init_noise is not None else draw()
        xpm = ops.affine(x, 2.0, -1.0)
        x = ops.renoise(xpm, xpm, n0, 0.0, 1.0, tb.sa_start, 0.0, tb.s1_start)
        fft = self.task == "deblur" or (self.task == "sr" and self.sr_mode == "blur")
        if fft:
            FB, _, F2B, FBFy = utils_sisr.pre_calculate(y, k.to(dev), self.sf)
            spec = torch.empty(B * C * H * (W // 2 + 1) * 2 * (2 if self.sf > 1 else 1), device=dev, dtype=torch.float32)
        elif self.task == "sr":
            degrade = Resizer((B, C, H, W), 1 / self.sf).to(dev)
        ts = self._t_vectors(B, dev)
        x0 = torch.empty_like(x)
        steps = tb.steps if max_steps is None else tb.steps[:max_steps]
        j = ju = 0
        U = self.iter_num_U
        for si, st in enumerate(steps):
            for u in range(1 if st["last"] else U):
                self.model.x0(x, ts[si], out=x0)
                if trace is not None:
                    trace.append(("x0_unet", st["i"], x0.clone()))
                if st["last"]:
                    break
                rho = st["rho"]
                if self.task == "inpaint":
                    ops.inpaint_prox(x0, y, mask, rho, self.guidance_scale, out=x0)
                elif fft:
                    utils_sisr.data_solution_fused(x0, FB, F2B, FBFy, rho, self.sf, 0.5, 0.5, 2.0, -1.0, self.guidance_scale,
                                                   out=x0, spec=spec)
                else:
                    for _ in range(self.in_iter):
                        down = degrade(ops.affine(x0, 0.5, 0.5))
                        ibp_update(x0, y, down, self.gamma, rho, self.sf, out=x0)
                if trace is not None:
                    trace.append(("x0_prox", st["i"], x0.clone()))
                nz = zeta_noise[j] if zeta_noise is not None else draw()
                ne = None
                if st["c_eta"] != 0.0:
                    ne = eta_noise[j] if eta_noise is not None else draw()
                ops.renoise(x, x0, nz, st["sa_t"], st["s1_t"], st["sa_p"], st["c_eps"], st["c_zeta"], noise_eta=ne,
                            c_eta=st["c_eta"], out=x)
                j += 1
                if u < U - 1:
                    # back from x_{t-1} to x_t (:460-467): x = a x + sqrt(s1_t^2 - a^2 s1_p^2) n, a = sa_t / sa_p - the same
                    # fused kernel with (x0 := x, sa_t = 0, s1_t = 1): sa_p' x + 0 + c_zeta' n
                    nu = u_noise[ju] if u_noise is not None else draw()
                    ops.renoise(x, x, nu, 0.0, 1.0, st["u_a"], 0.0, st["u_b"], out=x)
                    ju += 1
                if trace is not None:
                    trace.append(("x", st["i"], x.clone()))
            if st["last"]:
                break
        return ops.affine(x, 0.5, 0.5)
