"""GPU parity of the data-fidelity / re-noise / resize kernels against the CPU oracle (through the C-ABI)."""
import numpy as np
import pytest
import torch

from oracle import diffpir_oracle as orc
from dpir_testutil import rel_l2

pytestmark = pytest.mark.gpu


def _kernel(kh, kw, seed=0):
    g = torch.Generator().manual_seed(seed)
    k = torch.rand(1, 1, kh, kw, generator=g)
    return k / k.sum()


@pytest.mark.parametrize("sf,B,Bk,ksz", [(1, 2, 1, 25), (1, 3, 3, 13), (2, 2, 1, 25), (4, 2, 2, 25), (4, 1, 1, 61)])
@pytest.mark.parametrize("alpha", [3e-6, 1e-3, 0.3, 150.0])
def test_fft_solver_matches_fp64_oracle(sf, B, Bk, ksz, alpha):
    from diffpir_b200 import utils_sisr as sr
    g = torch.Generator().manual_seed(1)
    H = W = 64 if ksz <= 25 else 128
    y = torch.rand(B, 3, H // sf, W // sf, generator=g)
    x = torch.rand(B, 3, H, W, generator=g)
    k = torch.cat([_kernel(ksz, ksz, s) for s in range(Bk)], 0)
    # oracle: reference-literal formula in fp64 (the fp32 reference itself loses digits at small alpha)
    FB, FBC, F2B, FBFy = orc.pre_calculate(y.double(), k.double().expand(B, 1, ksz, ksz) if Bk == 1 else k.double(), sf)
    ref64 = orc.data_solution(x.double(), FB, FBC, F2B, FBFy, torch.tensor(alpha, dtype=torch.float64).view(1, 1, 1, 1), sf)
    FB32, FBC32, F2B32, FBFy32 = orc.pre_calculate(y, k.expand(B, 1, ksz, ksz) if Bk == 1 else k, sf)
    ref32 = orc.data_solution(x, FB32, FBC32, F2B32, FBFy32, torch.tensor(alpha).view(1, 1, 1, 1), sf)
    t = sr.pre_calculate(y.cuda(), k.cuda(), sf)
    out = sr.data_solution(x.cuda(), *t, torch.tensor(alpha).view(1, 1, 1, 1).cuda(), sf)
    err = rel_l2(out, ref64)
    tol = max(1e-4, 1.5 * rel_l2(ref32, ref64))
    assert err <= tol, (err, tol)
    # the cancellation-free form should be at least as close to fp64 as the reference's own fp32 evaluation
    assert err <= max(2e-5, 1.5 * rel_l2(ref32, ref64))


def test_precalc_tensors_match_oracle():
    from diffpir_b200 import utils_sisr as sr
    g = torch.Generator().manual_seed(2)
    y = torch.rand(2, 3, 32, 32, generator=g)
    k = _kernel(25, 25)
    FB, FBC, F2B, FBFy = orc.pre_calculate(y, k.expand(2, 1, 25, 25), 2)
    fb, fbc, f2b, fbfy = sr.pre_calculate(y.cuda(), k.cuda(), 2)
    Wh = 64 // 2 + 1
    assert rel_l2(torch.view_as_real(fb.cpu()), torch.view_as_real(FB[:1, :, :, :Wh].contiguous())) < 1e-5
    assert rel_l2(f2b, F2B[:1, :, :, :Wh]) < 1e-5
    assert rel_l2(torch.view_as_real(fbfy.cpu()), torch.view_as_real(FBFy[..., :Wh].contiguous())) < 1e-5
    assert rel_l2(torch.view_as_real(fbc.cpu().resolve_conj()), torch.view_as_real(FBC[:1, :, :, :Wh].contiguous())) < 1e-5


def test_fused_solver_affine_and_blend():
    from diffpir_b200 import utils_sisr as sr
    g = torch.Generator().manual_seed(3)
    y = torch.rand(2, 3, 64, 64, generator=g)
    x0 = torch.rand(2, 3, 64, 64, generator=g) * 2 - 1
    k = _kernel(13, 13)
    rho, gs = 0.02, 0.7
    t = orc.pre_calculate(y.double(), k.double().expand(2, 1, 13, 13), 1)
    x0p = orc.data_solution(x0.double() / 2 + 0.5, *t, torch.tensor(rho, dtype=torch.float64).view(1, 1, 1, 1), 1) * 2 - 1
    ref = x0.double() + gs * (x0p - x0.double())
    FB, _, F2B, FBFy = sr.pre_calculate(y.cuda(), k.cuda(), 1)
    out = sr.data_solution_fused(x0.cuda(), FB, F2B, FBFy, rho, 1, 0.5, 0.5, 2.0, -1.0, gs)
    assert rel_l2(out, ref) < 2e-5


def test_inpaint_prox_bit_exact_on_known_pixels_and_close_elsewhere():
    from diffpir_b200 import ops, synth
    g = torch.Generator().manual_seed(4)
    x0 = torch.rand(2, 3, 64, 64, generator=g) * 2 - 1
    y = torch.rand(2, 3, 64, 64, generator=g)
    mask = synth.random_mask(2, 64, 0.5, seed=1)
    for tau in (7e-7, 0.01, 3.0):
        ref = orc.inpaint_prox(x0, 2 * y - 1, mask, torch.tensor(tau).view(1, 1, 1, 1), 1.0)
        out = ops.inpaint_prox(x0.cuda(), y.cuda(), mask.cuda(), tau).cpu()
        # masked-out pixels (mask == 0) must return x0 exactly: (0*y + tau*x0)/(0+tau)
        hole = mask == 0
        assert torch.equal(out[hole], ref[hole])
        assert rel_l2(out, ref) < 1e-6
    # reference single-image driver passes y already in [-1,1] (main_ddpir_inpainting.py:267)
    ref = orc.inpaint_prox(x0, y, mask, torch.tensor(0.5).view(1, 1, 1, 1), 0.8)
    out = ops.inpaint_prox(x0.cuda(), y.cuda(), mask.cuda(), 0.5, guidance_scale=0.8, y_scale=1.0, y_bias=0.0)
    assert rel_l2(out, ref) < 1e-6


@pytest.mark.parametrize("eta,zeta", [(0.0, 0.3), (0.0, 1.0), (0.5, 0.25)])
def test_renoise_matches_oracle(eta, zeta):
    from diffpir_b200 import ops
    from diffpir_b200.sampler import StepTable
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 3, 64, 64, generator=g)
    x0 = torch.rand(2, 3, 64, 64, generator=g) * 2 - 1
    n1 = torch.randn(2, 3, 64, 64, generator=g)
    n2 = torch.randn(2, 3, 64, 64, generator=g)
    tb = StepTable(sigma=0.05, lambda_=7.0, zeta=zeta, eta=eta, iter_num=20)
    sched = orc.driver_schedule()
    for st in (tb.steps[0], tb.steps[7], tb.steps[-2]):
        ref = orc.renoise(x, x0, st["t_i"], st["t_im1"], sched, eta, zeta, n1, n2)
        out = ops.renoise(x.cuda(), x0.cuda(), n2.cuda(), st["sa_t"], st["s1_t"], st["sa_p"], st["c_eps"], st["c_zeta"],
                          noise_eta=n1.cuda(), c_eta=st["c_eta"])
        assert rel_l2(out, ref) < 1e-6


def test_eps_to_x0_matches_oracle():
    from diffpir_b200 import ops
    g = torch.Generator().manual_seed(6)
    mo = torch.randn(3, 6, 32, 32, generator=g)
    x = torch.randn(3, 3, 32, 32, generator=g)
    t = torch.tensor([999, 500, 3])
    tabs = orc.diffusion_tables()
    ref = orc.pred_xstart(mo, x, t, tabs)
    coef = torch.stack([torch.from_numpy(tabs["sqrt_recip_alphas_cumprod"])[t].float(),
                        torch.from_numpy(tabs["sqrt_recipm1_alphas_cumprod"])[t].float()], 1)
    out = ops.eps_to_x0(mo.cuda(), x.cuda(), coef.cuda())
    assert rel_l2(out, ref) < 1e-6
    assert float(out.max()) <= 1.0 and float(out.min()) >= -1.0


@pytest.mark.parametrize("scale,size", [(0.25, 256), (0.5, 64), (2.0, 32), (1 / 3, 96)])
def test_resizer_matches_oracle(scale, size):
    from diffpir_b200.utils_resizer import Resizer
    g = torch.Generator().manual_seed(7)
    x = torch.rand(2, 3, size, size, generator=g)
    ref = orc.resizer_forward(x, scale)
    out = Resizer(x.shape, scale).to("cuda")(x.cuda())
    assert tuple(out.shape) == tuple(ref.shape)
    assert rel_l2(out, ref) < 1e-6


def test_ibp_update_matches_oracle():
    from diffpir_b200 import ops
    from diffpir_b200.utils_resizer import Resizer, ibp_update
    g = torch.Generator().manual_seed(8)
    x0 = torch.rand(2, 3, 64, 64, generator=g) * 2 - 1
    y = torch.rand(2, 3, 16, 16, generator=g)
    ref = orc.ibp_step(x0, y, 4, 0.01, torch.tensor(0.3), 1)
    down = Resizer(x0.shape, 0.25).to("cuda")(ops.affine(x0.cuda(), 0.5, 0.5))
    out = ibp_update(x0.cuda(), y.cuda(), down, 0.01, 0.3, 4)
    assert rel_l2(out, ref) < 1e-6


def test_blur_operator_forward_transpose_kernel():
    import torch.nn.functional as F
    from diffpir_b200.utils_deblur import GaussialBlurOperator
    op = GaussialBlurOperator(kernel_size=21, intensity=3.0, device="cuda")
    k = op.get_kernel()
    assert tuple(k.shape) == (1, 1, 21, 21) and abs(float(k.sum()) - 1) < 1e-3
    g = torch.Generator().manual_seed(9)
    x = torch.rand(2, 3, 48, 40, generator=g)
    ref = F.conv2d(F.pad(x, (10, 10, 10, 10), mode="reflect"), k.float().expand(3, 1, 21, 21), groups=3)
    out = op.forward(x.cuda())
    assert rel_l2(out, ref) < 1e-5
    assert op.transpose(x) is x


def test_tensor2uint_and_psnr_tail_match_reference_formulas():
    """tensor2uint_batch / calculate_psnr_batch / Y-PSNR (utils_image.py:238-242, 601-610, 470-490; main_ddpir.py:482-515)."""
    import numpy as np
    from diffpir_b200 import utils_image as ui
    g = torch.Generator().manual_seed(10)
    x0 = torch.rand(3, 3, 64, 48, generator=g) * 1.2 - 0.1
    ref_u8 = np.uint8((np.transpose(x0.clamp(0, 1).numpy(), (0, 2, 3, 1)) * 255.0).round())
    assert np.array_equal(ui.tensor2uint_batch(x0.cuda()), ref_u8)             # byte output: bit-exact
    a, b = x0 * 2 - 1, torch.rand(3, 3, 64, 48, generator=g) * 2 - 1
    mse = torch.mean((a - b) ** 2, axis=(1, 2, 3))
    ref = torch.mean(20 * torch.log10(2.0 / torch.sqrt(mse + 1e-10))).item()
    assert abs(ui.calculate_psnr_batch(a.cuda(), b.cuda()) - ref) < 1e-4
    ya = torch.zeros_like(a); yb = torch.zeros_like(b)
    ya[:, 0] = 0.299 * a[:, 0] + 0.587 * a[:, 1] + 0.114 * a[:, 2]
    yb[:, 0] = 0.299 * b[:, 0] + 0.587 * b[:, 1] + 0.114 * b[:, 2]
    mse = torch.mean((ya - yb) ** 2, axis=(1, 2, 3))
    ref = torch.mean(20 * torch.log10(2.0 / torch.sqrt(mse + 1e-10))).item()
    assert abs(ui.calculate_psnr_y_batch(a.cuda(), b.cuda()) - ref) < 1e-4
    assert ui.calculate_psnr_batch(a.cuda(), a.cuda()) == float("inf")


@pytest.mark.parametrize("ksz", [13, 27, 61, 24])
def test_measurement_blur_uint8_bit_exact_vs_scipy(ksz):
    """The uint8 circular blur of CustomDataset.__getitem__ (main_ddpir.py:99): bit-exact against scipy.ndimage.convolve."""
    import numpy as np
    from scipy import ndimage
    from diffpir_b200 import synth, utils_image as ui
    imgs = synth.synthetic_images(2, 96, seed=31)
    ks = [synth.gaussian_kernel(ksz, 2.5, dtype=np.float64) if ksz != 27 else synth.motion_kernel(27, 0.6, 3).astype(np.float64),
          synth.motion_kernel(ksz, 0.5, 4).astype(np.float64)]
    rng = np.random.RandomState(0)
    noise = rng.normal(0, 0.1, imgs.shape).astype(np.float32)
    y, u8 = ui.blur_measurement_batch(imgs, np.stack(ks), noise=noise, return_u8=True)
    for b in range(2):
        ref = ndimage.convolve(imgs[b], np.expand_dims(ks[b], axis=2), mode="wrap")
        assert np.array_equal(u8[b].cpu().numpy(), ref)
        yl = (np.float32(ref / 255.0) * 2 - 1 + noise[b]) / 2 + 0.5
        assert np.abs(y[b].cpu().numpy() - np.transpose(yl, (2, 0, 1))).max() < 1e-6


def test_motion_blur_operator_forward_and_kernel():
    """MotionBlurOperator (utils/utils_deblur.py:683-705 surface; BASELINE config 5 depends on it): get_kernel() is a normalised
    non-negative (1,1,k,k) PSF that is reproducible per seed, forward = ReflectionPad2d(k//2) + depthwise correlation with it
    (checked against torch in fp64), transpose returns its input like the reference's."""
    import torch.nn.functional as F
    from diffpir_b200.utils_deblur import MotionBlurOperator
    op = MotionBlurOperator(kernel_size=61, intensity=0.5, device="cuda", seed=1234)
    k = op.get_kernel()
    assert tuple(k.shape) == (1, 1, 61, 61) and float(k.min()) >= 0 and abs(float(k.sum()) - 1) < 1e-5
    assert torch.equal(k, MotionBlurOperator(61, 0.5, "cuda", seed=1234).get_kernel())
    assert not torch.equal(k, MotionBlurOperator(61, 0.5, "cuda", seed=1235).get_kernel())
    x = torch.rand(2, 3, 96, 80, generator=torch.Generator().manual_seed(1))
    out = op.forward(x.cuda())
    ref = F.conv2d(F.pad(x.double(), (30, 30, 30, 30), mode="reflect"), k.double().repeat(3, 1, 1, 1), groups=3)
    assert rel_l2(out, ref) < 1e-6
    assert op.transpose(out) is out


def test_weight_arena_clone_computes_identical_results():
    """A rank that RECEIVES the packed weight arena (the one NCCL broadcast, parallel.broadcast_model) never calls
    load_state_dict: emulate it on one GPU by copying the arena into a second engine - outputs must be bit-identical (fp16s
    throughput mode and the fp16x3 split mode, whose accumulator scales also live in the arena)."""
    from dpir_testutil import make_model, small_config
    from diffpir_b200.unet import UNetModel
    cfg = small_config()
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 3, 64, 64, generator=g).cuda()
    t = torch.tensor([999, 123], device="cuda")
    for precision in ("fp16s", "fp16x3"):
        a = make_model(cfg, precision)
        b = UNetModel(cfg, precision=precision)
        assert b.missing_weights() > 0
        b.weight_arena().copy_(a.weight_arena())
        b.mark_weights_loaded()
        assert b.missing_weights() == 0
        assert torch.equal(a.x0(x, t), b.x0(x, t))
        del a, b
