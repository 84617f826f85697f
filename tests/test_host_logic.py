"""Host-side logic (no GPU): architecture replay, schedule tables, resizer tables, measurement synthesis."""
import numpy as np
import torch

from diffpir_b200 import arch, synth
from diffpir_b200.sampler import StepTable, make_seq
from oracle import diffpir_oracle as orc
from oracle import unet_oracle


def test_flops_match_survey():
    assert abs(arch.flops_per_image(arch.ffhq_config()) / 1e9 - 387.93) < 0.01
    assert abs(arch.flops_per_image(arch.imagenet_config()) / 1e9 - 2239.67) < 0.01


def test_param_specs_counts():
    f = arch.param_specs(arch.ffhq_config())
    assert len(f) == 362 and sum(int(np.prod(s)) for _, s in f) == 93563910
    i = arch.param_specs(arch.imagenet_config())
    assert len(i) == 566 and sum(int(np.prod(s)) for _, s in i) == 552814086


def test_layer_program_agrees_with_oracle_replay():
    for cfg in (arch.ffhq_config(), arch.imagenet_config()):
        a = arch.build_layers(cfg)
        b = unet_oracle.block_program(cfg.as_dict())
        for blocks_a, blocks_b in zip(a, (b[0], [b[1]], b[2])):
            assert len(blocks_a) == len(blocks_b)
            for la, lb in zip(blocks_a, blocks_b):
                assert [(l.kind, l.cin, l.cout) for l in la] == [(x[0], x[1], x[2] if x[0] != "attn" else x[1]) for x in lb]
                for l, x in zip(la, lb):
                    if l.kind == "res":
                        assert (l.updown or None) == x[3]
                    if l.kind == "attn":
                        assert l.heads == x[2]


def test_step_table_matches_oracle_schedule():
    for iters, skip in ((100, "quad"), (20, "quad"), (20, "uniform")):
        tb = StepTable(sigma=0.05, lambda_=7.0, zeta=0.3, iter_num=iters, skip_type=skip)
        sched = orc.driver_schedule()
        rhos, sigmas = orc.rhos_sigmas(sched, 7.0, 0.05)
        seq = orc.make_seq(1000, iters, skip)
        assert seq == make_seq(1000, iters, skip) == tb.seq
        red = sched["reduced_alpha_cumprod"].numpy()
        assert len(tb.steps) == len(seq)
        for st in tb.steps:
            i = st["i"]
            t_i = orc.find_nearest(red, sigmas[seq[i]].numpy())
            assert st["t_i"] == t_i == 999 - seq[i]
            assert np.float32(st["rho"]) == np.float32(rhos[t_i])
            if not st["last"]:
                assert st["t_im1"] == orc.find_nearest(red, sigmas[seq[i + 1]].numpy())
        assert tb.steps[-1]["last"] and sum(s["last"] for s in tb.steps) == 1
    tb = StepTable(sigma=0.0, lambda_=1.0, zeta=1.0, iter_num=100)
    assert tb.steps[0]["t_i"] == 999 and tb.n_renoise == 99


def test_resizer_tables_match_oracle():
    from diffpir_b200.utils_resizer import Resizer
    for n, scale in ((256, 0.25), (64, 0.5), (32, 2.0), (96, 1 / 3)):
        r = Resizer((2, 3, n, n), scale)
        w, f = orc.resizer_contributions(n, int(np.ceil(n * scale)), scale, antialiasing=scale < 1)
        for d in (0, 1):
            assert torch.equal(r._tabs[d][0], torch.tensor(w.T, dtype=torch.float32))
            assert np.array_equal(r._tabs[d][1].numpy(), f.T.astype(np.int32))   # index tables: bit-exact


def test_measurement_synthesis_truncates_like_scipy_uint8():
    img = synth.synthetic_images(1, 64, seed=5)[0]
    k = synth.gaussian_kernel(9, 1.5)
    y = synth.blur_measurement(img, k, 0.0, seed=0)
    q = y * 255.0
    assert np.allclose(q, np.round(q), atol=1e-3)      # uint8-valued: the reference blurs in uint8 (main_ddpir.py:99)


def test_masks():
    m = synth.random_mask(2, 64, 0.5, seed=1)
    assert m.shape == (2, 3, 64, 64) and set(np.unique(m.numpy())) == {0.0, 1.0}
    assert int((m[0, 0] == 0).sum()) == 64 * 64 // 2 and torch.equal(m[:, 0], m[:, 2])
    b = synth.box_mask(2, 256, 128, seed=2)
    assert int((b[0, 0] == 0).sum()) == 128 * 128


def test_test_mode_wrappers_match_reference_golden():
    """utils_model.test_mode modes 0-5 (pad / split / x8 / split+x8 / split once) against outputs of the reference's own
    test_mode driven by the same deterministic stand-in model_fn (oracle/gen_golden.py::gen_test_mode)."""
    import os
    from diffpir_b200 import utils_model
    from oracle import gen_golden
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "test_mode.npz"))
    x = gen_golden.test_mode_input()
    for i, kw in enumerate(gen_golden.TEST_MODE_CASES):
        out = utils_model.test_mode(gen_golden.fake_model_fn, None, x, noise_level=12.0, **kw)
        ref = torch.from_numpy(g[f"case{i}"])
        assert out.shape == ref.shape, (i, out.shape, ref.shape)
        assert float((out - ref).abs().max()) < 1e-6, i


def test_find_nearest_and_model_fn_argument_errors():
    from diffpir_b200 import utils_model
    arr = torch.tensor([0.1, 0.5, 0.9])
    assert utils_model.find_nearest(arr, 0.45) == 1 and utils_model.find_nearest(arr.numpy(), 2.0) == 2
    import pytest
    with pytest.raises(RuntimeError):
        utils_model.model_fn(torch.zeros(1, 3, 8, 8), 10.0, model_diffusion=object(), model_out_type="pred_x_prev",
                             diffusion=None, alphas_cumprod=torch.linspace(0.99, 0.01, 1000))


def test_mask_generator_matches_reference_goldens():
    """diffpir_b200.utils_inpaint.mask_generator draws from the global numpy stream in the reference's order: bit-exact
    against masks produced by the unmodified utils.utils_inpaint.mask_generator (tests/golden/misc.npz, oracle/gen_golden.py)."""
    import os
    from diffpir_b200.utils_inpaint import mask_generator
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "misc.npz"))
    img = torch.zeros(1, 3, 64, 64)
    np.random.seed(123)
    m = mask_generator("random", mask_prob_range=(0.5, 0.5), image_size=64)(img)
    assert m.shape == img.shape and m.dtype == torch.float32
    assert np.array_equal(m.numpy().astype(np.uint8), g["mask_random"])
    assert int((m[0, 0] == 0).sum()) == 64 * 64 // 2
    np.random.seed(123)
    b = mask_generator("box", mask_len_range=(16, 17), image_size=64, margin=(4, 4))(img)
    assert np.array_equal(b.numpy().astype(np.uint8), g["mask_box"])
    np.random.seed(123)
    e = mask_generator("extreme", mask_len_range=(16, 17), image_size=64, margin=(4, 4))(img)
    assert torch.equal(e, 1 - b)
    assert set(torch.unique(m).tolist()) <= {0.0, 1.0}


def test_imresize_tables_reproduce_reference_goldens():
    """The gather tables diffpir_b200.utils_image.imresize_tables hands to the resize kernel, applied here with numpy, give
    utils_image.imresize_np of the unmodified reference (tests/golden/imresize.npz) - including its symmetric padding."""
    import os
    from diffpir_b200.utils_image import imresize_tables
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "imresize.npz"))

    def apply(x, scale, aa=True):
        x = x if x.ndim == 3 else x[:, :, None]
        H, W, _ = x.shape
        wh, fh = imresize_tables(H, int(np.ceil(H * scale)), scale, aa)
        ww, fw = imresize_tables(W, int(np.ceil(W * scale)), scale, aa)
        y = np.einsum("to,tohc->ohc", wh, x[fh].astype(np.float32))                 # H pass first, as the reference
        return np.einsum("to,htoc->hoc", ww, y[:, fw].astype(np.float32))
    for key, scale, aa in (("a_down2", 1 / 2, True), ("a_down3", 1 / 3, True), ("a_down4", 1 / 4, True), ("a_up2", 2, True),
                           ("b_down2", 1 / 2, True), ("a_up3_noaa", 3, False)):
        src = g[key[0]]
        out = apply(src, scale, aa)
        ref = g[key] if g[key].ndim == 3 else g[key][:, :, None]
        assert out.shape == ref.shape, key
        assert np.abs(out - ref).max() < 2e-6, (key, np.abs(out - ref).max())
        assert np.abs(orc.imresize_np(src, scale, aa).reshape(ref.shape) - ref).max() < 1e-6      # the oracle is pinned too
    w, f = imresize_tables(256, 64, 0.25)
    assert w.dtype == np.float32 and f.dtype == np.int32 and f.min() >= 0 and f.max() < 256
    assert np.allclose(w.sum(0), 1, atol=1e-6)


def test_kernel_design_checks():
    """The numpy replays of the kernels' index arithmetic (no GPU): operand-producer thread mappings of k_conv3x3_v3 and the tiling /
    horizontal-taps-in-N formulation of the narrow-output kernel k_conv3x3_dxn against a direct convolution."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for name in ("check_dxn_formulation", "check_producer_mapping"):
        spec = importlib.util.spec_from_file_location(name, os.path.join(root, "tools", name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.main()
