"""Live re-check of the golden fixtures against the reference (only where /root/reference exists)."""
import pytest


@pytest.mark.reference
def test_goldens_reproduce_from_reference():
    from oracle import gen_golden
    assert gen_golden.main(check=True)


@pytest.mark.reference
def test_install_rebinds_and_restores_reference_symbols():
    """install() must swap exactly the operator surface of SURVEY section 8(b) and uninstall() must restore it (no compute is
    executed here: without a GPU the patched factory has to fail loudly instead of falling back)."""
    import torch
    from oracle import ref_shim
    ref_shim.install()
    import importlib
    sisr = importlib.import_module("utils.utils_sisr")
    um = importlib.import_module("utils.utils_model")
    rz = importlib.import_module("utils.utils_resizer")
    su = importlib.import_module("guided_diffusion.script_util")
    orig = (sisr.pre_calculate, sisr.data_solution, um.model_fn, um.test_mode, rz.Resizer, su.create_model_and_diffusion)
    from diffpir_b200 import install as inst
    from diffpir_b200 import utils_model as dum, utils_sisr as dus
    inst.install(precision="fp16")
    try:
        assert sisr.pre_calculate is dus.pre_calculate and sisr.data_solution is dus.data_solution
        assert um.model_fn is dum.model_fn and um.test_mode is dum.test_mode
        assert rz.Resizer is not orig[4] and su.create_model_and_diffusion is not orig[5]
        if not torch.cuda.is_available():
            args = um.create_argparser(dict(model_path="")).parse_args([])
            with pytest.raises(RuntimeError):
                su.create_model_and_diffusion(**su.args_to_dict(args, su.model_and_diffusion_defaults().keys()))
    finally:
        inst.uninstall()
    now = (sisr.pre_calculate, sisr.data_solution, um.model_fn, um.test_mode, rz.Resizer, su.create_model_and_diffusion)
    assert all(a is b for a, b in zip(orig, now))


@pytest.mark.reference
def test_native_diffusion_mirrors_the_reference_tables():
    """diffpir_b200.diffusion.GaussianDiffusion.from_reference: the float tables dpir_sample_step indexes (p_sample and
    ddim_sample coefficients, gaussian_diffusion.py:133-169, 569-580) equal the reference object's own arrays cast to float, for
    the plain 1000-step schedule DiffPIR builds and for a respaced one; the timestep map is mirrored."""
    import numpy as np
    import torch
    from oracle import ref_shim
    ref_shim.install()
    from guided_diffusion import script_util as su
    from diffpir_b200.diffusion import GaussianDiffusion
    for respacing in ("", "50"):
        gd = su.create_gaussian_diffusion(steps=1000, learn_sigma=True, noise_schedule="linear", use_kl=False, predict_xstart=False,
                                          rescale_timesteps=False, rescale_learned_sigmas=False, timestep_respacing=respacing)
        nat = GaussianDiffusion.from_reference(gd)
        T = gd.num_timesteps
        assert nat.num_timesteps == T
        tab = nat._table("cpu", eta=0.0)
        f32 = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float64)).float()        # noqa: E731
        cols = [gd.sqrt_recip_alphas_cumprod, gd.sqrt_recipm1_alphas_cumprod, gd.posterior_mean_coef1, gd.posterior_mean_coef2,
                gd.posterior_log_variance_clipped, np.log(gd.betas)]
        for j, c in enumerate(cols):
            assert torch.equal(tab[:, j], f32(c)), j
        abp = f32(gd.alphas_cumprod_prev)
        assert torch.equal(tab[:, 6], torch.sqrt(abp)) and torch.equal(tab[:, 7], torch.sqrt(1 - abp)) and float(tab[:, 8].abs().max()) == 0
        t = torch.arange(T)
        want = torch.tensor(gd.timestep_map)[t] if respacing else t
        assert torch.equal(nat.model_timesteps(t), want)
