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
