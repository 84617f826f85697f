"""Interpose the B200 hot path under the UNMODIFIED reference drivers.

    import diffpir_b200.install as inst
    inst.install(precision="fp16")             # before the driver module is executed
    runpy.run_path("main_ddpir.py", run_name="__main__")      # with the reference root on sys.path / as cwd

What gets rebound (attributes are resolved at call time because the drivers bind the modules, main_ddpir.py:12-15):
    utils.utils_sisr.pre_calculate / data_solution            -> diffpir_b200.utils_sisr
    utils.utils_model.model_fn / test_mode / find_nearest     -> diffpir_b200.utils_model
    utils.utils_resizer.Resizer                                -> diffpir_b200.utils_resizer.Resizer
    utils.utils_inpaint.mask_generator                         -> diffpir_b200.utils_inpaint.mask_generator (same numpy draws)
    guided_diffusion.script_util.create_model_and_diffusion   -> returns (diffpir_b200.UNetModel, reference SpacedDiffusion)
`create_model_and_diffusion` is imported by name in the drivers (main_ddpir.py:31-36), so install() must run before
the driver module is imported.  uninstall() restores everything.
"""
import importlib

_saved = {}


def _patch(mod, name, new):
    key = (mod.__name__, name)
    if key not in _saved:
        _saved[key] = (mod, getattr(mod, name))
    setattr(mod, name, new)


def install(precision="fp16", graph=True):
    from . import utils_model as um, utils_resizer as ur, utils_sisr as us
    from .arch import UNetConfig
    from .unet import UNetModel
    ref_sisr = importlib.import_module("utils.utils_sisr")
    ref_model = importlib.import_module("utils.utils_model")
    ref_resizer = importlib.import_module("utils.utils_resizer")
    ref_script = importlib.import_module("guided_diffusion.script_util")
    _patch(ref_sisr, "pre_calculate", us.pre_calculate)
    _patch(ref_sisr, "data_solution", us.data_solution)
    _patch(ref_model, "model_fn", um.model_fn)
    _patch(ref_model, "test_mode", um.test_mode)
    _patch(ref_model, "find_nearest", um.find_nearest)
    _patch(ref_resizer, "Resizer", ur.Resizer)
    try:        # data edges (host logic with the reference's numpy draws; utils_inpaint pulls in matplotlib through utils_image)
        from . import utils_inpaint as ui
        ref_inpaint = importlib.import_module("utils.utils_inpaint")
        _patch(ref_inpaint, "mask_generator", ui.mask_generator)
    except ImportError:
        pass
    ref_create = _saved.get((ref_script.__name__, "create_model_and_diffusion"), (None, ref_script.create_model_and_diffusion))[1]

    def create_model_and_diffusion(**kw):
        """script_util.py:74-127 with the UNet replaced; the diffusion object is the reference's own."""
        gd = ref_script.create_gaussian_diffusion(
            steps=kw["diffusion_steps"], learn_sigma=kw["learn_sigma"], noise_schedule=kw["noise_schedule"],
            use_kl=kw["use_kl"], predict_xstart=kw["predict_xstart"], rescale_timesteps=kw["rescale_timesteps"],
            rescale_learned_sigmas=kw["rescale_learned_sigmas"], timestep_respacing=kw["timestep_respacing"])
        if kw.get("class_cond") or kw.get("use_fp16"):
            raise NotImplementedError("class-conditional / use_fp16 models are outside the DiffPIR hot path")
        cfg = UNetConfig(image_size=kw["image_size"], model_channels=kw["num_channels"], num_res_blocks=kw["num_res_blocks"],
                         attention_resolutions=kw["attention_resolutions"], channel_mult=kw.get("channel_mult", "") or (),
                         out_channels=(6 if kw["learn_sigma"] else 3), num_heads=kw["num_heads"],
                         num_head_channels=kw["num_head_channels"], num_heads_upsample=kw["num_heads_upsample"],
                         use_scale_shift_norm=kw["use_scale_shift_norm"], resblock_updown=kw["resblock_updown"],
                         use_new_attention_order=kw["use_new_attention_order"])
        return UNetModel(cfg, precision=precision, diffusion_steps=kw["diffusion_steps"], graph=graph), gd

    create_model_and_diffusion.reference = ref_create
    _patch(ref_script, "create_model_and_diffusion", create_model_and_diffusion)
    return True


def uninstall():
    for (_, name), (mod, old) in list(_saved.items()):
        setattr(mod, name, old)
    _saved.clear()
