"""Import the UNMODIFIED reference from /root/reference (build container only; test infra).

Two in-process shims are needed: `matplotlib` (utils/utils_image.py:11) and
`motionblur` (utils/utils_deblur.py:7) are not installed.  Nothing here is used on the GPU box, where
/root/reference does not exist - `available()` gates every caller.
"""
import os
import sys
import types

REF = "/root/reference"


def available():
    return os.path.isdir(os.path.join(REF, "guided_diffusion"))


def install():
    if not available():
        raise RuntimeError("reference not present")
    if REF not in sys.path:
        sys.path.insert(0, REF)
    for m in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(m, types.ModuleType(m))
    if "motionblur" not in sys.modules:
        mb = types.ModuleType("motionblur")
        mb2 = types.ModuleType("motionblur.motionblur")
        mb2.Kernel = object
        mb.motionblur = mb2
        sys.modules["motionblur"] = mb
        sys.modules["motionblur.motionblur"] = mb2


def create_reference_model(model_config):
    """Build the reference UNetModel + SpacedDiffusion exactly as main_ddpir.py:216-240 does.
    `model_config` overrides utils_model.create_argparser defaults (num_channels, num_res_blocks,
    attention_resolutions, image_size, num_head_channels, channel_mult ...)."""
    install()
    from guided_diffusion.script_util import create_model_and_diffusion, model_and_diffusion_defaults, args_to_dict
    from utils import utils_model
    mc = dict(model_path="")
    mc.update(model_config)
    args = utils_model.create_argparser(mc).parse_args([])
    model, diffusion = create_model_and_diffusion(**args_to_dict(args, model_and_diffusion_defaults().keys()))
    model.eval()
    for p in model.parameters():
        p.requires_grad = False
    return model, diffusion


def ref_model_config(cfg):
    """UNetConfig (diffpir_b200.arch) or oracle cfg dict -> create_argparser overrides."""
    d = cfg if isinstance(cfg, dict) else cfg.as_dict()
    return dict(num_channels=d["model_channels"], num_res_blocks=d["num_res_blocks"],
                attention_resolutions=d["attention_resolutions"], image_size=d["image_size"],
                num_head_channels=d["num_head_channels"],
                channel_mult=",".join(str(int(c)) for c in d["channel_mult"]))
