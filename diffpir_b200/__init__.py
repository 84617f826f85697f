"""diffpir_b200 - Blackwell-native (sm_100a) DiffPIR sampling hot path.

Host side is Python mirroring the reference's operator surface (utils_model.model_fn / test_mode,
utils_sisr.pre_calculate / data_solution, Resizer, blur operators, the UNet model object); all
compute is hand-written CUDA reached through the C-ABI in include/diffpir_b200.h
(libdiffpir_b200.so, loaded with ctypes by diffpir_b200._lib).  There is no CPU fallback:
every operator raises if the CUDA library is missing or no GPU is visible.
"""
from .arch import UNetConfig, ffhq_config, imagenet_config  # noqa: F401

__version__ = "0.1.0"
