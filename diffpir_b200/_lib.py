"""ctypes binding of libdiffpir_b200.so (the C-ABI declared in include/diffpir_b200.h).

No fallback: if the shared library is missing it is built with nvcc; if that fails, or a compute entry
point reports an error (e.g. no GPU), a RuntimeError is raised.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int32, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdiffpir_b200.so")

PRECISION_FP32 = 0
PRECISION_FP16 = 1
PRECISION_FP16S = 2
PRECISION_FP16X3 = 3


class UNetConfigC(ctypes.Structure):
    _fields_ = [("image_size", c_int32), ("in_channels", c_int32), ("model_channels", c_int32),
                ("out_channels", c_int32), ("num_res_blocks", c_int32), ("n_channel_mult", c_int32),
                ("channel_mult", c_int32 * 8), ("n_attention_ds", c_int32), ("attention_ds", c_int32 * 8),
                ("num_heads", c_int32), ("num_head_channels", c_int32), ("num_heads_upsample", c_int32),
                ("use_new_attention_order", c_int32), ("precision", c_int32), ("diffusion_steps", c_int32),
                ("channel_mult_den", c_int32)]


# name -> (restype, argtypes); must list every symbol include/diffpir_b200.h declares (tests check this)
SIGNATURES = {
    "dpir_last_error": (c_char_p, []),
    "dpir_version": (c_int32, []),
    "dpir_launch_count": (c_int64, []),
    "dpir_launch_count_reset": (None, []),
    "dpir_create": (c_int32, [POINTER(UNetConfigC), POINTER(c_void_p)]),
    "dpir_destroy": (c_int32, [c_void_p]),
    "dpir_load_weight": (c_int32, [c_void_p, c_char_p, c_void_p, POINTER(c_int64), c_int32]),
    "dpir_missing_weights": (c_int32, [c_void_p]),
    "dpir_weight_arena": (c_int32, [c_void_p, POINTER(c_void_p), POINTER(c_size_t)]),
    "dpir_mark_weights_loaded": (c_int32, [c_void_p]),
    "dpir_nccl_version": (c_int32, [POINTER(c_int32)]),
    "dpir_comm_unique_id": (c_int32, [c_void_p, c_size_t]),
    "dpir_comm_create": (c_int32, [c_void_p, c_int32, c_int32, POINTER(c_void_p)]),
    "dpir_comm_adopt": (c_int32, [c_void_p, c_int32, c_int32, POINTER(c_void_p)]),
    "dpir_comm_destroy": (c_int32, [c_void_p]),
    "dpir_bcast_weights": (c_int32, [c_void_p, c_void_p, c_int32, c_void_p]),
    "dpir_set_tables": (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_int32]),
    "dpir_unet_forward": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p]),
    "dpir_unet_x0": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p]),
    "dpir_workspace_bytes": (c_int64, [c_void_p]),
    "dpir_set_graph_mode": (c_int32, [c_void_p, c_int32]),
    "dpir_set_film_table": (c_int32, [c_void_p, c_int32]),
    "dpir_profile_begin": (c_int32, [c_void_p]),
    "dpir_profile_end": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32]),
    "dpir_fft_plan_create": (c_int32, [c_int32, c_int32, c_int32, POINTER(c_void_p)]),
    "dpir_fft_plan_destroy": (c_int32, [c_void_p]),
    "dpir_rfft2": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "dpir_irfft2": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "dpir_fft_precalc": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32,
                                   c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dpir_fft_solve": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_int32, c_int32, c_int32,
                                 c_int32, c_float, c_float, c_float, c_float, c_float, c_void_p, c_void_p, c_void_p]),
    "dpir_fft_solve_dev": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32,
                                     c_int32, c_float, c_float, c_float, c_float, c_float, c_void_p, c_void_p, c_void_p]),
    "dpir_inpaint_prox_dev": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float, c_float, c_void_p, c_int64,
                                        c_void_p]),
    "dpir_ibp_update_dev": (c_int32, [c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_int32, c_int32, c_int32, c_int32,
                                      c_void_p, c_void_p]),
    "dpir_inpaint_prox": (c_int32, [c_void_p, c_void_p, c_void_p, c_float, c_float, c_float, c_float, c_void_p, c_int64,
                                    c_void_p]),
    "dpir_renoise": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float, c_float, c_float, c_float,
                               c_float, c_void_p, c_int64, c_void_p]),
    "dpir_eps_to_x0": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "dpir_sample_step": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_int32, c_int32,
                                   c_int32, c_int32, c_int32, c_void_p]),
    "dpir_affine": (c_int32, [c_void_p, c_float, c_float, c_void_p, c_int64, c_void_p]),
    "dpir_resize_separable": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_int32,
                                        c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p]),
    "dpir_ibp_update": (c_int32, [c_void_p, c_void_p, c_void_p, c_float, c_float, c_int32, c_int32, c_int32, c_int32,
                                  c_void_p, c_void_p]),
    "dpir_blur_reflect": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p]),
    "dpir_tensor2uint": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "dpir_sqdiff_batch": (c_int32, [c_void_p, c_void_p, c_float, c_float, c_float, c_float, c_int32, c_int32, c_int32, c_int32,
                                    c_void_p, c_void_p]),
    "dpir_measure_blur_u8": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32,
                                       c_int32, c_int32, c_void_p]),
    "dpir_conv_tc": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32,
                               c_int32, c_int32, c_int32, c_void_p]),
    "dpir_conv_tc_f16io": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32,
                                     c_int32, c_int32, c_int32, c_void_p]),
    "dpir_attention": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "dpir_conv_tc_variant": (c_int32, [c_int32]),
    "dpir_gn_coef": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32,
                               c_void_p, c_void_p]),
    "dpir_conv_gnfuse_f16io": (c_int32, [c_void_p, c_int32, c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                         c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "dpir_set_gn_fusion": (c_int32, [c_void_p, c_int32]),
    "dpir_conv_simt": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32,
                                 c_int32, c_int32, c_void_p]),
}

_lib = None


def load(build_if_missing=True):
    """Load (building first if needed) the shared library.  Loading works without a GPU; compute calls do not."""
    global _lib
    if _lib is not None:
        return _lib
    if build_if_missing:
        from . import build as _build
        try:
            stale = _build.needs_build()
        except OSError:
            stale = not os.path.exists(LIB_PATH)
        if stale:
            try:
                _build.build()
            except Exception as exc:
                # no nvcc on this box.  A shipped library may only be used when its stamp matches the sources (a stale binary
                # bound with the current ctypes SIGNATURES gives silent wrong results); DPIR_ALLOW_STALE_LIB=1 overrides.
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(f"libdiffpir_b200.so is missing and could not be built: {exc}") from exc
                if os.environ.get("DPIR_ALLOW_STALE_LIB") != "1":
                    raise RuntimeError(
                        f"libdiffpir_b200.so does not match the sources (stamp {_build.stamp_hash()} vs sources "
                        f"{_build.source_hash()}) and could not be rebuilt: {exc}; set DPIR_ALLOW_STALE_LIB=1 to load it anyway"
                    ) from exc
                import warnings
                warnings.warn(f"diffpir_b200: loading a STALE libdiffpir_b200.so (stamp {_build.stamp_hash()} vs sources "
                              f"{_build.source_hash()}): {exc}")
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} not found: run `python -m diffpir_b200.build` (there is no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().dpir_last_error()
        raise RuntimeError(f"diffpir_b200 {what} failed ({rc}): {msg.decode() if msg else ''}")


def require_cuda(*tensors):
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("diffpir_b200: no CUDA device visible - this engine has no CPU fallback")
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("diffpir_b200: tensors must live on the GPU (no CPU fallback)")


def device_scalar(v, device):
    """A scalar argument that lives on the GPU (the reference drivers pass `rhos[t_i]` as a CUDA tensor, main_ddpir.py:389):
    returns a 1-element contiguous fp32 CUDA tensor whose pointer the *_dev entry points read inside the kernel - no
    .item(), no host synchronisation.  Host numbers / CPU tensors return None (the caller passes them by value)."""
    import torch
    if torch.is_tensor(v) and v.is_cuda:
        v = v.reshape(-1)[:1]
        if v.dtype != torch.float32:
            v = v.float()
        return v.contiguous()
    return None


def ptr(t):
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(0)


def stream_ptr():
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def launch_count():
    return int(load().dpir_launch_count())


def launch_count_reset():
    load().dpir_launch_count_reset()
