"""ctypes binding of the C ABI declared in include/seunet_b200.h.

Loading fails loudly when the shared library has not been built: the hot path has no fallback.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# SEUNET_LIB_PATH: developer override for A/B timing of two builds of the same ABI
LIB_PATH = os.environ.get("SEUNET_LIB_PATH") or os.path.join(HERE, "libseunet_b200.so")

_c = ctypes
_vp, _i, _i64, _sz = _c.c_void_p, _c.c_int, _c.c_int64, _c.c_size_t

# name -> (restype, argtypes); mirrors include/seunet_b200.h one to one
SIGNATURES = {
    "seunet_version": (_i, []),
    "seunet_act_dtype": (_i, []),
    "seunet_last_error": (_c.c_char_p, []),
    "seunet_param_count": (_i64, [_i, _i]),
    "seunet_param_offset": (_i64, [_i, _i, _c.c_char_p]),
    "seunet_param_tensors": (_i, [_i, _i]),
    "seunet_param_name": (_c.c_char_p, [_i, _i, _i]),
    "seunet_param_numel": (_i64, [_i, _i, _i]),
    "seunet_plan_create": (_i, [_c.POINTER(_vp), _i, _i, _i, _i, _i, _i, _i, _i]),
    "seunet_plan_destroy": (None, [_vp]),
    "seunet_plan_workspace_bytes": (_sz, [_vp]),
    "seunet_plan_wimg_bytes": (_sz, [_vp]),
    "seunet_plan_bind": (_i, [_vp, _vp, _vp, _vp]),
    "seunet_pack_weights": (_i, [_vp, _vp, _vp]),
    "seunet_forward": (_i, [_vp, _vp, _c.POINTER(_i64), _c.POINTER(_i64), _vp, _vp, _vp, _vp, _vp, _vp]),
    "seunet_forward_window": (_i, [_vp, _vp, _c.POINTER(_i64), _c.POINTER(_i64), _vp, _vp, _vp, _c.POINTER(_i), _vp, _i, _i, _i, _i, _vp]),
    "seunet_backward": (_i, [_vp, _vp, _c.POINTER(_i64), _c.POINTER(_i64), _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "seunet_plan_set_timing": (_i, [_vp, _i]),
    "seunet_plan_timing_count": (_i, [_vp]),
    "seunet_plan_timing_get": (_i, [_vp, _i, _c.POINTER(_c.c_char_p), _c.POINTER(_c.c_float), _c.POINTER(_c.c_double)]),
    "seunet_plan_debug_buffer": (_i, [_vp, _c.c_char_p, _c.POINTER(_vp), _c.POINTER(_i), _c.POINTER(_i)]),
    "seunet_debug_poison_smem": (_i, [_vp]),
    "seunet_conv_scratch_bytes": (_sz, [_i, _i, _i, _i]),
    "seunet_conv_fprop": (_i, [_vp, _i, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "seunet_wgrad_scratch_bytes": (_sz, [_i, _i, _i]),
    "seunet_conv_wgrad": (_i, [_vp, _i, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "seunet_to_chunks": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _i, _i, _vp]),
    "seunet_from_chunks": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "seunet_loss_sums": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _i, _i64, _vp, _vp, _vp]),
    "seunet_loss_grad": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp]),
    "seunet_adamw_step": (_i, [_vp, _vp, _vp, _vp, _i64, _c.c_float, _c.c_float, _c.c_float, _c.c_float, _c.c_float, _i,
                               _c.c_float, _i64, _i64, _vp]),
    "seunet_hu_windows": (_i, [_vp, _i, _i64, _c.c_double, _vp, _vp]),
    "seunet_hu_windows_slab": (_i, [_vp, _i, _i64, _i64, _c.c_double, _vp, _vp]),
    "seunet_window_accumulate": (_i, [_vp, _c.POINTER(_i), _i, _i, _i, _i, _vp, _i, _i, _i, _i, _i, _vp]),
    "seunet_window_finalize": (_i, [_vp, _vp, _i, _i, _i, _c.c_float, _vp, _i, _i, _vp]),
    "seunet_postproc_scratch_bytes": (_sz, [_i, _i, _i, _i64]),
    "seunet_postproc_dti": (_i, [_vp, _i, _i, _i, _c.c_double, _c.c_double, _c.c_double, _vp, _vp, _i64, _vp]),
    "seunet_postproc_largest_component": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _i64, _vp]),
}

_lib = None


class SeunetError(RuntimeError):
    pass


def lib():
    """Returns the loaded library, raising if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SeunetError(
            f"{LIB_PATH} not found: build it with `python -m se_unet_airseg_b200.build` "
            "(or __graft_entry__.build()); the sm_100a CUDA extension is required, there is no fallback")
    l = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(l, name)  # AttributeError if header and library drift apart
        fn.restype = res
        fn.argtypes = args
    _lib = l
    return l


def check(rc, what=""):
    if rc != 0:
        msg = lib().seunet_last_error()
        raise SeunetError(f"{what} failed: {msg.decode() if msg else 'unknown error'}")


def ptr(t):
    """Device/host pointer of a torch tensor (or None)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr(stream=None):
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return ctypes.c_void_p(s.cuda_stream)
