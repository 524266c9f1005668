"""ctypes binding of librgie.so (C ABI declared in include/rgie.h).

The product path has NO fallback: if the shared library is missing or a call fails, we raise.  The library is built
in-tree by `__graft_entry__.build()` / `make -C regressor_guided_image_editing_b200/csrc`.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librgie.so")

F_EXPOSURE, F_SATURATION, F_TONE, F_COLOR, F_CONTRAST, F_SHARP, F_BLUR, F_SCALE = range(8)
F_GAMMA, F_BRIGHT, F_BW, F_HUE, F_WB, F_AFFINE = range(8, 14)
FILTER_KINDS = {"exposure": F_EXPOSURE, "saturation": F_SATURATION, "tone": F_TONE, "color": F_COLOR,
                "contrast": F_CONTRAST, "sharp": F_SHARP, "blur": F_BLUR, "scale": F_SCALE,
                "gamma": F_GAMMA, "bright": F_BRIGHT, "bw": F_BW, "hue": F_HUE, "wb": F_WB, "affine": F_AFFINE}
FILTER_NPARAM = {"exposure": 1, "saturation": 1, "tone": 8, "color": 24, "contrast": 1, "sharp": 1, "blur": 1,
                 "scale": 4}
PREC_FP32, PREC_BF16, PREC_BF16_SIMT, PREC_FP32_SIMT = 0, 1, 2, 3
PRECISIONS = {"fp32": PREC_FP32, "bf16": PREC_BF16, "bf16_simt": PREC_BF16_SIMT, "fp32_simt": PREC_FP32_SIMT}


class RgieError(RuntimeError):
    pass


_vp, _i, _l, _f = C.c_void_p, C.c_int, C.c_long, C.c_float

# name -> (restype, argtypes); every symbol include/rgie.h declares
SIGNATURES = {
    "rgie_version": (_i, []),
    "rgie_last_error": (C.c_char_p, []),
    "rgie_filter_param_count": (_i, [_i]),
    "rgie_filter_ws_floats": (_l, [_i, _i, _i]),
    "rgie_filter_fwd": (_i, [_i, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "rgie_filter_bwd": (_i, [_i, _vp, _vp, _vp, _vp, _i, _vp, _i, _i, _i, _i, _vp, _vp]),
    "rgie_filter_prefix_fwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "rgie_filter_prefix_bwd": (_i, [_vp, _vp, _vp, _i, _vp, _i, _i, _i, _i, _vp, _vp]),
    "rgie_params_default_fwd": (_i, [_vp, _vp, _i, _f, _vp]),
    "rgie_params_default_bwd": (_i, [_vp, _vp, _i, _f, _vp]),
    "rgie_resize_create": (_i, [_i, _i, _i, _i, C.POINTER(_vp)]),
    "rgie_resize_destroy": (None, [_vp]),
    "rgie_resize_fwd": (_i, [_vp, _vp, _vp, _i, _vp, _vp]),
    "rgie_resize_bwd": (_i, [_vp, _vp, _vp, _i, _vp, _vp]),
    "rgie_regressor_create": (_i, [C.POINTER(_vp), _i, _i, _i, _i, _i, C.POINTER(_vp)]),
    "rgie_regressor_destroy": (None, [_vp]),
    "rgie_regressor_workspace_bytes": (_l, [_vp]),
    "rgie_regressor_forward": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _i, _vp, _vp]),
    "rgie_regressor_forward_ex": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _l, _i, _i, _vp, _vp]),
    "rgie_regressor_backward": (_i, [_vp, _vp, _vp, _vp]),
    "rgie_regressor_set_input_transform": (_i, [_vp, C.c_float, C.c_float, _vp, _vp]),
    "rgie_regressor_set_profiling": (_i, [_vp, _i]),
    "rgie_regressor_num_ops": (_i, [_vp]),
    "rgie_regressor_get_profile": (_i, [_vp, _vp, _vp, _vp, _vp, _i, C.POINTER(_i)]),
    "rgie_launch_count": (_l, []),
    "rgie_regressor_tap": (_i, [_vp, C.c_char_p, _vp, _l, C.POINTER(_l), _vp]),
    "rgie_va_head": (_i, [_vp, _i, _i, _i, _i, _vp, _f, _f, _i, _f, _vp, _vp, _vp, _vp]),
    "rgie_adam_step": (_i, [_vp, _vp, _vp, _vp, _i, _i, _f, _f, _f, _f, _f, _f, _vp, _vp, _vp, _vp, _i, _vp]),
    "rgie_adam_step_sched": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _f, _f, _f, _f, _vp, _vp, _vp, _vp, _vp]),
    "rgie_record": (_i, [_vp, _vp, _vp, _i, _vp]),
    "rgie_counter_add": (_i, [_vp, _i, _vp]),
    "rgie_guidance_update": (_i, [_vp, _vp, _i, _l, _f, _i, _vp, _vp]),
    "rgie_midu_create": (_i, [C.POINTER(_vp), _i, _i, _i, _i, _i, C.POINTER(_vp)]),
    "rgie_midu_destroy": (None, [_vp]),
    "rgie_midu_forward": (_i, [_vp, _vp, _i, _vp, _vp]),
    "rgie_midu_backward": (_i, [_vp, _vp, _vp, _vp]),
    "rgie_gemm_selftest": (_i, [_i, _vp, _l, _i, _vp, _i, _i, C.POINTER(_l), _l, _l, _i, _vp, _vp, _i, _vp, _i, _vp]),
    "rgie_gemm_selftest_fp32": (_i, [_i, _vp, _l, _i, _i, _vp, _l, _i, _vp, _i, C.POINTER(_l), _l, _l, _i, _vp, _vp, _vp, _i,
                                     _vp, _vp]),
    "rgie_gemm_selftest_ex": (_i, [_i, _vp, _l, _i, _vp, _l, _i, _vp, _i, _i, C.POINTER(_l), _l, _l, _i, _vp, _vp, _vp, _i,
                                   _vp, _i, _vp, _vp]),
}

_lib = None


def load() -> C.CDLL:
    """Load librgie.so and attach the signatures.  Raises (never falls back) when the extension is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build the sm_100a extension first (python -c 'import __graft_entry__ as g; "
            f"g.build()' or make -C regressor_guided_image_editing_b200/csrc).  There is no CPU/PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)       # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    return load().rgie_last_error().decode("utf-8", "replace")


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise RgieError(f"{what}: {last_error()}" if what else last_error())


def ptr(t) -> C.c_void_p:
    """Device (or host) pointer of a torch tensor; None -> NULL."""
    if t is None:
        return C.c_void_p(0)
    return C.c_void_p(t.data_ptr())


def stream_ptr(device=None) -> C.c_void_p:
    import torch
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
