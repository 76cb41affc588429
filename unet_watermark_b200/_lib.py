"""ctypes binding of libuwm_b200.so (C ABI declared in include/uwm.h).

The library is built in-tree by ``__graft_entry__.build()`` (or ``python -m
unet_watermark_b200.build``).  There is no CPU fallback: if the shared object is missing,
``load()`` raises, and every compute entry point raises ``RuntimeError`` with the library's
own message when CUDA is unavailable.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
# UWM_LIB_PATH: A/B of two builds; UWM_TOOLS=1: the measurement build (python -m unet_watermark_b200.build --tools)
LIB_PATH = os.environ.get("UWM_LIB_PATH") or os.path.join(
    _HERE, "lib", "libuwm_b200_tools.so" if os.environ.get("UWM_TOOLS") == "1" else "libuwm_b200.so")

IN_F32_NCHW = 0
IN_U8_NHWC = 1
PACK_TAPS = 0
PACK_STEM_S2D = 1
PACK_UP2X_SHUFFLE = 2
PACK_UPCAT_SUBPIXEL = 3
PACK_S2D_CONV = 4
PACK_S2_PLANES = 5
PACK_TAPS_SKIP_PART = 6
PACK_UP2X_SHUFFLE_X_PART = 7


class ImageDesc(C.Structure):
    """uwm_image_desc: one image of a ragged batch inside a packed buffer."""
    _fields_ = [("offset", C.c_int64), ("width", C.c_int32), ("height", C.c_int32), ("pitch", C.c_int32),
                ("reserved", C.c_int32)]


class LayerDesc(C.Structure):
    _fields_ = [
        ("conv_key", C.c_char * 96),
        ("bn_key", C.c_char * 96),
        ("cin", C.c_int32), ("cout", C.c_int32), ("cout_pad", C.c_int32),
        ("kh", C.c_int32), ("kw", C.c_int32), ("stride", C.c_int32), ("pad", C.c_int32),
        ("pack", C.c_int32), ("relu", C.c_int32), ("has_residual", C.c_int32),
        ("w_elems", C.c_int64), ("b_elems", C.c_int64),
        ("flops_per_image", C.c_double),
        ("cin_skip", C.c_int32), ("reserved", C.c_int32),
    ]


# name -> (restype, argtypes).  Must list every symbol include/uwm.h declares
# (tests/test_abi.py checks the header against this table and against the .so).
_P = C.c_void_p
SIGNATURES = {
    "uwm_last_error": (C.c_char_p, []),
    "uwm_abi_version": (C.c_int, []),
    "uwm_kernel_launch_count": (C.c_uint64, []),
    "uwm_conv2d_nhwc_bf16": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, C.c_int,
                                       C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_int, C.c_int, _P,
                                       C.c_int, _P]),
    "uwm_conv2d_upcat_nhwc_bf16": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_int,
                                             C.c_int, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P,
                                             C.c_int, _P]),
    "uwm_conv2d_up2x_shuffle_nhwc_bf16": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, C.c_int,
                                                    C.c_int, _P, C.c_int, _P]),
    "uwm_conv2d_up2x_shuffle_res_nhwc_bf16": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, C.c_int,
                                                        _P, C.c_int, C.c_int, _P, C.c_int, _P]),
    "uwm_conv2d_upcat_subpixel_nhwc_bf16": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_int,
                                                      C.c_int, _P, _P, C.c_int, C.c_int, _P, C.c_int, _P]),
    "uwm_conv2d_s2_planes_nhwc_bf16": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, C.c_int,
                                                 C.c_int, _P, C.c_int, _P]),
    "uwm_conv2d_s2d_nhwc_bf16": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, C.c_int, _P, C.c_int, _P]),
    "uwm_head_s2d_nhwc_bf16": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P, C.c_int, _P, C.c_float, _P]),
    "uwm_head_nhwc_bf16": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P, C.c_int,
                                     _P, C.c_float, _P]),
    "uwm_maxpool3x3s2_nhwc_bf16": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_int, _P]),
    "uwm_upsample2x_nhwc_bf16": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_int, _P]),
    "uwm_prep_input": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P]),
    "uwm_model_create": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int, C.POINTER(_P)]),
    "uwm_model_destroy": (C.c_int, [_P]),
    "uwm_model_num_layers": (C.c_int, [_P]),
    "uwm_model_layer_desc": (C.c_int, [_P, C.c_int, C.POINTER(LayerDesc)]),
    "uwm_model_set_layer": (C.c_int, [_P, C.c_int, _P, C.c_int64, _P, C.c_int64]),
    "uwm_model_workspace_bytes": (C.c_size_t, [_P]),
    "uwm_model_num_kernels": (C.c_int, [_P]),
    "uwm_model_flops_per_image": (C.c_double, [_P]),
    "uwm_model_forward": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, C.c_int, _P, C.c_float, C.c_int, _P]),
    "uwm_model_read_tensor": (C.c_int, [_P, C.c_char_p, C.c_int, _P, C.c_int64, C.POINTER(C.c_int),
                                        C.POINTER(C.c_int), C.POINTER(C.c_int), _P]),
    "uwm_model_profile": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, _P, C.c_float, C.c_char_p,
                                    C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                    C.c_int, _P]),
    "uwm_resize_bilinear_u8": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P]),
    "uwm_mask_upscale_threshold": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, _P, C.c_float, _P, _P, _P]),
    "uwm_mask_postprocess_workspace": (C.c_size_t, [_P, C.c_int]),
    "uwm_mask_postprocess": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, _P, C.c_size_t, _P]),
    "uwm_mask_text_features": (C.c_int, [_P, _P, _P, C.c_int, C.c_uint, _P, _P, C.c_size_t, _P]),
    "uwm_mask_morphology": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_size_t, _P]),
    "uwm_mask_components": (C.c_int, [_P, _P, _P, C.c_int, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "uwm_mask_component_summary": (C.c_int, [_P, _P, _P, C.c_int, _P, _P, C.c_size_t, _P]),
    "uwm_bn_train_forward_nhwc_bf16": (C.c_int, [_P, C.c_longlong, C.c_int, _P, _P, _P, _P, C.c_float, C.c_float, _P,
                                                 C.c_int, _P, _P, _P, _P]),
    "uwm_bn_train_backward_nhwc_bf16": (C.c_int, [_P, _P, _P, C.c_longlong, C.c_int, _P, C.c_int, C.c_int, _P, _P, _P,
                                                  _P, _P, _P, _P]),
    "uwm_upsample2x_backward_nhwc_bf16": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_int, _P]),
    "uwm_maxpool3x3s2_backward_nhwc_bf16": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P]),
    "uwm_pack_train_weights": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "uwm_copy_channels_nhwc_bf16": (C.c_int, [_P, C.c_longlong, C.c_int, C.c_int, _P, C.c_int, _P]),
}

# exported only by the tools build (-DUWM_BENCH_TOOLS); bound when present
TOOLS_SIGNATURES = {
    "uwm_debug_set_trace": (C.c_int, [_P]),
    "uwm_debug_prim_cost": (C.c_int, [C.c_int, C.c_int, _P, _P]),
    "uwm_debug_handshake": (C.c_int, [C.c_int, C.c_int, C.c_int, _P, _P]),
    "uwm_debug_mma_rate": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P]),
}

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load libuwm_b200.so; raises (never falls back) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build the CUDA extension first "
            "(python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    for name, (res, args) in TOOLS_SIGNATURES.items():
        fn = getattr(lib, name, None)
        if fn is not None:
            fn.restype = res
            fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> int:
    """Raise RuntimeError(uwm_last_error()) on a negative status."""
    if rc < 0:
        msg = load().uwm_last_error().decode("utf-8", "replace")
        # smp raises RuntimeError for shapes not divisible by 32; keep the same exception type
        raise RuntimeError(f"{what + ': ' if what else ''}{msg}")
    return rc
