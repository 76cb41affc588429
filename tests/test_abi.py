"""C-ABI surface: every symbol include/uwm.h declares is exported by libuwm_b200.so and bound in
unet_watermark_b200/_lib.py; without a GPU the compute entries fail loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

from unet_watermark_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "uwm.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"#ifdef UWM_BENCH_TOOLS.*?#endif", "", src, flags=re.S)      # tools build only (checked below)
    return sorted(set(re.findall(r"\b(uwm_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_expected_surface():
    syms = header_symbols()
    for must in ("uwm_conv2d_nhwc_bf16", "uwm_head_nhwc_bf16", "uwm_maxpool3x3s2_nhwc_bf16",
                 "uwm_upsample2x_nhwc_bf16", "uwm_prep_input", "uwm_model_create", "uwm_model_forward",
                 "uwm_model_set_layer", "uwm_model_destroy", "uwm_last_error"):
        assert must in syms


def test_library_exports_every_header_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in header_symbols():
        assert hasattr(lib, s), f"{s} declared in include/uwm.h but not exported"


def test_binding_table_matches_header():
    assert sorted(_lib.SIGNATURES) == header_symbols()
    lib = _lib.load()
    assert lib.uwm_abi_version() == 1


def test_product_library_has_no_bench_tools():
    """uwm_debug_* and the UWM_DBG switches live only in the tools build (-DUWM_BENCH_TOOLS)."""
    src = open(os.path.join(ROOT, "include", "uwm.h")).read()
    tools = re.search(r"#ifdef UWM_BENCH_TOOLS(.*?)#endif", src, flags=re.S).group(1)
    declared = sorted(set(re.findall(r"\b(uwm_debug_[a-z0-9_]+)\s*\(", tools)))
    assert declared == sorted(_lib.TOOLS_SIGNATURES)
    if os.path.basename(_lib.LIB_PATH) == "libuwm_b200.so":
        lib = ctypes.CDLL(_lib.LIB_PATH)
        for s in declared:
            assert not hasattr(lib, s), f"{s} exported by the product library"
        blob = open(_lib.LIB_PATH, "rb").read()
        assert b"UWM_DBG" not in blob and b"UWM_TRACE_KH" not in blob


def test_layer_desc_struct_matches_header_layout():
    # char[96]*2 + 10*int32 + 2*int64 + double + 2*int32 (cin_skip, reserved)
    assert ctypes.sizeof(_lib.LayerDesc) == 96 * 2 + 10 * 4 + 2 * 8 + 8 + 2 * 4


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_cpu_fallback_model_create_fails_loudly():
    lib = _lib.load()
    h = ctypes.c_void_p()
    dec = (ctypes.c_int * 5)(256, 128, 64, 32, 16)
    rc = lib.uwm_model_create(34, dec, 64, 64, 1, ctypes.byref(h))
    assert rc < 0
    with pytest.raises(RuntimeError):
        _lib.check(rc, "uwm_model_create")


def test_bad_shape_message_matches_smp_family():
    lib = _lib.load()
    h = ctypes.c_void_p()
    dec = (ctypes.c_int * 5)(256, 128, 64, 32, 16)
    rc = lib.uwm_model_create(34, dec, 100, 64, 1, ctypes.byref(h))
    assert rc == -1
    assert b"divisible by 32" in lib.uwm_last_error()
    rc = lib.uwm_model_create(18, dec, 64, 64, 1, ctypes.byref(h))
    assert rc == -1 and b"resnet18" in lib.uwm_last_error()
