"""ctypes binding of libmore4d_sm100.so (the C ABI in include/more4d_b200.h).

The library is built in-tree by more4d_b200/build.py (``__graft_entry__.build()``).  Loading
works on a CPU-only box (the symbol tests need that); every compute call requires an sm_100
device and raises ``RuntimeError`` otherwise — there is no fallback path.
"""
from __future__ import annotations

import ctypes
import os
import re
from ctypes import c_float, c_int, c_longlong, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmore4d_sm100.so")
HEADER_PATH = os.path.join(HERE, "..", "include", "more4d_b200.h")

# epilogue ids (include/more4d_b200.h)
(EPI_BF16, EPI_GELU_TANH, EPI_GELU_ERF, EPI_F32, EPI_GATE_RESIDUAL_F32, EPI_ADD_BF16,
 EPI_F32_RAW) = range(7)

_P, _I, _L, _F = c_void_p, c_int, c_longlong, c_float

_SIGNATURES = {
    "m4d_version": (c_int, []),
    "m4d_error_string": (ctypes.c_char_p, [_I]),
    "m4d_device_check": (c_int, []),
    "m4d_gemm_bf16": (c_int, [_P, _L, _P, _L, _P, _P, _L, _I, _I, _I, _I, _P, _L, _P, _L, _I, _P]),
    "m4d_attention_fwd": (c_int, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _L, _L, _L, _L, _L, _L, _P,
                                  _F, _I, _P]),
    "m4d_attention_fwd_seg2": (c_int, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _L, _L, _L, _L, _L, _L, _F, _P]),
    "m4d_attention_fwd_scatter": (c_int, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _L, _L, _L, _L, _L, _L, _P,
                                          _F, _P]),
    "m4d_layernorm_modulate": (c_int, [_P, _I, _P, _P, _P, _P, _L, _L, _I, _I, _F, _P, _I, _P, _L,
                                       _I, _P, _P]),
    "m4d_rmsnorm_rope": (c_int, [_P, _L, _P, _P, _P, _P, _I, _I, _I, _I, _F, _P]),
    "m4d_rmsnorm_scatter": (c_int, [_P, _L, _P, _I, _I, _I, _F, _P, _I, _L, _L, _P]),
    "m4d_small_linear_f32": (c_int, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "m4d_timestep_embedding": (c_int, [_P, _I, _I, _P, _P]),
    "m4d_add_bcast_f32": (c_int, [_P, _P, _P, _I, _L, _L, _P]),
    "m4d_patchify": (c_int, [_P, _P, _I, _I, _I, _I, _I, _I, _P, _P]),
    "m4d_unpatchify": (c_int, [_P, _L, _I, _I, _I, _I, _I, _I, _P, _P]),
    "m4d_widen_rows": (c_int, [_P, _P, _I, _I, _I, _L, _I, _P]),
    "m4d_cfg_euler_step": (c_int, [_P, _P, _P, _F, _F, _L, _P]),
    "m4d_silu_bf16": (c_int, [_P, _P, _L, _P]),
    "m4d_conv_cl": (c_int, [_P, _I, _I, _I, _I, _P, _I, _I, _P] + [_I] * 12 + [_P, _I, _I, _I, _I, _P,
                            _I, _I, _P, _P]),
    "m4d_conv3x3_rmsnorm_cl": (c_int, [_P, _I, _I, _I, _I, _P, _I, _P, _I, _P, _P, _P, _P, _I, _P]),
    "m4d_rmsnorm_silu_cl": (c_int, [_P, _P, _P, _L, _I, _I, _P]),
    "m4d_upsample2x_cl": (c_int, [_P, _P, _I, _I, _I, _I, _P]),
    "m4d_planar_to_cl": (c_int, [_P, _P, _L, _I, _I, _P, _P, _P]),
    "m4d_cl_to_planar": (c_int, [_P, _P, _L, _I, _I, _P, _P, _P]),
    "m4d_groupnorm_swish_cl": (c_int, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _P]),
    "m4d_conv3x3_gnstats_cl": (c_int, [_P, _I, _I, _I, _I, _P, _I, _P, _P, _P, _P, _P, _P]),
    "m4d_groupnorm_apply_cl": (c_int, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _P]),
    "m4d_softmax_rows": (c_int, [_P, _P, _I, _I, _L, _L, _F, _P]),
    "m4d_transpose_bf16": (c_int, [_P, _P, _I, _I, _L, _L, _P]),
    "m4d_im2col3x3_cl": (c_int, [_P, _P, _I, _I, _I, _I, _I, _P]),
    "m4d_bilinear_repeat_cl": (c_int, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "m4d_gs_render_workspace": (c_longlong, [_L, _I, _I, _I, _L]),
    "m4d_gs_render": (c_int, [_P, _L, _P, _L, _P, _P, _P, _L, _I, _I, _I, _F, _F, _F, _P, _P, _P, _L, _L, _P, _P]),
    "m4d_project_points_workspace": (c_longlong, [_L, _I, _I]),
    "m4d_project_points": (c_int, [_P, _P, _P, _P, _L, _I, _I, _P, _P, _P, _L, _P]),
    "m4d_project_views_workspace": (c_longlong, [_L, _I, _I, _I]),
    "m4d_project_views": (c_int, [_P, _L, _P, _L, _P, _P, _L, _I, _I, _I, _P, _P, _P, _L, _P]),
}

_lib = None


def header_symbols() -> list[str]:
    """Every function name declared in include/more4d_b200.h."""
    with open(HEADER_PATH) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(m4d_[a-z0-9_]+)\s*\(", text)))


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback for the more4d_b200 kernels)")
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def dev_set_flags(flags: int) -> None:
    """Kernel-variant selection for A/B measurements.  Exists only in a development build
    (`python more4d_b200/build.py --force -DM4D_DEV`); the product library exports no such symbol
    and keeps no global mutable state."""
    l = lib()
    if not hasattr(l, "m4d_dev_set_flags"):
        raise RuntimeError("this libmore4d_sm100.so is the product build: no development variants "
                           "(rebuild with `python more4d_b200/build.py --force -DM4D_DEV`)")
    l.m4d_dev_set_flags.restype = None
    l.m4d_dev_set_flags.argtypes = [c_int]
    l.m4d_dev_set_flags(int(flags))


def is_dev_build() -> bool:
    return hasattr(lib(), "m4d_dev_set_flags")


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().m4d_error_string(rc).decode()
        raise RuntimeError(f"more4d_b200: {what} failed: {msg} (code {rc})")


_device_ok = None


def require_device() -> None:
    """Fail loudly unless an sm_100 CUDA device is current."""
    global _device_ok
    if _device_ok is None:
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("more4d_b200 needs a CUDA sm_100 (B200) device; none is visible "
                               "and there is no CPU fallback")
        rc = lib().m4d_device_check()
        check(rc, "device check (compute capability 10.x required)")
        _device_ok = True
