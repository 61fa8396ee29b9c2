"""ctypes binding of libfnst.so (the C ABI declared in include/fnst.h).

The library handle lives at module scope so nn.Modules that use it stay picklable
(the reference pickles the whole module, train.py:297).  There is no fallback: if the shared
library is missing or fails to load, importing this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

MAX_TAPS = 96
F32, F16, BF16 = 0, 1, 2
EPI_NHWC, EPI_D2S, EPI_NCHW_F32, EPI_ROWSUM9 = 0, 1, 2, 3
PAD_NONE, PAD_REFLECT, PAD_ZERO = 0, 1, 2
DESC_PREZEROED = 1
DESC_LINEAR = 2

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FNST_LIB", os.path.join(_HERE, "libfnst.so"))


class ConvDesc(C.Structure):
    """Mirror of `fnst_conv_desc` (include/fnst.h)."""
    _fields_ = [
        ("a", C.c_void_p),
        ("a_stride_w", C.c_int64), ("a_stride_h", C.c_int64), ("a_stride_n", C.c_int64),
        ("a_w", C.c_int32), ("a_h", C.c_int32), ("a_n", C.c_int32), ("a_c", C.c_int32),
        ("ntaps", C.c_int32), ("kc", C.c_int32),
        ("h0", C.c_int32), ("w0", C.c_int32),
        ("tap_dh", C.c_int8 * MAX_TAPS),
        ("tap_dw", C.c_int8 * MAX_TAPS),
        ("tap_c0", C.c_int16 * MAX_TAPS),
        ("b", C.c_void_p),
        ("n_gemm", C.c_int32),
        ("out_n", C.c_int32), ("out_h", C.c_int32), ("out_w", C.c_int32),
        ("epilogue", C.c_int32),
        ("c_out", C.c_int32),
        ("relu", C.c_int32),
        ("dtype", C.c_int32),
        ("out_dtype", C.c_int32),
        ("out", C.c_void_p),
        ("bias", C.c_void_p),
        ("stats", C.c_void_p),
        ("addend", C.c_void_p),
        ("mask", C.c_void_p),
        ("mask_dtype", C.c_int32),
        ("b_image_rows", C.c_int32),
        ("g_stride_w", C.c_int64), ("g_stride_h", C.c_int64), ("g_stride_n", C.c_int64),
        ("flags", C.c_int32),
    ]


EXPORTS = {
    # name: (restype, argtypes)
    "fnst_version": (C.c_int, []),
    "fnst_last_error": (C.c_char_p, []),
    "fnst_set_tuning": (C.c_int, [C.c_char_p, C.c_int]),
    "fnst_set_debug_buffer": (C.c_int, [C.c_void_p]),
    "fnst_device_supports_tc": (C.c_int, [C.c_int]),
    "fnst_conv_tc": (C.c_int, [C.POINTER(ConvDesc), C.c_int, C.c_void_p]),
    "fnst_finalconv_tc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "fnst_conv_simt": (C.c_int, [C.POINTER(ConvDesc), C.c_int, C.c_void_p]),
    "fnst_conv_first": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                  C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "fnst_inorm_apply": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                   C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "fnst_image_to_halo": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 11 + [C.c_void_p]),
    "fnst_maxpool2": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "fnst_gram": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "fnst_sse": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "fnst_tv": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "fnst_nhwc_to_nchw": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "fnst_nchw_to_nhwc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "fnst_u8_to_nchw": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int, C.c_void_p]),
    "fnst_nchw_to_u8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int, C.c_void_p]),
    "fnst_wgrad_simt": (C.c_int, [C.POINTER(ConvDesc), C.c_int, C.c_int, C.c_void_p]),
    "fnst_wgrad_tc": (C.c_int, [C.POINTER(ConvDesc), C.c_int, C.c_int, C.c_void_p]),
    "fnst_conv_first_wgrad": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "fnst_inorm_bwd_reduce": (C.c_int, [C.c_void_p] * 10 + [C.c_int] * 7 + [C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "fnst_inorm_bwd_apply": (C.c_int, [C.c_void_p] * 7 + [C.c_int] * 6 + [C.c_float, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "fnst_inorm_bwd_fused_parts": (C.c_int, [C.c_int] * 9),
    "fnst_inorm_bwd_fused": (C.c_int, [C.c_void_p] * 10 + [C.c_int] * 7 + [C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "fnst_affine_grads": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "fnst_loss_workspace_bytes": (C.c_int64, []),
    "fnst_sse_scaled": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p,
                                  C.c_int, C.c_int, C.c_void_p]),
    "fnst_tv_scaled": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "fnst_maxpool2_bwd": (C.c_int, [C.c_void_p] * 4 + [C.c_int] * 7 + [C.c_void_p]),
    "fnst_sse_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_float, C.c_void_p,
                               C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "fnst_tv_bwd": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_float, C.c_void_p, C.c_int, C.c_void_p]),
    "fnst_relu_mask": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "fnst_cast": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "fnst_gram_diff_sym": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_float, C.c_void_p,
                                     C.c_int, C.c_int, C.c_void_p]),
    "fnst_gather_cast": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_void_p]),
    "fnst_channel_sum": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "fnst_grad_norm_workspace_bytes": (C.c_int64, []),
    "fnst_grad_norm": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_int, C.c_void_p, C.c_float, C.c_void_p, C.c_int, C.c_void_p]),
    "fnst_grad_scale": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "fnst_resize_to_tensor": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                        C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int, C.c_void_p]),
    "fnst_resize_batch_to_tensor": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                              C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int, C.c_void_p]),
    "fnst_resize_window_host": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int]),
    "fnst_adam_step": (C.c_int, [C.POINTER(C.c_void_p)] * 4 + [C.POINTER(C.c_int64), C.c_int] + [C.c_double] * 5
                       + [C.c_int64, C.c_void_p, C.c_int, C.c_void_p]),
}


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"libfnst.so not found at {LIB_PATH}: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C fast_neural_style_transfer_b200/csrc`.  There is no CPU / library fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)      # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib.fnst_last_error()
        raise RuntimeError(f"libfnst {what} failed (rc={rc}): {msg.decode() if msg else '?'}")
