"""ctypes front-end of oracle/pil_resize.c (TEST INFRASTRUCTURE ONLY: tests/, smoke(), bench cpu_baseline).

`resize_bilinear_u8` restates Pillow's Image.resize(BILINEAR) for RGB uint8 images, `to_tensor` restates torchvision's
ToTensor (+ Normalize) -- the per-image transform of the reference's input pipeline (train.py:92-102,
inference.py:28-31, data/dataset.py:21-27).  Pinned against Pillow / torchvision by tests/test_oracle_resize.py."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libfnst_oracle.so")


def build() -> str:
    subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _SO


def _lib() -> C.CDLL:
    if not os.path.exists(_SO):
        build()
    lib = C.CDLL(_SO)
    lib.fnst_oracle_resize_coeffs.restype = C.c_int
    return lib


def resize_bilinear_u8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """img: (H, W, 3) uint8 (rows may be strided); returns (out_h, out_w, 3) uint8."""
    assert img.dtype == np.uint8 and img.ndim == 3 and img.shape[2] == 3 and img.strides[2] == 1 and img.strides[1] == 3
    out = np.empty((out_h, out_w, 3), np.uint8)
    _lib().fnst_oracle_resize_bilinear_u8(C.c_void_p(img.ctypes.data), img.shape[0], img.shape[1], C.c_int64(img.strides[0]),
                                         C.c_void_p(out.ctypes.data), out_h, out_w)
    return out


def to_tensor(img_u8: np.ndarray, mean=None, std=None) -> np.ndarray:
    """(H, W, 3) uint8 -> (3, H, W) float32 = u8/255 [then (x-mean)/std]."""
    img_u8 = np.ascontiguousarray(img_u8)
    h, w, _ = img_u8.shape
    out = np.empty((3, h, w), np.float32)
    m = None if mean is None else (C.c_float * 3)(*mean)
    s = None if std is None else (C.c_float * 3)(*std)
    _lib().fnst_oracle_to_tensor(C.c_void_p(img_u8.ctypes.data), h, w, m, s, C.c_void_p(out.ctypes.data))
    return out


def windows(in_size: int, out_size: int):
    """Per-output (first, length, fixed-point weights) of one axis, as Pillow's precompute_coeffs / normalize_coeffs_8bpc."""
    lib = _lib()
    b, k = C.POINTER(C.c_int)(), C.POINTER(C.c_int)()
    ks = lib.fnst_oracle_resize_coeffs(in_size, out_size, C.byref(b), C.byref(k))
    res = [(b[2 * i], b[2 * i + 1], [k[i * ks + x] for x in range(b[2 * i + 1])]) for i in range(out_size)]
    libc = C.CDLL(None)
    libc.free(b); libc.free(k)
    return ks, res
