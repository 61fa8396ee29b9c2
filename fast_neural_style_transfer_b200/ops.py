"""Operator wrappers: torch tensors in, libfnst launches on torch's current CUDA stream.

Layouts (see include/fnst.h): activations are NHWC tensors of the path's element type
(torch.float32 for the CUDA-core fp32 path, torch.float16 / bfloat16 for the tcgen05 path);
`ConvSpec` carries the tap table and packed weights of one gather-GEMM convolution.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import lib, check, ConvDesc

_DT = {torch.float32: _lib.F32, torch.float16: _lib.F16, torch.bfloat16: _lib.BF16}

# launches issued through this module since import (bench.py reports it as gpu_launches)
launch_count = 0


class KernelTimer:
    """CUDA-event pairs around tagged launches on the launching stream (bench.py roofline leg)."""

    def __init__(self, tag_prefix: str):
        self.prefix = tag_prefix
        self.pairs = []

    def wants(self, tag: str) -> bool:
        return bool(tag) and tag.startswith(self.prefix)

    def start(self):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def stop(self, e0) -> None:
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        self.pairs.append((e0, e1))

    def count(self) -> int:
        return len(self.pairs)

    def mean_ms(self):
        if not self.pairs:
            return None
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in self.pairs) / len(self.pairs)


kernel_timer: Optional[KernelTimer] = None


def dt(t: torch.dtype) -> int:
    return _DT[t]


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _ctx(t: torch.Tensor) -> Tuple[int, C.c_void_p]:
    if not t.is_cuda:
        raise RuntimeError("libfnst operators need CUDA tensors (no CPU fallback)")
    dev = t.device.index if t.device.index is not None else torch.cuda.current_device()
    return dev, C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


_TC_OK = {}


def require_tensor_cores(device: torch.device) -> None:
    """The tcgen05 / TMA kernels exist for sm_100a only: fail loudly on any other device (no fallback)."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _TC_OK:
        _TC_OK[idx] = bool(lib.fnst_device_supports_tc(idx))
    if not _TC_OK[idx]:
        raise RuntimeError(f"cuda:{idx} is not a compute-capability 10.x (Blackwell B200) device: the tcgen05/TMA kernels of "
                           "libfnst cannot run on it; use precision='fp32' (CUDA-core path) or a B200")


def _count(n: int = 1) -> None:
    global launch_count
    launch_count += n


@dataclass
class ConvSpec:
    """One gather-GEMM convolution: taps (dh, dw, c0), channels per tap, packed weights [n_gemm, ntaps*kc]."""
    taps: Sequence[Tuple[int, int, int]]
    kc: int
    weight: torch.Tensor
    n_gemm: int
    c_out: int
    h0: int = 0
    w0: int = 0
    epilogue: int = _lib.EPI_NHWC
    bias: Optional[torch.Tensor] = None
    relu: bool = False
    tag: str = ""
    addend: Optional[torch.Tensor] = None      # dgrad epilogue: out = (acc + addend) * (mask > 0)
    mask: Optional[torch.Tensor] = None
    per_image_weights: bool = False            # weight is (n, n_gemm, K): image i uses weight[i]


def _fill_desc(spec: ConvSpec, a: torch.Tensor, a_dims, a_strides, out_hw) -> ConvDesc:
    d = ConvDesc()
    n, ah, aw, ac = a_dims
    d.a = a.data_ptr()
    d.a_stride_n, d.a_stride_h, d.a_stride_w = a_strides
    d.a_n, d.a_h, d.a_w, d.a_c = n, ah, aw, ac
    d.ntaps, d.kc = len(spec.taps), spec.kc
    d.h0, d.w0 = spec.h0, spec.w0
    if len(spec.taps) > _lib.MAX_TAPS:
        raise RuntimeError(f"gather-GEMM with {len(spec.taps)} taps exceeds FNST_MAX_TAPS = {_lib.MAX_TAPS}")
    for i, (dh, dw, c0) in enumerate(spec.taps):
        if not (-128 <= dh <= 127 and -128 <= dw <= 127 and 0 <= c0 <= 32767):      # int8 / int16 fields of fnst_conv_desc
            raise RuntimeError(f"tap {i} = ({dh}, {dw}, {c0}) does not fit the descriptor (image too wide for the window view?)")
        d.tap_dh[i], d.tap_dw[i], d.tap_c0[i] = dh, dw, c0
    d.n_gemm = spec.n_gemm
    d.out_n, d.out_h, d.out_w = n, out_hw[0], out_hw[1]
    d.epilogue, d.c_out, d.relu = spec.epilogue, spec.c_out, int(spec.relu)
    d.dtype = dt(a.dtype)
    return d


class ZeroArena:
    """One zero-filled fp32 buffer handed out in slices: accumulators (InstanceNorm statistics, reduction sums) that
    kernels add into.  A single memset per forward / backward instead of one in front of every kernel keeps the
    kernels adjacent on the stream, so each can be chained to its predecessor by programmatic dependent launch."""

    def __init__(self, numel: int, device):
        self.buf = torch.zeros(numel, dtype=torch.float32, device=device)
        self.used = 0

    def take(self, *shape: int) -> torch.Tensor:
        n = 1
        for s in shape:
            n *= s
        n_al = (n + 3) // 4 * 4                      # keep every slice 16-byte aligned
        if self.used + n_al > self.buf.numel():
            raise RuntimeError("ZeroArena exhausted")
        t = self.buf[self.used:self.used + n].view(*shape)
        self.used += n_al
        return t


def conv_gather(spec: ConvSpec, a: torch.Tensor, a_dims: Tuple[int, int, int, int], a_strides: Tuple[int, int, int],
                out: torch.Tensor, out_hw: Tuple[int, int], stats: Optional[torch.Tensor], use_tc: bool,
                stats_zeroed: bool = False, linear: bool = False) -> None:
    """a_dims = (n, h, w, c) logical extents of the activation view; a_strides = (n, h, w) element strides.
    stats_zeroed: `stats` is already zero (slice of a ZeroArena); the call then issues no memset.
    linear: pixel-stream form (fnst.h FNST_DESC_LINEAR; tensor cores only): a_dims = (1, 1, pixels, c), out_hw = (1, pixels),
    tap (dh, dw) = pixel shift dh * (a_strides[1] // a_strides[2]) + dw."""
    d = _fill_desc(spec, a, a_dims, a_strides, out_hw)
    if stats_zeroed:
        d.flags = _lib.DESC_PREZEROED
    if linear:
        assert use_tc, "the pixel-stream form exists on the tensor-core kernel only"
        d.flags |= _lib.DESC_LINEAR
    wshape = (spec.n_gemm, len(spec.taps) * spec.kc)
    if spec.per_image_weights:
        wshape = (a_dims[0],) + wshape
        d.b_image_rows = spec.n_gemm
    assert spec.weight.is_contiguous() and tuple(spec.weight.shape) == wshape, (spec.weight.shape, wshape)
    assert spec.weight.dtype == a.dtype
    d.b = spec.weight.data_ptr()
    d.out_dtype = dt(out.dtype)
    if spec.addend is not None:
        assert spec.addend.dtype == out.dtype and spec.addend.is_contiguous()
        d.addend = spec.addend.data_ptr()
    if spec.mask is not None:
        assert spec.mask.is_contiguous()
        d.mask, d.mask_dtype = spec.mask.data_ptr(), dt(spec.mask.dtype)
    d.out = out.data_ptr()
    d.bias = None if spec.bias is None else spec.bias.data_ptr()
    d.stats = None if stats is None else stats.data_ptr()
    dev, st = _ctx(a)
    fn = lib.fnst_conv_tc if use_tc else lib.fnst_conv_simt
    timed = kernel_timer is not None and kernel_timer.wants(spec.tag)
    e0 = kernel_timer.start() if timed else None
    check(fn(C.byref(d), dev, st), "conv_tc" if use_tc else "conv_simt")
    if timed:
        kernel_timer.stop(e0)
    _count(2 if stats is not None and not stats_zeroed else 1)


def finalconv_stream(act_flat: torch.Tensor, n: int, h: int, w: int, wpacked: torch.Tensor, bias: torch.Tensor,
                     out: torch.Tensor) -> None:
    """final_conv forward (32 -> 3, 9x9) as the row-streaming tensor-core kernel.  act_flat: the (n, h+8, w+8, 32) reflect-halo
    buffer (flat or shaped); wpacked: engine.pack_final_stream; bias: >= 3 fp32 values on the device; out (n,3,h,w) fp32."""
    assert act_flat.numel() >= n * (h + 8) * (w + 8) * 32 and act_flat.dtype == wpacked.dtype and wpacked.numel() == 9 * 2 * 2 * 32 * 8
    assert out.dtype == torch.float32 and out.is_contiguous() and tuple(out.shape) == (n, 3, h, w) and bias.dtype == torch.float32
    dev, st = _ctx(act_flat)
    check(lib.fnst_finalconv_tc(_ptr(act_flat), _ptr(wpacked), _ptr(bias), _ptr(out), n, h, w, dt(act_flat.dtype), dev, st),
          "finalconv_tc")
    _count()


def conv_first(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], k: int, stride: int, pad: int,
               pad_mode: int, relu: bool, out: torch.Tensor, stats: Optional[torch.Tensor]) -> None:
    n, c, h, w = x.shape
    assert c == 3 and x.dtype == torch.float32 and x.is_contiguous()
    assert weight.dtype == torch.float32 and weight.is_contiguous() and weight.shape[0] == 3 * k * k   # tap-major
    dev, st = _ctx(x)
    check(lib.fnst_conv_first(_ptr(x), n, h, w, _ptr(weight), _ptr(bias), weight.shape[1], k, stride, pad, pad_mode,
                              int(relu), _ptr(out), dt(out.dtype), _ptr(stats), dev, st), "conv_first")
    _count(2 if stats is not None else 1)


def inorm_apply(raw: torch.Tensor, stats: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, out: torch.Tensor,
                relu: bool, pad: int = 0, pad_mode: int = _lib.PAD_NONE, s2d: bool = False,
                drop: Optional[torch.Tensor] = None, res: Optional[torch.Tensor] = None, res_pad: int = 0,
                eps: float = 1e-5, split: bool = False, out2: Optional[torch.Tensor] = None) -> None:
    """split: raw is fp32, out/res are fp16 buffers with 2c channels per pixel [hi | lo] (fp16x3 path).
    out2: optional bfloat16 twin of `out` (same geometry, c channels per pixel) for the weight-gradient GEMM."""
    n, h, w, c = raw.shape
    dev, st = _ctx(raw)
    if out2 is not None:
        assert out2.dtype == torch.bfloat16 and out2.is_contiguous() and out2.numel() * (2 if split else 1) >= out.numel()
    check(lib.fnst_inorm_apply(_ptr(raw), _ptr(stats), _ptr(gamma), _ptr(beta), _ptr(drop), _ptr(res), res_pad, _ptr(out), _ptr(out2),
                               n, h, w, c, dt(out.dtype), int(relu), eps, pad, pad_mode, int(s2d), dt(raw.dtype), int(split),
                               dev, st), "inorm_apply")
    _count()


def maxpool2(x: torch.Tensor) -> torch.Tensor:
    n, h, w, c = x.shape
    out = torch.empty((n, h // 2, w // 2, c), dtype=x.dtype, device=x.device)
    dev, st = _ctx(x)
    check(lib.fnst_maxpool2(_ptr(x), _ptr(out), n, h, w, c, dt(x.dtype), dev, st), "maxpool2")
    _count()
    return out


def gram(feat_nhwc: torch.Tensor, use_tc: bool, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out: optional ZEROED fp32 (n,c,c) tensor (e.g. a ZeroArena slice): the call then issues no memset."""
    n, h, w, c = feat_nhwc.shape
    assert feat_nhwc.is_contiguous()
    prezeroed = out is not None
    if out is None:
        out = torch.empty((n, c, c), dtype=torch.float32, device=feat_nhwc.device)
    assert out.dtype == torch.float32 and out.is_contiguous() and tuple(out.shape) == (n, c, c)
    dev, st = _ctx(feat_nhwc)
    check(lib.fnst_gram(_ptr(feat_nhwc), _ptr(out), n, h * w, c, dt(feat_nhwc.dtype), int(use_tc), int(prezeroed), dev, st), "gram")
    _count(1 if prezeroed else 2)
    return out


def sse(a: torch.Tensor, b: torch.Tensor, acc: torch.Tensor) -> None:
    """acc (float64 scalar tensor) += sum (a - b)^2 ; b is broadcast with period b.numel()."""
    assert a.is_contiguous() and b.is_contiguous() and acc.dtype == torch.float64
    assert a.numel() % b.numel() == 0
    dev, st = _ctx(a)
    check(lib.fnst_sse(_ptr(a), _ptr(b), a.numel(), b.numel(), dt(a.dtype), dt(b.dtype), _ptr(acc), dev, st), "sse")
    _count()


def tv(img: torch.Tensor, acc: torch.Tensor) -> None:
    b, c, h, w = img.shape
    assert img.dtype == torch.float32 and img.is_contiguous() and acc.dtype == torch.float64
    dev, st = _ctx(img)
    check(lib.fnst_tv(_ptr(img), b * c, h, w, _ptr(acc), dev, st), "tv")
    _count()


def nhwc_to_nchw(x: torch.Tensor) -> torch.Tensor:
    n, h, w, c = x.shape
    out = torch.empty((n, c, h, w), dtype=torch.float32, device=x.device)
    dev, st = _ctx(x)
    check(lib.fnst_nhwc_to_nchw(_ptr(x), _ptr(out), n, h, w, c, dt(x.dtype), dev, st), "nhwc_to_nchw")
    _count()
    return out


def nchw_to_nhwc(x: torch.Tensor, dtype: torch.dtype, c_pad: Optional[int] = None) -> torch.Tensor:
    n, c, h, w = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous()
    c_pad = c_pad or c
    out = torch.empty((n, h, w, c_pad), dtype=dtype, device=x.device)
    dev, st = _ctx(x)
    check(lib.fnst_nchw_to_nhwc(_ptr(x), _ptr(out), n, h, w, c, c_pad, dt(dtype), dev, st), "nchw_to_nhwc")
    _count()
    return out


# ---- backward operators -------------------------------------------------------------------------------

def wgrad(spec: ConvSpec, a: torch.Tensor, a_dims, a_strides, g: torch.Tensor, out_hw, use_tc: bool = False,
          g_strides: Optional[Tuple[int, int, int]] = None, out: Optional[torch.Tensor] = None,
          out_zeroed: bool = False) -> torch.Tensor:
    """dB fp32 [n_gemm, ntaps*kc] of the gather-GEMM `spec` (weights unused); g NHWC [n, oh, ow, n_gemm], or a
    strided view of it given by g_strides = (n, h, w) element strides.  out_zeroed: `out` is already zero."""
    d = _fill_desc(spec, a, a_dims, a_strides, out_hw)
    if out_zeroed:
        assert out is not None
        d.flags = _lib.DESC_PREZEROED
    if g_strides is None:
        assert g.is_contiguous() and g.shape == (a_dims[0], out_hw[0], out_hw[1], spec.n_gemm), (g.shape, spec.n_gemm)
    else:
        d.g_stride_n, d.g_stride_h, d.g_stride_w = g_strides
    if out is None:
        out = torch.empty((spec.n_gemm, len(spec.taps) * spec.kc), dtype=torch.float32, device=a.device)
    assert out.is_contiguous() and out.dtype == torch.float32 and out.numel() == spec.n_gemm * len(spec.taps) * spec.kc
    d.b, d.out = g.data_ptr(), out.data_ptr()
    dev, st = _ctx(a)
    fn = lib.fnst_wgrad_tc if use_tc else lib.fnst_wgrad_simt
    timed = kernel_timer is not None and kernel_timer.wants(spec.tag)
    e0 = kernel_timer.start() if timed else None
    check(fn(C.byref(d), dt(g.dtype), dev, st), "wgrad_tc" if use_tc else "wgrad_simt")
    if timed:
        kernel_timer.stop(e0)
    _count(1 if out_zeroed else 2)
    return out


def conv_first_wgrad(x: torch.Tensor, g: torch.Tensor, k: int, stride: int, pad: int, pad_mode: int,
                     out: Optional[torch.Tensor] = None) -> torch.Tensor:
    n, c, h, w = x.shape
    c_out = g.shape[-1]
    dw = torch.empty((3 * k * k, c_out), dtype=torch.float32, device=x.device) if out is None else out
    assert dw.is_contiguous() and dw.dtype == torch.float32 and dw.numel() == 3 * k * k * c_out
    dev, st = _ctx(x)
    check(lib.fnst_conv_first_wgrad(_ptr(x), n, h, w, _ptr(g), dt(g.dtype), c_out, k, stride, pad, pad_mode, _ptr(dw), dev, st),
          "conv_first_wgrad")
    _count(2)
    return dw


def inorm_bwd_reduce(gsrc, extra, raw, stats, gamma, beta, drop, gdtype, relu, pad=0, pad_mode=_lib.PAD_NONE, s2d=False,
                     eps: float = 1e-5, arena: Optional[ZeroArena] = None, sums: Optional[torch.Tensor] = None,
                     gsrc_slack: int = 0):
    """arena / sums: take the (zeroed) reduction buffer from the arena, or use the given zeroed [n,c,2] tensor, instead of
    having the call memset a fresh one.  gsrc_slack: gsrc is allocated that many rows and columns larger than its halo
    extent (output of a pixel-stream data-gradient GEMM)."""
    n, h, w, c = raw.shape
    gy = torch.empty((n, h, w, c), dtype=gdtype, device=raw.device)
    dgb = None                   # d gamma / d beta come out of pass 2 (inorm_bwd_apply): no same-address atomics in pass 1
    prezeroed = arena is not None or sums is not None
    if sums is not None:
        assert sums.dtype == torch.float32 and sums.is_contiguous() and sums.numel() == n * c * 2
    elif arena is not None:
        sums = arena.take(n, c, 2)
    else:
        sums = torch.empty((n, c, 2), dtype=torch.float32, device=raw.device)
    for t in (gsrc, extra):
        assert t is None or (t.dtype == gdtype and t.is_contiguous())
    dev, st = _ctx(raw)
    check(lib.fnst_inorm_bwd_reduce(_ptr(gsrc), _ptr(extra), _ptr(raw), _ptr(stats), _ptr(gamma), _ptr(beta), _ptr(drop), _ptr(gy),
                                    _ptr(sums), _ptr(dgb), n, h, w, c, dt(raw.dtype), dt(gdtype), int(relu), eps, pad, pad_mode, int(s2d),
                                    int(gsrc_slack), int(prezeroed), dev, st), "inorm_bwd_reduce")
    _count(1 if prezeroed else 2)
    return gy, sums


def inorm_bwd_apply(gy, raw, stats, sums, gamma, out_s2d=False, eps: float = 1e-5, want_dgb: bool = True, out_pad: int = 0):
    """Returns (d_raw, dgb) with dgb = [d gamma; d beta] (2, c) fp32 (None unless want_dgb).
    out_pad: d_raw is (n, h + 2*out_pad, w + 2*out_pad, c) with a zero halo written by the same launch."""
    n, h, w, c = raw.shape
    shape = (n, h // 2, w // 2, 4 * c) if out_s2d else (n, h + 2 * out_pad, w + 2 * out_pad, c)
    draw = torch.empty(shape, dtype=gy.dtype, device=raw.device)
    dgb = torch.empty((2, c), dtype=torch.float32, device=raw.device) if want_dgb else None
    dev, st = _ctx(raw)
    check(lib.fnst_inorm_bwd_apply(_ptr(gy), _ptr(raw), _ptr(stats), _ptr(sums), _ptr(gamma), _ptr(draw), _ptr(dgb), n, h, w, c,
                                   dt(raw.dtype), dt(gy.dtype), eps, int(out_s2d), int(out_pad), dev, st), "inorm_bwd_apply")
    _count()
    return draw, dgb


def inorm_bwd_fused_parts(raw: torch.Tensor, gdtype: torch.dtype, has_gsrc: bool = True, has_extra: bool = False, s2d: bool = False) -> int:
    """Cluster size the one-pass kernel needs for this configuration; 0 = unsupported / does not fit (use inorm_bwd_reduce +
    inorm_bwd_apply)."""
    n, h, w, c = raw.shape
    return int(lib.fnst_inorm_bwd_fused_parts(n, h, w, c, dt(raw.dtype), dt(gdtype), int(has_gsrc), int(has_extra), int(s2d)))


def inorm_bwd_fused(gsrc, extra, raw, stats, gamma, beta, drop, gdtype, relu, pad=0, pad_mode=_lib.PAD_NONE, s2d=False,
                    out_s2d=False, want_gy=False, sums: Optional[torch.Tensor] = None, eps: float = 1e-5):
    """InstanceNorm backward in one pass (reduce + apply).  Returns (d_raw, gy or None, sums [n,c,2])."""
    n, h, w, c = raw.shape
    shape = (n, h // 2, w // 2, 4 * c) if out_s2d else (n, h, w, c)
    draw = torch.empty(shape, dtype=gdtype, device=raw.device)
    gy = torch.empty((n, h, w, c), dtype=gdtype, device=raw.device) if want_gy else None
    if sums is None:
        sums = torch.empty((n, c, 2), dtype=torch.float32, device=raw.device)
    for t in (gsrc, extra):
        assert t is None or (t.dtype == gdtype and t.is_contiguous())
    dev, st = _ctx(raw)
    check(lib.fnst_inorm_bwd_fused(_ptr(gsrc), _ptr(extra), _ptr(raw), _ptr(stats), _ptr(gamma), _ptr(beta), _ptr(drop), _ptr(draw),
                                   _ptr(gy), _ptr(sums), n, h, w, c, dt(raw.dtype), dt(gdtype), int(relu), eps, pad, pad_mode, int(s2d),
                                   int(out_s2d), dev, st), "inorm_bwd_fused")
    _count()
    return draw, gy, sums


_AFFINE_TABLES = {}


def affine_grads(sums_flat: torch.Tensor, entries: Sequence[Tuple[int, int, int, int]], n: int, out: torch.Tensor) -> None:
    """entries: (offset of the layer's [n][C][2] block in sums_flat, C, offset of d gamma in out, offset of d beta in out).
    out[dgamma + c] = sum over images of sum(gy*xhat), out[dbeta + c] = sum over images of sum(gy)."""
    key = (tuple(entries), sums_flat.device)
    table = _AFFINE_TABLES.get(key)
    if table is None:
        rows = []
        for src, c, dg, db in entries:
            rows += [src, dg, db, c]              # int64 src, dgamma, dbeta; {int32 C, int32 pad} packed into one int64
        table = _AFFINE_TABLES[key] = torch.tensor(rows, dtype=torch.int64).to(sums_flat.device)
    assert sums_flat.dtype == torch.float32 and out.dtype == torch.float32 and out.is_contiguous()
    dev, st = _ctx(sums_flat)
    check(lib.fnst_affine_grads(_ptr(sums_flat), _ptr(table), len(entries), n, max(e[1] for e in entries), _ptr(out), dev, st), "affine_grads")
    _count()


def maxpool2_bwd(inp, gout, extra):
    n, h, w, c = inp.shape
    gin = torch.empty((n, h, w, c), dtype=gout.dtype, device=inp.device)
    assert gout.is_contiguous() and (extra is None or (extra.is_contiguous() and extra.dtype == gout.dtype))
    dev, st = _ctx(inp)
    check(lib.fnst_maxpool2_bwd(_ptr(inp), _ptr(gout), _ptr(extra), _ptr(gin), n, h, w, c, dt(inp.dtype), dt(gout.dtype), dev, st),
          "maxpool2_bwd")
    _count()
    return gin


def sse_bwd(a, b, scale, gdtype, relu_mask=False, coef: float = 1.0, out: Optional[torch.Tensor] = None):
    """2*coef*scale*(a-b) (b broadcast); scale is a 1-element fp32 CUDA tensor, coef a host constant.  out: write here."""
    assert a.is_contiguous() and b.is_contiguous() and scale.dtype == torch.float32
    da = torch.empty(a.shape, dtype=gdtype, device=a.device) if out is None else out
    assert da.dtype == gdtype and da.is_contiguous() and da.shape == a.shape
    dev, st = _ctx(a)
    check(lib.fnst_sse_bwd(_ptr(a), _ptr(b), a.numel(), b.numel(), dt(a.dtype), dt(b.dtype), _ptr(scale), float(coef), _ptr(da),
                           dt(gdtype), int(relu_mask), dev, st), "sse_bwd")
    _count()
    return da


def tv_bwd(img, scale, coef: float = 1.0):
    b, c, h, w = img.shape
    out = torch.empty_like(img)
    dev, st = _ctx(img)
    check(lib.fnst_tv_bwd(_ptr(img), b * c, h, w, _ptr(scale), float(coef), _ptr(out), dev, st), "tv_bwd")
    _count()
    return out


def channel_sum(x, out: Optional[torch.Tensor] = None):
    n, c, h, w = x.shape
    if out is None:
        out = torch.empty(c, dtype=torch.float32, device=x.device)
    assert out.dtype == torch.float32 and out.numel() == c and out.is_contiguous()
    dev, st = _ctx(x)
    check(lib.fnst_channel_sum(_ptr(x), n, c, h * w, _ptr(out), dev, st), "channel_sum")
    _count(2)
    return out


def relu_mask(g, extra, act):
    """(g + extra) * (act > 0)."""
    assert g.is_contiguous() and act.is_contiguous() and g.shape == act.shape
    out = torch.empty_like(g)
    dev, st = _ctx(g)
    check(lib.fnst_relu_mask(_ptr(g), _ptr(extra), _ptr(act), _ptr(out), g.numel(), dt(act.dtype), dt(g.dtype), dev, st), "relu_mask")
    _count()
    return out


def cast(x: torch.Tensor, dtype: torch.dtype, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    assert x.is_contiguous()
    if out is None:
        out = torch.empty(x.shape, dtype=dtype, device=x.device)
    assert out.dtype == dtype and out.is_contiguous() and out.numel() == x.numel()
    dev, st = _ctx(x)
    check(lib.fnst_cast(_ptr(x), _ptr(out), x.numel(), dt(x.dtype), dt(dtype), dev, st), "cast")
    _count()
    return out


def image_to_halo(x: torch.Tensor, pad: int, pad_mode: int, c_pad: int, rows: int, pitch: int, dtype: torch.dtype,
                  split: bool = False) -> torch.Tensor:
    """(n,3,h,w) fp32 -> flat 2-byte halo buffer viewed as (n, rows, pitch, c_pad) plus 128 zero elements of slack
    (window views of the last pixels read a little past the end)."""
    n, c, h, w = x.shape
    assert c == 3 and x.dtype == torch.float32 and x.is_contiguous()
    numel = n * rows * pitch * c_pad
    flat = torch.empty(numel + 128, dtype=dtype, device=x.device)
    flat[numel:].zero_()
    dev, st = _ctx(x)
    check(lib.fnst_image_to_halo(_ptr(x), _ptr(flat), n, h, w, pad, pad_mode, c_pad, rows, pitch, dt(dtype), int(split), dev, st),
          "image_to_halo")
    _count()
    return flat


def gram_diff_sym(g: torch.Tensor, gt: torch.Tensor, scale: torch.Tensor, coef: float, dtype: torch.dtype) -> torch.Tensor:
    """scale[0]*coef*((g-gt) + (g-gt)^T) per image, in `dtype`; gt is (C,C) or (B,C,C)."""
    n, c, _ = g.shape
    assert g.dtype == torch.float32 and gt.dtype == torch.float32 and g.is_contiguous() and gt.is_contiguous()
    out = torch.empty((n, c, c), dtype=dtype, device=g.device)
    dev, st = _ctx(g)
    check(lib.fnst_gram_diff_sym(_ptr(g), _ptr(gt), n, c, gt.numel(), _ptr(scale), float(coef), _ptr(out), dt(dtype), dev, st),
          "gram_diff_sym")
    _count()
    return out


_GATHER_MAPS = {}


def gather_pack(key: str, layout_fn, src: torch.Tensor, out_dtype: torch.dtype, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out = layout_fn(src).to(out_dtype) as ONE kernel (written into `out` when given: a contiguous tensor of that size).  `layout_fn` must be a pure re-layout (every output element is
    one input element or zero); its index map is derived once per (key, shape, device) by running it on a CPU tensor of
    element numbers and cached on the device.  (First use does host work: warm up before CUDA-graph capture.)"""
    ck = (key, tuple(src.shape), src.device)
    ent = _GATHER_MAPS.get(ck)
    if ent is None:
        probe = torch.arange(1, src.numel() + 1, dtype=torch.float64).reshape(src.shape)
        res = layout_fn(probe)
        idx = (res.reshape(-1).round().to(torch.int64) - 1).to(torch.int32)
        ent = _GATHER_MAPS[ck] = (idx.to(src.device), tuple(res.shape))
    idx, shape = ent
    s = src.detach()
    if not s.is_contiguous():
        s = s.contiguous()
    if out is None:
        out = torch.empty(shape, dtype=out_dtype, device=src.device)
    else:
        assert out.is_contiguous() and out.dtype == out_dtype and out.numel() == idx.numel()
    dev, st = _ctx(s)
    check(lib.fnst_gather_cast(_ptr(s), dt(s.dtype), _ptr(idx), _ptr(out), dt(out_dtype), idx.numel(), dev, st), "gather_cast")
    _count()
    return out


def gather_index(src: torch.Tensor, idx: torch.Tensor, out: torch.Tensor) -> None:
    """out[i] = idx[i] < 0 ? 0 : src[idx[i]] with a caller-built int32 index map (gradient assembly)."""
    assert idx.dtype == torch.int32 and idx.numel() == out.numel() and out.is_contiguous() and src.is_contiguous()
    dev, st = _ctx(src)
    check(lib.fnst_gather_cast(_ptr(src), dt(src.dtype), _ptr(idx), _ptr(out), dt(out.dtype), idx.numel(), dev, st), "gather_cast")
    _count()


_LOSS_WS = {}


def _loss_workspace(device: torch.device) -> torch.Tensor:
    """Zeroed once; the loss kernels leave it zeroed.  One per (device, stream)."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    key = (idx, torch.cuda.current_stream(idx).cuda_stream)
    ws = _LOSS_WS.get(key)
    if ws is None:
        ws = _LOSS_WS[key] = torch.zeros(int(lib.fnst_loss_workspace_bytes()) // 8, dtype=torch.float64, device=device)
    return ws


def sse_scaled(a: torch.Tensor, b: torch.Tensor, scale: float, out: torch.Tensor, accumulate: bool = False) -> None:
    """out[0] (+)= scale * sum (a - b)^2 ; b broadcast with period b.numel().  One launch, fp32 result on the device."""
    assert a.is_contiguous() and b.is_contiguous() and out.dtype == torch.float32 and a.numel() % b.numel() == 0
    dev, st = _ctx(a)
    check(lib.fnst_sse_scaled(_ptr(a), _ptr(b), a.numel(), b.numel(), dt(a.dtype), dt(b.dtype), float(scale),
                              _ptr(_loss_workspace(a.device)), _ptr(out), int(accumulate), dev, st), "sse_scaled")
    _count()


def tv_scaled(img: torch.Tensor, scale: float, out: torch.Tensor) -> None:
    b, c, h, w = img.shape
    assert img.dtype == torch.float32 and img.is_contiguous() and out.dtype == torch.float32
    dev, st = _ctx(img)
    check(lib.fnst_tv_scaled(_ptr(img), b * c, h, w, float(scale), _ptr(_loss_workspace(img.device)), _ptr(out), dev, st), "tv_scaled")
    _count()


def _f3(v):
    return (C.c_float * 3)(*[float(t) for t in v])


def u8_to_nchw(img_u8: torch.Tensor, mean=(0.0, 0.0, 0.0), std=(1.0, 1.0, 1.0)) -> torch.Tensor:
    """(n,h,w,3) uint8 -> (n,3,h,w) fp32: (u8/255 - mean)/std."""
    n, h, w, c = img_u8.shape
    assert c == 3 and img_u8.dtype == torch.uint8 and img_u8.is_contiguous()
    out = torch.empty((n, 3, h, w), dtype=torch.float32, device=img_u8.device)
    dev, st = _ctx(img_u8)
    check(lib.fnst_u8_to_nchw(_ptr(img_u8), _ptr(out), n, h, w, _f3(mean), _f3(std), dev, st), "u8_to_nchw")
    _count()
    return out


def nchw_to_u8(y: torch.Tensor, mean, std) -> torch.Tensor:
    """(n,3,h,w) fp32 -> (n,h,w,3) uint8: trunc(clamp(y*std + mean, 0, 1)*255), i.e. ToPILImage's mul(255).byte() (inference.py:57-60)."""
    n, c, h, w = y.shape
    assert c == 3 and y.dtype == torch.float32 and y.is_contiguous()
    out = torch.empty((n, h, w, 3), dtype=torch.uint8, device=y.device)
    dev, st = _ctx(y)
    check(lib.fnst_nchw_to_u8(_ptr(y), _ptr(out), n, h, w, _f3(mean), _f3(std), dev, st), "nchw_to_u8")
    _count()
    return out
