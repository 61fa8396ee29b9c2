"""Forward plans for the style-transfer hot path: weight packing, halo-buffer management and the
operator sequence of StyleTransferNet (models/model.py:49-65) and VGG-19 features[0:25]
(models/vgg19_net.py:56-65) expressed in libfnst operators.

Two precisions share the same plan:
  "fp16"/"bf16"  tcgen05 tensor-core path: 2-byte NHWC activations, fp32 accumulation/statistics
  "fp32"         CUDA-core path: fp32 NHWC activations (the 1e-4 parity path)
"""
from __future__ import annotations

import contextlib

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import os

import torch

from . import ops
from ._lib import EPI_NHWC, EPI_D2S, EPI_NCHW_F32, EPI_ROWSUM9, PAD_NONE, PAD_REFLECT, PAD_ZERO
from .ops import ConvSpec

# "fp16x3": error-compensated tensor-core path for the 1e-4 class (inference forward): every activation and weight is
# an fp16 pair (hi, lo) and each product is hi*hi + hi*lo + lo*hi accumulated in fp32 -- expressed on the ordinary
# tcgen05 gather-GEMM as three "virtual taps" per real tap that read different channel windows of [hi | lo] buffers.
PRECISIONS = {"fp32": torch.float32, "fp16": torch.float16, "bf16": torch.bfloat16, "fp16x3": torch.float16}


def act_dtype(precision: str) -> torch.dtype:
    if precision not in PRECISIONS:
        raise ValueError(f"unknown precision {precision!r}; expected one of {sorted(PRECISIONS)}")
    return PRECISIONS[precision]


# ---------------------------------------------------------------------------------------------
# Weight packing: PyTorch parameter layouts -> gather-GEMM B operands [n_gemm, ntaps*kc]
# ---------------------------------------------------------------------------------------------

def taps_kxk(k: int, origin: int = 0) -> List[Tuple[int, int, int]]:
    return [(kh + origin, kw + origin, 0) for kh in range(k) for kw in range(k)]


def pack_conv(w: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """Conv2d weight (O, C, k, k) -> (O, k*k*C), K index = (kh*k + kw)*C + c."""
    o, c, k, _ = w.shape
    return w.permute(0, 2, 3, 1).reshape(o, k * k * c).to(dtype).contiguous()


def pack_first(w: torch.Tensor) -> torch.Tensor:
    """First-layer weight (O, 3, k, k) -> tap-major fp32 (3*k*k, O) for fnst_conv_first."""
    o = w.shape[0]
    return w.detach().float().permute(1, 2, 3, 0).reshape(-1, o).contiguous()


def pack_first_tc(w: torch.Tensor, c_pad: int, dtype: torch.dtype) -> torch.Tensor:
    """First-layer weight (O, 3, k, k) -> (O, k*64): K index kh*64 + kw*c_pad + c (window of 64/c_pad pixels per
    kernel row; pixels >= k and channels >= 3 are zero)."""
    o, c, k, _ = w.shape
    b = torch.zeros((o, k, 64 // c_pad, c_pad), dtype=w.dtype, device=w.device)
    b[:, :, :k, :c] = w.permute(0, 2, 3, 1)
    return b.reshape(o, k * 64).to(dtype).contiguous()


def taps_s2d_3x3(c_in: int, cm: int = 1) -> List[Tuple[int, int, int]]:
    """3x3 stride-2 conv on a space-to-depth halo buffer: tap (kh,kw) reads spatial offset
    (kh>>1, kw>>1) of phase (kh&1, kw&1), i.e. channel window ((kh&1)*2 + (kw&1)) * c_in * cm
    (cm = 2 when every phase holds a [hi | lo] pair)."""
    return [(kh >> 1, kw >> 1, ((kh & 1) * 2 + (kw & 1)) * c_in * cm) for kh in range(3) for kw in range(3)]


# ---- fp16x3 helpers ------------------------------------------------------------------------------------------------

def split_hi_lo(w: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    hi = w.float().to(torch.float16)
    return hi, (w.float() - hi.float()).to(torch.float16)


def taps_x3(taps: Sequence[Tuple[int, int, int]], lo_offset: int) -> List[Tuple[int, int, int]]:
    """Each tap becomes (A_hi x B_hi), (A_hi x B_lo), (A_lo x B_hi): the third reads the lo channel window."""
    out = []
    for dh, dw, c0 in taps:
        out += [(dh, dw, c0), (dh, dw, c0), (dh, dw, c0 + lo_offset)]
    return out


def pack_x3(packed_f32: torch.Tensor, ntaps: int, kc: int) -> torch.Tensor:
    """fp32 operand [n, ntaps*kc] -> fp16 operand [n, ntaps*3*kc] with per-tap blocks (hi, lo, hi)."""
    n = packed_f32.shape[0]
    hi, lo = split_hi_lo(packed_f32)
    h, l = hi.view(n, ntaps, 1, kc), lo.view(n, ntaps, 1, kc)
    return torch.cat([h, l, h], dim=2).reshape(n, ntaps * 3 * kc).contiguous()


def pack_first_x3(w: torch.Tensor) -> torch.Tensor:
    """conv1 (O,3,9,9) for the split 8-channel image [hi0..2, 0, lo0..2, 0]: taps (kh, window 0/1, block 0/1), 8-pixel
    windows; block 0 multiplies hi and lo channels by w_hi, block 1 multiplies the hi channels by w_lo."""
    o, c, k, _ = w.shape
    hi, lo = split_hi_lo(w)
    b = torch.zeros((o, k, 2, 2, 8, 8), dtype=torch.float16, device=w.device)       # (o, kh, win, blk, px, ch)
    for kw in range(k):
        win, px = divmod(kw, 8)
        b[:, :, win, 0, px, 0:3] = hi[:, :, :, kw].permute(0, 2, 1)
        b[:, :, win, 0, px, 4:7] = hi[:, :, :, kw].permute(0, 2, 1)
        b[:, :, win, 1, px, 0:3] = lo[:, :, :, kw].permute(0, 2, 1)
    return b.reshape(o, k * 4 * 64).contiguous()


def pack_final_rowsum_x3(w: torch.Tensor) -> torch.Tensor:
    """final_conv (3,32,9,9) for split activations (one pixel = [hi32 | lo32] = one 128-byte row): 18 taps (kw, block),
    GEMM column kh*3 + o; block 0 = [w_hi | w_hi], block 1 = [w_lo | 0]."""
    hi, lo = split_hi_lo(w)
    b = torch.zeros((9, 3, 9, 2, 2, 32), dtype=torch.float16, device=w.device)       # (kh, o, kw, blk, half, c)
    hp, lp = hi.permute(2, 0, 3, 1), lo.permute(2, 0, 3, 1)                          # (kh, o, kw, c)
    b[:, :, :, 0, 0, :] = hp
    b[:, :, :, 0, 1, :] = hp
    b[:, :, :, 1, 0, :] = lp
    out = torch.zeros((32, 18 * 64), dtype=torch.float16, device=w.device)
    out[:27] = b.reshape(27, 18 * 64)
    return out.contiguous()


TAPS_2X2 = [(0, 0, 0), (0, 1, 0), (1, 0, 0), (1, 1, 0)]
# ConvTranspose2d(k=3, s=2, p=1, op=1): output row 2i+ph takes input row i+dh with kernel row _KT[ph][dh]
_KT = {0: {0: 1}, 1: {0: 2, 1: 0}}


def pack_conv_transpose(w: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """ConvTranspose2d weight (Cin, Cout, 3, 3) -> (4*Cout, 4*Cin): row (ph*2+pw)*Cout + o,
    K index (dh*2+dw)*Cin + c; unused (phase, tap) pairs stay zero (models/model.py:13-19)."""
    cin, cout = w.shape[0], w.shape[1]
    b = torch.zeros((2, 2, cout, 2, 2, cin), dtype=w.dtype, device=w.device)
    for ph in (0, 1):
        for dh, kh in _KT[ph].items():
            for pw in (0, 1):
                for dw, kw in _KT[pw].items():
                    b[ph, pw, :, dh, dw, :] = w[:, :, kh, kw].t()
    return b.reshape(4 * cout, 4 * cin).to(dtype).contiguous()


TAPS_ROWSUM = [(0, 2 * j, 0) for j in range(5)]


def pack_final_rowsum(w: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """final_conv weight (3, 32, 9, 9) -> (32, 5*64) for the separable ROWSUM9 form: GEMM column kh*3 + o, K index
    j*64 + jj*32 + c with kw = 2j + jj (pixel-pair view); the phantom tap kw = 9 and columns 27..31 are zero."""
    o, c, k, _ = w.shape
    assert (o, c, k) == (3, 32, 9)
    b = torch.zeros((9, 3, 5, 2, 32), dtype=w.dtype, device=w.device)       # (kh, o, j, jj, c)
    for j in range(5):
        for jj in range(2):
            kw = 2 * j + jj
            if kw < 9:
                b[:, :, j, jj, :] = w[:, :, :, kw].permute(2, 0, 1)
    out = torch.zeros((32, 5 * 64), dtype=w.dtype, device=w.device)
    out[:27] = b.reshape(27, 5 * 64)
    return out.to(dtype).contiguous()


def pack_final_stream(w: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """final_conv weight (3, 32, 9, 9) -> operand image of fnst_finalconv_tc: [kw][channel half][k-chunk][row][8 channels],
    row = kh*3 + o (27 used of 32), channel = half*16 + chunk*8 + e  (K-major core matrices of 8 rows x 16 bytes)."""
    o, c, k, _ = w.shape
    assert (o, c, k) == (3, 32, 9)
    b = torch.zeros((9, 2, 2, 32, 8), dtype=w.dtype, device=w.device)        # (kw, half, chunk, row, e)
    src = w.permute(3, 2, 0, 1).reshape(9, 27, 2, 2, 8)                        # (kw, kh*3+o, half, chunk, e)
    b[:, :, :, :27, :] = src.permute(0, 2, 3, 1, 4)
    return b.reshape(-1).to(dtype).contiguous()


def pack_final_plain(w: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    o = w.shape[0]
    b = torch.zeros((16, w.shape[2] * w.shape[3] * w.shape[1]), dtype=w.dtype, device=w.device)
    b[:o] = w.permute(0, 2, 3, 1).reshape(o, -1)
    return b.to(dtype).contiguous()


# ---------------------------------------------------------------------------------------------
# StyleTransferNet
# ---------------------------------------------------------------------------------------------

# final_conv forward through the row-streaming kernel (fnst_finalconv_tc); FNST_FINAL_STREAM=0 keeps the gather-GEMM ROWSUM9 form
FINAL_STREAM = os.environ.get("FNST_FINAL_STREAM", "1") != "0"
# measured (tools/check_finalconv.py): 808 vs 2078 us at 256 x 256x256, 924 vs 1993 us at 8 x 1080x1920, 25 vs 40 us at batch 4,
# 25 vs 21 us for a single 256x256 image -> the streaming kernel from two 256x256 images' worth of pixels upwards
FINAL_STREAM_MIN_PIXELS = 2 * 256 * 256

# channels of the 14 InstanceNorm layers (models/model.py:29,32,41,44,81,83): norm1, norm2, 5 x (in1, in2), norm3, norm4
STATS_CHANNELS = 64 + 256 + 10 * 256 + 64 + 32


def _half_up(v: int) -> int:
    return (v + 1) // 2


def _nhwc_strides(t: torch.Tensor) -> Tuple[int, int, int]:
    n, h, w, c = t.shape
    return (h * w * c, w * c, c)


class StyleNetPlan:
    """Packed weights + forward for one parameter set.  `params` maps the reference state-dict
    names (models/model.py:25-47) to CUDA tensors; re-pack (cheap) after every optimizer step."""

    def __init__(self, precision: str = "fp16"):
        self.precision = precision
        self.dtype = act_dtype(precision)
        self.use_tc = precision != "fp32"
        self.split = precision == "fp16x3"
        self.w: Dict[str, torch.Tensor] = {}
        self.params: Dict[str, torch.Tensor] = {}

    # -- weights ---------------------------------------------------------------------------------
    def _pack_x3(self, p) -> Dict[str, torch.Tensor]:
        f32 = torch.float32
        w = {"conv1": pack_first_x3(p["conv1.conv.weight"]),
             "conv2": pack_x3(pack_conv(p["conv2.conv.weight"], f32), 9, 64)}
        for i in range(5):
            w[f"res{i}a"] = pack_x3(pack_conv(p[f"res_blocks.{i}.conv1.conv.weight"], f32), 9, 256)
            w[f"res{i}b"] = pack_x3(pack_conv(p[f"res_blocks.{i}.conv2.conv.weight"], f32), 9, 256)
        w["up1"] = pack_x3(pack_conv_transpose(p["up1.upsample_conv.weight"], f32), 4, 256)
        w["up2"] = pack_x3(pack_conv_transpose(p["up2.upsample_conv.weight"], f32), 4, 64)
        w["final"] = pack_final_rowsum_x3(p["final_conv.conv.weight"])
        return w

    def pack(self, params: Dict[str, torch.Tensor], for_backward: bool = False) -> "StyleNetPlan":
        """for_backward: also pack the data-gradient operands (plan.wd) -- issued on a side stream that forward(tape=...)
        joins at its end, so these small re-layout kernels run next to the forward instead of in front of the backward."""
        p = {k: v.detach() for k, v in params.items()}
        self.params = p
        self.wd = None
        self._wd_stream = None
        dt = self.dtype
        if self.split:
            self.w = self._pack_x3(p)
            self.final_bias = torch.zeros(16, dtype=torch.float32, device=self.w["final"].device)
            self.final_bias[:3] = p["final_conv.conv.bias"].float()
            self._pack_for_backward(for_backward)
            return self
        # irregular re-layouts run as one gather kernel each (ops.gather_pack: cached index map of the layout function)
        gp, f64 = ops.gather_pack, torch.float64
        w = {"conv1": gp("first_tc4", lambda t: pack_first_tc(t, 4, f64), p["conv1.conv.weight"], dt) if self.use_tc
             else pack_first(p["conv1.conv.weight"]),
             "conv2": gp("conv", lambda t: pack_conv(t, f64), p["conv2.conv.weight"], dt)}
        # the ten 3x3 256->256 weights are packed by two kernels (stack, permute+cast) into one (10, 256, 2304) tensor
        names = [f"res_blocks.{i}.{c}.conv.weight" for i in range(5) for c in ("conv1", "conv2")]
        dev = p[names[0]].device
        res_all = torch.empty((10, 256, 9 * 256), dtype=dt, device=dev)
        # Training re-packs every step (the optimizer just changed the weights): only conv1 / conv2 are needed at once, so the
        # trunk / decoder weights (~25 us of re-layout kernels) are packed on a side branch that forward() joins in front of
        # the first residual block -- they run under conv1 / norm1 / conv2 / norm2 instead of in front of them.
        side = None
        if for_backward and dev.type == "cuda":
            main = torch.cuda.current_stream(dev)
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(main)
        with (torch.cuda.stream(side) if side is not None else contextlib.nullcontext()):
            stacked = torch.stack([p[n] for n in names])                          # (10, O, C, 3, 3)
            res_all.view(10, 256, 3, 3, 256).copy_(stacked.permute(0, 1, 3, 4, 2))
            w["res_all"] = res_all
            for i in range(5):
                w[f"res{i}a"], w[f"res{i}b"] = res_all[2 * i], res_all[2 * i + 1]
            w["up1"] = gp("convT", lambda t: pack_conv_transpose(t, f64), p["up1.upsample_conv.weight"], dt)
            w["up2"] = gp("convT", lambda t: pack_conv_transpose(t, f64), p["up2.upsample_conv.weight"], dt)
            if self.use_tc and FINAL_STREAM:
                w["final_stream"] = gp("final_stream", lambda t: pack_final_stream(t, f64), p["final_conv.conv.weight"], dt)
            w["final"] = (gp("final_rowsum", lambda t: pack_final_rowsum(t, f64), p["final_conv.conv.weight"], dt) if self.use_tc
                          else gp("final_plain", lambda t: pack_final_plain(t, f64), p["final_conv.conv.weight"], dt))
            self.final_bias = torch.zeros(16, dtype=torch.float32, device=w["final"].device)
            self.final_bias[:3] = p["final_conv.conv.bias"].float()
        self._wpack_stream = side
        self.w = w
        self._pack_for_backward(for_backward)
        return self

    def _join_weight_pack(self, dev) -> None:
        side = getattr(self, "_wpack_stream", None)
        if side is not None:
            torch.cuda.current_stream(dev).wait_stream(side)
            self._wpack_stream = None

    def _pack_for_backward(self, for_backward: bool) -> None:
        if for_backward and self.w["final"].is_cuda:
            from . import backward
            dev = self.w["final"].device
            main = torch.cuda.current_stream(dev)
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                self.wd = backward.pack_dgrad_operands(self)
            self._wd_stream = side

    def _affine(self, name: str) -> Tuple[torch.Tensor, torch.Tensor]:
        return self.params[name + ".weight"].float().contiguous(), self.params[name + ".bias"].float().contiguous()

    # -- forward ---------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor, drop_scales: Optional[Sequence[torch.Tensor]] = None,
                tape: Optional[dict] = None) -> torch.Tensor:
        """x: (B,3,H,W) fp32 CUDA.  Returns (B,3,H',W') fp32, H' = 4*ceil(ceil(H/2)/2).
        Conv biases in front of an InstanceNorm are mathematically cancelled by the mean
        subtraction and are skipped; final_conv's bias is applied."""
        assert x.dim() == 4 and x.shape[1] == 3
        x = x.contiguous().float()
        B, _, H, W = x.shape
        dev, dt, tc = x.device, self.dtype, self.use_tc
        w = self.w
        new = lambda *s: torch.empty(s, dtype=dt, device=dev)
        # training on the fp16 path: every saved activation is also written as bfloat16 by the kernel that produces it (the
        # weight-gradient GEMM multiplies it with a bf16 gradient and kind::f16 MMAs take one 16-bit format)
        twin = tape is not None and tc and dt != torch.bfloat16
        new2 = (lambda *s: torch.empty(s, dtype=torch.bfloat16, device=dev)) if twin else (lambda *s: None)
        arena = ops.ZeroArena(B * 2 * STATS_CHANNELS, dev)          # all InstanceNorm statistics: one memset per forward
        stats = lambda c: arena.take(B, c, 2)
        if H <= 4 or W <= 4:
            raise RuntimeError("StyleTransferNet needs H, W >= 5 (ReflectionPad2d(4))")
        if tc:
            ops.require_tensor_cores(dev)
        if self.split:
            return self._forward_x3(x, drop_scales, tape)

        # conv1: 9x9 stride 2, reflect 4 -> raw1 (B,H1,W1,64)
        H1, W1 = _half_up(H), _half_up(W)
        H2, W2 = _half_up(H1), _half_up(W1)
        H4, W4 = 4 * H2, 4 * W2
        # every zero fill of the forward is issued here, ahead of the first kernel, so that the kernels below follow each
        # other directly on the stream (programmatic dependent launch chains them; a fill in between would not)
        Hp, Wp = H1 + 2, W1 + 2
        Hs, Ws = _half_up(Hp), _half_up(Wp)
        # (the space-to-depth buffer is written completely by inorm_apply when the padded extent is even; with an odd
        # extent the last half-filled row / column of phases must read as zero)
        buf2 = (torch.empty if Hp % 2 == 0 and Wp % 2 == 0 else torch.zeros)((B, Hs, Ws, 256), dtype=dt, device=dev)
        buf2_b = (torch.empty if Hp % 2 == 0 and Wp % 2 == 0 else torch.zeros)((B, Hs, Ws, 256), dtype=torch.bfloat16, device=dev) if twin else None
        Hq, Wq = H4 + 8, W4 + 8
        flat = torch.empty(B * Hq * Wq * 32 + 128, dtype=dt, device=dev)   # slack: paired / 4-pixel window views
        flat[-128:].zero_()
        flat_b = None
        if twin:
            flat_b = torch.empty(B * Hq * Wq * 32 + 128, dtype=torch.bfloat16, device=dev)
            flat_b[-128:].zero_()
        raw1, st1 = new(B, H1, W1, 64), stats(64)
        if tc:
            # tensor-core form: 4-channel reflect-halo image; kernel row kh = tap (row pair kh>>1, row parity via the
            # channel offset), the 9 kernel columns x 4 channels sit inside one 64-element window (16 pixels)
            rows, pitch = 2 * (H1 + 4), (W + 8 + 1) // 2 * 2
            img = ops.image_to_halo(x, 4, PAD_REFLECT, 4, rows, pitch, dt)
            taps = [(kh >> 1, 0, (kh & 1) * pitch * 4) for kh in range(9)]
            ops.conv_gather(ConvSpec(taps, 64, w["conv1"], 64, 64), img, (B, H1 + 4, W1, pitch * 4 + 64),
                            (rows * pitch * 4, 2 * pitch * 4, 8), raw1, (H1, W1), st1, True, stats_zeroed=True)
        else:
            ops.conv_first(x, w["conv1"], None, 9, 2, 4, PAD_REFLECT, False, raw1, st1)
        # norm1 + relu -> space-to-depth halo buffer for the stride-2 conv2
        g, b = self._affine("norm1")
        ops.inorm_apply(raw1, st1, g, b, buf2, relu=True, pad=1, pad_mode=PAD_REFLECT, s2d=True, out2=buf2_b)
        # conv2: 3x3 stride 2 as a 9-tap gather over the s2d buffer
        raw2, st2 = new(B, H2, W2, 256), stats(256)
        spec = ConvSpec(taps_s2d_3x3(64), 64, w["conv2"], 256, 256)
        ops.conv_gather(spec, buf2, (B, Hs, Ws, 256), _nhwc_strides(buf2), raw2, (H2, W2), st2, tc, stats_zeroed=True)
        cur, cur_b = new(B, H2 + 2, W2 + 2, 256), new2(B, H2 + 2, W2 + 2, 256)
        g, b = self._affine("norm2")
        ops.inorm_apply(raw2, st2, g, b, cur, relu=True, pad=1, pad_mode=PAD_REFLECT, out2=cur_b)
        if tape is not None:
            tape.update(raw1=raw1, st1=st1, buf2=buf2, raw2=raw2, st2=st2, trunk=[cur])
            # bf16 twins (None when the activations already are bf16 / fp32): operands of the weight-gradient GEMMs
            tape["w"] = dict(buf2=buf2_b, trunk=[cur_b], mid=[], act4_flat=flat_b)

        # residual trunk
        self._join_weight_pack(dev)          # trunk / decoder weights packed on the side branch of pack(for_backward=True)
        taps9 = taps_kxk(3)
        for i in range(5):
            raw_a, st_a = new(B, H2, W2, 256), stats(256)
            ops.conv_gather(ConvSpec(taps9, 256, w[f"res{i}a"], 256, 256, tag=f"res{i}a"), cur, (B, H2 + 2, W2 + 2, 256),
                            _nhwc_strides(cur), raw_a, (H2, W2), st_a, tc, stats_zeroed=True)
            mid, mid_b = new(B, H2 + 2, W2 + 2, 256), new2(B, H2 + 2, W2 + 2, 256)
            g, b = self._affine(f"res_blocks.{i}.in1")
            drop = None if drop_scales is None else drop_scales[i].float().contiguous()
            ops.inorm_apply(raw_a, st_a, g, b, mid, relu=True, pad=1, pad_mode=PAD_REFLECT, drop=drop, out2=mid_b)
            raw_b, st_b = new(B, H2, W2, 256), stats(256)
            ops.conv_gather(ConvSpec(taps9, 256, w[f"res{i}b"], 256, 256, tag=f"res{i}b"), mid, (B, H2 + 2, W2 + 2, 256),
                            _nhwc_strides(mid), raw_b, (H2, W2), st_b, tc, stats_zeroed=True)
            last = i == 4
            nxt = new(B, H2, W2, 256) if last else new(B, H2 + 2, W2 + 2, 256)
            nxt_b = new2(B, H2, W2, 256) if last else new2(B, H2 + 2, W2 + 2, 256)
            g, b = self._affine(f"res_blocks.{i}.in2")
            ops.inorm_apply(raw_b, st_b, g, b, nxt, relu=False, pad=0 if last else 1,
                            pad_mode=PAD_NONE if last else PAD_REFLECT, res=cur, res_pad=1, out2=nxt_b)
            if tape is not None:
                tape.setdefault("blocks", []).append(dict(raw_a=raw_a, st_a=st_a, mid=mid, raw_b=raw_b, st_b=st_b, drop=drop))
                tape["trunk"].append(nxt)
                tape["w"]["trunk"].append(nxt_b)
                tape["w"]["mid"].append(mid_b)
            cur = nxt

        # up1: ConvTranspose2d(256->64) = 2x2-tap gather, 256 columns, depth-to-space
        H3, W3 = 2 * H2, 2 * W2
        raw3, st3 = new(B, H3, W3, 64), stats(64)
        ops.conv_gather(ConvSpec(TAPS_2X2, 256, w["up1"], 256, 64, epilogue=EPI_D2S), cur, (B, H2, W2, 256),
                        _nhwc_strides(cur), raw3, (H2, W2), st3, tc, stats_zeroed=True)
        act3, act3_b = new(B, H3, W3, 64), new2(B, H3, W3, 64)
        g, b = self._affine("norm3")
        ops.inorm_apply(raw3, st3, g, b, act3, relu=True, out2=act3_b)
        # up2: ConvTranspose2d(64->32)
        raw4, st4 = new(B, H4, W4, 32), stats(32)
        ops.conv_gather(ConvSpec(TAPS_2X2, 64, w["up2"], 128, 32, epilogue=EPI_D2S), act3, (B, H3, W3, 64),
                        _nhwc_strides(act3), raw4, (H3, W3), st4, tc, stats_zeroed=True)
        # norm4 + relu -> reflect-4 halo buffer (+ slack so the paired view of the last pixel stays in bounds)
        act4 = flat[:B * Hq * Wq * 32].view(B, Hq, Wq, 32)
        g, b = self._affine("norm4")
        ops.inorm_apply(raw4, st4, g, b, act4, relu=True, pad=4, pad_mode=PAD_REFLECT, out2=flat_b)
        # final_conv 9x9 -> NCHW fp32
        y = torch.empty((B, 3, H4, W4), dtype=torch.float32, device=dev)
        if tc and FINAL_STREAM and B * H4 * W4 >= FINAL_STREAM_MIN_PIXELS:
            # row-streaming kernel: every input row is staged in shared memory once, the 9 horizontal taps are address shifts
            ops.finalconv_stream(flat, B, H4, W4, w["final_stream"], self.final_bias, y)
        elif tc:
            # separable form: 5 pixel-pair taps along w, the 9 kernel rows live in the GEMM columns (ROWSUM9 epilogue)
            spec = ConvSpec(TAPS_ROWSUM, 64, w["final"], 32, 3, epilogue=EPI_ROWSUM9, bias=self.final_bias)
            ops.conv_gather(spec, act4, (B, Hq, Wq, 64), (Hq * Wq * 32, Wq * 32, 32), y, (H4, W4), None, True)
        else:
            spec = ConvSpec(taps_kxk(9), 32, w["final"], 16, 3, epilogue=EPI_NCHW_F32, bias=self.final_bias)
            ops.conv_gather(spec, act4, (B, Hq, Wq, 32), _nhwc_strides(act4), y, (H4, W4), None, False)
        if tape is not None:
            tape.update(raw3=raw3, st3=st3, act3=act3, raw4=raw4, st4=st4, act4=act4, act4_flat=flat, x=x)
            tape["w"]["act3"] = act3_b
        if getattr(self, "_wd_stream", None) is not None:
            torch.cuda.current_stream(dev).wait_stream(self._wd_stream)      # join the data-gradient operand packing branch
            self._wd_stream = None
        return y


    def _forward_x3(self, x: torch.Tensor, drop_scales, tape: Optional[dict] = None) -> torch.Tensor:
        """fp16x3 forward: same operator sequence; activations are fp16 [hi | lo] pairs (2C channels per pixel), raw conv
        outputs are fp32, every gather-GEMM runs three virtual taps per real tap (hi*hi, hi*lo, lo*hi).

        With a tape (training): the fp32-class forward is what the gradients need -- ReLU / InstanceNorm make the gradient a
        discontinuous function of the forward values, and rounding ANY forward operand to 16 bits moves the early layers'
        gradients by ~5e-2 (tools/exp_grad_rounding_points.py), whereas bf16 rounding in the backward GEMMs costs ~1e-2.  So
        the tape holds the fp32 raw outputs (masks, statistics) and, beside each split activation, a plain bfloat16 twin
        written by the same inorm_apply launch: the operand of the bf16 weight-gradient GEMM.  The backward is the ordinary
        bf16 tensor-core backward (backward.stylenet_backward_core)."""
        B, _, H, W = x.shape
        dev, w = x.device, self.w
        twin = tape is not None
        new2 = (lambda *s: torch.empty(s, dtype=torch.bfloat16, device=dev)) if twin else (lambda *s: None)
        act = lambda *s: torch.empty(s, dtype=torch.float16, device=dev)
        raw = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
        arena = ops.ZeroArena(B * 2 * STATS_CHANNELS, dev)          # all InstanceNorm statistics: one memset per forward
        stats = lambda c: arena.take(B, c, 2)

        def conv(taps, kc, lo_off, weight, n_gemm, c_out, a, a_dims, out, out_hw, st, epilogue=EPI_NHWC, strides=None, bias=None):
            spec = ConvSpec(taps_x3(taps, lo_off), kc, weight, n_gemm, c_out, epilogue=epilogue, bias=bias)
            ops.conv_gather(spec, a, a_dims, strides or _nhwc_strides(a), out, out_hw, st, True, stats_zeroed=st is not None)

        # conv1: split 8-channel image (hi0..2,0,lo0..2,0); 8-pixel windows, two windows per kernel row, two weight blocks
        H1, W1 = _half_up(H), _half_up(W)
        rows, pitch = 2 * (H1 + 4), (W + 16 + 1) // 2 * 2
        img = ops.image_to_halo(x, 4, PAD_REFLECT, 8, rows, pitch, torch.float16, split=True)
        taps = [(kh >> 1, win * 4, (kh & 1) * pitch * 8) for kh in range(9) for win in (0, 1) for _blk in (0, 1)]
        raw1, st1 = raw(B, H1, W1, 64), stats(64)
        ops.conv_gather(ConvSpec(taps, 64, w["conv1"], 64, 64), img, (B, H1 + 4, W1 + 4, pitch * 8 + 64),
                        (rows * pitch * 8, 2 * pitch * 8, 16), raw1, (H1, W1), st1, True, stats_zeroed=True)
        Hp, Wp = H1 + 2, W1 + 2
        Hs, Ws = _half_up(Hp), _half_up(Wp)
        buf2 = torch.zeros((B, Hs, Ws, 512), dtype=torch.float16, device=dev)
        buf2_b = torch.zeros((B, Hs, Ws, 256), dtype=torch.bfloat16, device=dev) if twin else None
        g, b = self._affine("norm1")
        ops.inorm_apply(raw1, st1, g, b, buf2, relu=True, pad=1, pad_mode=PAD_REFLECT, s2d=True, split=True, out2=buf2_b)
        H2, W2 = _half_up(H1), _half_up(W1)
        raw2, st2 = raw(B, H2, W2, 256), stats(256)
        conv(taps_s2d_3x3(64, 2), 64, 64, w["conv2"], 256, 256, buf2, (B, Hs, Ws, 512), raw2, (H2, W2), st2)
        cur, cur_b = act(B, H2 + 2, W2 + 2, 512), new2(B, H2 + 2, W2 + 2, 256)
        g, b = self._affine("norm2")
        ops.inorm_apply(raw2, st2, g, b, cur, relu=True, pad=1, pad_mode=PAD_REFLECT, split=True, out2=cur_b)
        if twin:
            tape.update(raw1=raw1, st1=st1, buf2=buf2, raw2=raw2, st2=st2, trunk=[cur], blocks=[])
            tape["w"] = dict(buf2=buf2_b, trunk=[cur_b], mid=[])
        taps9 = taps_kxk(3)
        for i in range(5):
            raw_a, st_a = raw(B, H2, W2, 256), stats(256)
            conv(taps9, 256, 256, w[f"res{i}a"], 256, 256, cur, (B, H2 + 2, W2 + 2, 512), raw_a, (H2, W2), st_a)
            mid, mid_b = act(B, H2 + 2, W2 + 2, 512), new2(B, H2 + 2, W2 + 2, 256)
            g, b = self._affine(f"res_blocks.{i}.in1")
            drop = None if drop_scales is None else drop_scales[i].float().contiguous()
            ops.inorm_apply(raw_a, st_a, g, b, mid, relu=True, pad=1, pad_mode=PAD_REFLECT, drop=drop, split=True, out2=mid_b)
            raw_b, st_b = raw(B, H2, W2, 256), stats(256)
            conv(taps9, 256, 256, w[f"res{i}b"], 256, 256, mid, (B, H2 + 2, W2 + 2, 512), raw_b, (H2, W2), st_b)
            last = i == 4
            nxt = act(B, H2, W2, 512) if last else act(B, H2 + 2, W2 + 2, 512)
            nxt_b = new2(B, H2, W2, 256) if last else new2(B, H2 + 2, W2 + 2, 256)
            g, b = self._affine(f"res_blocks.{i}.in2")
            ops.inorm_apply(raw_b, st_b, g, b, nxt, relu=False, pad=0 if last else 1,
                            pad_mode=PAD_NONE if last else PAD_REFLECT, res=cur, res_pad=1, split=True, out2=nxt_b)
            if twin:
                tape["blocks"].append(dict(raw_a=raw_a, st_a=st_a, mid=mid, raw_b=raw_b, st_b=st_b, drop=drop))
                tape["trunk"].append(nxt)
                tape["w"]["trunk"].append(nxt_b)
                tape["w"]["mid"].append(mid_b)
            cur = nxt
        H3, W3 = 2 * H2, 2 * W2
        raw3, st3 = raw(B, H3, W3, 64), stats(64)
        conv(TAPS_2X2, 256, 256, w["up1"], 256, 64, cur, (B, H2, W2, 512), raw3, (H2, W2), st3, epilogue=EPI_D2S)
        act3, act3_b = act(B, H3, W3, 128), new2(B, H3, W3, 64)
        g, b = self._affine("norm3")
        ops.inorm_apply(raw3, st3, g, b, act3, relu=True, split=True, out2=act3_b)
        H4, W4 = 2 * H3, 2 * W3
        raw4, st4 = raw(B, H4, W4, 32), stats(32)
        conv(TAPS_2X2, 64, 64, w["up2"], 128, 32, act3, (B, H3, W3, 128), raw4, (H3, W3), st4, epilogue=EPI_D2S)
        Hq, Wq = H4 + 8, W4 + 8
        act4 = act(B, Hq, Wq, 64)                           # one pixel = [hi32 | lo32] = one 128-byte row
        flat_b = None
        if twin:                                            # plain 32-channel bf16 copy (+ slack for the pixel-pair window view)
            flat_b = torch.empty(B * Hq * Wq * 32 + 128, dtype=torch.bfloat16, device=dev)
            flat_b[-128:].zero_()
        g, b = self._affine("norm4")
        ops.inorm_apply(raw4, st4, g, b, act4, relu=True, pad=4, pad_mode=PAD_REFLECT, split=True, out2=flat_b)
        y = torch.empty((B, 3, H4, W4), dtype=torch.float32, device=dev)
        taps = [(0, kw, 0) for kw in range(9) for _blk in (0, 1)]
        ops.conv_gather(ConvSpec(taps, 64, w["final"], 32, 3, epilogue=EPI_ROWSUM9, bias=self.final_bias), act4,
                        (B, Hq, Wq, 64), _nhwc_strides(act4), y, (H4, W4), None, True)
        if twin:
            tape.update(raw3=raw3, st3=st3, act3=act3, raw4=raw4, st4=st4, act4=act4, act4_flat=flat_b, x=x)
            tape["w"].update(act3=act3_b, act4_flat=flat_b)
        if getattr(self, "_wd_stream", None) is not None:
            torch.cuda.current_stream(dev).wait_stream(self._wd_stream)      # join the data-gradient operand packing branch
            self._wd_stream = None
        return y


# ---------------------------------------------------------------------------------------------
# VGG-19 features[0:25]
# ---------------------------------------------------------------------------------------------

VGG_LAYERS = ("slice1.0", "slice1.2", "slice2.5", "slice2.7", "slice3.10", "slice3.12", "slice3.14",
              "slice4.16", "slice4.19", "slice4.21", "slice5.23")


class VGGPlan:
    """Frozen VGG-19 feature stack.  forward() returns the five NHWC feature maps
    [relu1_2, relu2_2, relu3_3, relu4_2, relu4_3] (models/vgg19_net.py:56-65; element 3 is
    observed post-ReLU because torchvision's ReLUs are in-place)."""

    def __init__(self, precision: str = "fp16"):
        self.precision = precision
        self.dtype = act_dtype(precision)
        self.use_tc = precision != "fp32"
        self._out_buffers = None
        self.w: Dict[str, torch.Tensor] = {}
        self.b: Dict[str, torch.Tensor] = {}

    def pack(self, params: Dict[str, torch.Tensor]) -> "VGGPlan":
        self.params = {k: v.detach() for k, v in params.items()}
        self.derived: Dict[str, torch.Tensor] = {}      # operands derived from the (frozen) weights, e.g. data-gradient forms
        for name in VGG_LAYERS:
            wt = params[name + ".weight"].detach()
            self.b[name] = params[name + ".bias"].detach().float().contiguous()
            if name == "slice1.0":
                self.w[name] = pack_first_tc(wt, 8, self.dtype) if self.use_tc else pack_first(wt)
            else:
                self.w[name] = pack_conv(wt, self.dtype)
        return self

    def _conv(self, name: str, a: torch.Tensor, tape: Optional[dict]) -> torch.Tensor:
        B, H, W, C = a.shape
        cout = self.w[name].shape[0]
        out = (self._out_buffers or {}).get(name)           # caller-provided output (the five feature maps in one flat buffer)
        if out is None:
            out = torch.empty((B, H, W, cout), dtype=self.dtype, device=a.device)
        assert tuple(out.shape) == (B, H, W, cout) and out.dtype == self.dtype and out.is_contiguous()
        spec = ConvSpec(taps_kxk(3), C, self.w[name], cout, cout, h0=-1, w0=-1, bias=self.b[name], relu=True)
        ops.conv_gather(spec, a, (B, H, W, C), _nhwc_strides(a), out, (H, W), None, self.use_tc)
        if tape is not None:
            tape[name] = (a, out)
        return out

    FEATURE_LAYERS = ("slice1.2", "slice2.7", "slice3.14", "slice4.21", "slice5.23")

    @staticmethod
    def feature_shapes(B: int, H: int, W: int):
        """NHWC shapes of the five returned feature maps for a (B,3,H,W) input."""
        return [(B, H, W, 64), (B, H // 2, W // 2, 128), (B, H // 4, W // 4, 256), (B, H // 8, W // 8, 512), (B, H // 8, W // 8, 512)]

    def forward(self, x: torch.Tensor, tape: Optional[dict] = None, out_buffers: Optional[Dict[str, torch.Tensor]] = None) -> List[torch.Tensor]:
        """out_buffers: optional {layer name in FEATURE_LAYERS: preallocated NHWC output} (e.g. views of one flat buffer)."""
        assert x.dim() == 4 and x.shape[1] == 3
        x = x.contiguous().float()
        B, _, H, W = x.shape
        self._out_buffers = out_buffers
        if H < 8 or W < 8:
            raise RuntimeError("VGG19 feature stack needs H, W >= 8 (three 2x2 max-pools)")
        if self.use_tc:
            ops.require_tensor_cores(x.device)
        h = torch.empty((B, H, W, 64), dtype=self.dtype, device=x.device)
        if self.use_tc:
            # conv1_1 on tensor cores: 8-channel zero-halo image, 3 taps (kernel rows), 8-pixel windows
            rows, pitch = H + 2, W + 2
            img = ops.image_to_halo(x, 1, PAD_ZERO, 8, rows, pitch, self.dtype)
            spec = ConvSpec([(kh, 0, 0) for kh in range(3)], 64, self.w["slice1.0"], 64, 64, bias=self.b["slice1.0"], relu=True)
            ops.conv_gather(spec, img, (B, rows, W, 64), (rows * pitch * 8, pitch * 8, 8), h, (H, W), None, True)
        else:
            ops.conv_first(x, self.w["slice1.0"], self.b["slice1.0"], 3, 1, 1, PAD_ZERO, True, h, None)
        if tape is not None:
            tape["x"] = x
            tape["slice1.0"] = (x, h)
        f0 = self._conv("slice1.2", h, tape)
        h = self._conv("slice2.7", self._conv("slice2.5", ops.maxpool2(f0), tape), tape)
        f1 = h
        h = ops.maxpool2(f1)
        h = self._conv("slice3.14", self._conv("slice3.12", self._conv("slice3.10", h, tape), tape), tape)
        f2 = h
        h = self._conv("slice4.16", f2, tape)
        h = self._conv("slice4.21", self._conv("slice4.19", ops.maxpool2(h), tape), tape)
        f3 = h
        f4 = self._conv("slice5.23", f3, tape)
        return [f0, f1, f2, f3, f4]
