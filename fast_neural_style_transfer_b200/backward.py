"""Backward plans: gradients of StyleTransferNet (58 parameters) and of the frozen VGG-19 feature
stack (data gradient only), plus the loss-reduction backward helpers, in libfnst operators.

Every data gradient of a convolution is again a gather-GEMM (ops.conv_gather) on the output
gradient with negated taps and transposed packed weights; weight gradients use ops.wgrad; the
InstanceNorm / ReLU / Dropout2d / residual / ReflectionPad2d backward is fused in
ops.inorm_bwd_reduce + ops.inorm_bwd_apply (reference autograd: train.py:200).

Gradient element type: fp32 on the "fp32" path; bfloat16 on the tensor-core paths (gradient norms
reach 1e7-5e8 at random init, SURVEY 7.2, which overflows fp16).
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence

import torch

from . import engine, ops
from ._lib import EPI_NHWC, EPI_NCHW_F32, PAD_NONE, PAD_REFLECT, PAD_ZERO
from .engine import TAPS_2X2, _KT, _nhwc_strides, taps_kxk, taps_s2d_3x3
from .ops import ConvSpec


def grad_dtype(precision: str) -> torch.dtype:
    return torch.float32 if precision == "fp32" else torch.bfloat16


def grad_dtype_of(act: torch.dtype) -> torch.dtype:
    """Gradient element type for activations of type `act`: fp32 stays fp32; both 16-bit activation types carry bf16
    gradients (fp16's range is too small for the un-normalised Gram / style gradients)."""
    return torch.float32 if act == torch.float32 else torch.bfloat16


def _neg(taps):
    return [(-dh, -dw, 0) for dh, dw, _ in taps]


def pack_dgrad(bf: torch.Tensor, ntaps: int, kc: int, dtype: torch.dtype) -> torch.Tensor:
    """Forward operand [n_gemm, ntaps*kc] -> data-gradient operand [kc, ntaps*n_gemm] (same tap order)."""
    n_gemm = bf.shape[0]
    return bf.view(n_gemm, ntaps, kc).permute(2, 1, 0).reshape(kc, ntaps * n_gemm).to(dtype).contiguous()


def pack_dgrad_s2d(w: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """3x3 stride-2 conv weight (O, C, 3, 3) -> data-gradient operand for the space-to-depth input:
    rows (ph_h, ph_w, c), K = (dh, dw, o); kernel tap (kh, kw) = (2*dh + ph_h, 2*dw + ph_w) when <= 2."""
    o, c = w.shape[0], w.shape[1]
    b = torch.zeros((2, 2, c, 2, 2, o), dtype=w.dtype, device=w.device)
    for kh in range(3):
        for kw in range(3):
            b[kh & 1, kw & 1, :, kh >> 1, kw >> 1, :] = w[:, :, kh, kw].t()
    return b.reshape(4 * c, 4 * o).to(dtype).contiguous()


def unpack_conv(db: torch.Tensor, o: int, c: int, k: int) -> torch.Tensor:
    return db[:o].view(o, k, k, c).permute(0, 3, 1, 2).contiguous()


def unpack_conv_transpose(db: torch.Tensor, cin: int, cout: int) -> torch.Tensor:
    """Inverse of engine.pack_conv_transpose for gradients: (4*Cout, 4*Cin) -> (Cin, Cout, 3, 3)."""
    b = db.view(2, 2, cout, 2, 2, cin)
    dw = torch.zeros((cin, cout, 3, 3), dtype=db.dtype, device=db.device)
    for ph in (0, 1):
        for dh, kh in _KT[ph].items():
            for pw in (0, 1):
                for dwi, kw in _KT[pw].items():
                    dw[:, :, kh, kw] = b[ph, pw, :, dh, dwi, :].t()
    return dw


def _tc_ok(use_tc: bool, kc: int, n_gemm: int) -> bool:
    return use_tc and kc % 64 == 0 and (n_gemm == 16 or n_gemm % 32 == 0)


# -------------------------------------------------------------------------------------------------------
# StyleTransferNet backward
# -------------------------------------------------------------------------------------------------------
#
# Gradient plumbing.  Every weight-gradient GEMM writes its packed result into a slice of ONE zeroed fp32 staging buffer;
# every InstanceNorm backward writes its per-plane sums into a slice of ONE sums buffer.  Two launches then assemble the
# 58 gradients in the reference's parameter layouts inside ONE flat fp32 buffer (`assemble_gradients`): a gather through
# a cached index map (weights: packed -> OIHW / IOHW; conv biases in front of an InstanceNorm: zero) and a fixed-order
# reduction over the images for d gamma / d beta.  The per-parameter gradients are views of that buffer (so the
# data-parallel all-reduce and the optimizer tail see one contiguous bucket).

# FNST_INORM_BWD_FUSED=1 selects the one-pass InstanceNorm backward (TMA-staged cluster kernel) wherever the plane fits.
# Default off -- measured on B200 at batch 4 (profiles/r02_inorm_bwd_fused.md): alone the one-pass kernel takes 16.6 us
# against 14.5 + 7.3 us for the two passes, but inside the training step it LOSES 0.14 ms: its CTAs need a whole SM
# (~200 KB of shared memory) in co-scheduled clusters, so they cannot share SMs with the weight-gradient GEMMs of the
# side stream the way the small two-pass kernels do.
FUSED_INORM_BWD = os.environ.get("FNST_INORM_BWD_FUSED", "0") not in ("", "0")

# Pixel-stream data gradients (fnst.h FNST_DESC_LINEAR) for the residual trunk: the InstanceNorm backward writes d_raw into a
# buffer with a 2-pixel ZERO halo, and the data gradient of the 3x3 convolution behind ReflectionPad2d(1) runs over the linear
# pixel stream of that buffer -- 145 full tiles at 4 x 64 x 64 (one wave on 148 SMs; 21 us) instead of 180 ragged 8 x 16 boxes
# on the 66 x 66 domain (two waves; 35.5 us).  FNST_LINEAR_DGRAD=0 restores the box form.
LINEAR_DGRAD = os.environ.get("FNST_LINEAR_DGRAD", "1") not in ("", "0")

NORM_LAYERS = (["norm1", "norm2"] + [f"res_blocks.{i}.{n}" for i in range(5) for n in ("in1", "in2")] + ["norm3", "norm4"])


def _staging_layout(tc: bool):
    """name -> (element offset, shape) of the packed weight gradients inside the staging buffer; total elements."""
    items = [("res", (10, 256, 2304)), ("up1", (256, 1024)), ("up2", (128, 256)), ("conv2", (256, 576)),
             ("conv1", (64, 576) if tc else (243, 64)), ("final", (128, 576) if tc else (16, 81 * 32)), ("final_bias", (3,))]
    off, out = 0, {}
    for name, shape in items:
        n = 1
        for d in shape:
            n *= d
        out[name] = (off, shape)
        off += (n + 3) // 4 * 4
    return out, off


_FINAL_KW = torch.arange(8, -1, -1)          # final_conv weight-gradient window: kw -> pixel index i = 8 - kw


def _unpack_fns(tc: bool):
    """parameter name -> (staging slice name, slice selector, layout function packed -> parameter layout)."""
    fns = {}
    for k in range(10):
        name = f"res_blocks.{k // 2}.conv{k % 2 + 1}.conv.weight"
        fns[name] = ("res", k, lambda t: t.view(256, 3, 3, 256).permute(0, 3, 1, 2))
    fns["up1.upsample_conv.weight"] = ("up1", None, lambda t: unpack_conv_transpose(t, 256, 64))
    fns["up2.upsample_conv.weight"] = ("up2", None, lambda t: unpack_conv_transpose(t, 64, 32))
    fns["conv2.conv.weight"] = ("conv2", None, lambda t: unpack_conv(t, 256, 64, 3))
    if tc:
        fns["conv1.conv.weight"] = ("conv1", None, lambda t: t.view(64, 9, 16, 4)[:, :, :9, :3].permute(0, 3, 1, 2))
        # (i, j, kh, jj, c) -> (j, c, kh, kw), kw = 8 - i from the jj = 0 entries
        fns["final_conv.conv.weight"] = ("final", None, lambda t: t.view(16, 8, 9, 2, 32)[_FINAL_KW, :3, :, 0, :].permute(1, 3, 2, 0))
    else:
        fns["conv1.conv.weight"] = ("conv1", None, lambda t: t.view(3, 9, 9, 64).permute(3, 0, 1, 2))
        fns["final_conv.conv.weight"] = ("final", None, lambda t: unpack_conv(t, 3, 32, 9))
    fns["final_conv.conv.bias"] = ("final_bias", None, lambda t: t)
    return fns


_ASSEMBLY = {}


def _assembly(names: Sequence[str], shapes: Sequence[torch.Size], tc: bool, device) -> dict:
    """Cached per (parameter list, path, device): flat offsets, the int32 gather map staging -> flat (-1 = zero)."""
    key = (tuple(names), tc, str(device))
    ent = _ASSEMBLY.get(key)
    if ent is not None:
        return ent
    layout, _ = _staging_layout(tc)
    fns = _unpack_fns(tc)
    offsets, off = {}, 0
    pieces = []
    for name, shape in zip(names, shapes):
        numel = 1
        for d in shape:
            numel *= d
        offsets[name] = off
        off += numel
        if name in fns:
            slot, sel, fn = fns[name]
            base, sshape = layout[slot]
            cnt = 1
            for d in sshape:
                cnt *= d
            probe = (torch.arange(cnt, dtype=torch.float64) + (base + 1)).reshape(sshape)      # element number + 1 (0 = zero fill)
            probe = probe if sel is None else probe[sel]
            res = fn(probe)
            assert tuple(res.shape) == tuple(shape), (name, res.shape, shape)
            pieces.append((res.reshape(-1).round().to(torch.int64) - 1).to(torch.int32))
        else:
            pieces.append(torch.full((numel,), -1, dtype=torch.int32))       # zero (conv bias under InstanceNorm) or written by affine_grads
    ent = _ASSEMBLY[key] = dict(offsets=offsets, total=off, idx=torch.cat(pieces).to(device))
    return ent


def assemble_gradients(core: dict, names: Sequence[str], params: Dict[str, torch.Tensor], flat: Optional[torch.Tensor] = None,
                       first: Optional[str] = None, last: Optional[str] = None) -> torch.Tensor:
    """Flat fp32 gradient of all parameters (in `names` order) from the staging / sums buffers of `stylenet_backward_core`.
    first / last: assemble only the contiguous run of parameters from `first` up to (not including) `last` into the given
    `flat` (bucket-wise assembly while later stages of the backward still run; `bucket_bounds` names the cuts)."""
    asm = _assembly(names, [params[n].shape for n in names], core["tc"], core["staging"].device)
    if flat is None:
        flat = torch.empty(asm["total"], dtype=torch.float32, device=core["staging"].device)
    lo = 0 if first is None else asm["offsets"][first]
    hi = asm["total"] if last is None else asm["offsets"][last]
    ops.gather_index(core["staging"], asm["idx"][lo:hi], flat[lo:hi])
    entries = [(src, c, asm["offsets"][ln + ".weight"], asm["offsets"][ln + ".bias"]) for ln, (src, c) in core["norm_sums"].items()
               if lo <= asm["offsets"][ln + ".weight"] < hi]
    if entries:
        ops.affine_grads(core["sums"], entries, core["batch"], flat)
    return flat


# Stage cuts of the backward for bucket-wise gradient exchange (parallel.GradientAllReduce): after residual block 2 every
# gradient from `res_blocks.2` to the end of the parameter list is final (60 % of the bytes), after block 0 those of blocks 0
# and 1; conv1 / conv2 and their norms (2.6 %) come with the last stage.  Parameter order = module order (models/model.py:27-47).
STAGE_CUTS = (2, 0)


def bucket_bounds(stage: int):
    """(first, last) parameter names delimiting the gradients completed by stage 0, 1, 2 of the staged backward."""
    return [("res_blocks.2.conv1.conv.weight", None), ("res_blocks.0.conv1.conv.weight", "res_blocks.2.conv1.conv.weight"),
            (None, "res_blocks.0.conv1.conv.weight")][stage]


def _final_dgrad_layout(wt):                                              # (3, 32, 9, 9) -> (64, 9*64)
    """final_conv data-gradient operand for the PIXEL-PAIR form: one GEMM row computes the two neighbouring pixels (2m, 2m+1)
    of d_act4 (64 columns = parity x 32 channels = exactly two NHWC pixels) from ONE 128-byte window of the 4-channel
    zero-halo gradient image, 16 pixels x 4 channels starting at pixel 2m: window pixel i holds dy column u - kw + 8 for
    kw = parity + 8 - i.  Nine taps (kernel rows) instead of the 18 of the 8-channel single-pixel form: half the executed
    FLOPs and a quarter of the L2 -> shared-memory traffic."""
    wd = torch.zeros((2, 32, 9, 16, 4), dtype=wt.dtype)               # (parity, c, kh, i, j)
    wperm = wt.permute(1, 2, 3, 0)                                    # (c, kh, kw, j)
    for par in (0, 1):
        for i in range(16):
            kw = par + 8 - i
            if 0 <= kw <= 8:
                wd[par, :, :, i, :3] = wperm[:, :, kw, :]
    return wd.reshape(64, 9 * 64)


def pack_dgrad_operands(plan: "engine.StyleNetPlan") -> Dict[str, torch.Tensor]:
    """Data-gradient forms of all weights (packed / transposed per layer, gradient dtype): one gather kernel each.  They only
    depend on the parameters, so the training forward issues them on a side stream next to its own kernels."""
    p, tc = plan.params, plan.use_tc
    gdt = grad_dtype(plan.precision)
    gp, f64 = ops.gather_pack, torch.float64
    convT_dgrad = lambda kc: (lambda t: pack_dgrad(engine.pack_conv_transpose(t, f64), 4, kc, f64))
    # the ten 3x3 weights as one stacked tensor (10, O, C, 3, 3) -> (10, C, 9*O): one gather for all data-gradient operands
    stacked = torch.stack([p[f"res_blocks.{i}.{c}.conv.weight"] for i in range(5) for c in ("conv1", "conv2")])
    wd = {
        "res": gp("res_dgrad", lambda t: t.permute(0, 2, 3, 4, 1).reshape(10, 256, 9 * 256), stacked, gdt),
        "up1": gp("convT_dgrad", convT_dgrad(256), p["up1.upsample_conv.weight"], gdt),
        "up2": gp("convT_dgrad", convT_dgrad(64), p["up2.upsample_conv.weight"], gdt),
        "conv2": gp("s2d_dgrad", lambda t: pack_dgrad_s2d(t, f64), p["conv2.conv.weight"], gdt),
    }
    wfin = p["final_conv.conv.weight"]
    if tc:
        wd["final"] = gp("final_dgrad", _final_dgrad_layout, wfin, gdt)
    else:
        wd["final"] = gp("final_plain_dgrad", lambda t: pack_dgrad(engine.pack_final_plain(t, f64), 81, 32, f64), wfin, gdt)
    return wd


def stylenet_backward(plan: "engine.StyleNetPlan", tape: dict, dy: torch.Tensor) -> Dict[str, torch.Tensor]:
    """dy: (B,3,H',W') fp32.  Returns gradients for all 58 reference parameter names (views of one flat fp32 buffer)."""
    names = list(plan.params)
    flat = assemble_gradients(stylenet_backward_core(plan, tape, dy), names, plan.params)
    return dict(zip(names, [t.view_as(plan.params[n]) for t, n in zip(torch.split(flat, [plan.params[n].numel() for n in names]), names)]))


def stylenet_backward_core(plan: "engine.StyleNetPlan", tape: dict, dy: torch.Tensor) -> dict:
    """Everything of the backward pass except the final assembly: returns the staging buffer (packed weight gradients), the
    InstanceNorm sums buffer and where each norm layer's sums live -- static tensors when captured in a CUDA graph."""
    gen = stylenet_backward_stages(plan, tape, dy, cuts=())
    try:
        while True:
            next(gen)
    except StopIteration as done:
        return done.value


def stylenet_backward_stages(plan: "engine.StyleNetPlan", tape: dict, dy: torch.Tensor, cuts: Sequence[int] = STAGE_CUTS):
    """Generator form of the backward: yields the (partially filled) result dict after the residual blocks listed in `cuts`
    (weight-gradient side stream joined at every yield, so each stage can be captured as its own CUDA graph), returns it at
    the end.  Between two stages the caller may assemble and exchange the gradients that are already final."""
    p = plan.params
    gdt = grad_dtype(plan.precision)
    tc = plan.use_tc
    dev = dy.device
    dy = dy.contiguous().float()
    B, _, H4, W4 = dy.shape
    layout, staging_total = _staging_layout(tc)
    # split-K targets of every weight gradient: ONE memset.  On the tensor-core path every writer of this buffer runs on the
    # weight-gradient side branch, so the 25 MB fill is issued there too (first side block) instead of in front of the first
    # data gradient.
    side_zero = tc and os.environ.get("FNST_WGRAD_STREAM", "1") != "0"
    staging = (torch.empty if side_zero else torch.zeros)(staging_total, dtype=torch.float32, device=dev)

    def slot(name):
        off, shape = layout[name]
        n = 1
        for d in shape:
            n *= d
        return staging[off:off + n].view(shape)

    # per-plane sums of the 14 InstanceNorm backward passes, [B,C,2] each (zeroed once: the two-pass fallback accumulates)
    arena = ops.ZeroArena(2 * B * engine.STATS_CHANNELS + 64 * 14, dev)
    norm_sums: Dict[str, tuple] = {}
    wtw = tape.get("w") or {}                    # bf16 twins of the saved activations (None entries: use the activation itself)

    def inorm_backward(layer, gsrc, extra, raw, stats, drop, relu, pad=0, pad_mode=PAD_NONE, s2d=False, out_s2d=False, want_gy=False,
                       gsrc_slack=0, out_pad=0):
        """d_raw (and gy when asked) of one InstanceNorm layer; its sums slice is registered for the affine gradients.
        gsrc_slack / out_pad: geometry of the pixel-stream data gradients (see LINEAR_DGRAD)."""
        ga, ba = plan._affine(layer)
        src_off = arena.used
        sums = arena.take(B, raw.shape[-1], 2)
        norm_sums[layer] = (src_off, raw.shape[-1])
        if FUSED_INORM_BWD and not (gsrc_slack or out_pad) and ops.inorm_bwd_fused_parts(raw, gdt, gsrc is not None, extra is not None, s2d) > 0:
            d_raw, gy, _ = ops.inorm_bwd_fused(gsrc, extra, raw, stats, ga, ba, drop, gdt, relu, pad, pad_mode, s2d, out_s2d, want_gy, sums)
            return d_raw, gy
        gy, _ = ops.inorm_bwd_reduce(gsrc, extra, raw, stats, ga, ba, drop, gdt, relu, pad, pad_mode, s2d, sums=sums, gsrc_slack=gsrc_slack)
        d_raw, _ = ops.inorm_bwd_apply(gy, raw, stats, sums, ga, out_s2d=out_s2d, want_dgb=False, out_pad=out_pad)
        return d_raw, gy

    # Weight gradients are off the critical path (nothing downstream in this backward consumes them): they run on a
    # side stream forked from the main stream right after their operands are produced and joined at the end.  Under
    # CUDA-graph capture this becomes a parallel branch; operands are kept alive until the join so the allocator cannot
    # hand their memory to a later main-stream tensor while the side branch still reads it.
    main_stream = torch.cuda.current_stream(dev)
    side_stream = torch.cuda.Stream(device=dev) if os.environ.get("FNST_WGRAD_STREAM", "1") != "0" else None
    keep_alive = []

    class _Side:
        def __enter__(self_inner):
            if side_stream is not None:
                side_stream.wait_stream(main_stream)
                self_inner.ctx = torch.cuda.stream(side_stream)
                self_inner.ctx.__enter__()
            return self_inner

        def __exit__(self_inner, *exc):
            if side_stream is not None:
                self_inner.ctx.__exit__(*exc)
            return False

    def on_side(*tensors):
        keep_alive.extend(tensors)
        return _Side()

    def wgrad(spec, a, a_twin, a_dims, g, out_hw, out, g_pad=0):
        """Weight gradient of `spec` into a staging slice; tensor cores whenever the channel window is a multiple of 64
        (operand = the activation's bf16 twin when the forward wrote one, else a cast).  g_pad: g carries a halo of that width."""
        use_tc = tc and spec.kc % 64 == 0
        if use_tc and a.dtype != g.dtype:
            a = a_twin if a_twin is not None else ops.cast(a, g.dtype)
            keep_alive.append(a)
        g_strides = None
        if g_pad:
            g_strides = (g.stride(0), g.stride(1), g.stride(2))
            g = g[:, g_pad:g.shape[1] - g_pad, g_pad:g.shape[2] - g_pad, :]
        ops.wgrad(spec, a, a_dims, _nhwc_strides(a), g, out_hw, use_tc=use_tc, g_strides=g_strides, out=out, out_zeroed=True)

    wd_all = getattr(plan, "wd", None) or pack_dgrad_operands(plan)     # packed by the training forward (side stream) when it ran

    def dgrad(g, g_dims, wd, fwd_taps, fwd_kc, out_shape, out_hw, h0=0, w0=0):
        """Data gradient of a forward gather-GEMM with plain taps (c0 == 0); wd = its data-gradient operand
        [fwd_kc, ntaps * n_gemm_fwd] (pack_dgrad of the forward operand)."""
        n_gemm_f = wd.shape[1] // len(fwd_taps)
        spec = ConvSpec(_neg(fwd_taps), n_gemm_f, wd, fwd_kc, fwd_kc, h0=-h0, w0=-w0)
        out = torch.empty(out_shape, dtype=gdt, device=dev)
        ops.conv_gather(spec, g, g_dims, _nhwc_strides(g), out, out_hw, None, _tc_ok(tc, n_gemm_f, fwd_kc))
        return out

    # ---- final_conv (9x9, 32 -> 3) -------------------------------------------------------------------
    act4 = tape["act4"]
    Hq, Wq = act4.shape[1], act4.shape[2]
    taps81 = taps_kxk(9)
    wfin = p["final_conv.conv.weight"]
    if tc:
        # Tensor-core forms on zero-halo copies of dy (halo 8):
        #  wgrad (8-channel copy: one pixel = 16 bytes, 8 pixels = one 128-byte row): contraction over halo positions p; M side = 16-pixel dy window (128 = 16 px x 8 ch) at p, N side = act4
        #         pixel-pair window at p + (kh-8, 0): D[(i,j)][kh*64 + jj*32 + c] = dW[j][c][kh][jj+8-i]  (9 taps)
        rows_g, pitch_g = H4 + 16, W4 + 16
        g_str = (rows_g * pitch_g * 8, pitch_g * 8, 8)
        flat_act = tape["act4_flat"]
        taps9 = [(kh - 8, 0, 0) for kh in range(9)]
        with on_side(flat_act, dy, staging):
            if side_zero:
                staging.zero_()
            ops.channel_sum(dy, out=slot("final_bias"))          # d bias of final_conv: nothing downstream needs it
            g8 = ops.image_to_halo(dy, 8, PAD_ZERO, 8, rows_g, pitch_g, gdt)
            keep_alive.append(g8)
            a_g = flat_act
            if a_g.dtype != gdt:
                a_g = wtw.get("act4_flat") if wtw.get("act4_flat") is not None else ops.cast(flat_act, gdt)
            ops.wgrad(ConvSpec(taps9, 64, None, 128, 128), a_g, (B, Hq, Wq, 64), (Hq * Wq * 32, Wq * 32, 32), g8,
                      (rows_g, pitch_g), use_tc=True, g_strides=g_str, out=slot("final"), out_zeroed=True)
            keep_alive.append(a_g)

        # dgrad, pixel-pair form (see _final_dgrad_layout): 4-channel zero-halo image, windows start at every second pixel
        g4 = ops.image_to_halo(dy, 8, PAD_ZERO, 4, rows_g, pitch_g, gdt)
        d_act4 = torch.empty((B, Hq, Wq, 32), dtype=gdt, device=dev)
        ops.conv_gather(ConvSpec([(8 - kh, 0, 0) for kh in range(9)], 64, wd_all["final"], 64, 64), g4,
                        (B, rows_g, pitch_g // 2, 64), (rows_g * pitch_g * 4, pitch_g * 4, 8), d_act4.view(B, Hq, Wq // 2, 64),
                        (Hq, Wq // 2), None, True)
    else:
        ops.channel_sum(dy, out=slot("final_bias"))
        g16 = ops.nchw_to_nhwc(dy, gdt, c_pad=16)
        wgrad(ConvSpec(taps81, 32, None, 16, 3), act4, None, (B, Hq, Wq, 32), g16, (H4, W4), slot("final"))
        d_act4 = dgrad(g16, (B, H4, W4, 16), wd_all["final"], taps81, 32, (B, Hq, Wq, 32), (Hq, Wq))

    # ---- norm4 + up2 ------------------------------------------------------------------------------------
    d_raw4, _ = inorm_backward("norm4", d_act4, None, tape["raw4"], tape["st4"], None, True, 4, PAD_REFLECT, out_s2d=True)   # (B,H3,W3,128)
    act3 = tape["act3"]
    H3, W3 = act3.shape[1], act3.shape[2]
    with on_side(act3, d_raw4):
        wgrad(ConvSpec(TAPS_2X2, 64, None, 128, 32), act3, wtw.get("act3"), (B, H3, W3, 64), d_raw4, (H3, W3), slot("up2"))
    d_act3 = dgrad(d_raw4, (B, H3, W3, 128), wd_all["up2"], TAPS_2X2, 64, (B, H3, W3, 64), (H3, W3))

    # ---- norm3 + up1 ------------------------------------------------------------------------------------
    d_raw3, _ = inorm_backward("norm3", d_act3, None, tape["raw3"], tape["st3"], None, True, out_s2d=True)     # (B,H2,W2,256)
    trunk = tape["trunk"]
    trunk_w = wtw.get("trunk") or [None] * len(trunk)
    mid_w = wtw.get("mid") or [None] * 5
    last = trunk[5]
    H2, W2 = last.shape[1], last.shape[2]
    with on_side(last, d_raw3):
        wgrad(ConvSpec(TAPS_2X2, 256, None, 256, 64), last, trunk_w[5], (B, H2, W2, 256), d_raw3, (H2, W2), slot("up1"))
    g_plain = dgrad(d_raw3, (B, H2, W2, 256), wd_all["up1"], TAPS_2X2, 256, (B, H2, W2, 256), (H2, W2))

    # ---- residual trunk -----------------------------------------------------------------------------------
    taps9 = taps_kxk(3)
    pdims = (B, H2 + 2, W2 + 2, 256)
    gsrc, extra = None, g_plain          # gradient of the block output = fold(gsrc) + extra
    res_dg = wd_all["res"]                      # (10, 256, 9*256): the ten 3x3 data-gradient operands as one stacked tensor
    res_db = slot("res")

    lin = tc and LINEAR_DGRAD and not FUSED_INORM_BWD
    Z = 2 if lin else 0                         # zero halo of the trunk's d_raw buffers = slack of the gradients computed from them
    Hz, Wz = H2 + 2 * Z, W2 + 2 * Z

    def res_dgrad(g, idx):
        """Gradient w.r.t. the reflect-padded input of a trunk convolution from d_raw: (B, H2+2, W2+2, 256), or, in the
        pixel-stream form, (B, H2+4, W2+4, 256) whose [:, :H2+2, :W2+2] corner holds it (g then has a 2-pixel zero halo)."""
        if lin:
            m = B * Hz * Wz
            out = torch.empty((B, Hz, Wz, 256), dtype=gdt, device=dev)
            ops.conv_gather(ConvSpec([(Z - dh, Z - dw, 0) for dh, dw, _ in taps9], 256, res_dg[idx], 256, 256), g, (1, 1, m, 256),
                            (m * 256, Wz * 256, 256), out, (1, m), None, True, linear=True)
            return out
        out = torch.empty(pdims, dtype=gdt, device=dev)
        ops.conv_gather(ConvSpec(_neg(taps9), 256, res_dg[idx], 256, 256), g, (B, H2, W2, 256), _nhwc_strides(g), out,
                        (H2 + 2, W2 + 2), None, tc)
        return out

    for i in range(4, -1, -1):
        blk = tape["blocks"][i]
        pre = f"res_blocks.{i}"
        # in2 (no ReLU); the total output gradient also feeds the skip connection
        d_raw_b, g_out = inorm_backward(pre + ".in2", gsrc, extra, blk["raw_b"], blk["st_b"], None, False,
                                        1 if gsrc is not None else 0, PAD_REFLECT if gsrc is not None else PAD_NONE, want_gy=True,
                                        gsrc_slack=Z if gsrc is not None else 0, out_pad=Z)
        mid = blk["mid"]
        with on_side(mid, d_raw_b):
            wgrad(ConvSpec(taps9, 256, None, 256, 256, tag=f"wgrad_res{i}b"), mid, mid_w[i], pdims, d_raw_b, (H2, W2), res_db[2 * i + 1], g_pad=Z)
        d_mid = res_dgrad(d_raw_b, 2 * i + 1)
        # in1 + ReLU + Dropout2d
        d_raw_a, _ = inorm_backward(pre + ".in1", d_mid, None, blk["raw_a"], blk["st_a"], blk["drop"], True, 1, PAD_REFLECT,
                                    gsrc_slack=Z, out_pad=Z)
        cur = trunk[i]
        with on_side(cur, d_raw_a):
            wgrad(ConvSpec(taps9, 256, None, 256, 256, tag=f"wgrad_res{i}a"), cur, trunk_w[i], pdims, d_raw_a, (H2, W2), res_db[2 * i], g_pad=Z)
        gsrc = res_dgrad(d_raw_a, 2 * i)
        extra = g_out
        if i in cuts:
            if side_stream is not None:
                main_stream.wait_stream(side_stream)         # the stage's weight gradients are complete when the stage ends
            yield dict(staging=staging, sums=arena.buf, norm_sums=dict(norm_sums), batch=B, tc=tc)

    # ---- norm2 + conv2 (stride 2 on the space-to-depth buffer) ---------------------------------------------
    # conv2's data gradient lives on the (H2+1) x (W2+1) space-to-depth domain (65 x 65 at 256x256: 180 ragged boxes); same
    # pixel-stream form as the trunk with a 1-pixel zero halo around d_raw2: 4 * 66 * 66 / 128 = 137 full tiles
    Z2 = 1 if lin else 0
    d_raw2, _ = inorm_backward("norm2", gsrc, extra, tape["raw2"], tape["st2"], None, True, 1, PAD_REFLECT, gsrc_slack=Z, out_pad=Z2)
    buf2 = tape["buf2"]
    Hs, Ws = buf2.shape[1], buf2.shape[2]
    with on_side(buf2, d_raw2):
        wgrad(ConvSpec(taps_s2d_3x3(64), 64, None, 256, 256), buf2, wtw.get("buf2"), (B, Hs, Ws, 256), d_raw2, (H2, W2), slot("conv2"), g_pad=Z2)
    if lin and Hs == H2 + 1 and Ws == W2 + 1:
        m2 = B * (H2 + 2) * (W2 + 2)
        d_buf2 = torch.empty((B, H2 + 2, W2 + 2, 256), dtype=gdt, device=dev)          # [:, :Hs, :Ws] holds the gradient
        ops.conv_gather(ConvSpec([(1 - dh, 1 - dw, 0) for dh, dw, _ in TAPS_2X2], 256, wd_all["conv2"], 256, 256), d_raw2, (1, 1, m2, 256),
                        (m2 * 256, (W2 + 2) * 256, 256), d_buf2, (1, m2), None, True, linear=True)
        slack1 = 1
    else:
        g2 = d_raw2[:, Z2:Z2 + H2, Z2:Z2 + W2, :].contiguous() if Z2 else d_raw2
        d_buf2 = torch.empty((B, Hs, Ws, 256), dtype=gdt, device=dev)
        ops.conv_gather(ConvSpec(_neg(TAPS_2X2), 256, wd_all["conv2"], 256, 256), g2, (B, H2, W2, 256), _nhwc_strides(g2), d_buf2,
                        (Hs, Ws), None, tc)
        slack1 = 0

    # ---- norm1 + conv1 -----------------------------------------------------------------------------------------
    raw1 = tape["raw1"]
    d_raw1, _ = inorm_backward("norm1", d_buf2, None, raw1, tape["st1"], None, True, 1, PAD_REFLECT, s2d=True, gsrc_slack=slack1)
    if tc:
        # same window view as the forward (engine.StyleNetPlan.forward): taps = kernel rows, 16-pixel x 4-channel windows
        x = tape["x"]
        H1, W1 = raw1.shape[1], raw1.shape[2]
        rows, pitch = 2 * (H1 + 4), (x.shape[3] + 8 + 1) // 2 * 2
        img = ops.image_to_halo(x, 4, PAD_REFLECT, 4, rows, pitch, gdt)
        taps = [(kh >> 1, 0, (kh & 1) * pitch * 4) for kh in range(9)]
        with on_side(img, d_raw1):           # (on the side branch like every other writer of the staging buffer)
            ops.wgrad(ConvSpec(taps, 64, None, 64, 64), img, (B, H1 + 4, W1, pitch * 4 + 64), (rows * pitch * 4, 2 * pitch * 4, 8),
                      d_raw1, (H1, W1), use_tc=True, out=slot("conv1"), out_zeroed=True)
    else:
        ops.conv_first_wgrad(tape["x"], d_raw1, 9, 2, 4, PAD_REFLECT, out=slot("conv1"))             # tap-major (243, 64)
    if side_stream is not None:
        main_stream.wait_stream(side_stream)             # join the weight-gradient branch
    del keep_alive[:]
    return dict(staging=staging, sums=arena.buf, norm_sums=norm_sums, batch=B, tc=tc)


# -------------------------------------------------------------------------------------------------------
# VGG-19 backward (frozen weights: data gradient only)
# -------------------------------------------------------------------------------------------------------

def vgg_backward(plan: "engine.VGGPlan", tape: dict, dfeats: Sequence[Optional[torch.Tensor]]) -> torch.Tensor:
    """dfeats: gradients of the five NHWC feature maps (None where unused).  Returns dx (B,3,H,W) fp32.
    The gradient element type is independent of the activation type: bf16 on both tensor-core precisions, so the
    drop-in's default pair (fp16 net + fp16 VGG) back-propagates without overflow."""
    gdt = grad_dtype(plan.precision)
    tc = plan.use_tc
    taps9 = taps_kxk(3, origin=-1)
    dfe = [None if g is None else g.contiguous().to(gdt) for g in dfeats]
    if all(g is None for g in dfe):
        raise RuntimeError("vgg_backward called without any feature gradient")

    def dgrad(name, g, addend=None, masked=True):
        a, _ = tape[name]
        B, H, W, cin = a.shape
        cout = g.shape[-1]
        wd = plan.derived.get("dgrad." + name)                         # [cin, 9*cout]; the weights are frozen: packed once
        if wd is None:
            wd = plan.derived["dgrad." + name] = pack_dgrad(plan.w[name], 9, cin, gdt)
        out = torch.empty((B, H, W, cin), dtype=gdt, device=g.device)
        spec = ConvSpec(_neg(taps9), cout, wd, cin, cin, addend=addend, mask=a if masked else None)
        ops.conv_gather(spec, g, (B, H, W, cout), _nhwc_strides(g), out, (H, W), None, _tc_ok(tc, cout, cin))
        return out

    f4 = tape["slice5.23"][1]
    f3 = tape["slice5.23"][0]
    if dfe[4] is not None:
        g = ops.relu_mask(dfe[4], None, f4)
        g = dgrad("slice5.23", g, addend=dfe[3])
    elif dfe[3] is not None:
        g = ops.relu_mask(dfe[3], None, f3)
    else:
        g = None
    f2 = tape["slice4.16"][0]
    if g is not None:
        g = dgrad("slice4.21", g)
        g = dgrad("slice4.19", g, masked=False)
        g = ops.maxpool2_bwd(tape["slice4.16"][1], g, None)
        g = dgrad("slice4.16", g, addend=dfe[2])
    elif dfe[2] is not None:
        g = ops.relu_mask(dfe[2], None, f2)
    f1 = tape["slice2.7"][1]
    if g is not None:
        g = dgrad("slice3.14", g)
        g = dgrad("slice3.12", g)
        g = dgrad("slice3.10", g, masked=False)
        g = ops.maxpool2_bwd(f1, g, dfe[1])
    elif dfe[1] is not None:
        g = ops.relu_mask(dfe[1], None, f1)
    f0 = tape["slice1.2"][1]
    if g is not None:
        g = dgrad("slice2.7", g)
        g = dgrad("slice2.5", g, masked=False)
        g = ops.maxpool2_bwd(f0, g, dfe[0])
    else:
        g = ops.relu_mask(dfe[0], None, f0)
    g = dgrad("slice1.2", g)
    # conv1_1: 64 -> 3 image channels, NCHW fp32 output
    x = tape["x"]
    B, _, H, W = x.shape
    wd = plan.derived.get("dgrad.slice1.0")
    if wd is None:
        w0 = plan.params["slice1.0.weight"]                               # (64, 3, 3, 3)
        wd = torch.zeros((16, 9, 64), dtype=torch.float32, device=x.device)
        wd[:3] = w0.permute(1, 2, 3, 0).reshape(3, 9, 64)
        wd = plan.derived["dgrad.slice1.0"] = wd.reshape(16, 576).to(gdt).contiguous()
    dx = torch.empty((B, 3, H, W), dtype=torch.float32, device=x.device)
    spec = ConvSpec(_neg(taps9), 64, wd, 16, 3, epilogue=EPI_NCHW_F32)
    ops.conv_gather(spec, g, (B, H, W, 64), _nhwc_strides(g), dx, (H, W), None, tc)
    return dx


# -------------------------------------------------------------------------------------------------------
# Loss reductions
# -------------------------------------------------------------------------------------------------------

def gram_backward(f: torch.Tensor, dg: torch.Tensor) -> torch.Tensor:
    """f NHWC (B,H,W,C); dg (B,C,C) fp32.  dF[p,i] = sum_j F[p,j] * (dG + dG^T)[i,j]: a 1x1 gather-GEMM with per-image weights."""
    return gram_apply(f, (dg + dg.transpose(1, 2)).to(grad_dtype_of(f.dtype)).contiguous())


def gram_apply(f: torch.Tensor, s: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dF[n] = F[n] S[n] for per-image symmetric factors S (B,C,C) in the gradient dtype of the features (fp16
    features are multiplied as bf16: kind::f16 MMAs take one 16-bit format, and the products exceed fp16's range).
    out: write the result here (same shape, dtype of s)."""
    B, H, W, C = f.shape
    if f.dtype != s.dtype:
        f = ops.cast(f, s.dtype)
    if out is None:
        out = torch.empty_like(f)
    assert out.shape == f.shape and out.dtype == f.dtype and out.is_contiguous()
    use_tc = f.dtype != torch.float32 and C % 64 == 0
    spec = ConvSpec([(0, 0, 0)], C, s, C, C, per_image_weights=True)      # one launch, image n multiplies by s[n]
    ops.conv_gather(spec, f, (B, H, W, C), _nhwc_strides(f), out, (H, W), None, use_tc)
    return out


def sse_backward(a: torch.Tensor, b: torch.Tensor, g: torch.Tensor, coef: float = 1.0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    return ops.sse_bwd(a, b, g.reshape(1).float(), grad_dtype_of(a.dtype), coef=coef, out=out)


def tv_backward(x: torch.Tensor, g: torch.Tensor, coef: float = 1.0) -> torch.Tensor:
    return ops.tv_bwd(x, g.reshape(1).float(), coef=coef)
