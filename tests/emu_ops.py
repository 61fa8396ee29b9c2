"""Pure-torch CPU emulation of the libfnst operator semantics documented in include/fnst.h.

TEST INFRASTRUCTURE: lets the CPU test-suite check the host logic of the product (tap tables,
weight packing, halo / space-to-depth buffers, depth-to-space epilogue, operator order) against
the oracle without a GPU, by substituting these functions for fast_neural_style_transfer_b200.ops.
Never imported by the product.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

EPI_NHWC, EPI_D2S, EPI_NCHW_F32, EPI_ROWSUM9 = 0, 1, 2, 3
PAD_NONE, PAD_REFLECT, PAD_ZERO = 0, 1, 2


def _reflect(i: torch.Tensor, n: int) -> torch.Tensor:
    i = i.abs()
    return torch.where(i >= n, 2 * (n - 1) - i, i)


def conv_gather(spec, a, a_dims, a_strides, out, out_hw, stats, use_tc, stats_zeroed=False):
    n, ah, aw, ac = a_dims
    sn, sh, sw = a_strides
    view = a.as_strided((n, ah, aw, ac), (sn, sh, sw, 1), a.storage_offset()).double()
    oh, ow = out_hw
    wgt = spec.weight.double()
    if spec.epilogue == EPI_ROWSUM9:
        oh = oh + 8                     # the GEMM row space is the input rows (8-row halo below the output)
    acc = torch.zeros((n, oh, ow, spec.n_gemm), dtype=torch.float64)
    for t, (dh, dw, c0) in enumerate(spec.taps):
        hs = torch.arange(oh) + spec.h0 + dh
        ws = torch.arange(ow) + spec.w0 + dw
        hm = ((hs >= 0) & (hs < ah)).double().view(1, -1, 1, 1)
        wm = ((ws >= 0) & (ws < aw)).double().view(1, 1, -1, 1)
        patch = view[:, hs.clamp(0, ah - 1)][:, :, ws.clamp(0, aw - 1)][..., c0:c0 + spec.kc] * hm * wm
        if getattr(spec, "per_image_weights", False):
            acc += torch.einsum("nhwc,njc->nhwj", patch, wgt[:, :, t * spec.kc:(t + 1) * spec.kc])
        else:
            acc += patch @ wgt[:, t * spec.kc:(t + 1) * spec.kc].t()
    c = spec.c_out
    if spec.epilogue == EPI_ROWSUM9:
        oh -= 8
        v = sum(acc[:, kh:kh + oh, :, kh * c:(kh + 1) * c] for kh in range(9))
        if spec.bias is not None:
            v = v + spec.bias.double()[:c]
        out.copy_(v.permute(0, 3, 1, 2).to(out.dtype))
        return
    if spec.epilogue == EPI_D2S:
        v = acc.view(n, oh, ow, 2, 2, c).permute(0, 1, 3, 2, 4, 5).reshape(n, 2 * oh, 2 * ow, c)
    else:
        v = acc[..., :c]
    if spec.bias is not None:
        v = v + spec.bias.double()[:c]
    if spec.relu:
        v = v.clamp_min(0)
    if stats is not None:
        if not stats_zeroed:
            stats.zero_()                       # the call zeroes the accumulators unless the caller already did
        stats[:, :, 0] += v.sum(dim=(1, 2)).float()
        stats[:, :, 1] += (v * v).sum(dim=(1, 2)).float()
    if spec.epilogue == EPI_NCHW_F32:
        out.copy_(v.permute(0, 3, 1, 2).to(out.dtype))
    else:
        out.copy_(v.to(out.dtype))


def conv_first(x, weight, bias, k, stride, pad, pad_mode, relu, out, stats):
    xp = F.pad(x.double(), (pad,) * 4, mode="reflect" if pad_mode == PAD_REFLECT else "constant")
    w_oihw = weight.double().view(3, k, k, -1).permute(3, 0, 1, 2)      # tap-major (c,kh,kw,o) -> OIHW
    v = F.conv2d(xp, w_oihw, None if bias is None else bias.double(), stride=stride)
    if relu:
        v = v.clamp_min(0)
    v = v.permute(0, 2, 3, 1)
    if stats is not None:
        stats[:, :, 0] = v.sum(dim=(1, 2)).float()
        stats[:, :, 1] = (v * v).sum(dim=(1, 2)).float()
    out.copy_(v.to(out.dtype))


def image_to_halo(x, pad, pad_mode, c_pad, rows, pitch, dtype, split=False):
    n, c, h, w = x.shape
    xp = F.pad(x, (pad,) * 4, mode="reflect" if pad_mode == PAD_REFLECT else "constant")
    buf = torch.zeros((n, rows, pitch, c_pad), dtype=torch.float32)
    buf[:, :h + 2 * pad, :w + 2 * pad, :3] = xp.permute(0, 2, 3, 1)
    if split:
        hi = buf[..., :3].to(dtype).float()
        buf[..., 4:7] = buf[..., :3] - hi
        buf[..., :3] = hi
    flat = torch.zeros(buf.numel() + 128, dtype=dtype)
    flat[:buf.numel()] = buf.reshape(-1).to(dtype)
    return flat


def inorm_apply(raw, stats, gamma, beta, out, relu, pad=0, pad_mode=PAD_NONE, s2d=False, drop=None, res=None,
                res_pad=0, eps=1e-5, split=False):
    n, h, w, c = raw.shape
    cnt = h * w
    mean = stats[:, :, 0].double() / cnt
    var = (stats[:, :, 1].double() / cnt - mean * mean).clamp_min(0)
    a = gamma.double() / torch.sqrt(var + eps)
    b = beta.double() - mean * a
    y = raw.double() * a.view(n, 1, 1, c) + b.view(n, 1, 1, c)
    if relu:
        y = y.clamp_min(0)
    if drop is not None:
        y = y * drop.double().view(n, 1, 1, c)
    if res is not None:
        r = res.double()[:, res_pad:res_pad + h, res_pad:res_pad + w, :]
        y = y + (r[..., :c] + r[..., c:] if split else r)
    if split:                                              # [hi | lo] fp16 pair per pixel
        hi = y.to(out.dtype).double()
        y = torch.cat([hi, y - hi], dim=-1)
        c = 2 * c
    if pad:
        if pad_mode == PAD_REFLECT:
            hi = _reflect(torch.arange(-pad, h + pad), h)
            wi = _reflect(torch.arange(-pad, w + pad), w)
            y = y[:, hi][:, :, wi]
        else:
            y = F.pad(y, (0, 0, pad, pad, pad, pad))
    if s2d:
        hp, wp = y.shape[1], y.shape[2]
        y = F.pad(y, (0, 0, 0, wp % 2, 0, hp % 2))
        hs, ws = y.shape[1] // 2, y.shape[2] // 2
        y = y.view(n, hs, 2, ws, 2, c).permute(0, 1, 3, 2, 4, 5).reshape(n, hs, ws, 4 * c)
    out.copy_(y.to(out.dtype))


def maxpool2(x):
    n, h, w, c = x.shape
    v = x[:, :h // 2 * 2, :w // 2 * 2].reshape(n, h // 2, 2, w // 2, 2, c)
    return v.amax(dim=(2, 4))


def gram(feat_nhwc, use_tc):
    n, h, w, c = feat_nhwc.shape
    f = feat_nhwc.reshape(n, h * w, c).double()
    return (f.transpose(1, 2) @ f).float()


def sse(a, b, acc):
    d = a.double().reshape(-1, b.numel()) - b.double().reshape(1, -1)
    acc += (d * d).sum()


def tv(img, acc):
    acc += ((img[:, :, 1:] - img[:, :, :-1]).double() ** 2).sum() + ((img[:, :, :, 1:] - img[:, :, :, :-1]).double() ** 2).sum()


def nhwc_to_nchw(x):
    return x.permute(0, 3, 1, 2).float().contiguous()


def nchw_to_nhwc(x, dtype):
    return x.permute(0, 2, 3, 1).to(dtype).contiguous()


def finalconv_stream(act_flat, n, h, w, wpacked, bias, out):
    """fnst_finalconv_tc: decode the operand image back to OIHW and convolve the halo buffer (valid 9x9)."""
    act = act_flat.reshape(-1)[: n * (h + 8) * (w + 8) * 32].view(n, h + 8, w + 8, 32).double()
    b = wpacked.view(9, 2, 2, 32, 8).double()                                   # (kw, half, chunk, row, e)
    wt = b[:, :, :, :27, :].permute(3, 1, 2, 4, 0).reshape(9, 3, 32, 9).permute(1, 2, 0, 3)   # (o, c, kh, kw)
    out.copy_(F.conv2d(act.permute(0, 3, 1, 2), wt, bias[:3].double()))


def gather_pack(key, layout_fn, src, out_dtype):
    return layout_fn(src.detach().double()).to(out_dtype)


def require_tensor_cores(device):
    return None


def install(monkeypatch, ops_module):
    """Substitute every operator of `ops_module` by its emulation."""
    for name in ("require_tensor_cores", "conv_gather", "conv_first", "image_to_halo", "inorm_apply", "maxpool2", "gram", "sse", "tv", "nhwc_to_nchw", "nchw_to_nhwc", "gather_pack", "finalconv_stream"):
        monkeypatch.setattr(ops_module, name, globals()[name])
