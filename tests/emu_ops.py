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


def conv_gather(spec, a, a_dims, a_strides, out, out_hw, stats, use_tc, stats_zeroed=False, linear=False):
    n, ah, aw, ac = a_dims
    sn, sh, sw = a_strides
    if linear:
        # pixel-stream form (fnst.h FNST_DESC_LINEAR): one row of aw pixels, tap (dh, dw) = linear shift dh * pitch + dw
        assert n == 1 and ah == 1 and out_hw[0] == 1 and spec.epilogue == EPI_NHWC and stats is None
        pitch = sh // sw
        view = a.as_strided((aw, ac), (sw, 1), a.storage_offset()).double()
        acc = torch.zeros((out_hw[1], spec.n_gemm), dtype=torch.float64)
        for t, (dh, dw, c0) in enumerate(spec.taps):
            q = torch.arange(out_hw[1]) + dh * pitch + dw
            ok = ((q >= 0) & (q < aw)).double().view(-1, 1)
            acc += (view[q.clamp(0, aw - 1)][:, c0:c0 + spec.kc] * ok) @ spec.weight.double()[:, t * spec.kc:(t + 1) * spec.kc].t()
        out.view(-1, spec.c_out).copy_(acc[:, :spec.c_out].to(out.dtype))
        return
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
    if getattr(spec, "addend", None) is not None:          # dgrad epilogue extras: out = (acc + addend) * (mask > 0)
        v = v + spec.addend.double()
    if getattr(spec, "mask", None) is not None:
        v = v * (spec.mask.double() > 0)
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
                res_pad=0, eps=1e-5, split=False, out2=None):
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

    def geometry(t):                                       # halo + space-to-depth, any channel count
        cc = t.shape[-1]
        if pad:
            if pad_mode == PAD_REFLECT:
                hi = _reflect(torch.arange(-pad, h + pad), h)
                wi = _reflect(torch.arange(-pad, w + pad), w)
                t = t[:, hi][:, :, wi]
            else:
                t = F.pad(t, (0, 0, pad, pad, pad, pad))
        if s2d:
            hp, wp = t.shape[1], t.shape[2]
            t = F.pad(t, (0, 0, 0, wp % 2, 0, hp % 2))
            hs, ws = t.shape[1] // 2, t.shape[2] // 2
            t = t.view(n, hs, 2, ws, 2, cc).permute(0, 1, 3, 2, 4, 5).reshape(n, hs, ws, 4 * cc)
        return t

    if out2 is not None:                                   # bf16 twin: same geometry, c channels per pixel, never split
        t2 = geometry(y)
        out2.view(-1)[:t2.numel()].copy_(t2.reshape(-1).to(out2.dtype))
    if split:                                              # [hi | lo] fp16 pair per pixel
        hi = y.to(out.dtype).double()
        y = torch.cat([hi, y - hi], dim=-1)
    out.copy_(geometry(y).to(out.dtype))


def maxpool2(x):
    n, h, w, c = x.shape
    v = x[:, :h // 2 * 2, :w // 2 * 2].reshape(n, h // 2, 2, w // 2, 2, c)
    return v.amax(dim=(2, 4))


def gram(feat_nhwc, use_tc, out=None):
    n, h, w, c = feat_nhwc.shape
    f = feat_nhwc.reshape(n, h * w, c).double()
    res = (f.transpose(1, 2) @ f).float()
    if out is None:
        return res
    out += res                                             # split-K partials accumulate into the zeroed tensor
    return out


def sse(a, b, acc):
    d = a.double().reshape(-1, b.numel()) - b.double().reshape(1, -1)
    acc += (d * d).sum()


def sse_scaled(a, b, scale, out, accumulate=False):
    d = a.double().reshape(-1, b.numel()) - b.double().reshape(1, -1)
    r = (scale * (d * d).sum()).float()
    out.copy_(out + r if accumulate else r)


def tv_scaled(img, scale, out):
    acc = torch.zeros((), dtype=torch.float64)
    tv(img, acc)
    out.copy_((scale * acc).float())


def tv(img, acc):
    acc += ((img[:, :, 1:] - img[:, :, :-1]).double() ** 2).sum() + ((img[:, :, :, 1:] - img[:, :, :, :-1]).double() ** 2).sum()


def nhwc_to_nchw(x):
    return x.permute(0, 3, 1, 2).float().contiguous()


def nchw_to_nhwc(x, dtype, c_pad=None):
    v = x.permute(0, 2, 3, 1)
    if c_pad is not None and c_pad > v.shape[-1]:          # zero channels up to c_pad
        v = F.pad(v, (0, c_pad - v.shape[-1]))
    return v.to(dtype).contiguous()


def finalconv_stream(act_flat, n, h, w, wpacked, bias, out):
    """fnst_finalconv_tc: decode the operand image back to OIHW and convolve the halo buffer (valid 9x9)."""
    act = act_flat.reshape(-1)[: n * (h + 8) * (w + 8) * 32].view(n, h + 8, w + 8, 32).double()
    b = wpacked.view(9, 2, 2, 32, 8).double()                                   # (kw, half, chunk, row, e)
    wt = b[:, :, :, :27, :].permute(3, 1, 2, 4, 0).reshape(9, 3, 32, 9).permute(1, 2, 0, 3)   # (o, c, kh, kw)
    out.copy_(F.conv2d(act.permute(0, 3, 1, 2), wt, bias[:3].double()))


def gather_pack(key, layout_fn, src, out_dtype, out=None):
    res = layout_fn(src.detach().double()).to(out_dtype)
    if out is None:
        return res
    out.view(-1).copy_(res.reshape(-1))
    return out


def gather_index(src, idx, out):
    flat = src.reshape(-1)
    i = idx.long()
    out.view(-1).copy_(torch.where(i < 0, torch.zeros((), dtype=flat.dtype), flat[i.clamp_min(0)]).to(out.dtype))


def require_tensor_cores(device):
    return None


def install(monkeypatch, ops_module):
    """Substitute every operator of `ops_module` by its emulation."""
    for name in ("require_tensor_cores", "conv_gather", "conv_first", "image_to_halo", "inorm_apply", "maxpool2", "gram", "sse", "tv", "sse_scaled", "tv_scaled", "nhwc_to_nchw", "nchw_to_nhwc", "gather_pack", "gather_index", "finalconv_stream"):
        monkeypatch.setattr(ops_module, name, globals()[name])


# ---------------------------------------------------------------------------------------------------------
# Backward operators (include/fnst.h, "Backward operators" section)
# ---------------------------------------------------------------------------------------------------------

def wgrad(spec, a, a_dims, a_strides, g, out_hw, use_tc=False, g_strides=None, out=None, out_zeroed=False):
    """dB[j][t*kc+c] = sum_{n,h,w} g[n,h,w,j] * A[n, h+h0+dh[t], w+w0+dw[t], c0[t]+c] (A reads as zero out of range)."""
    n, ah, aw, ac = a_dims
    sn, sh, sw = a_strides
    oh, ow = out_hw
    view = a.as_strided((n, ah, aw, ac), (sn, sh, sw, 1), a.storage_offset()).double()
    if g_strides is None:
        gv = g.double()
    else:
        gv = g.as_strided((n, oh, ow, spec.n_gemm), (g_strides[0], g_strides[1], g_strides[2], 1), g.storage_offset()).double()
    blocks = []
    for dh, dw, c0 in spec.taps:
        hs = torch.arange(oh) + spec.h0 + dh
        ws = torch.arange(ow) + spec.w0 + dw
        hm = ((hs >= 0) & (hs < ah)).double().view(1, -1, 1, 1)
        wm = ((ws >= 0) & (ws < aw)).double().view(1, 1, -1, 1)
        patch = view[:, hs.clamp(0, ah - 1)][:, :, ws.clamp(0, aw - 1)][..., c0:c0 + spec.kc] * hm * wm
        blocks.append(torch.einsum("nhwj,nhwc->jc", gv, patch))
    db = torch.cat(blocks, dim=1).float()
    if out is None:
        return db
    if out_zeroed:
        out += db.view_as(out)
    else:
        out.copy_(db.view_as(out))
    return out


def conv_first_wgrad(x, g, k, stride, pad, pad_mode, out=None):
    res = _conv_first_wgrad(x, g, k, stride, pad, pad_mode)
    if out is None:
        return res
    out.view(-1).copy_(res.reshape(-1))
    return out


def _conv_first_wgrad(x, g, k, stride, pad, pad_mode):
    xp = F.pad(x.double(), (pad,) * 4, mode="reflect" if pad_mode == PAD_REFLECT else "constant")
    n, c, _, _ = x.shape
    _, ho, wo, co = g.shape
    cols = F.unfold(xp, k, stride=stride).view(n, c * k * k, ho * wo)            # rows ordered (c, kh, kw) = tap-major
    return torch.einsum("nkp,npo->ko", cols, g.double().reshape(n, ho * wo, co)).float()


def _norm_consts(raw, stats, eps):
    n, h, w, c = raw.shape
    cnt = h * w
    mean = stats[:, :, 0].double() / cnt
    var = (stats[:, :, 1].double() / cnt - mean * mean).clamp_min(0)
    rstd = 1.0 / torch.sqrt(var + eps)
    xhat = (raw.double() - mean.view(n, 1, 1, c)) * rstd.view(n, 1, 1, c)
    return mean, rstd, xhat


def inorm_bwd_fused_parts(raw, gdtype, has_gsrc=True, has_extra=False, s2d=False):
    return 1 if raw.shape[-1] % 16 == 0 and raw.shape[1] * raw.shape[2] <= 64 * 64 and not s2d else 0     # exercise both paths on CPU


def inorm_bwd_fused(gsrc, extra, raw, stats, gamma, beta, drop, gdtype, relu, pad=0, pad_mode=PAD_NONE, s2d=False,
                    out_s2d=False, want_gy=False, sums=None, eps=1e-5):
    tmp = torch.zeros(raw.shape[0], raw.shape[-1], 2, dtype=torch.float32)
    gy, tmp = inorm_bwd_reduce(gsrc, extra, raw, stats, gamma, beta, drop, gdtype, relu, pad, pad_mode, s2d, eps, sums=tmp)
    d, _ = inorm_bwd_apply(gy, raw, stats, tmp, gamma, out_s2d, eps)
    if sums is None:
        sums = tmp
    else:
        sums.copy_(tmp)                                    # written, not accumulated
    return d, (gy if want_gy else None), sums


def affine_grads(sums_flat, entries, n, out):
    for src, c, dg, db in entries:
        blk = sums_flat[src:src + n * c * 2].view(n, c, 2).double()
        out[dg:dg + c] = blk[:, :, 1].sum(0).float()
        out[db:db + c] = blk[:, :, 0].sum(0).float()


def inorm_bwd_reduce(gsrc, extra, raw, stats, gamma, beta, drop, gdtype, relu, pad=0, pad_mode=PAD_NONE, s2d=False, eps=1e-5,
                     arena=None, sums=None, gsrc_slack=0):
    n, h, w, c = raw.shape
    g = torch.zeros((n, h, w, c), dtype=torch.float64)
    if gsrc is not None:
        hp, wp = h + 2 * pad, w + 2 * pad
        gs = gsrc.double()
        if s2d:                                      # undo the space-to-depth layout written by inorm_apply
            hs, ws = (hp + 1) // 2 + gsrc_slack, (wp + 1) // 2 + gsrc_slack
            gs = gs.view(n, hs, ws, 2, 2, c).permute(0, 1, 3, 2, 4, 5).reshape(n, 2 * hs, 2 * ws, c)[:, :hp, :wp]
        else:
            gs = gs.view(n, hp + gsrc_slack, wp + gsrc_slack, c)[:, :hp, :wp]
        if pad and pad_mode == PAD_REFLECT:          # ReflectionPad2d backward: every halo position adds into its source
            hi = _reflect(torch.arange(-pad, h + pad), h)
            wi = _reflect(torch.arange(-pad, w + pad), w)
            tmp = torch.zeros((n, h, wp, c), dtype=torch.float64).index_add_(1, hi, gs)
            g = torch.zeros((n, h, w, c), dtype=torch.float64).index_add_(2, wi, tmp)
        else:
            g = gs[:, pad:pad + h, pad:pad + w].clone()
    if extra is not None:
        g = g + extra.double()
    _, _, xhat = _norm_consts(raw, stats, eps)
    if drop is not None:
        g = g * drop.double().view(n, 1, 1, c)
    if relu:
        y = xhat * gamma.double() + beta.double()
        g = g * (y > 0)
    if sums is None:
        sums = arena.take(n, c, 2) if arena is not None else torch.zeros((n, c, 2), dtype=torch.float32)
    gy = g.to(gdtype)
    sums[:, :, 0] += g.sum(dim=(1, 2)).float()
    sums[:, :, 1] += (g * xhat).sum(dim=(1, 2)).float()
    return gy, sums


def inorm_bwd_apply(gy, raw, stats, sums, gamma, out_s2d=False, eps=1e-5, want_dgb=True, out_pad=0):
    n, h, w, c = raw.shape
    cnt = h * w
    _, rstd, xhat = _norm_consts(raw, stats, eps)
    m1 = (sums[:, :, 0].double() / cnt).view(n, 1, 1, c)
    m2 = (sums[:, :, 1].double() / cnt).view(n, 1, 1, c)
    d = gamma.double() * rstd.view(n, 1, 1, c) * (gy.double() - m1 - xhat * m2)
    if out_s2d:
        d = d.view(n, h // 2, 2, w // 2, 2, c).permute(0, 1, 3, 2, 4, 5).reshape(n, h // 2, w // 2, 4 * c)
    if out_pad:
        d = F.pad(d, (0, 0, out_pad, out_pad, out_pad, out_pad))
    dgb = torch.stack([sums[:, :, 1].sum(0), sums[:, :, 0].sum(0)]).float()
    return d.to(gy.dtype), dgb


def maxpool2_bwd(inp, gout, extra):
    n, h, w, c = inp.shape
    x = inp.double().permute(0, 3, 1, 2)
    _, idx = F.max_pool2d(x, 2, 2, return_indices=True)                         # first maximal element (PyTorch tie rule)
    route = F.max_unpool2d(gout.double().permute(0, 3, 1, 2), idx, 2, 2, output_size=(h, w)).permute(0, 2, 3, 1)
    if extra is not None:
        route = route + extra.double()
    return (route * (inp.double() > 0)).to(gout.dtype)


def relu_mask(g, extra, act):
    v = g.double() if extra is None else g.double() + extra.double()
    return (v * (act.double() > 0)).to(g.dtype)


def sse_bwd(a, b, scale, gdtype, relu_mask=False, coef=1.0, out=None):
    d = 2.0 * coef * scale.double()[0] * (a.double().reshape(-1, b.numel()) - b.double().reshape(1, -1))
    d = d.view(a.shape)
    if relu_mask:
        d = d * (a.double() > 0)
    if out is not None:
        out.copy_(d.to(gdtype))
        return out
    return d.to(gdtype)


def tv_bwd(img, scale, coef=1.0):
    x = img.double()
    d = torch.zeros_like(x)
    dh = x[:, :, 1:] - x[:, :, :-1]
    dw = x[:, :, :, 1:] - x[:, :, :, :-1]
    d[:, :, 1:] += 2 * dh
    d[:, :, :-1] -= 2 * dh
    d[:, :, :, 1:] += 2 * dw
    d[:, :, :, :-1] -= 2 * dw
    return (coef * scale.double()[0] * d).float()


def channel_sum(x, out=None):
    res = x.double().sum(dim=(0, 2, 3)).float()
    if out is None:
        return res
    out.copy_(res)
    return out


def cast(x, dtype):
    return x.to(dtype)


def gram_diff_sym(g, gt, scale, coef, dtype):
    d = g.double() - gt.double().reshape(-1, g.shape[1], g.shape[2])
    return (scale.double()[0] * coef * (d + d.transpose(1, 2))).to(dtype)


def install_backward(monkeypatch, ops_module):
    """Substitute the backward operators too, and make the stream plumbing of backward.py a no-op on CPU."""
    install(monkeypatch, ops_module)
    for name in ("wgrad", "conv_first_wgrad", "inorm_bwd_reduce", "inorm_bwd_apply", "inorm_bwd_fused", "inorm_bwd_fused_parts",
                 "affine_grads", "maxpool2_bwd", "relu_mask", "sse_bwd", "tv_bwd", "channel_sum", "cast", "gram_diff_sym"):
        monkeypatch.setattr(ops_module, name, globals()[name])
    monkeypatch.setenv("FNST_WGRAD_STREAM", "0")                   # no side stream for the weight-gradient branch
    monkeypatch.setattr(torch.cuda, "current_stream", lambda device=None: None)
