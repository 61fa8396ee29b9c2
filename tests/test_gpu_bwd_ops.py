"""GPU parity of the backward operators (through the C ABI) against torch autograd in float64 on CPU.
Operator-level checks are tight (no chaotic mask flips through a deep chain)."""
import pytest
import torch
import torch.nn.functional as F

import emu_ops

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from fast_neural_style_transfer_b200 import backward, engine, ops
    from fast_neural_style_transfer_b200.ops import ConvSpec
    from fast_neural_style_transfer_b200 import _lib

DEV = "cuda"


def rel_l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _emu_conv(spec, a, a_dims, out_shape, out_hw):
    """Differentiable float64 CPU evaluation of a gather-GEMM (NHWC / D2S epilogues)."""
    n, ah, aw, ac = a_dims
    oh, ow = out_hw
    acc = torch.zeros((n, oh, ow, spec.n_gemm), dtype=torch.float64)
    for t, (dh, dw, c0) in enumerate(spec.taps):
        hs = torch.arange(oh) + spec.h0 + dh
        ws = torch.arange(ow) + spec.w0 + dw
        hm = ((hs >= 0) & (hs < ah)).double().view(1, -1, 1, 1)
        wm = ((ws >= 0) & (ws < aw)).double().view(1, 1, -1, 1)
        patch = a[:, hs.clamp(0, ah - 1)][:, :, ws.clamp(0, aw - 1)][..., c0:c0 + spec.kc] * hm * wm
        acc = acc + patch @ spec.weight[:, t * spec.kc:(t + 1) * spec.kc].t()
    return acc


@pytest.mark.parametrize("kind", ["res3x3", "s2d", "convT", "final"])
@pytest.mark.parametrize("gdtype", [torch.float32, torch.bfloat16])
def test_wgrad_and_dgrad(kind, gdtype):
    g = torch.Generator().manual_seed(21)
    rnd = lambda *s: torch.randn(*s, generator=g)
    adt = torch.float32 if gdtype == torch.float32 else torch.float16
    B = 2
    if kind == "res3x3":
        H, W, cin, cout = 12, 10, 64, 128
        a = rnd(B, H + 2, W + 2, cin); wt = rnd(cout, cin, 3, 3) / 24
        taps, kc, n_gemm, packed = engine.taps_kxk(3), cin, cout, engine.pack_conv(wt, torch.float64)
        a_dims, ohw = (B, H + 2, W + 2, cin), (H, W)
    elif kind == "s2d":
        H, W, cin, cout = 9, 11, 64, 128
        hs, ws = (H + 3) // 2, (W + 3) // 2
        a = rnd(B, hs, ws, 4 * cin); wt = rnd(cout, cin, 3, 3) / 24
        taps, kc, n_gemm, packed = engine.taps_s2d_3x3(cin), cin, cout, engine.pack_conv(wt, torch.float64)
        a_dims, ohw = (B, hs, ws, 4 * cin), ((H + 1) // 2, (W + 1) // 2)
    elif kind == "convT":
        H, W, cin, cout = 7, 9, 64, 32
        a = rnd(B, H, W, cin); wt = rnd(cin, cout, 3, 3) / 24
        taps, kc, n_gemm, packed = engine.TAPS_2X2, cin, 4 * cout, engine.pack_conv_transpose(wt, torch.float64)
        a_dims, ohw = (B, H, W, cin), (H, W)
    else:
        H, W, cin, cout = 10, 12, 32, 3
        a = rnd(B, H + 8, W + 8, cin); wt = rnd(cout, cin, 9, 9) / 50
        taps, kc, n_gemm, packed = engine.taps_kxk(9), cin, 16, engine.pack_final_plain(wt, torch.float64)
        a_dims, ohw = (B, H + 8, W + 8, cin), (H, W)
    a = a.to(adt)
    gout = rnd(B, ohw[0], ohw[1], n_gemm).to(gdtype)
    if kind == "final":
        gout[..., 3:] = 0
    # reference: autograd through the float64 gather-GEMM
    a64 = a.double().requires_grad_(True)
    w64 = packed.clone().requires_grad_(True)
    spec64 = ConvSpec(taps, kc, w64, n_gemm, n_gemm)
    (_emu_conv(spec64, a64, a_dims, None, ohw) * gout.double()).sum().backward()
    # wgrad
    db = ops.wgrad(ConvSpec(taps, kc, None, n_gemm, n_gemm), a.to(DEV), a_dims, engine._nhwc_strides(a), gout.to(DEV), ohw)
    tol = 1e-5 if gdtype == torch.float32 else 2e-3
    assert rel_l2(db, w64.grad) < tol
    if gdtype != torch.float32 and kc % 64 == 0:      # tensor-core wgrad: fp16 activations x bf16 gradients, fp32 accumulate
        a_b = ops.cast(a.to(DEV), gdtype)                 # same 16-bit format on both operands (mixed formats are illegal)
        db_tc = ops.wgrad(ConvSpec(taps, kc, None, n_gemm, n_gemm), a_b, a_dims, engine._nhwc_strides(a), gout.to(DEV), ohw, use_tc=True)
        (_emu_conv(ConvSpec(taps, kc, w64.detach().clone().requires_grad_(True), n_gemm, n_gemm), a_b.cpu().double(), a_dims, None, ohw)).sum()
        w65 = packed.clone().requires_grad_(True)
        (_emu_conv(ConvSpec(taps, kc, w65, n_gemm, n_gemm), a_b.cpu().double(), a_dims, None, ohw) * gout.double()).sum().backward()
        assert rel_l2(db_tc, w65.grad) < 1e-5, "tensor-core wgrad is not the exact fp32-accumulated product of its operands"
    # dgrad as a gather-GEMM on gout
    if kind == "s2d":
        wd = backward.pack_dgrad_s2d(wt, gdtype)
        spec_d = ConvSpec(backward._neg(engine.TAPS_2X2), n_gemm, wd.to(DEV), 4 * cin, 4 * cin)
    else:
        wd = backward.pack_dgrad(packed.float(), len(taps), kc, gdtype)
        spec_d = ConvSpec(backward._neg(taps), n_gemm, wd.to(DEV), kc, kc)
    da = torch.empty(a.shape, dtype=gdtype, device=DEV)
    for use_tc in ([False, True] if (gdtype != torch.float32 and n_gemm % 64 == 0) else [False]):
        da.fill_(float("nan"))
        ops.conv_gather(spec_d, gout.to(DEV), (B, ohw[0], ohw[1], n_gemm), engine._nhwc_strides(gout), da, (a.shape[1], a.shape[2]), None, use_tc)
        assert rel_l2(da, a64.grad) < (1e-5 if gdtype == torch.float32 else 6e-3), use_tc
    # unpack helpers round-trip to the PyTorch parameter layout
    if kind == "convT":
        dw = backward.unpack_conv_transpose(w64.grad, cin, cout)
        x = a.double().permute(0, 3, 1, 2).clone()
        wref = wt.double().clone().requires_grad_(True)
        y = F.conv_transpose2d(x, wref, stride=2, padding=1, output_padding=1)
        gimg = gout.double().view(B, H, W, 2, 2, cout).permute(0, 1, 3, 2, 4, 5).reshape(B, 2 * H, 2 * W, cout).permute(0, 3, 1, 2)
        (y * gimg).sum().backward()
        assert rel_l2(dw, wref.grad) < 1e-10


def test_conv_first_wgrad():
    g = torch.Generator().manual_seed(22)
    x = torch.rand((2, 3, 37, 45), generator=g)
    wt = (torch.randn((64, 3, 9, 9), generator=g) / 9).double().requires_grad_(True)
    y = F.conv2d(F.pad(x.double(), (4,) * 4, mode="reflect"), wt, stride=2)
    gout = torch.randn(y.shape, generator=g).double()
    (y * gout).sum().backward()
    dw = ops.conv_first_wgrad(x.to(DEV), gout.permute(0, 2, 3, 1).float().contiguous().to(DEV), 9, 2, 4, _lib.PAD_REFLECT)
    got = dw.view(3, 9, 9, 64).permute(3, 0, 1, 2)
    assert rel_l2(got, wt.grad) < 1e-5


@pytest.mark.parametrize("gdtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("cfg", [dict(c=256, pad=1, s2d=False, relu=True, drop=True, extra=False, out_s2d=False),
                                 dict(c=256, pad=1, s2d=False, relu=False, drop=False, extra=True, out_s2d=False),
                                 dict(c=64, pad=1, s2d=True, relu=True, drop=False, extra=False, out_s2d=False),
                                 dict(c=32, pad=4, s2d=False, relu=True, drop=False, extra=False, out_s2d=True),
                                 dict(c=64, pad=0, s2d=False, relu=True, drop=False, extra=False, out_s2d=True)])
@pytest.mark.parametrize("mode,hw", [("two_pass", (10, 12)), ("fused", (10, 12)), ("fused", (72, 48)), ("fused", (100, 98)), ("fused", (160, 160))])
def test_inorm_backward(cfg, gdtype, mode, hw):
    """Two-pass operators (reduce + apply) and the one-pass cluster kernel (1, 2, 4 and 8 CTAs per slab by plane size)."""
    g = torch.Generator().manual_seed(23)
    n, (h, w), c = 2, hw, cfg["c"]
    if h * w > 4096:
        c = min(c, 32)                          # large planes: fewer channels keep the float64 reference quick
    adt = torch.float32 if gdtype == torch.float32 else torch.float16
    raw = (torch.randn((n, h, w, c), generator=g) * 1.5 + 0.3).to(adt)
    gamma, beta = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g) * 0.3
    drop = (torch.rand((n, c), generator=g) < 0.9).float() / 0.9 if cfg["drop"] else None
    pad = cfg["pad"]
    # reference forward in float64 with autograd: out = halo(relu(IN(raw)) * drop)
    r64 = raw.double().requires_grad_(True)
    g64, b64 = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    x = r64.permute(0, 3, 1, 2)
    mu, var = x.mean((2, 3), keepdim=True), x.var((2, 3), unbiased=False, keepdim=True)
    y = (x - mu) / torch.sqrt(var + 1e-5) * g64.view(1, -1, 1, 1) + b64.view(1, -1, 1, 1)
    if cfg["relu"]:
        y = F.relu(y)
    if drop is not None:
        y = y * drop.double().view(n, c, 1, 1)
    plain = y
    if pad:
        y = F.pad(y, (pad,) * 4, mode="reflect")
    gbuf = torch.randn(y.shape, generator=g).to(gdtype)                     # gradient of the halo buffer (NCHW here)
    extra = torch.randn((n, h, w, c), generator=g).to(gdtype) if cfg["extra"] else None
    loss = (y * gbuf.double()).sum()
    if extra is not None:
        loss = loss + (plain * extra.double().permute(0, 3, 1, 2)).sum()
    loss.backward()
    # device: lay the buffer gradient out exactly as the forward buffer (NHWC halo, optionally space-to-depth)
    gs = gbuf.permute(0, 2, 3, 1).contiguous()
    if cfg["s2d"]:
        hp, wp = gs.shape[1], gs.shape[2]
        gs = F.pad(gs.float(), (0, 0, 0, wp % 2, 0, hp % 2)).to(gdtype)
        gs = gs.view(n, gs.shape[1] // 2, 2, gs.shape[2] // 2, 2, c).permute(0, 1, 3, 2, 4, 5).reshape(n, gs.shape[1] // 2, gs.shape[2] // 2, 4 * c).contiguous()
    st = torch.stack([raw.double().sum((1, 2)), (raw.double() ** 2).sum((1, 2))], dim=-1).float()
    args = (gs.to(DEV), None if extra is None else extra.to(DEV), raw.to(DEV), st.to(DEV), gamma.to(DEV),
            beta.to(DEV), None if drop is None else drop.to(DEV), gdtype, cfg["relu"], pad,
            _lib.PAD_REFLECT if pad else _lib.PAD_NONE, cfg["s2d"])
    if mode == "two_pass":
        gy, sums = ops.inorm_bwd_reduce(*args)
        draw, dgb = ops.inorm_bwd_apply(gy, raw.to(DEV), st.to(DEV), sums, gamma.to(DEV), out_s2d=cfg["out_s2d"])
    else:
        parts = ops.inorm_bwd_fused_parts(raw.to(DEV), gdtype, True, extra is not None, cfg["s2d"])
        esz = raw.element_size()
        cr = min(256 // w, h) if w <= 256 else 0                                 # rows per TMA chunk (<= 256 pixels)
        stride = -(-(cr * w) // 8) * 8 if cr else 0

        def choose():
            """Mirror of ibf_geometry: per slab width the smallest cluster that fits ~200 KB; most CTAs (cap 128); ties: a cluster
            of at most 4 beats one of 8, else the widest slab."""
            best, best_ctas = 0, 0
            for cw in (64, 32, 16):
                if c % cw:
                    continue
                for k in (1, 2, 4, 8):
                    rows = -(-(-(-h // k)) // cr) * cr                            # ceil(ceil(h / k) / cr) * cr: whole chunks
                    if (rows // cr) * stride * cw * esz * (2 + (extra is not None)) <= 200 * 1024 and rows // cr <= 32:
                        ctas = min(128, k * (c // cw) * n)
                        if ctas > best_ctas or (ctas == best_ctas and best > 4 and k <= 4):
                            best, best_ctas = k, ctas
                        break
            return best
        expect = 0 if (cfg["s2d"] or not cr) else choose()
        assert parts == expect, (parts, expect)
        if parts == 0:
            pytest.skip("space-to-depth gradient buffer / plane too large for 8 CTAs: the plan uses the two-pass operators here")
        draw, gy, sums = ops.inorm_bwd_fused(*args, out_s2d=cfg["out_s2d"], want_gy=True)
        draw2, gy2, _ = ops.inorm_bwd_fused(*args, out_s2d=cfg["out_s2d"], want_gy=False)
        assert gy2 is None and torch.equal(draw2, draw)                        # deterministic, with or without the gy output
        gy_ref, sums_ref = ops.inorm_bwd_reduce(*args)
        # (16-bit gradients: border pixels are rounded once more, when the folded halo contributions are added into the staged tile)
        assert rel_l2(gy, gy_ref) < (1e-6 if gdtype == torch.float32 else 4e-3) and rel_l2(sums, sums_ref) < (2e-5 if gdtype == torch.float32 else 4e-3)
        dgb = torch.empty((2, c), dtype=torch.float32, device=DEV)
        ops.affine_grads(sums.reshape(-1), [(0, c, 0, c)], n, dgb.view(-1))
    ref = r64.grad
    if cfg["out_s2d"]:
        ref = ref.view(n, h // 2, 2, w // 2, 2, c).permute(0, 1, 3, 2, 4, 5).reshape(n, h // 2, w // 2, 4 * c)
    tol = 2e-5 if gdtype == torch.float32 else 1e-2
    assert rel_l2(draw, ref) < tol
    dgam, dbet = dgb[0], dgb[1]
    assert rel_l2(sums.sum(0)[:, 1], dgam) < 1e-5 and rel_l2(sums.sum(0)[:, 0], dbet) < 1e-5
    assert rel_l2(dgam, g64.grad) < tol and rel_l2(dbet, b64.grad) < tol


@pytest.mark.parametrize("gdtype", [torch.float32, torch.bfloat16])
def test_pool_mask_sse_tv_backward(gdtype):
    g = torch.Generator().manual_seed(24)
    adt = torch.float32 if gdtype == torch.float32 else torch.bfloat16
    a = torch.relu(torch.randn((2, 9, 10, 64), generator=g)).to(adt)          # ReLU output, odd height
    gout = torch.randn((2, 4, 5, 64), generator=g).to(gdtype)
    extra = torch.randn((2, 9, 10, 64), generator=g).to(gdtype)
    a64 = a.double().requires_grad_(True)
    pooled = F.max_pool2d(a64.permute(0, 3, 1, 2), 2, 2)
    ((pooled * gout.double().permute(0, 3, 1, 2)).sum()).backward()
    ref = (a64.grad + extra.double()) * (a.double() > 0)
    got = ops.maxpool2_bwd(a.to(DEV), gout.to(DEV), extra.to(DEV))
    assert rel_l2(got, ref) < (1e-6 if gdtype == torch.float32 else 4e-3)
    got = ops.relu_mask(gout.to(DEV), None, a[:, :4, :5].contiguous().to(DEV))
    assert rel_l2(got, gout.double() * (a[:, :4, :5].double() > 0)) < 1e-6
    b = torch.randn((2, 9, 10, 64), generator=g).to(adt)
    scale = torch.tensor([0.37], device=DEV)
    da = ops.sse_bwd(a.to(DEV), b.to(DEV), scale, gdtype)
    assert rel_l2(da, 2 * 0.37 * (a.double() - b.double())) < (1e-6 if gdtype == torch.float32 else 4e-3)
    img = torch.randn((2, 3, 13, 17), generator=g)
    i64 = img.double().requires_grad_(True)
    (((i64[:, :, 1:] - i64[:, :, :-1]) ** 2).sum() + ((i64[:, :, :, 1:] - i64[:, :, :, :-1]) ** 2).sum()).backward()
    assert rel_l2(ops.tv_bwd(img.to(DEV), scale), 0.37 * i64.grad) < 1e-6
    assert rel_l2(ops.channel_sum(img.to(DEV)), img.double().sum((0, 2, 3))) < 1e-5


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("shape", [(2, 16, 16, 64), (1, 24, 8, 128), (2, 12, 6, 256), (1, 4, 4, 512)])
def test_gram_tc(shape, dtype):
    g = torch.Generator().manual_seed(25)
    f = torch.relu(torch.randn(shape, generator=g)).to(dtype)
    ref = emu_ops.gram(f, False)                                   # exact Gram of the rounded features
    got = ops.gram(f.to(DEV), use_tc=True)
    assert rel_l2(got, ref) < 1e-5
    assert rel_l2(ops.gram(f.to(DEV), use_tc=False), ref) < 1e-5


@pytest.mark.parametrize("shape", [(4, 64, 64), (2, 10, 12), (3, 37, 45)])
def test_pixel_stream_dgrad_matches_box_form(shape):
    """Pixel-stream data gradient of the trunk convolutions (fnst.h FNST_DESC_LINEAR; backward.LINEAR_DGRAD) against the box form:
    inorm_bwd_apply(out_pad=2) == dense d_raw inside a ZERO halo; the 3x3 data-gradient GEMM over the linear pixel stream of
    that buffer == the box form on the (H+2) x (W+2) domain, bit for bit (same accumulation order), in the [:H+2, :W+2] corner of
    an (H+4) x (W+4) buffer; inorm_bwd_reduce(gsrc_slack=2) on that buffer == the dense call."""
    B, H, W = shape
    C = 256
    g = torch.Generator().manual_seed(29)
    gdt, adt = torch.bfloat16, torch.float16
    raw = (torch.randn((B, H, W, C), generator=g) * 1.5 + 0.3).to(adt).to(DEV)
    gy = torch.randn((B, H, W, C), generator=g).to(gdt).to(DEV)
    gamma = (torch.rand(C, generator=g) + 0.5).to(DEV)
    beta = (torch.randn(C, generator=g) * 0.3).to(DEV)
    st = torch.stack([raw.double().sum((1, 2)), (raw.double() ** 2).sum((1, 2))], dim=-1).float()
    sums = torch.stack([gy.double().sum((1, 2)), torch.zeros((B, C), dtype=torch.float64, device=DEV)], dim=-1).float().contiguous()
    # poison the allocator's free blocks so that an unwritten halo cannot be zero by luck
    junk = torch.full((B * (H + 4) * (W + 4) * C * 2,), float("nan"), dtype=gdt, device=DEV)
    del junk
    d_pad, _ = ops.inorm_bwd_apply(gy, raw, st, sums, gamma, want_dgb=False, out_pad=2)
    d_dense, _ = ops.inorm_bwd_apply(gy, raw, st, sums, gamma, want_dgb=False)
    assert d_pad.shape == (B, H + 4, W + 4, C)
    assert torch.equal(d_pad[:, 2:-2, 2:-2], d_dense)
    halo = d_pad.clone()
    halo[:, 2:-2, 2:-2] = 0
    assert torch.equal(halo, torch.zeros_like(halo))

    wt = (torch.randn((C, C, 3, 3), generator=g) / 48)
    wd = backward.pack_dgrad(engine.pack_conv(wt, torch.float64), 9, C, gdt).to(DEV)
    taps9 = engine.taps_kxk(3)
    box = torch.empty((B, H + 2, W + 2, C), dtype=gdt, device=DEV)
    ops.conv_gather(ConvSpec(backward._neg(taps9), C, wd, C, C), d_dense, (B, H, W, C), engine._nhwc_strides(d_dense), box,
                    (H + 2, W + 2), None, True)
    m = B * (H + 4) * (W + 4)
    lin = torch.empty((B, H + 4, W + 4, C), dtype=gdt, device=DEV)
    ops.conv_gather(ConvSpec([(2 - dh, 2 - dw, 0) for dh, dw, _ in taps9], C, wd, C, C), d_pad, (1, 1, m, C),
                    (m * C, (W + 4) * C, C), lin, (1, m), None, True, linear=True)
    assert torch.equal(lin[:, :H + 2, :W + 2], box)

    drop = ((torch.rand((B, C), generator=g) < 0.9).float() / 0.9).to(DEV)
    common = (raw, st, gamma, beta, drop, gdt, True, 1, _lib.PAD_REFLECT, False)
    gy_a, sums_a = ops.inorm_bwd_reduce(box, None, *common)
    gy_b, sums_b = ops.inorm_bwd_reduce(lin, None, *common, gsrc_slack=2)
    assert torch.equal(gy_a, gy_b) and rel_l2(sums_b, sums_a) < 1e-5


@pytest.mark.parametrize("cfg", [dict(shape=(4, 64, 64, 256), pad=1, extra=False, relu=True, drop=True),
                                 dict(shape=(4, 64, 64, 256), pad=1, extra=True, relu=False, drop=False),
                                 dict(shape=(2, 10, 12, 256), pad=0, extra=True, relu=False, drop=False, gsrc=False),
                                 dict(shape=(1, 37, 45, 64), pad=0, extra=False, relu=True, drop=False),
                                 dict(shape=(2, 20, 24, 32), pad=4, extra=False, relu=True, drop=False),
                                 dict(shape=(3, 5, 7, 128), pad=1, extra=True, relu=True, drop=True)])
def test_inorm_bwd_reduce_tma_form_equals_register_form(cfg):
    """Pass 1 of the InstanceNorm backward with each image row staged by TMA box loads
    (tuning knob inorm_bwd_tma) against the register-pipelined form: the same operations in the same order -- gy bit for bit, the per-plane sums up to the
    order of their fp32 atomics."""
    B, H, W, C = cfg["shape"]
    pad = cfg["pad"]
    g = torch.Generator().manual_seed(41)
    gdt, adt = torch.bfloat16, torch.float16
    raw = (torch.randn((B, H, W, C), generator=g) * 1.5 + 0.3).to(adt).to(DEV)
    st = torch.stack([raw.double().sum((1, 2)), (raw.double() ** 2).sum((1, 2))], dim=-1).float()
    gamma, beta = (torch.rand(C, generator=g) + 0.5).to(DEV), (torch.randn(C, generator=g) * 0.3).to(DEV)
    drop = ((torch.rand((B, C), generator=g) < 0.9).float() / 0.9).to(DEV) if cfg["drop"] else None
    gsrc = torch.randn((B, H + 2 * pad, W + 2 * pad, C), generator=g).to(gdt).to(DEV) if cfg.get("gsrc", True) else None
    extra = torch.randn((B, H, W, C), generator=g).to(gdt).to(DEV) if cfg["extra"] else None
    res = []
    for tma in (1, 0):
        _lib.check(_lib.lib.fnst_set_tuning(b"inorm_bwd_tma", tma), "knob")
        res.append(ops.inorm_bwd_reduce(gsrc, extra, raw, st, gamma, beta, drop, gdt, cfg["relu"], pad,
                                        _lib.PAD_REFLECT if pad else _lib.PAD_NONE, False))
    _lib.check(_lib.lib.fnst_set_tuning(b"inorm_bwd_tma", 0), "knob")          # the default
    assert torch.isfinite(res[0][0].float()).all()
    assert torch.equal(res[0][0], res[1][0])
    assert rel_l2(res[0][1], res[1][1]) < 1e-5
