"""GPU parity tests of the libfnst operators (through the C ABI) against the operator semantics
emulated on CPU in float64 (tests/emu_ops.py) and against the oracle."""
import pytest
import torch

import emu_ops
from oracle import stylenet_oracle as O

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from fast_neural_style_transfer_b200 import engine, ops
    from fast_neural_style_transfer_b200.ops import ConvSpec
    from fast_neural_style_transfer_b200 import _lib

DEV = "cuda"


def rel_l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _run_conv_pair(spec_cpu, a_cpu, a_dims, a_strides, out_shape, out_hw, with_stats, use_tc, out_dtype=None):
    """Run one gather-GEMM on the GPU and in the CPU emulation; return (gpu_out, cpu_out, gpu_stats, cpu_stats)."""
    n = a_dims[0]
    out_dtype = out_dtype or a_cpu.dtype
    out_cpu = torch.zeros(out_shape, dtype=torch.float64)
    st_cpu = torch.zeros((n, spec_cpu.c_out, 2)) if with_stats else None
    emu_ops.conv_gather(spec_cpu, a_cpu, a_dims, a_strides, out_cpu, out_hw, st_cpu, False)
    spec_gpu = ConvSpec(spec_cpu.taps, spec_cpu.kc, spec_cpu.weight.to(DEV), spec_cpu.n_gemm, spec_cpu.c_out, spec_cpu.h0,
                        spec_cpu.w0, spec_cpu.epilogue, None if spec_cpu.bias is None else spec_cpu.bias.to(DEV), spec_cpu.relu,
                        per_image_weights=spec_cpu.per_image_weights)
    # keep any slack that follows the view in its storage (paired final-conv view reads 32 elements past the end)
    base = a_cpu._base if a_cpu._base is not None else a_cpu
    a_gpu = base.to(DEV).view(-1)[:a_cpu.numel()].view(a_cpu.shape)
    out_gpu = torch.full(out_shape, float("nan"), dtype=torch.float32 if spec_cpu.epilogue in (_lib.EPI_NCHW_F32, _lib.EPI_ROWSUM9) else out_dtype, device=DEV)
    st_gpu = torch.empty((n, spec_cpu.c_out, 2), device=DEV) if with_stats else None
    ops.conv_gather(spec_gpu, a_gpu, a_dims, a_strides, out_gpu, out_hw, st_gpu, use_tc)
    torch.cuda.synchronize()
    return out_gpu, out_cpu, st_gpu, st_cpu


# Paired-pixel view of a 9x9 / 32-channel conv with the plain NCHW-fp32 epilogue (the form final_conv used before the
# separable ROWSUM9 epilogue; kept as a test of overlapping-stride activation views + the NCHW epilogue).
def taps_final_pairs():
    return [(kh, 2 * j, 0) for kh in range(9) for j in range(5)]


def pack_final_pairs(w: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """final_conv weight (3, 32, 9, 9) -> (16, 45*64): the activation view pairs two adjacent
    pixels (64 = 2 x 32 channels) per tap so that every K block is a full 128-byte row; the
    phantom tap kw = 9 and output rows 3..15 are zero."""
    o, c, k, _ = w.shape
    assert (c, k) == (32, 9) and o <= 16
    b = torch.zeros((16, 9, 5, 2, 32), dtype=w.dtype, device=w.device)
    for j in range(5):
        for jj in range(2):
            kw = 2 * j + jj
            if kw < 9:
                b[:o, :, j, jj, :] = w[:, :, :, kw].permute(0, 2, 1)
    return b.reshape(16, 45 * 64).to(dtype).contiguous()



CONV_CASES = {
    # name: (B, H, W, Cin, Cout, kind)
    "res3x3_small": (2, 16, 16, 256, 256, "reflect3"),
    "res3x3_odd": (1, 13, 21, 256, 256, "reflect3"),
    "res3x3_64": (1, 64, 64, 256, 256, "reflect3"),
    "vgg_64_64": (1, 24, 40, 64, 64, "zero3"),
    "vgg_128_256": (1, 16, 16, 128, 256, "zero3"),
    "vgg_512_512_tiny": (2, 4, 4, 512, 512, "zero3"),
    "s2d": (2, 17, 23, 64, 256, "s2d"),
    "convT_256_64": (1, 9, 12, 256, 64, "convT"),
    "convT_64_32": (2, 16, 16, 64, 32, "convT"),
    "final_pairs": (1, 20, 28, 32, 3, "final"),
    "final_rowsum": (2, 21, 27, 32, 3, "final_rowsum"),
}


def _make_case(name, dtype, tc):
    B, H, W, cin, cout, kind = CONV_CASES[name]
    g = torch.Generator().manual_seed(sum(map(ord, name)))
    rnd = lambda *s: torch.randn(*s, generator=g)
    if kind == "reflect3":
        a = rnd(B, H + 2, W + 2, cin).to(dtype)
        w = (rnd(cout, cin, 3, 3) / (3 * cin ** 0.5))
        spec = ConvSpec(engine.taps_kxk(3), cin, engine.pack_conv(w, dtype), cout, cout)
        return spec, a, (B, H + 2, W + 2, cin), engine._nhwc_strides(a), (B, H, W, cout), (H, W), True
    if kind == "zero3":
        a = rnd(B, H, W, cin).to(dtype)
        w = (rnd(cout, cin, 3, 3) / (3 * cin ** 0.5))
        spec = ConvSpec(engine.taps_kxk(3), cin, engine.pack_conv(w, dtype), cout, cout, h0=-1, w0=-1, bias=rnd(cout) * 0.1, relu=True)
        return spec, a, (B, H, W, cin), engine._nhwc_strides(a), (B, H, W, cout), (H, W), False
    if kind == "s2d":
        hs, ws = (H + 3) // 2, (W + 3) // 2
        a = rnd(B, hs, ws, 4 * cin).to(dtype)
        w = (rnd(cout, cin, 3, 3) / (3 * cin ** 0.5))
        spec = ConvSpec(engine.taps_s2d_3x3(cin), cin, engine.pack_conv(w, dtype), cout, cout)
        ho, wo = (H + 1) // 2, (W + 1) // 2
        return spec, a, (B, hs, ws, 4 * cin), engine._nhwc_strides(a), (B, ho, wo, cout), (ho, wo), True
    if kind == "convT":
        a = rnd(B, H, W, cin).to(dtype)
        w = (rnd(cin, cout, 3, 3) / (3 * cin ** 0.5))
        spec = ConvSpec(engine.TAPS_2X2, cin, engine.pack_conv_transpose(w, dtype), 4 * cout, cout, epilogue=_lib.EPI_D2S)
        return spec, a, (B, H, W, cin), engine._nhwc_strides(a), (B, 2 * H, 2 * W, cout), (H, W), True
    if kind in ("final", "final_rowsum"):
        hq, wq = H + 8, W + 8
        flat = torch.zeros(B * hq * wq * 32 + 64, dtype=dtype)
        flat[:B * hq * wq * 32] = rnd(B * hq * wq * 32).to(dtype)
        a = flat[:B * hq * wq * 32].view(B, hq, wq, 32)
        w = (rnd(3, 32, 9, 9) / (9 * 32 ** 0.5))
        bias = torch.zeros(16); bias[:3] = rnd(3)
        if kind == "final_rowsum" and tc:
            spec = ConvSpec(engine.TAPS_ROWSUM, 64, engine.pack_final_rowsum(w, dtype), 32, 3, epilogue=_lib.EPI_ROWSUM9, bias=bias)
            return spec, a, (B, hq, wq, 64), (hq * wq * 32, wq * 32, 32), (B, 3, H, W), (H, W), False
        if tc:
            spec = ConvSpec(taps_final_pairs(), 64, pack_final_pairs(w, dtype), 16, 3, epilogue=_lib.EPI_NCHW_F32, bias=bias)
            return spec, a, (B, hq, wq, 64), (hq * wq * 32, wq * 32, 32), (B, 3, H, W), (H, W), False
        spec = ConvSpec(engine.taps_kxk(9), 32, engine.pack_final_plain(w, dtype), 16, 3, epilogue=_lib.EPI_NCHW_F32, bias=bias)
        return spec, a, (B, hq, wq, 32), engine._nhwc_strides(a), (B, 3, H, W), (H, W), False
    raise KeyError(kind)


@pytest.mark.parametrize("name", list(CONV_CASES))
def test_conv_simt_fp32(name):
    spec, a, dims, strides, oshape, ohw, with_stats = _make_case(name, torch.float32, tc=False)
    got, ref, st, st_ref = _run_conv_pair(spec, a, dims, strides, oshape, ohw, with_stats, use_tc=False)
    assert not torch.isnan(got).any()
    assert rel_l2(got, ref) < 2e-6
    if with_stats:
        assert rel_l2(st[:, :, 1], st_ref[:, :, 1]) < 1e-5
        assert float((st[:, :, 0].cpu() - st_ref[:, :, 0]).abs().max()) < 1e-3 * float(st_ref[:, :, 1].max()) ** 0.5


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("name", list(CONV_CASES))
def test_conv_tc(name, dtype):
    """tcgen05 kernel == exact fp32-accumulated GEMM of the (already rounded) 2-byte operands."""
    spec, a, dims, strides, oshape, ohw, with_stats = _make_case(name, dtype, tc=True)
    got, ref, st, st_ref = _run_conv_pair(spec, a, dims, strides, oshape, ohw, with_stats, use_tc=True)
    assert not torch.isnan(got.float()).any()
    tol = 2e-5 if got.dtype == torch.float32 else (1e-3 if dtype == torch.float16 else 6e-3)   # output rounding
    assert rel_l2(got, ref) < tol
    if with_stats:
        # statistics are those of the fp32 accumulators (generic epilogue) or of exactly the stored, rounded values (staged
        # shared-memory epilogue of one-tile-per-CTA launches): either is a valid InstanceNorm statistic of its tensor
        g64 = got.double().cpu()
        st_got = torch.stack([g64.sum(dim=(1, 2)), (g64 ** 2).sum(dim=(1, 2))], dim=-1)
        exact = rel_l2(st[:, :, 1], st_ref[:, :, 1]) < 1e-4 and \
            float((st[:, :, 0].cpu() - st_ref[:, :, 0]).abs().max()) < 2e-3 * float(st_ref[:, :, 1].max()) ** 0.5
        stored = rel_l2(st[:, :, 1], st_got[:, :, 1]) < 2e-5 and \
            float((st[:, :, 0].cpu().double() - st_got[:, :, 0]).abs().max()) < 1e-4 * float(st_got[:, :, 1].max()) ** 0.5
        assert exact or stored


def test_conv_tc_fp32_out():
    spec, a, dims, strides, oshape, ohw, with_stats = _make_case("res3x3_small", torch.float16, tc=True)
    got, ref, _, _ = _run_conv_pair(spec, a, dims, strides, oshape, ohw, True, use_tc=True, out_dtype=torch.float32)
    assert rel_l2(got, ref) < 1e-5      # fp32 accumulation over K = 2304


@pytest.mark.parametrize("stride,k,pad,mode,relu", [(2, 9, 4, 1, False), (1, 3, 1, 2, True)])
def test_conv_first(stride, k, pad, mode, relu):
    g = torch.Generator().manual_seed(4)
    x = torch.rand((2, 3, 37, 45), generator=g)
    w = engine.pack_first(torch.randn((64, 3, k, k), generator=g) / k)
    bias = torch.randn(64, generator=g) if relu else None
    ho, wo = (37 + 2 * pad - k) // stride + 1, (45 + 2 * pad - k) // stride + 1
    ref = torch.zeros((2, ho, wo, 64), dtype=torch.float64)
    st_ref = torch.zeros((2, 64, 2))
    emu_ops.conv_first(x, w, bias, k, stride, pad, mode, relu, ref, st_ref)
    for dtype, tol in ((torch.float32, 2e-6), (torch.float16, 1e-3)):
        out = torch.empty((2, ho, wo, 64), dtype=dtype, device=DEV)
        st = torch.empty((2, 64, 2), device=DEV)
        ops.conv_first(x.to(DEV), w.to(DEV), None if bias is None else bias.to(DEV), k, stride, pad, mode, relu, out, st)
        assert rel_l2(out, ref) < tol
        assert rel_l2(st[:, :, 1], st_ref[:, :, 1]) < 1e-5


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("cfg", [dict(c=256, pad=1, mode=1, s2d=False, res=True, drop=True, relu=False),
                                 dict(c=64, pad=1, mode=1, s2d=True, res=False, drop=False, relu=True),
                                 dict(c=32, pad=4, mode=1, s2d=False, res=False, drop=False, relu=True),
                                 dict(c=64, pad=0, mode=0, s2d=False, res=False, drop=False, relu=True)])
def test_inorm_apply(dtype, cfg):
    g = torch.Generator().manual_seed(5)
    n, h, w, c = 2, 9, 13, cfg["c"]
    raw = (torch.randn((n, h, w, c), generator=g) * 2 + 0.5).to(dtype)
    st = torch.stack([raw.double().sum(dim=(1, 2)), (raw.double() ** 2).sum(dim=(1, 2))], dim=-1).float()
    gamma, beta = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g)
    drop = (torch.rand((n, c), generator=g) < 0.9).float() / 0.9 if cfg["drop"] else None
    res = torch.randn((n, h + 2, w + 2, c), generator=g).to(dtype) if cfg["res"] else None
    pad = cfg["pad"]
    hp, wp = h + 2 * pad, w + 2 * pad
    oshape = (n, (hp + 1) // 2, (wp + 1) // 2, 4 * c) if cfg["s2d"] else (n, hp, wp, c)
    ref = torch.zeros(oshape, dtype=torch.float64)
    emu_ops.inorm_apply(raw, st, gamma, beta, ref, cfg["relu"], pad, cfg["mode"], cfg["s2d"], drop, res, 1)
    out = torch.zeros(oshape, dtype=dtype, device=DEV)
    ops.inorm_apply(raw.to(DEV), st.to(DEV), gamma.to(DEV), beta.to(DEV), out, cfg["relu"], pad, cfg["mode"], cfg["s2d"],
                    None if drop is None else drop.to(DEV), None if res is None else res.to(DEV), 1)
    tol = {torch.float32: 3e-6, torch.float16: 1e-3, torch.bfloat16: 6e-3}[dtype]
    assert rel_l2(out, ref) < tol
    # bf16 twin written by the same launch (operand of the weight-gradient GEMM): same geometry, bf16 rounding of the same values
    out_b = torch.zeros(oshape, dtype=dtype, device=DEV)
    twin = torch.zeros(oshape, dtype=torch.bfloat16, device=DEV)
    ops.inorm_apply(raw.to(DEV), st.to(DEV), gamma.to(DEV), beta.to(DEV), out_b, cfg["relu"], pad, cfg["mode"], cfg["s2d"],
                    None if drop is None else drop.to(DEV), None if res is None else res.to(DEV), 1, out2=twin)
    assert torch.equal(out_b, out)
    assert rel_l2(twin, ref) < 6e-3
    if dtype == torch.float32:
        assert torch.equal(twin, out.to(torch.bfloat16))


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_maxpool_gram_sse_tv_layout(dtype):
    g = torch.Generator().manual_seed(6)
    x = torch.randn((2, 10, 14, 64), generator=g).to(dtype)
    assert torch.equal(ops.maxpool2(x.to(DEV)).cpu(), emu_ops.maxpool2(x))
    gr = ops.gram(x.to(DEV), use_tc=False)
    assert rel_l2(gr, emu_ops.gram(x, False)) < 1e-5
    acc = torch.zeros((), dtype=torch.float64, device=DEV)
    y = torch.randn((2, 10, 14, 64), generator=g).to(dtype)
    ops.sse(x.to(DEV), y.to(DEV), acc)
    ref = torch.zeros((), dtype=torch.float64); emu_ops.sse(x, y, ref)
    assert abs(float(acc) / float(ref) - 1) < 1e-5
    # one-launch scaled forms (loss scalar incl. normalisation, deterministic block-ordered sum, self-cleaning workspace)
    o1 = torch.full((), 7.0, dtype=torch.float32, device=DEV)
    ops.sse_scaled(x.to(DEV), y.to(DEV), 0.25, o1)
    assert abs(float(o1) / (0.25 * float(ref)) - 1) < 1e-5
    ops.sse_scaled(x.to(DEV), y.to(DEV), 0.5, o1, accumulate=True)
    assert abs(float(o1) / (0.75 * float(ref)) - 1) < 1e-5
    o2 = torch.empty((), dtype=torch.float32, device=DEV)
    ops.sse_scaled(x.to(DEV), y.to(DEV), 0.25, o2)
    ops.sse_scaled(x.to(DEV), y.to(DEV), 0.5, o2, accumulate=True)
    assert torch.equal(o1, o2)                                   # run-to-run identical
    big = torch.randn((3, 3, 300, 301), generator=g)
    rtv = torch.zeros((), dtype=torch.float64); emu_ops.tv(big, rtv)
    o3 = torch.empty((), dtype=torch.float32, device=DEV)
    ops.tv_scaled(big.to(DEV), 1.0 / big.numel(), o3)
    assert abs(float(o3) / (float(rtv) / big.numel()) - 1) < 1e-5
    tgt = torch.randn((64, 64), generator=g)
    acc.zero_(); ops.sse(gr, tgt.to(DEV), acc)
    ref.zero_(); emu_ops.sse(gr.cpu(), tgt, ref)
    assert abs(float(acc) / float(ref) - 1) < 1e-5
    img = torch.randn((2, 3, 17, 19), generator=g)
    acc.zero_(); ops.tv(img.to(DEV), acc)
    ref.zero_(); emu_ops.tv(img, ref)
    assert abs(float(acc) / float(ref) - 1) < 1e-5
    nchw = ops.nhwc_to_nchw(x.to(DEV))
    assert torch.equal(nchw.cpu(), x.permute(0, 3, 1, 2).float())
    back = ops.nchw_to_nhwc(nchw, dtype)
    assert torch.equal(back.cpu(), x)


def test_error_reporting():
    spec, a, dims, strides, oshape, ohw, _ = _make_case("res3x3_small", torch.float32, tc=False)
    with pytest.raises(RuntimeError, match="fp16 or bf16"):
        _run_conv_pair(spec, a, dims, strides, oshape, ohw, True, use_tc=True)
    with pytest.raises(RuntimeError, match="CUDA tensors"):
        ops.maxpool2(torch.zeros(1, 4, 4, 8))


@pytest.mark.parametrize("dtype,use_tc", [(torch.float32, False), (torch.bfloat16, True), (torch.float16, False)])
def test_conv_per_image_weights(dtype, use_tc):
    """b_image_rows: image n multiplies by its own weight matrix (Gram backward: dF[n] = F[n] (dG[n] + dG[n]^T))."""
    g = torch.Generator().manual_seed(31)
    B, H, W, C = 3, 10, 12, 128
    a = torch.randn((B, H, W, C), generator=g).to(dtype)
    wts = (torch.randn((B, C, C), generator=g) / C ** 0.5).to(dtype)
    spec = ConvSpec([(0, 0, 0)], C, wts, C, C, per_image_weights=True)
    got, ref, _, _ = _run_conv_pair(spec, a, (B, H, W, C), engine._nhwc_strides(a), (B, H, W, C), (H, W), False, use_tc)
    assert rel_l2(got, ref) < {torch.float32: 2e-6, torch.float16: 1e-3, torch.bfloat16: 6e-3}[dtype]


def test_split_producers():
    """fp16 (hi | lo) split writers of the fp16x3 path: inorm_apply(split) and image_to_halo(split)."""
    g = torch.Generator().manual_seed(41)
    n, h, w, c = 2, 9, 13, 64
    raw = torch.randn((n, h, w, c), generator=g) * 2 + 0.5
    st = torch.stack([raw.double().sum(dim=(1, 2)), (raw.double() ** 2).sum(dim=(1, 2))], dim=-1).float()
    gamma, beta = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g)
    res_f = torch.randn((n, h + 2, w + 2, c), generator=g)
    res_hi = res_f.half()
    res = torch.cat([res_hi, (res_f - res_hi.float()).half()], dim=-1)
    for s2d in (False, True):
        hp, wp = h + 2, w + 2
        oshape = (n, (hp + 1) // 2, (wp + 1) // 2, 8 * c) if s2d else (n, hp, wp, 2 * c)
        ref = torch.zeros(oshape, dtype=torch.float16)
        r = None if s2d else res
        emu_ops.inorm_apply(raw, st, gamma, beta, ref, True, 1, 1, s2d, None, r, 1, split=True)
        out = torch.zeros(oshape, dtype=torch.float16, device=DEV)
        ops.inorm_apply(raw.to(DEV), st.to(DEV), gamma.to(DEV), beta.to(DEV), out, True, 1, 1, s2d, None,
                        None if r is None else r.to(DEV), 1, split=True)
        if s2d:
            got = out.cpu().float().view(n, oshape[1], oshape[2], 4, 2, c).sum(4)
            want = ref.float().view(n, oshape[1], oshape[2], 4, 2, c).sum(4)
        else:
            got = out.cpu().float().view(*oshape[:3], 2, c).sum(3)           # hi + lo recombined
            want = ref.float().view(*oshape[:3], 2, c).sum(3)
        assert rel_l2(got, want) < 3e-6
    x = torch.rand((2, 3, 11, 14), generator=g)
    ref = emu_ops.image_to_halo(x, 4, 1, 8, 20, 24, torch.float16, split=True)
    got = ops.image_to_halo(x.to(DEV), 4, 1, 8, 20, 24, torch.float16, split=True)
    assert torch.equal(got.cpu(), ref)
