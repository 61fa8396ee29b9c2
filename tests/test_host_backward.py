"""Host-side logic of the backward plans (fast_neural_style_transfer_b200/backward.py: data-gradient operands, negated
tap tables, window views, weight-gradient unpacking, ReflectionPad2d fold, residual / dropout routing) checked on CPU
against torch autograd of the oracle, with the libfnst operators replaced by the float64 emulations of tests/emu_ops.py."""
import pytest
import torch

import emu_ops
from oracle import stylenet_oracle as O
from fast_neural_style_transfer_b200 import backward, engine, ops


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def _oracle_net_grads(p, x, drop, dy):
    q = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    y = O.stylenet_forward(q, x, drop)
    y.backward(dy)
    return {k: v.grad for k, v in q.items()}


@pytest.mark.parametrize("precision,shape", [("fp32", (2, 24, 24)), ("fp32", (1, 20, 28)), ("fp16", (2, 24, 24)), ("fp16", (1, 36, 20)),
                                             ("fp16x3", (2, 24, 24)), ("fp16x3", (1, 36, 20))])
def test_stylenet_backward_matches_autograd(monkeypatch, precision, shape):
    """fp16x3: the split forward keeps its own fp16 (hi, lo) arithmetic and its bf16 activation twins here, so the bound is
    that of bf16 weight-gradient operands (the structure of the tape -- split buffers, twins, window views -- is what is checked)."""
    emu_ops.install_backward(monkeypatch, ops)
    monkeypatch.setattr(backward, "grad_dtype", lambda precision: torch.float32)      # plan structure in fp32 arithmetic
    p = O.make_net_params(seed=3, random_affine=True)
    b, h, w = shape
    x = O.make_image(b, h, w, seed=11)
    drop = O.make_dropout_scales(b, seed=9)
    plan = engine.StyleNetPlan(precision)
    if precision != "fp16x3":
        plan.dtype = torch.float32
    plan.pack(p)
    tape = {}
    y = plan.forward(x, drop, tape)
    if precision == "fp16x3":
        with torch.no_grad():
            assert rel_l2(y, O.stylenet_forward(p, x, drop)) < 1e-4
    dy = torch.randn(y.shape, generator=torch.Generator().manual_seed(1))
    grads = backward.stylenet_backward(plan, tape, dy)
    ref = _oracle_net_grads(p, x, drop, dy)
    assert set(grads) == set(ref)
    scale = max(float(g.abs().max()) for g in ref.values())
    for k, g in ref.items():
        assert grads[k].shape == g.shape, k
        if k.endswith("conv.bias") and not k.startswith("final_conv"):
            # a bias in front of InstanceNorm has zero gradient; autograd returns fp32 noise there (SURVEY 8c hazard i)
            assert float(grads[k].abs().max()) == 0.0 and float(g.abs().max()) < 1e-4 * scale, k
        else:
            assert rel_l2(grads[k], g) < (2e-4 if precision != "fp16x3" else 1e-2), (k, rel_l2(grads[k], g))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("used", [(0, 1, 2, 4), (4,), (0,), (1, 2), (3,)])
def test_vgg_backward_matches_autograd(monkeypatch, precision, used):
    emu_ops.install_backward(monkeypatch, ops)
    monkeypatch.setattr(backward, "grad_dtype", lambda precision: torch.float32)      # plan structure in fp32 arithmetic
    p = O.make_vgg_params(seed=1)
    x = O.make_image(2, 16, 24, seed=77, normalized=True)
    plan = engine.VGGPlan(precision)
    plan.dtype = torch.float32
    plan.pack(p)
    tape = {}
    feats = plan.forward(x, tape)
    gen = torch.Generator().manual_seed(2)
    dfe = [torch.randn(f.shape, generator=gen) if i in used else None for i, f in enumerate(feats)]      # NHWC, like the features
    dx = backward.vgg_backward(plan, tape, dfe)
    xr = x.clone().requires_grad_(True)
    ref_feats = O.vgg_forward(p, xr)
    total = sum((f * g.permute(0, 3, 1, 2)).sum() for f, g in zip(ref_feats, dfe) if g is not None)
    total.backward()
    assert dx.shape == x.shape
    assert rel_l2(dx, xr.grad) < 1e-4


def test_loss_backward_helpers(monkeypatch):
    emu_ops.install_backward(monkeypatch, ops)
    gen = torch.Generator().manual_seed(3)
    f = torch.randn((2, 5, 6, 8), generator=gen)                                          # NHWC features
    dg = torch.randn((2, 8, 8), generator=gen)
    fr = f.clone().requires_grad_(True)
    G = torch.einsum("nhwi,nhwj->nij", fr, fr)
    (G * dg).sum().backward()
    assert rel_l2(backward.gram_backward(f, dg), fr.grad) < 1e-5
    img = torch.randn((2, 3, 7, 9), generator=gen)
    ir = img.clone().requires_grad_(True)
    (3.0 * (((ir[:, :, 1:] - ir[:, :, :-1]) ** 2).sum() + ((ir[:, :, :, 1:] - ir[:, :, :, :-1]) ** 2).sum())).backward()
    assert rel_l2(backward.tv_backward(img, torch.tensor(3.0)), ir.grad) < 1e-6
    a, t = torch.randn((2, 4, 4), generator=gen), torch.randn((4, 4), generator=gen)
    ar = a.clone().requires_grad_(True)
    (0.5 * ((ar - t) ** 2).sum()).backward()
    assert rel_l2(backward.sse_backward(a, t, torch.tensor(0.5)), ar.grad) < 1e-6


def test_dgrad_operand_layouts():
    """pack_dgrad / pack_dgrad_s2d / unpack_conv_transpose against direct definitions."""
    gen = torch.Generator().manual_seed(4)
    bf = torch.randn((6, 3 * 5), generator=gen, dtype=torch.float64)                      # [n_gemm, ntaps*kc]
    wd = backward.pack_dgrad(bf, 3, 5, torch.float64)
    for j in range(6):
        for t in range(3):
            for c in range(5):
                assert wd[c, t * 6 + j] == bf[j, t * 5 + c]
    w = torch.randn((4, 8, 3, 3), generator=gen, dtype=torch.float64)                     # ConvTranspose2d weight (in, out, k, k)
    packed = engine.pack_conv_transpose(w, torch.float64)
    assert torch.equal(backward.unpack_conv_transpose(packed, 4, 8), w)                   # unpack inverts pack on the 9 used taps
    w2 = torch.randn((6, 2, 3, 3), generator=gen, dtype=torch.float64)
    x = torch.randn((1, 2, 9, 9), generator=gen, dtype=torch.float64, requires_grad=True)
    y = torch.nn.functional.conv2d(x, w2, stride=2)                                       # (1, 6, 4, 4) on an already padded image
    g = torch.randn(y.shape, generator=gen, dtype=torch.float64)
    y.backward(g)
    # data gradient through the space-to-depth operand: d_buf[n, hs, ws, (ph,pw,c)] = sum_{dh,dw,o} g[hs-dh, ws-dw, o] * wd[(ph,pw,c), (dh,dw,o)]
    wd = backward.pack_dgrad_s2d(w2, torch.float64).view(2, 2, 2, 2, 2, 6)                # (ph, pw, c, dh, dw, o)
    gp = torch.nn.functional.pad(g, (1, 1, 1, 1))                                          # zero outside
    hs, ws = 5, 5
    d = torch.zeros((hs, ws, 2, 2, 2), dtype=torch.float64)
    for dh in (0, 1):
        for dw in (0, 1):
            patch = gp[0, :, 1 - dh:1 - dh + hs, 1 - dw:1 - dw + ws]                      # g[hs - dh, ws - dw]
            d += torch.einsum("ohw,pqco->hwpqc", patch, wd[:, :, :, dh, dw, :])
    full = d.permute(4, 0, 2, 1, 3).reshape(2, 2 * hs, 2 * ws)[:, :9, :9]
    assert torch.allclose(full, x.grad[0], atol=1e-12)


def test_staged_backward_and_bucket_wise_assembly(monkeypatch):
    """The generator form of the backward (backward.stylenet_backward_stages: cut points after residual blocks 2 and 0, the
    data-parallel path captures one CUDA graph per stage) with the gradients assembled bucket by bucket
    (assemble_gradients(first, last) over backward.bucket_bounds): every bucket is final when its stage ends, the three
    buckets tile the flat buffer back to front, and the result equals the single-shot backward."""
    emu_ops.install_backward(monkeypatch, ops)
    monkeypatch.setattr(backward, "grad_dtype", lambda precision: torch.float32)
    p = O.make_net_params(seed=3, random_affine=True)
    x = O.make_image(2, 24, 24, seed=11)
    drop = O.make_dropout_scales(2, seed=9)
    plan = engine.StyleNetPlan("fp16")
    plan.dtype = torch.float32
    plan.pack(p)
    tape = {}
    y = plan.forward(x, drop, tape)
    dy = torch.randn(y.shape, generator=torch.Generator().manual_seed(1))
    names = list(plan.params)
    whole = backward.assemble_gradients(backward.stylenet_backward_core(plan, tape, dy), names, plan.params)

    gen = backward.stylenet_backward_stages(plan, tape, dy)
    asm = backward._assembly(names, [plan.params[n].shape for n in names], True, dy.device)
    flat, stage, ranges = None, 0, []
    while True:
        try:
            core, done = next(gen), False
        except StopIteration as end:
            core, done = end.value, True
        first, last = backward.bucket_bounds(stage)
        flat = backward.assemble_gradients(core, names, plan.params, flat, first, last)
        lo = 0 if first is None else asm["offsets"][first]
        hi = asm["total"] if last is None else asm["offsets"][last]
        ranges.append((lo, hi))
        assert torch.equal(flat[lo:hi], whole[lo:hi]), stage                    # final as soon as its stage has run
        stage += 1
        if done:
            break
    assert stage == 3 and ranges[0][1] == asm["total"] and ranges[0][0] == ranges[1][1] and ranges[1][0] == ranges[2][1] and ranges[2][0] == 0
    assert torch.equal(flat, whole)
