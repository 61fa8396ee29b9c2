"""Host-side logic of the product (tap tables, packing, halo buffers, operator order) checked on CPU
against the oracle by emulating the libfnst operator semantics (tests/emu_ops.py)."""
import pytest
import torch

import emu_ops
from oracle import stylenet_oracle as O
from fast_neural_style_transfer_b200 import engine, ops


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
@pytest.mark.parametrize("shape", [(1, 32, 32), (2, 21, 27)])
def test_stylenet_plan_matches_oracle(monkeypatch, precision, shape):
    emu_ops.install(monkeypatch, ops)
    p = O.make_net_params(seed=3, random_affine=True)
    b, h, w = shape
    x = O.make_image(b, h, w, seed=11)
    plan = engine.StyleNetPlan(precision)
    # exercise the plan structure in fp32 arithmetic (fp16 plan = paired final conv, same ops otherwise)
    plan.dtype = torch.float32
    plan.pack(p)
    y = plan.forward(x)
    with torch.no_grad():
        ref = O.stylenet_forward(p, x)
    assert y.shape == ref.shape
    assert rel_l2(y, ref) < 1e-5


def test_stylenet_plan_dropout(monkeypatch):
    emu_ops.install(monkeypatch, ops)
    p = O.make_net_params(seed=3, random_affine=True)
    x = O.make_image(2, 24, 24, seed=12)
    drop = O.make_dropout_scales(2, seed=9)
    plan = engine.StyleNetPlan("fp32").pack(p)
    y = plan.forward(x, drop_scales=drop)
    with torch.no_grad():
        ref = O.stylenet_forward(p, x, drop)
    assert rel_l2(y, ref) < 1e-5


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_vgg_plan_matches_oracle(monkeypatch, precision):
    emu_ops.install(monkeypatch, ops)
    p = O.make_vgg_params(seed=1)
    x = O.make_image(2, 16, 24, seed=77, normalized=True)
    plan = engine.VGGPlan(precision)
    plan.dtype = torch.float32          # tensor-core plan structure (windowed conv1_1) in fp32 arithmetic
    plan.pack(p)
    feats = plan.forward(x)
    with torch.no_grad():
        ref = O.vgg_forward(p, x)
    for f, r in zip(feats, ref):
        assert rel_l2(f.permute(0, 3, 1, 2), r) < 1e-5


def test_pack_conv_transpose_is_subpixel_form():
    # ConvTranspose2d(k3,s2,p1,op1) == 2x2-tap gather with 4*Cout columns + depth-to-space
    torch.manual_seed(0)
    w = torch.randn(8, 4, 3, 3, dtype=torch.float64)
    x = torch.randn(1, 8, 5, 6, dtype=torch.float64)
    ref = torch.nn.functional.conv_transpose2d(x, w, stride=2, padding=1, output_padding=1)
    b = engine.pack_conv_transpose(w, torch.float64)            # (16, 32)
    xp = torch.nn.functional.pad(x, (0, 1, 0, 1)).permute(0, 2, 3, 1)   # zero row/col = TMA OOB fill
    cols = torch.cat([xp[:, dh:dh + 5, dw:dw + 6, :] for dh, dw, _ in engine.TAPS_2X2], dim=-1)
    out = (cols @ b.t()).view(1, 5, 6, 2, 2, 4).permute(0, 1, 3, 2, 4, 5).reshape(1, 10, 12, 4).permute(0, 3, 1, 2)
    assert torch.allclose(out, ref, atol=1e-12)


@pytest.mark.parametrize("shape", [(1, 32, 32), (2, 21, 27)])
def test_stylenet_plan_fp16x3_is_fp32_accurate(monkeypatch, shape):
    """Error-compensated split: real fp16 (hi, lo) pairs, three virtual taps per tap -> 1e-4 class accuracy."""
    emu_ops.install(monkeypatch, ops)
    p = O.make_net_params(seed=3, random_affine=True)
    b, h, w = shape
    x = O.make_image(b, h, w, seed=11)
    y = engine.StyleNetPlan("fp16x3").pack(p).forward(x)
    with torch.no_grad():
        ref = O.stylenet_forward(p, x)
    assert y.shape == ref.shape
    err = rel_l2(y, ref)
    print("fp16x3 emulated rel_l2", err)
    assert err < 1e-4
