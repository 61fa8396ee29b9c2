"""GPU parity of the full forward plans against the CPU oracle (same seeded parameters/inputs).

Tolerances (BASELINE.json north_star): fp32 path 1e-4 relative L2; tensor-core path 1e-2 relative L2
and <= 1.0 absolute on the 0-255 pixel scale (pixel = clamp(y*std+mean, 0, 1)*255, inference.py:52-57).
"""
import pytest
import torch

from oracle import stylenet_oracle as O

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from fast_neural_style_transfer_b200 import engine

DEV = "cuda"


def rel_l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _cuda(p):
    return {k: v.to(DEV) for k, v in p.items()}


@pytest.mark.parametrize("shape", [(1, 64, 64), (2, 37, 45), (1, 256, 256)])
def test_stylenet_fp32_path(shape):
    p = O.make_net_params(seed=0, random_affine=True)
    b, h, w = shape
    x = O.make_image(b, h, w, seed=1234)
    with torch.no_grad():
        ref = O.stylenet_forward(p, x)
    y = engine.StyleNetPlan("fp32").pack(_cuda(p)).forward(x.to(DEV))
    assert y.shape == ref.shape
    assert rel_l2(y, ref) < 1e-4


@pytest.mark.parametrize("precision,tol_l2,tol_px", [("fp16", 1e-2, 1.0), ("bf16", 5e-2, 6.0)])
@pytest.mark.parametrize("shape", [(1, 64, 64), (2, 37, 45), (1, 256, 256)])
def test_stylenet_tensor_core_path(shape, precision, tol_l2, tol_px):
    """fp16 operands meet the north_star tolerance; bf16 operands are reported with their own
    (looser, measured) bound -- single-pass bf16 cannot meet 1e-2 / 1.0 px at random init."""
    p = O.make_net_params(seed=0)
    b, h, w = shape
    x = O.make_image(b, h, w, seed=1234)
    with torch.no_grad():
        ref = O.stylenet_forward(p, x)
    y = engine.StyleNetPlan(precision).pack(_cuda(p)).forward(x.to(DEV))
    assert y.shape == ref.shape
    err = rel_l2(y, ref)
    px = float((O.to_pixels(y.cpu()) - O.to_pixels(ref)).abs().max())
    print(f"{precision} {shape}: rel_l2={err:.3e} max_px={px:.3f}")
    assert err < tol_l2
    assert px <= tol_px


def test_stylenet_dropout_train_mode():
    p = O.make_net_params(seed=0, random_affine=True)
    x = O.make_image(2, 48, 48, seed=5)
    drop = O.make_dropout_scales(2, seed=7)
    with torch.no_grad():
        ref = O.stylenet_forward(p, x, drop)
    y = engine.StyleNetPlan("fp32").pack(_cuda(p)).forward(x.to(DEV), [d.to(DEV) for d in drop])
    assert rel_l2(y, ref) < 1e-4


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("fp16", 5e-3)])
def test_vgg_features(precision, tol):
    p = O.make_vgg_params(seed=1)
    x = O.make_image(2, 64, 48, seed=77, normalized=True)
    with torch.no_grad():
        ref = O.vgg_forward(p, x)
    feats = engine.VGGPlan(precision).pack(_cuda(p)).forward(x.to(DEV))
    for i, (f, r) in enumerate(zip(feats, ref)):
        assert f.shape == (r.shape[0], r.shape[2], r.shape[3], r.shape[1])
        err = rel_l2(f.permute(0, 3, 1, 2), r)
        print(f"vgg {precision} feat{i}: {err:.3e}")
        assert err < tol, i


@pytest.mark.parametrize("shape", [(1, 64, 64), (2, 37, 45), (1, 256, 256)])
def test_stylenet_fp16x3_tensor_core_path_meets_fp32_tolerance(shape):
    """Error-compensated fp16 (hi, lo) split on the tcgen05 kernel: the 1e-4 class on tensor cores."""
    p = O.make_net_params(seed=0, random_affine=True)
    b, h, w = shape
    x = O.make_image(b, h, w, seed=1234)
    with torch.no_grad():
        ref = O.stylenet_forward(p, x)
    y = engine.StyleNetPlan("fp16x3").pack(_cuda(p)).forward(x.to(DEV))
    err = rel_l2(y, ref)
    print(f"fp16x3 {shape}: rel_l2={err:.3e}")
    assert y.shape == ref.shape and err < 1e-4


@pytest.mark.parametrize("shape", [(1, 135, 240), (3, 50, 34)])
def test_stylenet_odd_aspect_and_batch(shape):
    """1080p-like aspect (1/8 scale) and an odd batch: ragged tiles on every layer of the tensor-core path."""
    p = O.make_net_params(seed=0, random_affine=True)
    b, h, w = shape
    x = O.make_image(b, h, w, seed=99)
    with torch.no_grad():
        ref = O.stylenet_forward(p, x)
    for precision, tol in (("fp16", 1e-2), ("fp16x3", 1e-4)):
        y = engine.StyleNetPlan(precision).pack(_cuda(p)).forward(x.to(DEV))
        assert y.shape == ref.shape
        assert rel_l2(y, ref) < tol, precision
