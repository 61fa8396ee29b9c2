"""GPU parity of the optimizer tail: fast_neural_style_transfer_b200.optim.clip_grad_norm_ / Adam
against torch.nn.utils.clip_grad_norm_ / torch.optim.Adam (the calls of train.py:203-205, optimizer of
train.py:135-139) on the same parameters and gradients."""
import copy

import pytest
import torch

from fast_neural_style_transfer_b200 import optim as fo

pytestmark = pytest.mark.gpu
DEV = "cuda"

# odd sizes: below one vector, not a multiple of 4, exactly one chunk, chunk + tail, several chunks
SHAPES = [(3,), (64,), (7, 5), (4096,), (4097,), (64, 3, 9, 9), (256, 64, 3, 3), (13, 1001)]


def _make(shapes, seed, grad_scale=1.0):
    g = torch.Generator().manual_seed(seed)
    ps = [torch.nn.Parameter(torch.randn(s, generator=g).to(DEV)) for s in shapes]
    for p in ps:
        p.grad = (torch.randn(p.shape, generator=g) * grad_scale).to(DEV)
    return ps


def _rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


@pytest.mark.parametrize("grad_scale", [10.0, 1e-4])          # clipping active / inactive (coef clamps to 1)
def test_clip_grad_norm_matches_torch(grad_scale):
    a = _make(SHAPES, 0, grad_scale)
    b = [torch.nn.Parameter(p.detach().clone()) for p in a]
    for pa, pb in zip(a, b):
        pb.grad = pa.grad.clone()
    n_ref = torch.nn.utils.clip_grad_norm_(b, max_norm=1.0)
    n_got = fo.clip_grad_norm_(a, max_norm=1.0)
    assert n_got.shape == () and n_got.is_cuda
    assert abs(float(n_got) / float(n_ref) - 1) < 1e-6
    for pa, pb in zip(a, b):
        assert _rel(pa.grad, pb.grad) < 1e-6
    # the self-cleaning workspace gives the same answer on a second call (norm of the clipped gradients)
    n2 = fo.clip_grad_norm_(a, max_norm=1.0)
    assert abs(float(n2) / float(torch.nn.utils.clip_grad_norm_(b, max_norm=1.0)) - 1) < 1e-6


def test_clip_skips_missing_grads_and_unaligned_views():
    flat = torch.randn(1000, device=DEV)
    p1 = torch.nn.Parameter(torch.randn(10, device=DEV)); p1.grad = flat[1:11]           # 4-byte aligned only
    p2 = torch.nn.Parameter(torch.randn(5, device=DEV))                                   # no gradient
    p3 = torch.nn.Parameter(torch.randn(301, device=DEV)); p3.grad = flat[13:314]
    ref = torch.sqrt((flat[1:11] ** 2).sum() + (flat[13:314] ** 2).sum())
    keep = flat.clone()
    n = fo.clip_grad_norm_([p1, p2, p3], max_norm=0.5)
    assert abs(float(n) / float(ref) - 1) < 1e-6
    coef = 0.5 / (float(ref) + 1e-6)
    assert torch.allclose(flat[1:11], keep[1:11] * coef, rtol=1e-6, atol=0)
    assert torch.allclose(flat[13:314], keep[13:314] * coef, rtol=1e-6, atol=0)
    assert torch.equal(flat[:1], keep[:1]) and torch.equal(flat[11:13], keep[11:13]) and torch.equal(flat[314:], keep[314:])


def test_adam_matches_torch_over_steps_with_scheduler():
    a = _make(SHAPES, 1)
    b = [torch.nn.Parameter(p.detach().clone()) for p in a]
    kw = dict(lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-5)                   # train.py:135-139
    oa, ob = fo.Adam(a, **kw), torch.optim.Adam(b, **kw)
    sa = torch.optim.lr_scheduler.CosineAnnealingLR(oa, T_max=10, eta_min=1e-7)          # train.py:141-145
    sb = torch.optim.lr_scheduler.CosineAnnealingLR(ob, T_max=10, eta_min=1e-7)
    g = torch.Generator().manual_seed(7)
    for it in range(6):
        for pa, pb in zip(a, b):
            gr = torch.randn(pa.shape, generator=g).to(DEV) * (10.0 if it % 2 else 0.01)
            pa.grad, pb.grad = gr.clone(), gr.clone()
        fo.clip_grad_norm_(a, 1.0); torch.nn.utils.clip_grad_norm_(b, 1.0)
        v0 = a[0]._version
        oa.step(); ob.step(); sa.step(); sb.step()
        assert a[0]._version > v0                                                        # packed-weight caches see the update
        for pa, pb in zip(a, b):
            assert _rel(pa, pb) < 2e-6, it
    for pa, pb in zip(a, b):
        assert _rel(oa.state[pa]["exp_avg"], ob.state[pb]["exp_avg"]) < 1e-5
        assert _rel(oa.state[pa]["exp_avg_sq"], ob.state[pb]["exp_avg_sq"]) < 1e-5
        assert float(oa.state[pa]["step"]) == float(ob.state[pb]["step"]) == 6.0


def test_adam_fused_clip_equals_separate_clip():
    a, b = _make(SHAPES, 2, 10.0), _make(SHAPES, 2, 10.0)
    oa, ob = fo.Adam(a, lr=1e-3, weight_decay=1e-5), fo.Adam(b, lr=1e-3, weight_decay=1e-5)
    nc = fo.compute_grad_norm([p.grad for p in a], 1.0)
    oa.step(grad_scale=nc[1:])
    fo.clip_grad_norm_(b, 1.0); ob.step()
    for pa, pb in zip(a, b):
        assert _rel(pa, pb) < 1e-6


def test_state_dict_interchangeable_with_torch_adam():
    a = _make(SHAPES, 3)
    b = [torch.nn.Parameter(p.detach().clone()) for p in a]
    for pa, pb in zip(a, b):
        pb.grad = pa.grad.clone()
    oa, ob = fo.Adam(a, lr=1e-3, weight_decay=1e-5), torch.optim.Adam(b, lr=1e-3, weight_decay=1e-5)
    oa.step(); ob.step()
    sd = copy.deepcopy(oa.state_dict())
    assert set(sd["state"][0].keys()) == set(ob.state_dict()["state"][0].keys())
    ob2 = torch.optim.Adam(b, lr=1e-3, weight_decay=1e-5); ob2.load_state_dict(sd)        # ours -> torch
    oa2 = fo.Adam(a, lr=1e-3, weight_decay=1e-5); oa2.load_state_dict(copy.deepcopy(ob.state_dict()))   # torch -> ours
    for pa, pb in zip(a, b):
        gr = torch.randn_like(pa)
        pa.grad, pb.grad = gr.clone(), gr.clone()
    oa2.step(); ob2.step()
    for pa, pb in zip(a, b):
        assert _rel(pa, pb) < 2e-6
    assert float(oa2.state[a[0]]["step"]) == 2.0


def test_more_than_64_tensors_and_errors():
    shapes = [(17,)] * 70 + [(5000,)]
    a = _make(shapes, 4, 5.0)
    b = [torch.nn.Parameter(p.detach().clone()) for p in a]
    for pa, pb in zip(a, b):
        pb.grad = pa.grad.clone()
    assert abs(float(fo.clip_grad_norm_(a, 1.0)) / float(torch.nn.utils.clip_grad_norm_(b, 1.0)) - 1) < 1e-6
    oa, ob = fo.Adam(a), torch.optim.Adam(b)
    oa.step(); ob.step()
    for pa, pb in zip(a, b):
        assert _rel(pa, pb) < 2e-6
    cpu = torch.nn.Parameter(torch.randn(4)); cpu.grad = torch.randn(4)
    with pytest.raises(RuntimeError):
        fo.clip_grad_norm_([cpu], 1.0)
    with pytest.raises(RuntimeError):
        fo.Adam([cpu]).step()
    half = torch.nn.Parameter(torch.randn(8, device=DEV).half()); half.grad = torch.randn(8, device=DEV).half()
    with pytest.raises(RuntimeError):
        fo.Adam([half]).step()
    with pytest.raises(ValueError):
        fo.Adam(a, amsgrad=True)
