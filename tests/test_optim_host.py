"""Host logic of fast_neural_style_transfer_b200.optim on CPU: the C-ABI calls are replaced by the numpy emulation
of tests/emu_optim.py (same documented semantics as include/fnst.h), everything above the ABI is the product's code.
Checked against torch.nn.utils.clip_grad_norm_ / torch.optim.Adam, i.e. train.py:135-145 and :203-206."""
import copy

import pytest
import torch

from fast_neural_style_transfer_b200 import optim as fo
import emu_optim

SHAPES = [(3,), (64,), (7, 5), (4097,), (16, 3, 9, 9)]


@pytest.fixture()
def emu(monkeypatch):
    lib = emu_optim.EmuLib()
    monkeypatch.setattr(fo, "lib", lib)
    monkeypatch.setattr(fo.ops, "_ctx", lambda t: (0, None))
    monkeypatch.setattr(fo, "_workspace", lambda device: torch.zeros(2, dtype=torch.float64))
    monkeypatch.setattr(fo, "_CLIP_LISTS", {})
    return lib


def _make(seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    ps = [torch.nn.Parameter(torch.randn(s, generator=g)) for s in SHAPES]
    for p in ps:
        p.grad = torch.randn(p.shape, generator=g) * scale
    return ps


def _twin(a):
    b = [torch.nn.Parameter(p.detach().clone()) for p in a]
    for pa, pb in zip(a, b):
        pb.grad = None if pa.grad is None else pa.grad.clone()
    return b


def test_no_cpu_fallback_without_the_library():
    p = torch.nn.Parameter(torch.randn(4)); p.grad = torch.randn(4)
    with pytest.raises(RuntimeError, match="CUDA"):
        fo.clip_grad_norm_([p], 1.0)
    with pytest.raises(RuntimeError, match="CUDA"):
        fo.Adam([p]).step()


def test_constructor_validation_matches_reference_usage():
    p = [torch.nn.Parameter(torch.randn(4))]
    fo.Adam(p, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-5)          # train.py:135-139
    for bad in (dict(amsgrad=True), dict(maximize=True), dict(decoupled_weight_decay=True), dict(lr=-1.0), dict(betas=(0.3, 0.999)),
                dict(eps=-1.0), dict(weight_decay=-1.0)):
        with pytest.raises(ValueError):
            fo.Adam(p, **bad)


@pytest.mark.parametrize("scale", [10.0, 1e-3])
def test_clip_grad_norm(emu, scale):
    a = _make(0, scale); a[2].grad = None
    b = _twin(a)
    n_ref = torch.nn.utils.clip_grad_norm_(b, 1.0)
    n_got = fo.clip_grad_norm_(a, 1.0)
    assert emu.calls == ["grad_norm", "grad_scale"]
    assert abs(float(n_got) / float(n_ref) - 1) < 1e-6
    for pa, pb in zip(a, b):
        if pb.grad is not None:
            assert torch.allclose(pa.grad, pb.grad, rtol=1e-6, atol=0)
    assert float(fo.clip_grad_norm_([torch.nn.Parameter(torch.zeros(2))], 1.0)) == 0.0     # no gradients at all
    with pytest.raises(RuntimeError):
        fo.clip_grad_norm_(a, 1.0, norm_type=1.0)


def test_training_tail_matches_torch_with_scheduler_and_resume(emu):
    a = _make(1); b = _twin(a)
    kw = dict(lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-5)
    oa, ob = fo.Adam(a, **kw), torch.optim.Adam(b, **kw)
    sa = torch.optim.lr_scheduler.CosineAnnealingLR(oa, T_max=8, eta_min=1e-7)
    sb = torch.optim.lr_scheduler.CosineAnnealingLR(ob, T_max=8, eta_min=1e-7)
    g = torch.Generator().manual_seed(5)

    def one_step(oa, ob, sa, sb, it):
        for pa, pb in zip(a, b):
            gr = torch.randn(pa.shape, generator=g) * (10.0 if it % 2 else 0.01)
            pa.grad, pb.grad = gr.clone(), gr.clone()
        fo.clip_grad_norm_(a, 1.0); torch.nn.utils.clip_grad_norm_(b, 1.0)
        v0 = a[0]._version
        oa.step(); ob.step(); sa.step(); sb.step()
        assert a[0]._version > v0
        for pa, pb in zip(a, b):
            assert torch.allclose(pa, pb, rtol=2e-6, atol=2e-7), it

    for it in range(4):
        one_step(oa, ob, sa, sb, it)
    assert oa.param_groups[0]["lr"] == ob.param_groups[0]["lr"] < 1e-3
    steps = {id(oa.state[p]["step"]) for p in a}
    assert len(steps) == 1 and float(oa.state[a[0]]["step"]) == 4.0            # one shared counter

    # checkpoint / resume across implementations (train.py:39-66, :269-283)
    sd_a, sd_b = copy.deepcopy(oa.state_dict()), copy.deepcopy(ob.state_dict())
    assert sd_a["state"].keys() == sd_b["state"].keys()
    assert set(sd_a["state"][0]) == set(sd_b["state"][0]) == {"step", "exp_avg", "exp_avg_sq"}
    assert set(sd_a["param_groups"][0]) == set(sd_b["param_groups"][0])
    oa2, ob2 = fo.Adam(a, **kw), torch.optim.Adam(b, **kw)
    oa2.load_state_dict(sd_b); ob2.load_state_dict(sd_a)                        # crossed on purpose
    sa2 = torch.optim.lr_scheduler.CosineAnnealingLR(oa2, T_max=4, eta_min=1e-7)
    sb2 = torch.optim.lr_scheduler.CosineAnnealingLR(ob2, T_max=4, eta_min=1e-7)
    for it in range(4, 7):
        one_step(oa2, ob2, sa2, sb2, it)
    assert float(oa2.state[a[0]]["step"]) == float(ob2.state[b[0]]["step"]) == 7.0


def test_fused_clip_scale_and_partial_grads(emu):
    a = _make(2, 10.0); b = _twin(a)
    a[1].grad = None; b[1].grad = None
    oa, ob = fo.Adam(a, lr=1e-3, weight_decay=1e-5), torch.optim.Adam(b, lr=1e-3, weight_decay=1e-5)
    nc = fo.compute_grad_norm([p.grad for p in a if p.grad is not None], 1.0)
    keep = a[0].grad.clone()
    oa.step(grad_scale=nc[1:])
    assert torch.equal(a[0].grad, keep)                                         # fused form leaves the gradients alone
    torch.nn.utils.clip_grad_norm_(b, 1.0); ob.step()
    for pa, pb in zip(a, b):
        assert torch.allclose(pa, pb, rtol=2e-6, atol=2e-7)
    assert len(oa.state[a[1]]) == 0                                             # untouched parameter: no state, like torch
