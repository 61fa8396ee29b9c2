"""The drop-in modules keep the reference's Python surface: class names, child names,
58 state-dict keys / shapes, default initialisation, picklability.  CPU only (no compute calls)."""
import io
import os
import pickle
import sys

import pytest
import torch

from oracle import stylenet_oracle as O

DROPIN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "fast_neural_style_transfer_b200", "dropin")


@pytest.fixture(scope="module")
def dropin():
    sys.path.insert(0, DROPIN)
    for m in [k for k in sys.modules if k == "models" or k.startswith("models.") or k == "losses" or k.startswith("losses.")]:
        del sys.modules[m]
    import models.model as mm
    import models.vgg19_net as mv
    import losses.losses as ll
    yield mm, mv, ll
    sys.path.remove(DROPIN)
    for m in [k for k in sys.modules if k == "models" or k.startswith("models.") or k == "losses" or k.startswith("losses.") or k == "config"]:
        del sys.modules[m]


def test_state_dict_contract(dropin):
    mm, _, _ = dropin
    net = mm.StyleTransferNet()
    ref = O.make_net_params(seed=0)
    sd = net.state_dict()
    assert list(sd.keys()) == list(ref.keys())           # same names, same order (models/model.py:25-47)
    for k in ref:
        assert sd[k].shape == ref[k].shape, k
    net.load_state_dict(ref)                             # strict
    assert sum(p.numel() for p in net.parameters()) == 6_243_843
    assert [n for n, _ in net.named_children()] == ["conv1", "norm1", "conv2", "norm2", "res_blocks", "up1", "norm3", "up2", "norm4", "final_conv"]
    assert isinstance(net.res_blocks[0].dropout, torch.nn.Dropout2d) and net.res_blocks[0].dropout.p == 0.1


def test_default_init_matches_torch_modules(dropin):
    """Same construction order + same torch default init => same weights as the reference under one seed."""
    mm, _, _ = dropin
    torch.manual_seed(0)
    net = mm.StyleTransferNet()
    torch.manual_seed(0)
    c1 = torch.nn.Conv2d(3, 64, 9, stride=2)             # first module the reference constructs (models/model.py:28)
    assert torch.equal(net.conv1.conv.weight, c1.weight) and torch.equal(net.conv1.conv.bias, c1.bias)
    assert torch.equal(net.norm1.weight, torch.ones(64))


def test_pickle_and_cpu_guard(dropin):
    mm, _, ll = dropin
    net = mm.StyleTransferNet()
    buf = io.BytesIO()
    torch.save(net, buf)                                 # whole-module pickle, train.py:297
    buf.seek(0)
    net2 = torch.load(buf, weights_only=False)
    assert torch.equal(net2.final_conv.conv.weight, net.final_conv.conv.weight)
    with pytest.raises(RuntimeError, match="CUDA"):
        net(torch.zeros(1, 3, 32, 32))                   # no CPU fallback
    with pytest.raises(RuntimeError, match="CUDA"):
        ll.total_variation_loss(torch.zeros(1, 3, 8, 8))


def test_every_declared_symbol_is_exported_and_bound():
    """include/fnst.h is the contract: every declared entry point must be exported by libfnst.so and bound in _lib.EXPORTS."""
    import re
    from fast_neural_style_transfer_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "fnst.h")).read()
    declared = set(re.findall(r"^(?:int|int64_t|const char\*)\s+(fnst_\w+)\s*\(", header, flags=re.M))
    assert len(declared) >= 25
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(_lib.lib, name), name


def test_vgg_contract(dropin):
    _, mv, _ = dropin
    vgg = mv.VGG19()
    ref = O.make_vgg_params(seed=1)
    assert sorted(vgg.state_dict().keys()) == sorted(ref.keys())
    vgg.load_state_dict(ref)
    assert all(not p.requires_grad for p in vgg.parameters())
    assert hasattr(vgg, "slice5")


def test_vgg_weight_policy(dropin, monkeypatch, tmp_path):
    """The reference builds vgg19(weights='DEFAULT') (models/vgg19_net.py:27).  The drop-in must never fall back to random
    weights silently: offline override file -> explicit random opt-in -> torchvision's download, else raise."""
    import torchvision.models as tvm
    _, mv, _ = dropin
    calls = []

    def fake_vgg19(weights=None, **kw):
        calls.append(weights)
        if weights is not None:
            raise OSError("no network in this test")
        return _REAL_VGG19(weights=None)

    global _REAL_VGG19
    _REAL_VGG19 = tvm.vgg19
    monkeypatch.setattr(tvm, "vgg19", fake_vgg19)
    monkeypatch.delenv("FNST_VGG19_RANDOM_INIT", raising=False)
    monkeypatch.delenv("FNST_VGG19_WEIGHTS", raising=False)
    with pytest.raises(RuntimeError, match="FNST_VGG19_WEIGHTS"):
        mv.VGG19()
    assert calls == ["DEFAULT"]                                   # tried the reference's own source first
    # offline override: a torchvision state dict on disk
    torch.manual_seed(5)
    src = _REAL_VGG19(weights=None)
    path = tmp_path / "vgg19.pth"
    torch.save(src.state_dict(), path)
    monkeypatch.setenv("FNST_VGG19_WEIGHTS", str(path))
    vgg = mv.VGG19()
    assert torch.equal(vgg.slice5[1].weight, src.features[23].weight)
    # explicit opt-in to random init
    monkeypatch.delenv("FNST_VGG19_WEIGHTS")
    monkeypatch.setenv("FNST_VGG19_RANDOM_INIT", "1")
    assert mv.VGG19().precision == "bf16"                         # loss network default: bf16 (gradients need its range)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree not present on this machine")
def test_same_seed_gives_the_reference_weights(dropin):
    """All 58 tensors: the drop-in constructed under a seed equals the unmodified reference module constructed under the
    same seed (same nn modules in the same construction order, models/model.py:25-47) -- SURVEY 8b 'same default init'."""
    import importlib.util
    mm, _, _ = dropin
    spec = importlib.util.spec_from_file_location("_reference_model", "/root/reference/models/model.py")
    ref_mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_mod)
    torch.manual_seed(123)
    ref = ref_mod.StyleTransferNet().state_dict()
    torch.manual_seed(123)
    got = mm.StyleTransferNet().state_dict()
    assert list(ref) == list(got) and len(ref) == 58
    for k in ref:
        assert torch.equal(ref[k], got[k]), k


def test_export_switch_traces_the_reference_graph(dropin, tmp_path):
    """model_scripting/torchscript_model.py:25 (`torch.jit.trace(net, torch.rand(1,3,256,256), strict=False)`) on the drop-in:
    ctypes calls cannot be traced, so while an export runs the modules evaluate through their own stock children -- the traced
    module computes the reference function (here: equal to the oracle on CPU) and loads without the package."""
    mm, _, _ = dropin
    torch.manual_seed(3)
    net = mm.StyleTransferNet().eval()
    x = torch.rand(1, 3, 64, 64)
    traced = torch.jit.trace(net, x, strict=False)
    path = tmp_path / "model_traced.pt"
    traced.save(str(path))
    again = torch.jit.load(str(path), map_location="cpu")
    x2 = torch.rand(1, 3, 64, 64)
    with torch.no_grad():
        ref = O.stylenet_forward({k: v.detach() for k, v in net.state_dict().items()}, x2)
        got = again(x2)
    assert float((got - ref).norm() / ref.norm()) < 1e-5
    assert net.to_reference_module() is net
    # outside an export the same call refuses CPU tensors (no CPU fallback at run time)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(x)
    for sub in (net.conv1, net.res_blocks[0], net.up1):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            sub(torch.rand(1, sub.conv.in_channels if hasattr(sub, "conv") else 256, 16, 16))
