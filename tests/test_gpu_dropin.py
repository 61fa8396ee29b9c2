"""GPU parity of the drop-in Python surface (models.model / models.vgg19_net / losses.losses) vs the oracle."""
import os
import sys

import pytest
import torch

from oracle import stylenet_oracle as O

pytestmark = pytest.mark.gpu
DROPIN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "fast_neural_style_transfer_b200", "dropin")
DEV = "cuda"


def rel_l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.fixture(scope="module")
def dropin():
    sys.path.insert(0, DROPIN)
    for m in [k for k in sys.modules if k.split(".")[0] in ("models", "losses", "config")]:
        del sys.modules[m]
    import models.model as mm
    import models.vgg19_net as mv
    import losses.losses as ll
    yield mm, mv, ll
    sys.path.remove(DROPIN)


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("fp16", 1e-2)])
def test_inference_like_reference_script(dropin, precision, tol):
    """inference.py:33-48: build, load_state_dict, eval, no_grad, forward."""
    mm, _, _ = dropin
    p = O.make_net_params(seed=0)
    net = mm.StyleTransferNet().to(DEV)
    net.load_state_dict(p)
    net.precision = precision
    net.eval()
    x = O.make_image(1, 256, 256, seed=1234)
    with torch.no_grad():
        y = net(x.to(DEV))
        ref = O.stylenet_forward(p, x)
    assert y.shape == ref.shape and y.dtype == torch.float32
    assert rel_l2(y, ref) < tol
    if precision == "fp16":
        assert float((O.to_pixels(y.cpu()) - O.to_pixels(ref)).abs().max()) <= 1.0


def test_train_mode_dropout_uses_torch_rng(dropin):
    mm, _, _ = dropin
    p = O.make_net_params(seed=0)
    net = mm.StyleTransferNet().to(DEV)
    net.load_state_dict(p)
    net.precision = "fp32"
    net.train()
    x = O.make_image(2, 48, 48, seed=3).to(DEV)
    with torch.no_grad():
        torch.manual_seed(11); y1 = net(x)
        torch.manual_seed(11)
        ones = torch.ones((2, 256, 1, 1), device=DEV)
        drop = [torch.nn.functional.dropout2d(ones, 0.1, True).view(2, 256).cpu() for _ in range(5)]
        ref = O.stylenet_forward(p, x.cpu(), drop)
        torch.manual_seed(12); y2 = net(x)
    assert rel_l2(y1, ref) < 1e-4
    assert not torch.equal(y1, y2)
    net.eval()
    with torch.no_grad():
        assert rel_l2(net(x), O.stylenet_forward(p, x.cpu())) < 1e-4


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("fp16", 1e-2)])
def test_vgg_and_losses_forward(dropin, precision, tol):
    _, mv, ll = dropin
    vp = O.make_vgg_params(seed=1)
    vgg = mv.VGG19().to(DEV)
    vgg.load_state_dict(vp)
    vgg.precision = precision
    vgg.eval()
    x = O.make_image(2, 64, 64, seed=77, normalized=True)
    y = O.make_image(2, 64, 64, seed=79, normalized=True)
    sty = O.make_image(1, 64, 64, seed=78, normalized=True)
    with torch.no_grad():
        feats, other = vgg(x.to(DEV)), vgg(y.to(DEV))
        targets = [ll.gram_matrix(f).squeeze(0) for f in vgg(sty.to(DEV))]        # train.py:25-37
        rf, ro = O.vgg_forward(vp, x), O.vgg_forward(vp, y)
        rt = O.style_targets(vp, sty)
        assert [tuple(f.shape) for f in feats] == [tuple(f.shape) for f in rf]
        for i in range(5):
            assert rel_l2(targets[i], rt[i]) < tol, i
        s, c, tv = ll.style_loss(other, targets), ll.content_loss(other, feats), ll.total_variation_loss(x.to(DEV))
        assert s.dim() == 0 and s.dtype == torch.float32
        assert abs(float(s) / float(O.style_loss(ro, rt)) - 1) < tol
        assert abs(float(c) / float(O.content_loss(ro, rf)) - 1) < tol
        assert abs(float(tv) / float(O.total_variation_loss(x)) - 1) < 1e-5
        # plain NCHW fp32 tensors (not produced by the drop-in VGG) are accepted as well
        g = ll.gram_matrix(rf[1].to(DEV))
        assert rel_l2(g, O.gram_matrix(rf[1])) < 1e-5


def test_cuda_graph_replay_tracks_inputs_and_weights(dropin):
    """Small no-grad forwards are replayed from a CUDA graph: new inputs and new weights must be honoured."""
    mm, _, _ = dropin
    p0, p1 = O.make_net_params(seed=0), O.make_net_params(seed=5, random_affine=True)
    net = mm.StyleTransferNet().to(DEV)
    net.load_state_dict(p0)
    net.precision = "fp32"
    net.eval()
    xa, xb = O.make_image(1, 64, 64, seed=1), O.make_image(1, 64, 64, seed=2)
    with torch.no_grad():
        ya = net(xa.to(DEV)); yb = net(xb.to(DEV)); ya2 = net(xa.to(DEV))
        assert len(net._graphs) == 1
        assert rel_l2(ya, O.stylenet_forward(p0, xa)) < 1e-4 and rel_l2(yb, O.stylenet_forward(p0, xb)) < 1e-4
        assert rel_l2(ya, ya2) < 1e-5          # (fp32 atomics in the statistics: not bitwise reproducible)
        net.load_state_dict(p1)
        yc = net(xa.to(DEV))
        assert rel_l2(yc, O.stylenet_forward(p1, xa)) < 1e-4
        assert len(net._graphs) == 1          # the stale capture was dropped


def test_stylize_uint8_matches_reference_postprocessing(dropin):
    """uint8 in / uint8 out on the GPU == inference.py:44-60 (ToTensor, forward, de-normalise, clamp, ToPILImage).
    ToPILImage on a float tensor is `pic.mul(255).byte()` -- truncation, not rounding -- so the reference pixel is
    floor(clamp(...) * 255); compared here against torchvision's own to_pil_image."""
    import numpy as np
    from torchvision.transforms.functional import to_pil_image
    mm, _, _ = dropin
    p = O.make_net_params(seed=0)
    net = mm.StyleTransferNet().to(DEV); net.load_state_dict(p); net.precision = "fp32"; net.eval()
    g = torch.Generator().manual_seed(8)
    img = torch.randint(0, 256, (2, 40, 48, 3), generator=g, dtype=torch.uint8)
    out = net.stylize_uint8(img.to(DEV)).cpu()
    x = img.permute(0, 3, 1, 2).float() / 255.0                            # transforms.ToTensor()
    with torch.no_grad():
        y01 = O.to_pixels(O.stylenet_forward(p, x)) / 255.0                # de-normalise + clamp to [0,1] (inference.py:52-56)
    ref = torch.stack([torch.from_numpy(np.asarray(to_pil_image(y01[i]))) for i in range(y01.shape[0])])     # inference.py:59
    assert out.shape == ref.shape and out.dtype == torch.uint8
    diff = (out.float() - ref.float()).abs()
    # values within fp32 noise of an integer boundary may land on either side; a rounding implementation would be off by
    # one on about half of the pixels
    assert float(diff.max()) <= 1.0 and float((diff > 0).float().mean()) < 0.01


def test_pinned_host_batch_is_pipelined_and_equal_to_the_device_path(dropin):
    """net(pinned host batch) (extension: a CPU input is an error in the reference): chunked H2D / forward / D2H on three
    streams; the result must be the device path's result, for a batch that spans several chunks with a ragged last one."""
    mm, _, _ = dropin
    p = O.make_net_params(seed=0)
    net = mm.StyleTransferNet().to(DEV); net.load_state_dict(p); net.precision = "fp16"; net.eval()
    net.HOST_CHUNK_PIXELS = 3 * 64 * 64                                   # 3 images per chunk -> chunks of 3, 3, 1
    x = O.make_image(7, 64, 64, seed=21)
    with torch.no_grad():
        want = net(x.to(DEV)).cpu()
        got = net(x.pin_memory())
        one = net(x[:2].pin_memory())                                      # single chunk
    assert not got.is_cuda and got.is_pinned() and got.shape == want.shape
    # (not bit-equal: the InstanceNorm statistics of a tile-parallel launch are accumulated in a batch-dependent order, and
    #  from the stored fp16 values or the fp32 accumulators depending on the launch shape -- both inside the fp16 path's 2e-3 class)
    assert rel_l2(got, want) < 5e-3 and rel_l2(one, want[:2]) < 5e-3
    with torch.no_grad():
        ref = O.stylenet_forward(p, x)
    assert rel_l2(got, ref) < 1e-2 and rel_l2(want, ref) < 1e-2
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(x)                                                             # pageable host memory is still refused


def test_vgg_graph_hands_out_views_and_takes_gradients_in_place(dropin):
    """Zero-copy hand-offs around the captured VGG graphs (autograd_fns.VGGGraph): the features are views of the graph's static
    output (two instances alternate while the caller still holds last step's features, as the reference loop does,
    train.py:178-185); from the second backward on the loss nodes write their gradients straight into the captured backward's
    input buffers.  Values: features of earlier steps stay intact while held, and every step's input gradient matches the
    oracle's autograd."""
    from fast_neural_style_transfer_b200 import autograd_fns
    _, mv, ll = dropin
    vp = O.make_vgg_params(seed=1)
    vgg = mv.VGG19().to(DEV)
    vgg.load_state_dict(vp)
    vgg.precision = "fp32"
    vgg.eval()
    sty = O.make_image(1, 32, 32, seed=78, normalized=True)
    with torch.no_grad():
        targets = [ll.gram_matrix(f).squeeze(0) for f in vgg(sty.to(DEV))]
    rt = O.style_targets(vp, sty)
    ptrs, kept = [], []
    feats = None
    for step in range(4):
        content = O.make_image(2, 32, 32, seed=100 + step, normalized=True)
        x = O.make_image(2, 32, 32, seed=200 + step, normalized=True)
        xd = x.to(DEV).requires_grad_(True)
        with torch.no_grad():
            cf = vgg(content.to(DEV))
        feats = vgg(xd)                                  # (the previous step's `feats` is alive during this call)
        ptrs.append(feats[0].data_ptr())
        loss = 1000.0 * ll.content_loss(feats, cf) + ll.style_loss(feats, targets)
        state = next(s for lst in vgg._graphs.values() for s in lst if s.with_tape and s.fwd.outputs.data_ptr() == feats[0].data_ptr())
        if step >= 2:                                    # this instance has captured its backward: all four gradient slots are claimed
            assert state.claimed == {0, 1, 2, 4}
        loss.backward()
        xr = x.clone().requires_grad_(True)
        rf = O.vgg_forward(vp, xr)
        (1000.0 * O.content_loss(rf, O.vgg_forward(vp, content)) + O.style_loss(rf, rt)).backward()
        assert rel_l2(xd.grad, xr.grad) < 2e-3, step                 # (a stale or doubly-written gradient buffer would be an O(1) error)
        for i in (0, 2, 4):
            assert rel_l2(feats[i], rf[i]) < 1e-4
        kept.append((feats[2], rf[2].detach()))
        if len(kept) > 1:                                # the features of the previous step, still referenced: untouched by this replay
            assert rel_l2(kept[-2][0], kept[-2][1]) < 1e-4
            kept.pop(0)
    taped = [s for lst in vgg._graphs.values() for s in lst if s.with_tape]
    assert len(taped) == 2 and ptrs[0] == ptrs[2] and ptrs[1] == ptrs[3] and ptrs[0] != ptrs[1]
