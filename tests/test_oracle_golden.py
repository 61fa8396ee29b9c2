"""Pin the CPU oracle against golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import stylenet_oracle as O


def rel_l2(a, b):
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.fixture(scope="module")
def net_p():
    return O.make_net_params(seed=0, random_affine=True)


@pytest.fixture(scope="module")
def vgg_p():
    return O.make_vgg_params(seed=1)


def test_param_inventory(net_p, vgg_p):
    # 58 tensors / 6 243 843 parameters; 11 VGG convs / 8 225 344 (2b)
    assert len(net_p) == 58
    assert sum(v.numel() for v in net_p.values()) == 6_243_843
    assert len(vgg_p) == 22
    assert sum(v.numel() for v in vgg_p.values()) == 8_225_344
    assert net_p["up1.upsample_conv.weight"].shape == (256, 64, 3, 3)
    assert net_p["final_conv.conv.bias"].shape == (3,)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_net_forward_matches_reference(golden_dir, net_p, tag):
    g = np.load(os.path.join(golden_dir, "net_forward.npz"))
    b, h, w = (int(v) for v in g[f"shape_{tag}"])
    x = O.make_image(b, h, w, seed=10 + b + h)
    with torch.no_grad():
        y = O.stylenet_forward(net_p, x)
    ref = g[f"y_{tag}"]
    assert tuple(y.shape) == ref.shape
    # output size law H' = 4*ceil(ceil(H/2)/2)  
    assert y.shape[2] == 4 * (-(-(-(-h // 2)) // 2)) and y.shape[3] == 4 * (-(-(-(-w // 2)) // 2))
    assert rel_l2(y, ref) < 2e-6


def test_vgg_and_losses_match_reference(golden_dir, vgg_p):
    g = np.load(os.path.join(golden_dir, "vgg_losses.npz"))
    x = O.make_image(2, 16, 24, seed=77, normalized=True)
    sty = O.make_image(1, 16, 16, seed=78, normalized=True)
    with torch.no_grad():
        feats = O.vgg_forward(vgg_p, x)
        targets = O.style_targets(vgg_p, sty)
        other = O.vgg_forward(vgg_p, O.make_image(2, 16, 24, seed=79, normalized=True))
    for i, f in enumerate(feats):
        assert rel_l2(f, g[f"feat{i}"]) < 2e-6, i
        assert float(f.min()) >= 0.0            # element 3 observed post-ReLU (in-place aliasing)
    for i in range(3):
        assert rel_l2(targets[i], g[f"target{i}"]) < 2e-6
    for i in (3, 4):
        assert rel_l2(targets[i][:16, :16], g[f"target{i}_corner"]) < 2e-6
        assert abs(float(targets[i].double().sum()) - float(g[f"target{i}_sum"])) <= 1e-5 * abs(float(g[f"target{i}_sum"]))
    assert rel_l2(O.gram_matrix(feats[0]), g["gram0"]) < 2e-6
    assert abs(float(O.style_loss(other, targets)) / float(g["style"]) - 1) < 1e-5
    assert abs(float(O.content_loss(other, feats)) / float(g["content"]) - 1) < 1e-5
    assert abs(float(O.total_variation_loss(x)) / float(g["tv"]) - 1) < 1e-5


def test_training_step_matches_reference(golden_dir, net_p, vgg_p):
    g = np.load(os.path.join(golden_dir, "train_step.npz"))
    b, h, w = 2, 32, 32
    content = O.make_image(b, h, w, seed=5, normalized=True)
    sty = O.make_image(1, h, w, seed=6, normalized=True)
    drop = O.make_dropout_scales(b, seed=7)
    targets = O.style_targets(vgg_p, sty)
    losses, grads = O.loss_and_grads(net_p, vgg_p, content, targets, drop)
    for k in ("total", "content", "style", "tv"):
        assert abs(float(losses[k]) / float(g[k]) - 1) < 2e-5, k
    assert rel_l2(losses["stylized"], g["stylized"]) < 2e-6
    gn_ref = float(g["grad_norm"])
    for k in grads:
        ref = float(g["gnorm/" + k])
        got = float(grads[k].double().norm())
        # conv biases under InstanceNorm have ~0 gradient (fp32 noise): absolute tolerance
        assert abs(got - ref) <= 2e-4 * ref + 1e-6 * gn_ref, (k, got, ref)
    for k in ("conv1.conv.weight", "norm2.weight", "up2.upsample_conv.weight", "final_conv.conv.weight", "final_conv.conv.bias"):
        assert rel_l2(grads[k], g["grad/" + k]) < 2e-4, k
    # clip + Adam (train.py:203-205)
    params = {k: v.clone() for k, v in net_p.items()}
    gnorm = O.clip_and_adam(params, grads, {}, step=1)
    assert abs(float(gnorm) / gn_ref - 1) < 2e-4
    for k in ("conv1.conv.weight", "norm2.weight", "norm2.bias", "final_conv.conv.weight", "res_blocks.4.in2.weight"):
        assert rel_l2(params[k], g["after/" + k]) < 1e-6, k
        # the update itself (lr-sized) must agree, not just the (dominant) old value
        upd, ref_upd = params[k] - net_p[k], torch.as_tensor(g["after/" + k]) - net_p[k]
        assert rel_l2(upd, ref_upd) < 5e-3, k


def test_dropout_values(golden_dir):
    g = np.load(os.path.join(golden_dir, "dropout.npz"))
    s = O.make_dropout_scales(3, seed=1)
    assert len(s) == 5 and s[0].shape == (3, 256)
    assert np.allclose(np.unique(torch.stack(s).numpy()), g["values"])


def test_reflection_index_law():
    # ReflectionPad2d: index -i -> i, (n-1)+i -> (n-1)-i, edge not repeated
    x = torch.arange(6.0).view(1, 1, 1, 6)
    y = torch.nn.functional.pad(x, (4, 4, 0, 0), mode="reflect").flatten().tolist()
    assert y == [4, 3, 2, 1, 0, 1, 2, 3, 4, 5, 4, 3, 2, 1]
