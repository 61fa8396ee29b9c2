"""CPU oracle for the style-transfer hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, as plain functional PyTorch-on-CPU code, the arithmetic of the
reference's hot path (HajarHAMDOUCH01/Fast-neural-style-transfer).  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py`
may import it; the product package never does (it fails loudly without its CUDA library).

Parity pin: the reference ships no tests, golden vectors or fixtures, so
this oracle is pinned against *outputs of the reference itself*: `tests/golden/make_golden.py`
imports the unmodified reference modules from /root/reference in the build container, feeds
them the parameters generated here, and stores their outputs in `tests/golden/*.npz`;
`tests/test_oracle_golden.py` checks every function below against those files.

Every function cites the reference file:line it follows.  The arithmetic itself lives in
third-party PyTorch (unpinned in the reference's requirements.txt:1-3; torch 2.11.0 /
torchvision 0.26.0 in this image); the functional `torch.nn.functional` calls used here are
the published semantics of the `nn.Module`s the reference instantiates.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

IMAGENET_MEAN = (0.485, 0.456, 0.406)   # train.py:93-96
IMAGENET_STD = (0.229, 0.224, 0.225)

# ----------------------------------------------------------------------------------------
# Parameter generators (deterministic, independent of nn.Module construction order)
# ----------------------------------------------------------------------------------------

# (state-dict prefix, kind, in_ch, out_ch, kernel) in models/model.py:25-47 construction order.
NET_LAYERS = (
    [("conv1.conv", "conv", 3, 64, 9), ("norm1", "in", 64, 64, 0),
     ("conv2.conv", "conv", 64, 256, 3), ("norm2", "in", 256, 256, 0)]
    + [item for i in range(5) for item in (
        (f"res_blocks.{i}.conv1.conv", "conv", 256, 256, 3), (f"res_blocks.{i}.in1", "in", 256, 256, 0),
        (f"res_blocks.{i}.conv2.conv", "conv", 256, 256, 3), (f"res_blocks.{i}.in2", "in", 256, 256, 0))]
    + [("up1.upsample_conv", "convT", 256, 64, 3), ("norm3", "in", 64, 64, 0),
       ("up2.upsample_conv", "convT", 64, 32, 3), ("norm4", "in", 32, 32, 0),
       ("final_conv.conv", "conv", 32, 3, 9)]
)

# torchvision vgg19().features indices of the 11 convs the reference runs, with the
# nn.Sequential slice that owns each (models/vgg19_net.py:38-51).
VGG_CONVS = (
    ("slice1.0", 3, 64), ("slice1.2", 64, 64),
    ("slice2.5", 64, 128), ("slice2.7", 128, 128),
    ("slice3.10", 128, 256), ("slice3.12", 256, 256), ("slice3.14", 256, 256),
    ("slice4.16", 256, 256), ("slice4.19", 256, 512), ("slice4.21", 512, 512),
    ("slice5.23", 512, 512),
)


def make_net_params(seed: int = 0, dtype=torch.float32, random_affine: bool = False) -> Dict[str, torch.Tensor]:
    """58 tensors with the reference's state-dict names/shapes (models/model.py:25-47).

    Values follow PyTorch's default conv init distribution (uniform +-1/sqrt(fan_in), both
    weight and bias; for ConvTranspose2d fan_in is computed from weight.size(1)*k*k) and
    InstanceNorm2d(affine=True) init (weight 1, bias 0) unless `random_affine`.
    """
    g = torch.Generator().manual_seed(seed)
    p: Dict[str, torch.Tensor] = {}
    for name, kind, cin, cout, k in NET_LAYERS:
        if kind == "in":
            if random_affine:
                p[name + ".weight"] = (0.5 + torch.rand(cout, generator=g, dtype=torch.float64)).to(dtype)
                p[name + ".bias"] = (torch.rand(cout, generator=g, dtype=torch.float64) - 0.5).to(dtype)
            else:
                p[name + ".weight"] = torch.ones(cout, dtype=dtype)
                p[name + ".bias"] = torch.zeros(cout, dtype=dtype)
            continue
        if kind == "conv":
            shape, fan_in = (cout, cin, k, k), cin * k * k
        else:  # ConvTranspose2d weight is (in, out, k, k); torch's fan_in uses size(1)
            shape, fan_in = (cin, cout, k, k), cout * k * k
        bound = 1.0 / math.sqrt(fan_in)
        p[name + ".weight"] = ((torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * bound).to(dtype)
        p[name + ".bias"] = ((torch.rand(cout, generator=g, dtype=torch.float64) * 2 - 1) * bound).to(dtype)
    return p


def make_vgg_params(seed: int = 1, dtype=torch.float32, random_bias: bool = True) -> Dict[str, torch.Tensor]:
    """22 tensors named like the reference VGG19 state dict (`slice1.0.weight`, ...).

    torchvision's random init is kaiming_normal_(fan_out, relu) with zero bias; a small random
    bias is added by default so the bias path is exercised by parity tests.
    """
    g = torch.Generator().manual_seed(seed)
    p: Dict[str, torch.Tensor] = {}
    for name, cin, cout in VGG_CONVS:
        std = math.sqrt(2.0 / (cout * 9))
        p[name + ".weight"] = (torch.randn((cout, cin, 3, 3), generator=g, dtype=torch.float64) * std).to(dtype)
        b = torch.randn(cout, generator=g, dtype=torch.float64) * 0.05 if random_bias else torch.zeros(cout, dtype=torch.float64)
        p[name + ".bias"] = b.to(dtype)
    return p


def make_image(batch: int, h: int, w: int, seed: int = 1234, normalized: bool = False, dtype=torch.float32) -> torch.Tensor:
    """Synthetic image batch: rand in [0,1] (inference.py:28-31) or ImageNet-normalised (train.py:92-102)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand((batch, 3, h, w), generator=g, dtype=torch.float64)
    if normalized:
        mean = torch.tensor(IMAGENET_MEAN, dtype=torch.float64).view(1, 3, 1, 1)
        std = torch.tensor(IMAGENET_STD, dtype=torch.float64).view(1, 3, 1, 1)
        x = (x - mean) / std
    return x.to(dtype)


def make_dropout_scales(batch: int, seed: int, p_drop: float = 0.1) -> List[torch.Tensor]:
    """Five per-(n,c) Dropout2d scale tensors of shape (B,256): 0 or 1/(1-p) (models/model.py:84,88)."""
    g = torch.Generator().manual_seed(seed)
    keep = 1.0 - p_drop
    return [(torch.rand((batch, 256), generator=g) < keep).to(torch.float32) / keep for _ in range(5)]


# ----------------------------------------------------------------------------------------
# StyleTransferNet (models/model.py)
# ----------------------------------------------------------------------------------------

def conv_layer(x, w, b, stride: int):
    """ConvLayer.forward, models/model.py:74-75: ReflectionPad2d(k//2) then Conv2d(padding=0)."""
    pad = w.shape[-1] // 2
    return F.conv2d(F.pad(x, (pad, pad, pad, pad), mode="reflect"), w, b, stride=stride)


def upsample_conv(x, w, b):
    """UpsampleConv.forward, models/model.py:21-22: ConvTranspose2d(k=3, stride=2, padding=1, output_padding=1)."""
    return F.conv_transpose2d(x, w, b, stride=2, padding=1, output_padding=1)


def instance_norm(x, gamma, beta, eps: float = 1e-5):
    """nn.InstanceNorm2d(C, affine=True), models/model.py:29,32,41,44,81,83: biased variance, eps 1e-5."""
    mu = x.mean(dim=(2, 3), keepdim=True)
    var = x.var(dim=(2, 3), unbiased=False, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * gamma.view(1, -1, 1, 1) + beta.view(1, -1, 1, 1)


def residual_block(p, prefix: str, x, drop_scale: Optional[torch.Tensor]):
    """ResidualBlock.forward, models/model.py:86-90 (dropout scale (B,C) or None for eval)."""
    y = F.relu(instance_norm(conv_layer(x, p[prefix + ".conv1.conv.weight"], p[prefix + ".conv1.conv.bias"], 1),
                             p[prefix + ".in1.weight"], p[prefix + ".in1.bias"]))
    if drop_scale is not None:
        y = y * drop_scale.to(y.dtype).view(y.shape[0], y.shape[1], 1, 1)
    y = instance_norm(conv_layer(y, p[prefix + ".conv2.conv.weight"], p[prefix + ".conv2.conv.bias"], 1),
                      p[prefix + ".in2.weight"], p[prefix + ".in2.bias"])
    return x + y


def stylenet_forward(p: Dict[str, torch.Tensor], x: torch.Tensor,
                     drop_scales: Optional[Sequence[torch.Tensor]] = None) -> torch.Tensor:
    """StyleTransferNet.forward, models/model.py:49-65.  `drop_scales`: five (B,256) tensors for
    `.train()` mode, None for `.eval()`."""
    h = F.relu(instance_norm(conv_layer(x, p["conv1.conv.weight"], p["conv1.conv.bias"], 2),
                             p["norm1.weight"], p["norm1.bias"]))
    h = F.relu(instance_norm(conv_layer(h, p["conv2.conv.weight"], p["conv2.conv.bias"], 2),
                             p["norm2.weight"], p["norm2.bias"]))
    for i in range(5):
        h = residual_block(p, f"res_blocks.{i}", h, None if drop_scales is None else drop_scales[i])
    h = F.relu(instance_norm(upsample_conv(h, p["up1.upsample_conv.weight"], p["up1.upsample_conv.bias"]),
                             p["norm3.weight"], p["norm3.bias"]))
    h = F.relu(instance_norm(upsample_conv(h, p["up2.upsample_conv.weight"], p["up2.upsample_conv.bias"]),
                             p["norm4.weight"], p["norm4.bias"]))
    return conv_layer(h, p["final_conv.conv.weight"], p["final_conv.conv.bias"], 1)


# ----------------------------------------------------------------------------------------
# VGG-19 features[0:25] (models/vgg19_net.py:56-65)
# ----------------------------------------------------------------------------------------

def vgg_forward(p: Dict[str, torch.Tensor], x: torch.Tensor) -> List[torch.Tensor]:
    """VGG19.forward, models/vgg19_net.py:56-65.  Returns [relu1_2, relu2_2, relu3_3, relu4_2, relu4_3].

    Element 3 is sliced as the pre-ReLU conv4_2 (features[21]) but torchvision's ReLU is
    inplace, so features[22] (first op of slice5) overwrites it: callers observe relu4_2."""
    def cr(t, name):
        return F.relu(F.conv2d(t, p[name + ".weight"], p[name + ".bias"], padding=1))
    h = cr(cr(x, "slice1.0"), "slice1.2")
    f0 = h
    h = cr(cr(F.max_pool2d(h, 2, 2), "slice2.5"), "slice2.7")
    f1 = h
    h = cr(cr(cr(F.max_pool2d(h, 2, 2), "slice3.10"), "slice3.12"), "slice3.14")
    f2 = h
    h = cr(h, "slice4.16")
    h = cr(cr(F.max_pool2d(h, 2, 2), "slice4.19"), "slice4.21")
    f3 = h
    f4 = cr(h, "slice5.23")
    return [f0, f1, f2, f3, f4]


# ----------------------------------------------------------------------------------------
# Losses (losses/losses.py)
# ----------------------------------------------------------------------------------------

def gram_matrix(feat: torch.Tensor) -> torch.Tensor:
    """losses/losses.py:6-13: un-normalised bmm(F, F^T) on (b, c, h*w)."""
    b, c, h, w = feat.shape
    f = feat.reshape(b, c, h * w)
    return torch.bmm(f, f.transpose(1, 2))


STYLE_LAYERS = ((0, 0.25), (1, 0.3), (2, 0.45))   # zip([0,1,2,4],[.25,.3,.45]) losses/losses.py:18-24


def style_loss(features: Sequence[torch.Tensor], target_grams: Sequence[torch.Tensor]):
    """losses/losses.py:15-44: sum_l w_l * SSE(G_l, G*_l) / c^2 (SSE summed over the batch)."""
    total = 0.0
    for idx, weight in STYLE_LAYERS:
        g = gram_matrix(features[idx])
        tgt = target_grams[idx]
        c = tgt.shape[0]                              # taken before unsqueeze, losses.py:30
        if tgt.dim() == 2:
            tgt = tgt.unsqueeze(0)
        tgt = tgt.expand_as(g)
        total = total + weight * ((g - tgt) ** 2).sum() / (c * c)
    return total


def content_loss(features: Sequence[torch.Tensor], target_features: Sequence[torch.Tensor]):
    """losses/losses.py:46-60: SSE(feat[4], target[4]) / (c*h*w) -- not divided by batch."""
    a, t = features[4], target_features[4]
    _, c, h, w = a.shape
    return ((a - t) ** 2).sum() / (c * h * w)


def total_variation_loss(img: torch.Tensor):
    """losses/losses.py:62-73."""
    b, c, h, w = img.shape
    tv_h = ((img[:, :, 1:, :] - img[:, :, :-1, :]) ** 2).sum()
    tv_w = ((img[:, :, :, 1:] - img[:, :, :, :-1]) ** 2).sum()
    return (tv_h + tv_w) / (b * c * h * w)


def style_targets(vgg_p, style_img: torch.Tensor) -> List[torch.Tensor]:
    """get_style_targets, train.py:25-37: five (C,C) Gram matrices of a 1xCxHxW style image."""
    with torch.no_grad():
        return [gram_matrix(f).squeeze(0) for f in vgg_forward(vgg_p, style_img)]


# ----------------------------------------------------------------------------------------
# Training step (train.py:168-206)
# ----------------------------------------------------------------------------------------

def perceptual_losses(net_p, vgg_p, content: torch.Tensor, targets, drop_scales=None,
                      content_weight: float = 1000.0, style_weight: float = 1.0, tv_weight: float = 10.0):
    """train.py:171-190: forward, clamp(-3,3), two VGG passes, weighted loss sum."""
    y = torch.clamp(stylenet_forward(net_p, content, drop_scales), -3, 3)
    with torch.no_grad():
        content_feats = vgg_forward(vgg_p, content)
    feats = vgg_forward(vgg_p, y)
    c = content_loss(feats, content_feats)
    s = style_loss(feats, targets)
    tv = total_variation_loss(y)
    total = content_weight * c + style_weight * s + tv_weight * tv
    return total, c, s, tv, y


def loss_and_grads(net_p, vgg_p, content, targets, drop_scales=None, **weights):
    """train.py:199-200: zero_grad + total_loss.backward(); returns losses and the 58 gradients."""
    leaf = {k: v.detach().clone().requires_grad_(True) for k, v in net_p.items()}
    total, c, s, tv, y = perceptual_losses(leaf, vgg_p, content, targets, drop_scales, **weights)
    total.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaf.items()}
    return {"total": total.detach(), "content": c.detach(), "style": s.detach(), "tv": tv.detach(),
            "stylized": y.detach()}, grads


def clip_and_adam(params, grads, state, step: int, lr: float = 1e-3, max_norm: float = 1.0,
                  betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-5):
    """train.py:203-205: clip_grad_norm_(1.0) then Adam(lr, betas, eps, weight_decay=1e-5) (coupled L2).

    `state` maps name -> (exp_avg, exp_avg_sq); `step` is 1-based.  Returns the global grad norm."""
    total_norm = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).to(torch.float32)
    coef = torch.clamp(max_norm / (total_norm + 1e-6), max=1.0)
    b1, b2 = betas
    for k, w in params.items():
        g = grads[k] * coef + weight_decay * w
        m, v = state.setdefault(k, (torch.zeros_like(w), torch.zeros_like(w)))
        m.mul_(b1).add_(g, alpha=1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = (v.sqrt() / math.sqrt(1 - b2 ** step)).add_(eps)
        w.addcdiv_(m, denom, value=-(lr / (1 - b1 ** step)))
    return total_norm


def cosine_lr(step: int, total_steps: int, base_lr: float = 1e-3, eta_min: float = 1e-7) -> float:
    """CosineAnnealingLR(T_max=total_steps, eta_min=1e-7) closed form, train.py:141-145."""
    return eta_min + (base_lr - eta_min) * (1 + math.cos(math.pi * step / total_steps)) / 2


def to_pixels(y: torch.Tensor) -> torch.Tensor:
    """inference.py:52-57: de-normalise, clamp to [0,1]; scaled to the 0-255 pixel range."""
    mean = torch.tensor(IMAGENET_MEAN, dtype=y.dtype).view(1, 3, 1, 1)
    std = torch.tensor(IMAGENET_STD, dtype=y.dtype).view(1, 3, 1, 1)
    return torch.clamp(y * std + mean, 0, 1) * 255.0
