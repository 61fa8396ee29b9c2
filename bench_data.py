"""Synthetic weights and images for bench.py's B200 arm (seeded, shape- and distribution-faithful to the
reference's defaults).  Kept outside `oracle/` so the measured product path never imports the checker;
`tests/test_bench_data.py` asserts these generators are bit-identical to the oracle's, so both arms of the
bench and the parity tests see the same numbers.

Weights: PyTorch's default Conv2d / ConvTranspose2d init distribution, uniform(-1/sqrt(fan_in), 1/sqrt(fan_in)) for
weight and bias (models/model.py:25-47 instantiates plain nn modules); InstanceNorm2d(affine=True) = (1, 0);
VGG-19: torchvision's kaiming_normal_(fan_out, relu) (models/vgg19_net.py:26-27 with no weights available) plus a
small random bias.  Images: rand in [0,1] (inference.py:28-31), optionally ImageNet-normalised (train.py:92-102).
"""
import math
from typing import Dict

import torch

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def _net_layout():
    """(state-dict prefix, kind, in, out, kernel) in the construction order of models/model.py:25-47."""
    rows = [("conv1.conv", "conv", 3, 64, 9), ("norm1", "in", 64, 64, 0), ("conv2.conv", "conv", 64, 256, 3), ("norm2", "in", 256, 256, 0)]
    for i in range(5):
        for j in (1, 2):
            rows.append((f"res_blocks.{i}.conv{j}.conv", "conv", 256, 256, 3))
            rows.append((f"res_blocks.{i}.in{j}", "in", 256, 256, 0))
    rows += [("up1.upsample_conv", "convT", 256, 64, 3), ("norm3", "in", 64, 64, 0), ("up2.upsample_conv", "convT", 64, 32, 3),
             ("norm4", "in", 32, 32, 0), ("final_conv.conv", "conv", 32, 3, 9)]
    return rows


_VGG_LAYOUT = (("slice1.0", 3, 64), ("slice1.2", 64, 64), ("slice2.5", 64, 128), ("slice2.7", 128, 128), ("slice3.10", 128, 256),
               ("slice3.12", 256, 256), ("slice3.14", 256, 256), ("slice4.16", 256, 256), ("slice4.19", 256, 512),
               ("slice4.21", 512, 512), ("slice5.23", 512, 512))


def net_state_dict(seed: int = 0) -> Dict[str, torch.Tensor]:
    gen = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    for prefix, kind, cin, cout, k in _net_layout():
        if kind == "in":
            sd[prefix + ".weight"], sd[prefix + ".bias"] = torch.ones(cout), torch.zeros(cout)
            continue
        # ConvTranspose2d stores (in, out, k, k) and torch derives fan_in from size(1)
        shape, fan_in = ((cout, cin, k, k), cin * k * k) if kind == "conv" else ((cin, cout, k, k), cout * k * k)
        bound = 1.0 / math.sqrt(fan_in)
        sd[prefix + ".weight"] = ((torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1) * bound).float()
        sd[prefix + ".bias"] = ((torch.rand(cout, generator=gen, dtype=torch.float64) * 2 - 1) * bound).float()
    return sd


def vgg_state_dict(seed: int = 1) -> Dict[str, torch.Tensor]:
    gen = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    for prefix, cin, cout in _VGG_LAYOUT:
        std = math.sqrt(2.0 / (cout * 9))
        sd[prefix + ".weight"] = (torch.randn((cout, cin, 3, 3), generator=gen, dtype=torch.float64) * std).float()
        sd[prefix + ".bias"] = (torch.randn(cout, generator=gen, dtype=torch.float64) * 0.05).float()
    return sd


def image_batch(batch: int, h: int, w: int, seed: int = 1234, normalized: bool = False) -> torch.Tensor:
    gen = torch.Generator().manual_seed(seed)
    x = torch.rand((batch, 3, h, w), generator=gen, dtype=torch.float64)
    if normalized:
        x = (x - torch.tensor(IMAGENET_MEAN, dtype=torch.float64).view(1, 3, 1, 1)) / torch.tensor(IMAGENET_STD, dtype=torch.float64).view(1, 3, 1, 1)
    return x.float()
