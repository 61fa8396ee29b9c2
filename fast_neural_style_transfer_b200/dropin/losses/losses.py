"""Drop-in for the reference's losses/losses.py: gram_matrix, style_loss, content_loss,
total_variation_loss with the reference's signatures and normalisations (losses/losses.py:6-73),
computed by libfnst reductions.  All are differentiable with respect to their first argument."""
import os
import sys

import torch

_PKG_PARENT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if _PKG_PARENT not in sys.path:
    sys.path.append(_PKG_PARENT)

from fast_neural_style_transfer_b200 import autograd_fns as _fn   # noqa: E402


def gram_matrix(input_feat):
    """(b,c,h,w) -> (b,c,c) fp32, un-normalised F F^T (losses/losses.py:6-13)."""
    return _fn.gram(input_feat)


def style_loss(input_features, target_grams):
    """sum over (idx, weight) in zip([0,1,2,4], [.25,.3,.45]) of weight * SSE(G, G*) / c^2
    (losses/losses.py:15-44; the zip stops after three layers)."""
    pairs = list(zip([0, 1, 2, 4], [0.25, 0.3, 0.45]))
    feats, targets, cs = [], [], []
    for idx, _ in pairs:
        target, feat = target_grams[idx], input_features[idx]
        if target.dim() == 3 and target.size(0) not in (1, feat.size(0)):
            raise RuntimeError("target gram batch does not match input batch")
        cs.append(target.shape[0])          # taken before the unsqueeze in the reference (losses.py:30): C for (C,C) targets
        feats.append(feat)
        targets.append(target.squeeze(0) if target.dim() == 3 and target.size(0) == 1 else target)
    return _fn.style_loss_fused(feats, targets, [w for _, w in pairs], cs)


def content_loss(input_features, target_features):
    """SSE(feat[4], target[4]) / (c*h*w) (losses/losses.py:46-60)."""
    a, t = input_features[4], target_features[4]
    _, c, h, w = a.size()
    return _fn.sse(a, t, scale=1.0 / (c * h * w))          # normalisation folded into the reduction kernel


def total_variation_loss(img):
    """(sum dh^2 + sum dw^2) / (b*c*h*w) (losses/losses.py:62-73)."""
    b, c, h, w = img.size()
    return _fn.tv(img, scale=1.0 / (b * c * h * w))
