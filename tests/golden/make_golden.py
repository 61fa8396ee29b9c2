"""Generate golden vectors from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py        # needs /root/reference; writes tests/golden/*.npz

The reference modules are imported from /root/reference (never copied).  Parameters and inputs
come from oracle.stylenet_oracle's seeded generators and are loaded into the reference modules
with load_state_dict, so the fixtures pin "reference(params, x)" for the oracle to reproduce.

VGG19: the shipped constructor cannot run (models/vgg19_net.py:27 downloads weights, :51 uses an
undefined slice5).  `_RefVGG` below overrides ONLY __init__ (same add_module names, same slices);
forward is inherited unmodified from the reference, so slicing / in-place-ReLU aliasing are the
reference's own.
"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)

from models.model import StyleTransferNet            # noqa: E402  (reference)
from models.vgg19_net import VGG19                   # noqa: E402  (reference)
from losses import losses as ref_losses              # noqa: E402  (reference)
from oracle import stylenet_oracle as O              # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


class _RefVGG(VGG19):
    def __init__(self):
        nn.Module.__init__(self)
        from torchvision.models import vgg19
        f = vgg19(weights=None).features
        for name, lo, hi in (("slice1", 0, 4), ("slice2", 4, 9), ("slice3", 9, 16), ("slice4", 16, 22), ("slice5", 22, 25)):
            seq = nn.Sequential()
            for i in range(lo, hi):
                seq.add_module(str(i), f[i])
            setattr(self, name, seq)
        for p in self.parameters():
            p.requires_grad = False


def ref_net(params):
    net = StyleTransferNet()
    net.load_state_dict(params)
    return net


def ref_vgg(params):
    vgg = _RefVGG()
    vgg.load_state_dict(params)
    return vgg.eval()


def main():
    torch.manual_seed(0)
    torch.set_num_threads(4)
    net_p = O.make_net_params(seed=0, random_affine=True)
    vgg_p = O.make_vgg_params(seed=1)

    # ---- 1. StyleTransferNet.eval() forward at small / odd sizes ------------------------------
    net = ref_net(net_p).eval()
    fwd = {}
    for tag, (b, h, w) in {"a": (1, 32, 32), "b": (2, 20, 28), "c": (1, 37, 45)}.items():
        x = O.make_image(b, h, w, seed=10 + b + h)
        with torch.no_grad():
            fwd[f"y_{tag}"] = net(x).numpy()
        fwd[f"shape_{tag}"] = np.array([b, h, w])
    np.savez_compressed(os.path.join(OUT, "net_forward.npz"), **fwd)

    # ---- 2. VGG features + losses -------------------------------------------------------------
    vgg = ref_vgg(vgg_p)
    x = O.make_image(2, 16, 24, seed=77, normalized=True)
    sty = O.make_image(1, 16, 16, seed=78, normalized=True)
    with torch.no_grad():
        feats = vgg(x)
        sfe = vgg(sty)
        targets = [ref_losses.gram_matrix(f).squeeze(0) for f in sfe]      # train.py:32-35
        other = vgg(O.make_image(2, 16, 24, seed=79, normalized=True))
        d = {f"feat{i}": f.numpy() for i, f in enumerate(feats)}
        d.update({f"target{i}": t.numpy() for i, t in enumerate(targets[:3])})
        for i in (3, 4):        # 512x512 targets: keep a corner + the total to bound fixture size
            d[f"target{i}_corner"] = targets[i][:16, :16].numpy()
            d[f"target{i}_sum"] = np.float64(targets[i].double().sum().item())
        d["gram0"] = ref_losses.gram_matrix(feats[0]).numpy()
        d["style"] = np.float64(ref_losses.style_loss(other, targets).item())
        d["content"] = np.float64(ref_losses.content_loss(other, feats).item())
        d["tv"] = np.float64(ref_losses.total_variation_loss(x).item())
    np.savez_compressed(os.path.join(OUT, "vgg_losses.npz"), **d)

    # ---- 3. One training step (train.py:168-206) in .train() mode, dropout pinned ------------
    b, h, w = 2, 32, 32
    content = O.make_image(b, h, w, seed=5, normalized=True)
    sty = O.make_image(1, h, w, seed=6, normalized=True)
    drop = O.make_dropout_scales(b, seed=7)
    net = ref_net(net_p).train()
    # Pin Dropout2d: replace each block's dropout by a module multiplying the pinned scale.
    class _Pinned(nn.Module):
        def __init__(self, s):
            super().__init__()
            self.s = s
        def forward(self, t):
            return t * self.s.view(t.shape[0], t.shape[1], 1, 1)
    for i, blk in enumerate(net.res_blocks):
        blk.dropout = _Pinned(drop[i])
    with torch.no_grad():
        targets = [ref_losses.gram_matrix(f).squeeze(0) for f in vgg(sty)]
    opt = torch.optim.Adam(net.parameters(), lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-5)
    stylized = torch.clamp(net(content), -3, 3)                                 # train.py:171-174
    with torch.no_grad():
        cf = vgg(content)
    sf = vgg(stylized)
    c = ref_losses.content_loss(sf, cf)
    s = ref_losses.style_loss(sf, targets)
    tv = ref_losses.total_variation_loss(stylized)
    total = 1000.0 * c + 1 * s + 10 * tv
    opt.zero_grad()
    total.backward()
    grads = {k: v.grad.detach().clone() for k, v in net.named_parameters()}
    gnorm = torch.nn.utils.clip_grad_norm_(net.parameters(), max_norm=1.0)
    opt.step()
    after = {k: v.detach().clone() for k, v in net.named_parameters()}
    t = {"total": np.float64(total.item()), "content": np.float64(c.item()), "style": np.float64(s.item()),
         "tv": np.float64(tv.item()), "grad_norm": np.float64(gnorm.item()),
         "stylized": stylized.detach().numpy()}
    for k in grads:
        t["gnorm/" + k] = np.float64(grads[k].double().norm().item())
    # full gradients / updated values for a few small tensors (keep the fixture small)
    for k in ("conv1.conv.weight", "norm2.weight", "norm2.bias", "up2.upsample_conv.weight", "final_conv.conv.weight",
              "final_conv.conv.bias", "res_blocks.4.in2.weight"):
        t["grad/" + k] = grads[k].numpy()
        t["after/" + k] = after[k].numpy()
    np.savez_compressed(os.path.join(OUT, "train_step.npz"), **t)

    # ---- 4. Dropout2d mask recipe -------------------------------------------------
    torch.manual_seed(123)
    ones = torch.ones(3, 256, 4, 4)
    m = nn.Dropout2d(0.1).train()(ones)[:, :, 0, 0]
    np.savez_compressed(os.path.join(OUT, "dropout.npz"), values=np.unique(m.numpy()))
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
