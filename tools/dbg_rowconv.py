"""Bring-up check of the row-streaming kernel: one tap at a time against the generic gather-GEMM (which taps / columns go wrong).
Used to establish that a row-shifted start address of a 128-byte-swizzled K-major operand needs base offset 0."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fast_neural_style_transfer_b200 import _lib, engine, ops
from fast_neural_style_transfer_b200.ops import ConvSpec
dev = "cuda"
B, H, W, cout = 1, 12, 40, 64
torch.manual_seed(0)
a = torch.randn((B, H, W, 64), device=dev).half()
taps = engine.taps_kxk(3, origin=-1)
for mode in (0,):
    for t in range(9):
        wt = torch.zeros((cout, 9, 64), device=dev)
        wt[:, t] = torch.randn((cout, 64), device=dev) / 8
        wt = wt.reshape(cout, 576).half()
        spec = ConvSpec(taps, 64, wt, cout, cout)
        outs = []
        for rs in (1, 0):
            _lib.check(_lib.lib.fnst_set_tuning(b"conv_rowstream", rs), "k")
            _lib.check(_lib.lib.fnst_set_tuning(b"dbg_mode", mode), "k")
            out = torch.full((B, H, W, cout), float("nan"), dtype=torch.half, device=dev)
            ops.conv_gather(spec, a, (B, H, W, 64), engine._nhwc_strides(a), out, (H, W), None, True)
            torch.cuda.synchronize()
            outs.append(out.float())
        err = (outs[0] - outs[1]).abs()
        bad = (err > 1e-2).any(-1)[0]          # (H, W)
        print("mode", mode, "tap", divmod(t, 3), "max err", float(err.max()), "bad pixels", int(bad.sum()), "of", H * W,
              "bad cols", sorted(set(bad.nonzero()[:, 1].tolist()))[:12])
_lib.lib.fnst_set_tuning(b"dbg_mode", 0)
