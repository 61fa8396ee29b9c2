#!/usr/bin/env python
"""CTA-pair (cta_group::2) gather-GEMM vs the single-CTA form of the same kernel: identical inputs, results compared
element-wise (same accumulation order, so they should agree to the last bit) plus timing.  B200 only."""
import sys, os, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fast_neural_style_transfer_b200 import engine, ops, _lib
from fast_neural_style_transfer_b200.ops import ConvSpec
from fast_neural_style_transfer_b200._lib import EPI_D2S

DEV = torch.device("cuda", 0)

def knob(k, v):
    _lib.check(_lib.lib.fnst_set_tuning(k.encode(), int(v)), "set_tuning")

def run(B, hw, cin, cout, block_n, pair, epilogue=0, dt=torch.float16, taps=None, n_gemm=None, w=None):
    torch.manual_seed(0)
    taps = taps or engine.taps_kxk(3)
    n_gemm = n_gemm or cout
    pad = 2 if len(taps) == 9 else 1
    a = torch.randn((B, hw + pad, hw + pad, cin), device=DEV).to(dt)
    wt = (torch.randn((n_gemm, len(taps) * cin), device=DEV) * 0.05).to(dt)
    oh = hw
    if epilogue == EPI_D2S:
        out = torch.zeros((B, 2 * oh, 2 * oh, cout), dtype=dt, device=DEV)
    else:
        out = torch.zeros((B, oh, oh, cout), dtype=dt, device=DEV)
    st = torch.zeros((B, cout, 2), dtype=torch.float32, device=DEV)
    knob("conv_block_n", block_n); knob("conv_pair", pair)
    spec = ConvSpec(taps, cin, wt, n_gemm, cout, epilogue=epilogue)
    ops.conv_gather(spec, a, tuple(a.shape), engine._nhwc_strides(a), out, (oh, oh), st, True)
    torch.cuda.synchronize()
    return out, st

def timeit(B, hw, cin, cout, block_n, pair, reps=30):
    torch.manual_seed(0)
    sets = 8
    taps = engine.taps_kxk(3)
    ins = [torch.randn((B, hw + 2, hw + 2, cin), device=DEV).half() for _ in range(sets)]
    outs = [torch.empty((B, hw, hw, cout), dtype=torch.float16, device=DEV) for _ in range(sets)]
    wt = (torch.randn((cout, 9 * cin), device=DEV) * 0.05).half()
    arena = ops.ZeroArena(sets * B * cout * 2, DEV)
    sts = [arena.take(B, cout, 2) for _ in range(sets)]
    knob("conv_block_n", block_n); knob("conv_pair", pair)
    spec = ConvSpec(taps, cin, wt, cout, cout)
    def go():
        for a, o, s in zip(ins, outs, sts):
            ops.conv_gather(spec, a, tuple(a.shape), engine._nhwc_strides(a), o, (hw, hw), s, True, stats_zeroed=True)
    go(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): go()
    e1.record(); torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / (reps * sets)
    return us, 2.0 * B * hw * hw * cout * 9 * cin / us / 1e6

ok = True
cases = [dict(B=4, hw=64, cin=256, cout=256, block_n=256), dict(B=4, hw=64, cin=256, cout=256, block_n=128),
         dict(B=3, hw=24, cin=64, cout=128, block_n=128),                     # odd tile count (phantom tile), ragged edges
         dict(B=1, hw=40, cin=128, cout=256, block_n=256),
         dict(B=2, hw=32, cin=256, cout=64, block_n=256, epilogue=EPI_D2S, taps=engine.TAPS_2X2, n_gemm=256),
         dict(B=5, hw=17, cin=64, cout=256, block_n=256, dt=torch.bfloat16)]
for c in cases:
    o0, s0 = run(pair=0, **c)
    o1, s1 = run(pair=2, **c)
    d = (o0.float() - o1.float()).abs().max().item()
    ds = ((s0 - s1).abs() / (s0.abs() + 1)).max().item()
    good = d == 0.0 and ds < 1e-4
    ok &= good
    print(("OK  " if good else "FAIL"), {k: (str(v) if not isinstance(v, (int,)) else v) for k, v in c.items() if k != "taps"}, "max|diff|", d, "stats rel", ds, flush=True)
for B in (4, 32, 256):
    for bn in (128, 256):
        for pair in (0, 2):
            us, tf = timeit(B, 64, 256, 256, bn, pair, reps=30 if B < 256 else 6)
            print(json.dumps(dict(batch=B, block_n=bn, pair=pair, us=round(us, 2), tflops=round(tf, 1))), flush=True)
knob("conv_block_n", 0); knob("conv_pair", 1)
sys.exit(0 if ok else 1)
