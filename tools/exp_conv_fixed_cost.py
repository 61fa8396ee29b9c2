#!/usr/bin/env python
"""Fixed cost vs per-k-block cost of the gather-GEMM kernel at batch 4: kernel time as a function of C_in
(9, 18, 36, 72 k-blocks of 64 channels per tile) and of the number of tiles per CTA.  B200 only."""
import sys, os, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fast_neural_style_transfer_b200 import engine, ops, _lib
from fast_neural_style_transfer_b200.ops import ConvSpec
DEV = torch.device("cuda", 0)
def knob(k, v): _lib.check(_lib.lib.fnst_set_tuning(k.encode(), int(v)), "set_tuning")

def timeit(B, hw, cin, cout, block_n, pair, stats=True, reps=20, sets=8):
    taps = engine.taps_kxk(3)
    ins = [torch.randn((B, hw + 2, hw + 2, cin), device=DEV).half() for _ in range(sets)]
    outs = [torch.empty((B, hw, hw, cout), dtype=torch.float16, device=DEV) for _ in range(sets)]
    wt = (torch.randn((cout, 9 * cin), device=DEV) * 0.05).half()
    arena = ops.ZeroArena(sets * B * cout * 2, DEV)
    sts = [arena.take(B, cout, 2) if stats else None for _ in range(sets)]
    knob("conv_block_n", block_n); knob("conv_pair", pair)
    spec = ConvSpec(taps, cin, wt, cout, cout)
    def go():
        for a, o, s in zip(ins, outs, sts):
            ops.conv_gather(spec, a, tuple(a.shape), engine._nhwc_strides(a), o, (hw, hw), s, True, stats_zeroed=True)
    side = torch.cuda.Stream(device=DEV); side.wait_stream(torch.cuda.current_stream(DEV))
    with torch.cuda.stream(side): go()
    torch.cuda.current_stream(DEV).wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): go()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): g.replay()
    e1.record(); torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / (reps * sets)
    return us, 2.0 * B * hw * hw * cout * 9 * cin / us / 1e6

for mode in (0, 1, 2, 3):
    knob("dbg_mode", mode)
    us, tf = timeit(4, 64, 256, 256, 256, 0, True)
    print(json.dumps(dict(exp="stats_parts", dbg_mode=mode, note="bit0: no column sums, bit1: no global atomics", us=round(us, 2))), flush=True)
knob("dbg_mode", 0)
for stats in (True, False):
    for cin in (64, 128, 256, 512):
        us, tf = timeit(4, 64, cin, 256, 256, 0, stats)
        print(json.dumps(dict(exp="cin", stats=stats, batch=4, cin=cin, kblocks=9 * cin // 64, us=round(us, 2), tflops=round(tf, 1))), flush=True)
for pair in (0, 2):
    for B in (1, 2, 4, 8, 9, 16, 18, 32):      # tiles = 32 * B on 148 CTAs
        us, tf = timeit(B, 64, 256, 256, 256, pair)
        print(json.dumps(dict(exp="tiles", pair=pair, batch=B, tiles=32 * B, us=round(us, 2), tflops=round(tf, 1))), flush=True)
for cin in (64, 128, 256, 512, 1024):
    us, tf = timeit(4, 64, cin, 256, 256, 2, False)
    print(json.dumps(dict(exp="cin", pair=1, stats=False, batch=4, cin=cin, kblocks=9 * cin // 64, us=round(us, 2), tflops=round(tf, 1))), flush=True)
us, tf = timeit(4, 64, 1024, 256, 256, 0, False)
print(json.dumps(dict(exp="cin", pair=0, stats=False, batch=4, cin=1024, kblocks=144, us=round(us, 2), tflops=round(tf, 1))), flush=True)
knob("conv_block_n", 0)
