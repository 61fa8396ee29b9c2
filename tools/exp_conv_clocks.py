#!/usr/bin/env python
"""SM clocks per k-block inside the gather-GEMM main loop (fnst_set_debug_buffer): is the tensor pipe the limiter?
Ideal: 4 MMAs of M128 x N256 x K16 = 512 clocks per k-block of 64 channels.  B200 only."""
import sys, os, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fast_neural_style_transfer_b200 import engine, ops, _lib
from fast_neural_style_transfer_b200.ops import ConvSpec
DEV = torch.device("cuda", 0)
def knob(k, v): _lib.check(_lib.lib.fnst_set_tuning(k.encode(), int(v)), "set_tuning")

dbg = torch.zeros(4 * 148, dtype=torch.int64, device=DEV)
for B in (4, 32, 256):
    for block_n in (128, 256):
        for pair in (0, 2):
            cin = cout = 256; hw = 64
            a = torch.randn((B, hw + 2, hw + 2, cin), device=DEV).half()
            out = torch.empty((B, hw, hw, cout), dtype=torch.float16, device=DEV)
            wt = (torch.randn((cout, 9 * cin), device=DEV) * 0.05).half()
            st = torch.zeros((B, cout, 2), device=DEV)
            knob("conv_block_n", block_n); knob("conv_pair", pair)
            spec = ConvSpec(engine.taps_kxk(3), cin, wt, cout, cout)
            for rep in range(3):
                if rep == 2:
                    dbg.zero_()
                    _lib.lib.fnst_set_debug_buffer(dbg.data_ptr())
                ops.conv_gather(spec, a, tuple(a.shape), engine._nhwc_strides(a), out, (hw, hw), st, True, stats_zeroed=True)
            torch.cuda.synchronize()
            _lib.lib.fnst_set_debug_buffer(None)
            d = dbg.view(148, 4).cpu()
            live = d[d[:, 2] > 0]
            clk_per_kb = (live[:, 0].double() / live[:, 2].double())
            ghz = (live[:, 0].double() / live[:, 1].double())
            ideal = 512 * block_n / 256
            print(json.dumps(dict(batch=B, block_n=block_n, pair=pair, ctas=int(live.shape[0]), clk_per_kblock=round(float(clk_per_kb.mean()), 1),
                                  clk_per_kblock_max=round(float(clk_per_kb.max()), 1), ideal=ideal,
                                  mma_util=round(ideal / float(clk_per_kb.mean()), 3), sm_ghz=round(float(ghz.mean()), 3))), flush=True)
knob("conv_block_n", 0); knob("conv_pair", 1)
