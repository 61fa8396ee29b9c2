#!/usr/bin/env python
"""Per-CTA timeline of the gather-GEMM kernel (fnst_set_debug_buffer): where the fixed cost of a one-tile-per-CTA launch
goes -- set-up, first operand stage, main loop, epilogue, exit.  globaltimer nanoseconds.  B200 only."""
import sys, os, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fast_neural_style_transfer_b200 import engine, ops, _lib
from fast_neural_style_transfer_b200.ops import ConvSpec
DEV = torch.device("cuda", 0)
def knob(k, v): _lib.check(_lib.lib.fnst_set_tuning(k.encode(), int(v)), "set_tuning")

knob("conv_rowstream", 0)          # this tool times the generic gather-GEMM (the 64-channel case would otherwise take rowconv_tc)
dbg = torch.zeros(12 * 148, dtype=torch.int64, device=DEV)
for B, cin, cout, hw, stats in ((4, 256, 256, 64, True), (4, 256, 256, 64, False), (1, 256, 256, 64, True), (4, 512, 512, 32, False), (4, 64, 64, 256, False)):
    a = torch.randn((B, hw + 2, hw + 2, cin), device=DEV).half()
    out = torch.empty((B, hw, hw, cout), dtype=torch.float16, device=DEV)
    wt = (torch.randn((cout, 9 * cin), device=DEV) * 0.05).half()
    st = torch.zeros((B, cout, 2), device=DEV) if stats else None
    spec = ConvSpec(engine.taps_kxk(3), cin, wt, cout, cout)
    for rep in range(4):
        if rep == 3:
            dbg.zero_()
            torch.cuda.synchronize()
            _lib.lib.fnst_set_debug_buffer(dbg.data_ptr())
        ops.conv_gather(spec, a, tuple(a.shape), engine._nhwc_strides(a), out, (hw, hw), st, True, stats_zeroed=True)
    torch.cuda.synchronize()
    _lib.lib.fnst_set_debug_buffer(None)
    tl = dbg[4 * 148:].view(148, 8).cpu().double()
    live = tl[tl[:, 0] > 0]
    t0 = live[:, 0].min()
    def med(c): return round(float(c.median()) / 1e3, 2)
    rec = dict(batch=B, cin=cin, cout=cout, hw=hw, stats=stats, ctas=int(live.shape[0]),
               entry_spread_us=round(float(live[:, 0].max() - t0) / 1e3, 2),
               setup_us=med(live[:, 1] - live[:, 0]), first_stage_us=med(live[:, 2] - live[:, 1]),
               main_loop_us=med(live[:, 3] - live[:, 2]), epilogue_us=med(live[:, 4] - live[:, 3]),
               exit_us=med(live[:, 5] - live[:, 4]), cta_total_us=med(live[:, 5] - live[:, 0]),
               grid_span_us=round(float(live[:, 5].max() - t0) / 1e3, 2))
    print(json.dumps(rec), flush=True)
