#!/usr/bin/env python
"""Per-CTA timeline of the one-pass InstanceNorm backward kernel (fnst_set_debug_buffer), globaltimer ns.  B200 only."""
import sys, os, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fast_neural_style_transfer_b200 import ops, _lib
from fast_neural_style_transfer_b200._lib import PAD_REFLECT
DEV = torch.device("cuda", 0)
B, hw, c = 4, 64, 256
adt, gdt = torch.float16, torch.bfloat16
raw = torch.randn((B, hw, hw, c), device=DEV).to(adt)
gsrc = torch.randn((B, hw + 2, hw + 2, c), device=DEV).to(gdt)
st = torch.rand((B, c, 2), device=DEV) * hw * hw
st[:, :, 1] += hw * hw
g, b = torch.ones(c, device=DEV), torch.zeros(c, device=DEV)
sums = torch.empty((B, c, 2), device=DEV)
dbg = torch.zeros(8 * 1024, dtype=torch.int64, device=DEV)
for rep in range(4):
    if rep == 3:
        torch.cuda.synchronize()
        _lib.lib.fnst_set_debug_buffer(dbg.data_ptr())
    ops.inorm_bwd_fused(gsrc, None, raw, st, g, b, None, gdt, True, 1, PAD_REFLECT, sums=sums)
torch.cuda.synchronize()
_lib.lib.fnst_set_debug_buffer(None)
tl = dbg.view(-1, 8).cpu().double()
live = tl[tl[:, 0] > 0]
t0 = live[:, 0].min()
names = ["setup_consts", "tma_wait", "fold", "phase1", "reduce", "phase2", "exit_barrier"]
rec = {"ctas": int(live.shape[0]), "entry_spread_us": round(float(live[:, 0].max() - t0) / 1e3, 2),
       "grid_span_us": round(float(live[:, 7].max() - t0) / 1e3, 2)}
for i in range(7):
    d = live[:, i + 1] - live[:, i]
    rec[names[i]] = [round(float(d.median()) / 1e3, 2), round(float(d.max()) / 1e3, 2)]
print(json.dumps(rec))
