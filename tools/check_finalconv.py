#!/usr/bin/env python
"""fnst_finalconv_tc (row-streaming final_conv) against torch conv2d on the same fp16 inputs, and its time against the
gather-GEMM ROWSUM9 form.  B200 only."""
import sys, os, json
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fast_neural_style_transfer_b200 import engine, ops, _lib
from fast_neural_style_transfer_b200.ops import ConvSpec
from fast_neural_style_transfer_b200._lib import EPI_ROWSUM9
DEV = torch.device("cuda", 0)
def knob(k, v): _lib.check(_lib.lib.fnst_set_tuning(k.encode(), int(v)), "set_tuning")

def case(B, H, W, swap):
    torch.manual_seed(1)
    act = torch.randn((B, H + 8, W + 8, 32), device=DEV).half()
    flat = torch.zeros(act.numel() + 128, dtype=torch.float16, device=DEV); flat[:act.numel()] = act.reshape(-1)
    w = (torch.randn((3, 32, 9, 9), device=DEV) * 0.05)
    bias = torch.zeros(16, device=DEV); bias[:3] = torch.tensor([0.1, -0.2, 0.3], device=DEV)
    ws = engine.pack_final_stream(w, torch.float16)
    y = torch.zeros((B, 3, H, W), device=DEV)
    knob("dbg_mode", 16 if swap else 0)
    ops.finalconv_stream(flat, B, H, W, ws, bias, y)
    torch.cuda.synchronize()
    knob("dbg_mode", 0)
    ref = F.conv2d(act.float().permute(0, 3, 1, 2), w.half().float(), bias[:3])
    return float((y - ref).norm() / ref.norm()), float((y - ref).abs().max())

good_swap = 0          # LBO = K-direction core-matrix stride, SBO = 8-row-group stride (the swapped reading faults)
for (B, H, W) in ((1, 256, 256), (3, 100, 131), (1, 1080, 1920), (5, 17, 9)):
    r, m = case(B, H, W, good_swap)
    print(json.dumps(dict(shape=(B, H, W), rel_l2=r, max_abs=m, ok=r < 2e-3)), flush=True)
# timing, batch 256 at 256x256 and 8 at 1080p
for (B, H, W) in ((256, 256, 256), (8, 1080, 1920), (4, 256, 256), (1, 256, 256)):
    act = torch.randn((B, H + 8, W + 8, 32), device=DEV).half()
    flat = torch.zeros(act.numel() + 128, dtype=torch.float16, device=DEV); flat[:act.numel()] = act.reshape(-1)
    w = (torch.randn((3, 32, 9, 9), device=DEV) * 0.05)
    bias = torch.zeros(16, device=DEV)
    ws = engine.pack_final_stream(w, torch.float16)
    wr = engine.pack_final_rowsum(w, torch.float16)
    y = torch.zeros((B, 3, H, W), device=DEV)
    def stream(): ops.finalconv_stream(flat, B, H, W, ws, bias, y)
    def rowsum():
        spec = ConvSpec(engine.TAPS_ROWSUM, 64, wr, 32, 3, epilogue=EPI_ROWSUM9, bias=bias)
        ops.conv_gather(spec, flat[:act.numel()].view(B, H + 8, W + 8, 32), (B, H + 8, W + 8, 64), ((H + 8) * (W + 8) * 32, (W + 8) * 32, 32), y, (H, W), None, True)
    res = {}
    for name, fn in (("stream", stream), ("rowsum9", rowsum)):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        res[name] = round(1e3 * e0.elapsed_time(e1) / reps, 1)
    print(json.dumps(dict(shape=(B, H, W), us=res)), flush=True)
