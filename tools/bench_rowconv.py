"""Times the 64-input-channel 3x3 layers (VGG conv1_2, conv2_1 forward; conv1_2 data gradient) with the row-streaming kernel
and with the generic gather-GEMM (tuning knob conv_rowstream).  CUDA events around 100 launches replayed from a CUDA graph."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fast_neural_style_transfer_b200 import _lib, backward, engine, ops
from fast_neural_style_transfer_b200.ops import ConvSpec

dev = "cuda"
res = {}
for name, (B, H, W, cout, form) in {"conv1_2_fwd": (4, 256, 256, 64, "f"), "conv2_1_fwd": (4, 128, 128, 128, "f"),
                                     "conv1_2_dgrad": (4, 256, 256, 64, "d")}.items():
    dt = torch.float16 if form == "f" else torch.bfloat16
    a = torch.randn((B, H, W, 64), device=dev).to(dt)
    wt = (torch.randn((cout, 576), device=dev) / 24).to(dt)
    taps = engine.taps_kxk(3, origin=-1)
    if form == "f":
        spec = ConvSpec(taps, 64, wt, cout, cout, bias=torch.randn(cout, device=dev), relu=True)
    else:
        spec = ConvSpec(backward._neg(taps), 64, wt, cout, cout, addend=torch.randn((B, H, W, cout), device=dev).to(dt),
                        mask=torch.randn((B, H, W, cout), device=dev).half())
    out = torch.empty((B, H, W, cout), dtype=dt, device=dev)
    for rs in (1, 0, 8):
        _lib.check(_lib.lib.fnst_set_tuning(b"conv_rowstream", 1 if rs else 0), "knob")
        _lib.check(_lib.lib.fnst_set_tuning(b"dbg_mode", rs if rs > 1 else 0), "knob")
        # 20 launches captured in one CUDA graph (a Python launch costs ~20 us of host time: longer than the kernel)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                ops.conv_gather(spec, a, (B, H, W, 64), engine._nhwc_strides(a), out, (H, W), None, True)
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for _ in range(20):
                ops.conv_gather(spec, a, (B, H, W, 64), engine._nhwc_strides(a), out, (H, W), None, True)
        graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 100
        gf = 2.0 * B * H * W * cout * 576 / 1e9
        res[f"{name}_{ {1: 'rowstream', 0: 'gather', 8: 'rowstream_without_epilogue_stores'}[rs] }"] = dict(us=round(us, 2), tflops=round(1e3 * gf / us, 1))
print(json.dumps(res, indent=1))
