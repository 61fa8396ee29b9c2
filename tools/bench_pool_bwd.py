"""MaxPool2d(2,2) backward kernel (fused ReLU mask + residual-branch gradient): device time per launch from CUDA-graph replays at the three VGG sizes of the batch-4 step."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fast_neural_style_transfer_b200 import ops
dev = "cuda"
for (B, H, W, C) in ((4, 256, 256, 64), (4, 128, 128, 128), (4, 64, 64, 256)):
    inp = torch.randn((B, H, W, C), device=dev).bfloat16()
    gout = torch.randn((B, H // 2, W // 2, C), device=dev).bfloat16()
    extra = torch.randn((B, H, W, C), device=dev).bfloat16()
    for _ in range(3): ops.maxpool2_bwd(inp, gout, extra)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s): ops.maxpool2_bwd(inp, gout, extra)
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        for _ in range(20): ops.maxpool2_bwd(inp, gout, extra)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): g.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 10
    mb = (inp.numel() * 3 + gout.numel()) * 2 / 1e6
    print((B, H, W, C), round(us, 2), "us per launch,", round(mb / us, 0), "GB/s (algorithmic bytes: input + residual gradient read, pooled gradient read, gradient written)")
