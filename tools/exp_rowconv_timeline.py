"""Per-row timeline of CTA 0 of the row-streaming 3x3 kernel (globaltimer stamps through fnst_set_debug_buffer)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fast_neural_style_transfer_b200 import _lib, engine, ops
from fast_neural_style_transfer_b200.ops import ConvSpec
dev = "cuda"
B, H, W, cout = 4, 256, 256, int(sys.argv[1]) if len(sys.argv) > 1 else 64
a = torch.randn((B, H, W, 64), device=dev).half()
wt = (torch.randn((cout, 576), device=dev) / 24).half()
spec = ConvSpec(engine.taps_kxk(3, origin=-1), 64, wt, cout, cout, bias=torch.randn(cout, device=dev), relu=True)
out = torch.empty((B, H, W, cout), dtype=torch.half, device=dev)
buf = torch.zeros(4 * 148 + 8 * 148, dtype=torch.int64, device=dev)
for it in range(3):
    buf.zero_()
    _lib.check(_lib.lib.fnst_set_debug_buffer(buf.data_ptr()), "dbg")
    ops.conv_gather(spec, a, (B, H, W, 64), engine._nhwc_strides(a), out, (H, W), None, True)
    torch.cuda.synchronize()
_lib.lib.fnst_set_debug_buffer(None)
t = buf.cpu().tolist()
t0 = min(x for x in t[:320] if x)
rel = lambda x: (x - t0) / 1e3 if x else float("nan")
print("row : tma_issue  loaded  mma_issued  acc_complete  epilogue_done   (us since the first stamp)")
for r in range(20):
    print(f"{r:3d} : {rel(t[r]):8.2f} {rel(t[64 + r]):8.2f} {rel(t[128 + r]):8.2f} {rel(t[192 + r]):8.2f} {rel(t[256 + r]):8.2f}")
