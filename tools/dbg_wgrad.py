import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import emu_ops
from fast_neural_style_transfer_b200 import ops, engine
from fast_neural_style_transfer_b200.ops import ConvSpec
case = sys.argv[1]
g = torch.Generator().manual_seed(1)
def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm())
if case.startswith("gram"):
    dt = torch.float16 if case.endswith("f16") else torch.bfloat16
    c = int(case.split("_")[1])
    f = torch.relu(torch.randn((2, 16, 16, c), generator=g)).to(dt)
    got = ops.gram(f.cuda(), use_tc=True); torch.cuda.synchronize()
    print(case, "rel", rel(got, emu_ops.gram(f, False)))
else:
    adt = {"hh": torch.float16, "bb": torch.bfloat16, "hb": torch.float16}[case.split("_")[1]]
    gdt = {"hh": torch.float16, "bb": torch.bfloat16, "hb": torch.bfloat16}[case.split("_")[1]]
    kc, n_gemm = int(case.split("_")[2]), int(case.split("_")[3])
    B, H, W = 2, 12, 10
    a = torch.randn((B, H + 2, W + 2, kc), generator=g).to(adt)
    go = torch.randn((B, H, W, n_gemm), generator=g).to(gdt)
    taps = engine.taps_kxk(3)
    spec = ConvSpec(taps, kc, None, n_gemm, n_gemm)
    ref = ops.wgrad(spec, a.cuda(), (B, H + 2, W + 2, kc), engine._nhwc_strides(a), go.cuda(), (H, W), use_tc=False); torch.cuda.synchronize()
    got = ops.wgrad(spec, a.cuda(), (B, H + 2, W + 2, kc), engine._nhwc_strides(a), go.cuda(), (H, W), use_tc=True); torch.cuda.synchronize()
    print(case, "rel vs simt", rel(got, ref))
