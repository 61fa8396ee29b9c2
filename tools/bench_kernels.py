#!/usr/bin/env python
"""Per-kernel device timings of the batch-4 training shapes (and tuning-knob sweeps), B200 only.

Each case is launched back to back from a captured CUDA graph over `sets` rotating buffer sets (2 = L2-hot, like
consecutive layers of a batch-4 step whose tensors stay in the 126 MB L2; many = L2-cold) and timed with CUDA events.

    python tools/bench_kernels.py [--batch 4] [--out gpurun_out/bench_kernels.json]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from fast_neural_style_transfer_b200 import engine, ops, _lib       # noqa: E402
from fast_neural_style_transfer_b200.ops import ConvSpec            # noqa: E402
from fast_neural_style_transfer_b200._lib import PAD_REFLECT         # noqa: E402

DEV = torch.device("cuda", 0)


def knob(name, value):
    _lib.check(_lib.lib.fnst_set_tuning(name.encode(), int(value)), "set_tuning")


def time_graph(fn, n_launch, reps=20):
    side = torch.cuda.Stream(device=DEV)
    side.wait_stream(torch.cuda.current_stream(DEV))
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream(DEV).wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / (reps * n_launch)       # microseconds per launch


def case_conv3x3(B, sets, dt=torch.float16, hw=64, cin=256, cout=256):
    ins = [torch.randn((B, hw + 2, hw + 2, cin), device=DEV).to(dt) for _ in range(sets)]
    outs = [torch.empty((B, hw, hw, cout), dtype=dt, device=DEV) for _ in range(sets)]
    wt = (torch.randn((cout, 9 * cin), device=DEV) * 0.02).to(dt)
    arena = ops.ZeroArena(sets * B * cout * 2, DEV)
    sts = [arena.take(B, cout, 2) for _ in range(sets)]
    spec = ConvSpec(engine.taps_kxk(3), cin, wt, cout, cout)

    def run():
        for a, o, st in zip(ins, outs, sts):
            ops.conv_gather(spec, a, tuple(a.shape), engine._nhwc_strides(a), o, (hw, hw), st, True, stats_zeroed=True)
    return run, sets, 2.0 * B * hw * hw * cout * 9 * cin


def case_wgrad3x3(B, sets, hw=64, c=256):
    gdt = torch.bfloat16
    acts = [torch.randn((B, hw + 2, hw + 2, c), device=DEV).to(gdt) for _ in range(sets)]
    gs = [torch.randn((B, hw, hw, c), device=DEV).to(gdt) for _ in range(sets)]
    outs = [torch.zeros((c, 9 * c), dtype=torch.float32, device=DEV) for _ in range(sets)]
    spec = ConvSpec(engine.taps_kxk(3), c, None, c, c)

    def run():
        for a, g, o in zip(acts, gs, outs):
            ops.wgrad(spec, a, tuple(a.shape), engine._nhwc_strides(a), g, (hw, hw), use_tc=True, out=o, out_zeroed=True)
    return run, sets, 2.0 * B * hw * hw * c * 9 * c


def case_inorm_apply(B, sets, hw=64, c=256, dt=torch.float16):
    raws = [torch.randn((B, hw, hw, c), device=DEV).to(dt) for _ in range(sets)]
    outs = [torch.empty((B, hw + 2, hw + 2, c), dtype=dt, device=DEV) for _ in range(sets)]
    st = torch.rand((B, c, 2), device=DEV) * hw * hw
    st[:, :, 1] += hw * hw
    g, b = torch.ones(c, device=DEV), torch.zeros(c, device=DEV)

    def run():
        for r, o in zip(raws, outs):
            ops.inorm_apply(r, st, g, b, o, relu=True, pad=1, pad_mode=PAD_REFLECT)
    return run, sets, B * hw * hw * c * 2 * dt.itemsize


def case_inorm_bwd(B, sets, which, hw=64, c=256):
    adt, gdt = torch.float16, torch.bfloat16
    raws = [torch.randn((B, hw, hw, c), device=DEV).to(adt) for _ in range(sets)]
    gsrc = [torch.randn((B, hw + 2, hw + 2, c), device=DEV).to(gdt) for _ in range(sets)]
    extra = [torch.randn((B, hw, hw, c), device=DEV).to(gdt) for _ in range(sets)]
    st = torch.rand((B, c, 2), device=DEV) * hw * hw
    st[:, :, 1] += hw * hw
    g, b = torch.ones(c, device=DEV), torch.zeros(c, device=DEV)
    arena = ops.ZeroArena(2 * sets * (2 * B + 2) * c + 64, DEV)
    res = {}

    def run_reduce():
        arena.used = 0
        for r, gs_, ex in zip(raws, gsrc, extra):
            res["last"] = ops.inorm_bwd_reduce(gs_, ex, r, st, g, b, None, gdt, True, 1, PAD_REFLECT, arena=arena)

    if which == "reduce":
        return run_reduce, sets, B * hw * hw * c * (2 + 2 + 2 + 2)
    run_reduce()
    gy, sums = res["last"]

    def run_apply():
        for r in raws:
            ops.inorm_bwd_apply(gy, r, st, sums, g)
    return run_apply, sets, B * hw * hw * c * (2 + 2 + 2)


def case_inorm_bwd_fused(B, sets, hw=64, c=256, want_gy=False, extra_on=False):
    adt, gdt = torch.float16, torch.bfloat16
    raws = [torch.randn((B, hw, hw, c), device=DEV).to(adt) for _ in range(sets)]
    gsrc = [torch.randn((B, hw + 2, hw + 2, c), device=DEV).to(gdt) for _ in range(sets)]
    extra = [torch.randn((B, hw, hw, c), device=DEV).to(gdt) if extra_on else None for _ in range(sets)]
    st = torch.rand((B, c, 2), device=DEV) * hw * hw
    st[:, :, 1] += hw * hw
    g, b = torch.ones(c, device=DEV), torch.zeros(c, device=DEV)
    sums = torch.empty((B, c, 2), device=DEV)

    def run():
        for r, gs_, ex in zip(raws, gsrc, extra):
            ops.inorm_bwd_fused(gs_, ex, r, st, g, b, None, gdt, True, 1, PAD_REFLECT, want_gy=want_gy, sums=sums)
    return run, sets, B * hw * hw * c * (2 + 2 + 2 + (2 if extra_on else 0) + (2 if want_gy else 0))


def case_grad_assemble(sets):
    from fast_neural_style_transfer_b200 import backward
    sys.path.insert(0, os.path.join(ROOT, "fast_neural_style_transfer_b200", "dropin"))
    from models.model import StyleTransferNet
    net = StyleTransferNet().to(DEV)
    names = [n for n, _ in net.named_parameters()]
    params = dict(net.named_parameters())
    _, total = backward._staging_layout(True)
    cores = [dict(staging=torch.randn(total, device=DEV), sums=torch.randn(2 * 4 * engine.STATS_CHANNELS + 64 * 14, device=DEV),
                  norm_sums={ln: (0, params[ln + ".weight"].numel()) for ln in backward.NORM_LAYERS}, batch=4, tc=True) for _ in range(sets)]

    def run():
        for core in cores:
            backward.assemble_gradients(core, names, params)
    return run, sets, 3 * 4.0 * sum(p.numel() for p in params.values())


def case_optimizer_tail(sets, impl):
    """clip_grad_norm_(1.0) + Adam.step over the 58 StyleTransferNet parameter tensors (train.py:203-205): libfnst's three
    multi-tensor kernels vs torch's foreach implementations.  GPU time per whole tail; bytes = 4 + 8 + 28 per element."""
    from fast_neural_style_transfer_b200 import optim as fo
    sys.path.insert(0, os.path.join(ROOT, "fast_neural_style_transfer_b200", "dropin"))
    from models.model import StyleTransferNet
    nets = [StyleTransferNet().to(DEV) for _ in range(min(sets, 4))]
    for net in nets:
        for p in net.parameters():
            p.grad = torch.randn_like(p) * 0.01
    adam, clip = (fo.Adam, fo.clip_grad_norm_) if impl == "fnst" else (torch.optim.Adam, torch.nn.utils.clip_grad_norm_)
    kw = dict(capturable=True) if impl == "torch" else {}
    opts = [adam(net.parameters(), lr=1e-3, weight_decay=1e-5, **kw) for net in nets]
    numel = sum(p.numel() for p in nets[0].parameters())

    def run():
        for net, opt in zip(nets, opts):
            clip(net.parameters(), 1.0)
            opt.step()
    return run, len(nets), numel * 40.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "bench_kernels.json"))
    ap.add_argument("--only", default="", help="substring filter on the case name")
    args = ap.parse_args()
    B = args.batch
    results = []

    def rec(name, maker, **knobs):
        if args.only and args.only not in name:
            return
        for k, v in knobs.items():
            knob(k, v)
        for sets in (2, 24):
            run, n, work = maker(sets)
            us = time_graph(run, n)
            row = dict(name=name, batch=B, sets=sets, us=round(us, 2), **knobs)
            if work > 1e9:
                row["tflops"] = round(work / us / 1e6, 1)
            else:
                row["gbs"] = round(work / us / 1e3, 1)
            results.append(row)
            print(json.dumps(row), flush=True)

    for pdl in (0, 1):
        for bn in (128, 256):
            rec("conv3x3_256", lambda s: case_conv3x3(B, s), pdl=pdl, conv_block_n=bn)
    knob("conv_block_n", 0)
    for waves in (2, 3, 4, 6):
        for bn in (128, 256):
            rec("wgrad3x3_256", lambda s: case_wgrad3x3(B, s), pdl=1, wgrad_waves_x2=waves, wgrad_bn=bn)
    knob("wgrad_waves_x2", 2)
    knob("wgrad_bn", 0)
    for pdl in (0, 1):
        rec("inorm_apply_256", lambda s: case_inorm_apply(B, s), pdl=pdl)
    for blocks in (1, 2):
        rec("inorm_bwd_reduce_256", lambda s: case_inorm_bwd(B, s, "reduce"), pdl=1, inorm_bwd_tma=0, inorm_bwd_blocks=blocks)
    rec("inorm_bwd_reduce_256", lambda s: case_inorm_bwd(B, s, "reduce"), pdl=1, inorm_bwd_tma=1)
    rec("inorm_bwd_apply_256", lambda s: case_inorm_bwd(B, s, "apply"), pdl=1)
    rec("inorm_bwd_fused_256", lambda s: case_inorm_bwd_fused(B, s), pdl=1)
    rec("inorm_bwd_fused_256_extra_gy", lambda s: case_inorm_bwd_fused(B, s, want_gy=True, extra_on=True), pdl=1)
    rec("inorm_bwd_fused_64ch_128px", lambda s: case_inorm_bwd_fused(B, s, hw=128, c=64), pdl=1)
    rec("grad_assemble", lambda s: case_grad_assemble(min(s, 4)), pdl=1)
    rec("vgg_conv1_2_64", lambda s: case_conv3x3(B, s, dt=torch.bfloat16, hw=256, cin=64, cout=64), pdl=1)
    rec("vgg_conv2_2_128", lambda s: case_conv3x3(B, s, dt=torch.bfloat16, hw=128, cin=128, cout=128), pdl=1)
    for impl in ("fnst", "torch"):
        try:
            rec("optimizer_tail_" + impl, lambda s, impl=impl: case_optimizer_tail(s, impl), pdl=1)
        except Exception as e:                       # torch's foreach path may refuse stream capture: keep the other rows
            print(json.dumps({"name": "optimizer_tail_" + impl, "error": repr(e)[:300]}), flush=True)
            torch.cuda.synchronize()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
