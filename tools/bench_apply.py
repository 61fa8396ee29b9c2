"""Microbench of the HBM-bound InstanceNorm-apply kernel: GB/s vs the measured copy peak."""
import json, os, sys, torch
sys.path.insert(0, '.')
from fast_neural_style_transfer_b200 import ops, _lib
dev = torch.device('cuda', 0)
peak = json.load(open('MEASURED_PEAKS.json'))['hbm_gbs'] if os.path.exists('MEASURED_PEAKS.json') else 6650.0
def run(n, h, w, c, pad, mode, s2d, res, relu, dtype=torch.float16, sets=3, reps=10):
    raws = [torch.randn((n, h, w, c), device=dev).to(dtype) for _ in range(sets)]
    st = torch.stack([raws[0].float().sum((1, 2)), (raws[0].float() ** 2).sum((1, 2))], -1).contiguous()
    gamma, beta = torch.ones(c, device=dev), torch.zeros(c, device=dev)
    hp, wp = h + 2 * pad, w + 2 * pad
    oshape = (n, (hp + 1) // 2, (wp + 1) // 2, 4 * c) if s2d else (n, hp, wp, c)
    outs = [torch.zeros(oshape, dtype=dtype, device=dev) for _ in range(sets)]
    ress = [torch.randn((n, h + 2, w + 2, c), device=dev).to(dtype) for _ in range(sets)] if res else [None] * sets
    def go():
        for r, o, rs in zip(raws, outs, ress):
            ops.inorm_apply(r, st, gamma, beta, o, relu, pad, mode, s2d, None, rs, 1)
    go(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): go()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): g.replay()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (reps * sets)
    es = raws[0].element_size()
    bytes_ = raws[0].numel() * es + outs[0].numel() * es + (ress[0].numel() * es if res else 0)
    gbs = bytes_ / ms / 1e6
    print(f"apply n={n} {h}x{w}x{c} pad={pad} s2d={s2d} res={res}: {ms*1e3:8.1f} us  {gbs:7.0f} GB/s  {gbs/peak:5.1%} of measured copy peak ({peak:.0f})")
    return dict(shape=[n, h, w, c], pad=pad, s2d=s2d, res=res, us=ms * 1e3, gbs=gbs, frac=gbs / peak, bytes=bytes_)
out = [run(256, 64, 64, 256, 1, 1, False, False, True), run(256, 64, 64, 256, 1, 1, False, True, False),
       run(256, 128, 128, 64, 1, 1, True, False, True), run(256, 256, 256, 32, 4, 1, False, False, True),
       run(256, 128, 128, 64, 0, 0, False, False, True), run(8, 270, 480, 256, 1, 1, False, True, False)]
json.dump(out, open('gpurun_out/bench_apply.json', 'w'), indent=1)
