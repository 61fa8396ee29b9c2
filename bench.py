#!/usr/bin/env python
"""bench.py -- style-transfer hot path on B200 (see DESIGN.md section "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload infer256|infer1080|train] [--impl reference]

One JSON line on stdout (rank 0).  A "step" is one pass of the hot path over one batch of synthetic
images that are already resident in HBM; `e2e` repeats the measurement through the drop-in module
with pinned HOST buffers (H2D of the inputs and D2H of the result inside the timed region).
`--impl reference` times the CPU restatement of the reference (oracle/, PyTorch CPU ops, all host
threads) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("FNST_VGG19_RANDOM_INIT", "1")     # BASELINE.json: no network, VGG-19 weights are random (seeded below)

WORKLOADS = {
    # name: (global batch, H, W, metric, unit, algorithmic GFLOP per image of the dominant kernel launch (one 3x3 256->256 conv))
    "infer256": dict(batch=256, h=256, w=256, metric="stylized images/sec (256x256)", unit="images/s", scaling="strong",
                     desc="BASELINE.json configs[3]: batched inference 256x3x256x256 sharded by batch"),
    "infer1080": dict(batch=8, h=1080, w=1920, metric="stylized images/sec (1080x1920)", unit="images/s", scaling="weak",
                      desc="BASELINE.json configs[2] shape, batch 8 per GPU"),
    "infer1080_b1": dict(batch=1, h=1080, w=1920, metric="stylized images/sec (1080x1920, batch 1)", unit="images/s", scaling="weak",
                         desc="BASELINE.json configs[2]: single 1x3x1080x1920 image (latency case)"),
    "infer256_b1": dict(batch=1, h=256, w=256, metric="stylized images/sec (256x256, batch 1)", unit="images/s", scaling="weak",
                        desc="BASELINE.json configs[0]: single 1x3x256x256 image (latency case)"),
    "train": dict(batch=4, h=256, w=256, metric="train steps/sec (batch 4 per GPU, 256x256)", unit="steps/s", scaling="weak",
                  desc="BASELINE.json configs[1]/[4]: perceptual-loss training step, batch 4 per GPU"),
}
NET_GFLOP_256 = 52.867          # algorithmic forward GFLOP per 256x256 image
TRAIN_GFLOP_IMG = 286.82        # algorithmic GFLOP per image of one training step


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in out.strip().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def dist_setup(n_gpus):
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local


def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(value, world, dev):
    if world == 1:
        return value
    import torch.distributed as dist
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def time_dominant_kernel(net, per_rank, h, w, dev):
    """Device time of the dominant kernel of the path: the 3x3 256->256 gather-GEMM of the residual trunk
    (`conv_tc_kernel`, ten forward launches per network pass; its data- and weight-gradient twins in training).
    Launched back to back from a captured CUDA graph on rotating buffer sets whose total footprint exceeds the
    126 MB L2, timed with CUDA events on the launching stream.  Returns (ms per launch, FLOP per launch)."""
    from fast_neural_style_transfer_b200 import engine, ops
    from fast_neural_style_transfer_b200.ops import ConvSpec
    if os.environ.get("FNST_BENCH_NO_ROOFLINE"):          # profiling runs (ncu launch lists) skip this extra leg
        return float("nan"), 0.0, 0
    plan = net._plan()
    dt = plan.dtype
    h2, w2 = (h + 3) // 4, (w + 3) // 4
    set_bytes = per_rank * ((h2 + 2) * (w2 + 2) + h2 * w2) * 256 * dt.itemsize     # (fp16x3 sets are larger still)
    n_sets = max(2, min(64, int(2.5 * 126e6 / set_bytes) + 1))
    split = getattr(plan, "split", False)                 # fp16x3: [hi | lo] activations, three virtual taps per tap, fp32 raw output
    cin = 512 if split else 256
    ins = [torch.randn((per_rank, h2 + 2, w2 + 2, cin), device=dev).to(dt) for _ in range(n_sets)]
    outs = [torch.empty((per_rank, h2, w2, 256), dtype=torch.float32 if split else dt, device=dev) for _ in range(n_sets)]
    arena = ops.ZeroArena(n_sets * per_rank * 256 * 2, dev)      # as in the product path: statistics zeroed once per forward
    stats = [arena.take(per_rank, 256, 2) for _ in range(n_sets)]
    taps = engine.taps_x3(engine.taps_kxk(3), 256) if split else engine.taps_kxk(3)
    spec = ConvSpec(taps, 256, plan.w["res0a"], 256, 256)
    def launch_all():
        for a, o, st in zip(ins, outs, stats):
            ops.conv_gather(spec, a, tuple(a.shape), engine._nhwc_strides(a), o, (h2, w2), st, plan.use_tc, stats_zeroed=True)
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        launch_all()
    torch.cuda.current_stream(dev).wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        launch_all()
    g.replay()
    torch.cuda.synchronize()
    reps = max(3, min(50, 2000 // n_sets))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (reps * n_sets)
    return ms, 2.0 * per_rank * h2 * w2 * 256 * 2304, n_sets * reps


# -------------------------------------------------------------------------------------------------------
def run_reference(args, wl):
    """CPU arm: the oracle (port of the reference's PyTorch modules) on all host threads."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    from oracle import stylenet_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    p = O.make_net_params(seed=0)
    if args.workload == "train":
        vp = O.make_vgg_params(seed=1)
        sample_b = 4
        content = O.make_image(sample_b, wl["h"], wl["w"], seed=1234, normalized=True)
        targets = O.style_targets(vp, O.make_image(1, wl["h"], wl["w"], seed=4321, normalized=True))
        state = {}
        params = {k: v.clone() for k, v in p.items()}
        def step(i):
            _, grads = O.loss_and_grads(params, vp, content, targets, O.make_dropout_scales(sample_b, seed=i))
            O.clip_and_adam(params, grads, state, step=i + 1)
        units = 1.0
        sample = f"one full training step, batch {sample_b} at {wl['h']}x{wl['w']} (the whole workload unit)"
    else:
        sample_b = 8 if wl["h"] <= 256 else 1
        x = O.make_image(sample_b, wl["h"], wl["w"], seed=1234)
        def step(i):
            with torch.no_grad():
                O.stylenet_forward(p, x)
        units = float(sample_b)
        sample = f"{sample_b} image(s) of {wl['h']}x{wl['w']} per step (bounded sample of the {wl['batch']}-image batch)"
    for i in range(args.warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(args.warmup + i)
    dt = time.perf_counter() - t0
    value = units * args.steps / dt
    line = {"impl": "reference", "metric": wl["metric"], "value": value, "unit": wl["unit"], "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": wl["scaling"],
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "desc": wl["desc"], "device": "host CPU"},
            "cpu_baseline": {"value": value, "unit": wl["unit"], "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": wl["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_baseline(workload, wl):
    """Bounded CPU sample of the same workload on rank 0 (oracle port, all host threads)."""
    from oracle import stylenet_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    p = O.make_net_params(seed=0)
    if workload == "train":
        vp = O.make_vgg_params(seed=1)
        content = O.make_image(4, wl["h"], wl["w"], seed=1234, normalized=True)
        targets = O.style_targets(vp, O.make_image(1, wl["h"], wl["w"], seed=4321, normalized=True))
        params = {k: v.clone() for k, v in p.items()}
        state = {}
        def step(i):
            _, grads = O.loss_and_grads(params, vp, content, targets, O.make_dropout_scales(4, seed=i))
            O.clip_and_adam(params, grads, state, step=i + 1)
        reps, units, sample = 3, 1.0, "3 full training steps (batch 4, 256x256) after 1 warm-up"
    elif wl["h"] <= 256:
        x = O.make_image(16, wl["h"], wl["w"], seed=1234)
        def step(i):
            with torch.no_grad():
                O.stylenet_forward(p, x)
        reps, units, sample = 4, 16.0, "4 forward passes over 16 of the 256 images after 1 warm-up"
    else:
        x = O.make_image(1, wl["h"], wl["w"], seed=1234)
        def step(i):
            with torch.no_grad():
                O.stylenet_forward(p, x)
        reps, units, sample = 2, 1.0, "2 forward passes over 1 image after 1 warm-up"
    step(0)
    t0 = time.perf_counter()
    for i in range(reps):
        step(i + 1)
    dt = time.perf_counter() - t0
    return {"value": units * reps / dt, "unit": wl["unit"], "cores": cores, "kind": "port", "sample": sample}


def time_optimizer_tail(net, dev):
    """Device time of the optimizer tail (clip_grad_norm_ + Adam.step, train.py:203-205) on libfnst's three multi-tensor
    kernels: a twin of the network's 58 parameter tensors (the benchmark's own weights are not touched), captured as a
    CUDA graph, replayed back to back, CUDA events on the launching stream.  The working set (p, g, m, v = 100 MB) fits
    the 126 MB L2, so 256 MB are written between replays.  Returns (microseconds per tail, algorithmic bytes per tail:
    4 + 8 + 28 bytes per parameter element)."""
    from fast_neural_style_transfer_b200 import optim as fnst_optim
    if os.environ.get("FNST_BENCH_NO_ROOFLINE"):
        return float("nan"), 0.0
    twins = [torch.nn.Parameter(p.detach().clone()) for p in net.parameters()]
    for p in twins:
        p.grad = torch.randn_like(p) * 1e-2
    opt = fnst_optim.Adam(twins, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-5)
    def tail():
        fnst_optim.clip_grad_norm_(twins, max_norm=1.0)
        opt.step()
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        tail()
    torch.cuda.current_stream(dev).wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        tail()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    total, reps = 0.0, 20
    for _ in range(reps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record()
        torch.cuda.synchronize()
        total += e0.elapsed_time(e1)
    return 1e3 * total / reps, 40.0 * sum(p.numel() for p in twins)


# -------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default=os.environ.get("FNST_BENCH_WORKLOAD", "train"), choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="fp16", choices=["fp16", "bf16", "fp32", "fp16x3"])
    ap.add_argument("--optimizer", default="fnst", choices=["fnst", "torch"],
                    help="train workload: clip_grad_norm_ + Adam on libfnst's multi-tensor kernels (default) or torch's foreach "
                         "implementations (A/B only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--l2-flush", action="store_true", help="write a 256 MB buffer between timed iterations (small workloads)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, wl)

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    rank, world, local = dist_setup(args.gpus)
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    args.warmup = max(args.warmup, 3)

    import bench_data
    from fast_neural_style_transfer_b200 import ops
    sys.path.insert(0, os.path.join(ROOT, "fast_neural_style_transfer_b200", "dropin"))
    from models.model import StyleTransferNet

    peaks = load_peaks()
    net = StyleTransferNet()
    net.load_state_dict(bench_data.net_state_dict(seed=0))
    net = net.to(dev).eval()
    net.precision = args.precision

    if args.workload == "train":
        import bench_train
        return bench_train.run(args, wl, net, rank, world, dev, peaks)

    per_rank = wl["batch"] // world if wl["scaling"] == "strong" else wl["batch"]
    total_images = per_rank * world
    g = torch.Generator().manual_seed(1234 + rank)
    x_host = torch.rand((per_rank, 3, wl["h"], wl["w"]), generator=g).pin_memory()
    x = x_host.to(dev)
    y_host = None

    # ---- device-resident throughput ------------------------------------------------------------------
    timer = ops.KernelTimer(tag_prefix="res")
    with torch.no_grad():
        for _ in range(args.warmup):
            y = net(x)
        barrier(world)
        sampler = ClockSampler(local) if rank == 0 else None
        l0 = ops.launch_count
        ops.kernel_timer = timer
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        small = per_rank * wl["h"] * wl["w"] * 64 * 2 * 4 < 2 * 126e6 or args.l2_flush    # working set near/below the 126 MB L2
        if small:
            flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
            pairs = []
            for _ in range(args.steps):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); y = net(x); b.record()
                pairs.append((a, b))
            barrier(world)
            ms_local = sum(a.elapsed_time(b) for a, b in pairs)
        else:
            e0.record()
            for _ in range(args.steps):
                y = net(x)
            e1.record()
            barrier(world)
            ms_local = e0.elapsed_time(e1)
        ops.kernel_timer = None
        launches = ops.launch_count - l0
        clocks = sampler.stop() if sampler else None
        ms = max_over_ranks(ms_local, world, dev)
        value = total_images * args.steps / (ms / 1e3)

        # ---- end to end through the drop-in module with host buffers -----------------------------------
        y_host = torch.empty(y.shape, dtype=y.dtype).pin_memory()
        for _ in range(2):
            y_host.copy_(net(x_host.to(dev, non_blocking=True)), non_blocking=True)
        barrier(world)
        e0.record()
        for _ in range(args.steps):
            y_host.copy_(net(x_host.to(dev, non_blocking=True)), non_blocking=True)
        e1.record()
        barrier(world)
        ms_e2e = max_over_ranks(e0.elapsed_time(e1), world, dev)

        # ---- same, through the uint8 extension: uint8 HWC host buffers, pre/post-processing on the GPU
        u8_host = (x_host.permute(0, 2, 3, 1) * 255).round().to(torch.uint8).contiguous().pin_memory()
        u8_out = torch.empty((per_rank, y.shape[2], y.shape[3], 3), dtype=torch.uint8).pin_memory()
        for _ in range(2):
            u8_out.copy_(net.stylize_uint8(u8_host.to(dev, non_blocking=True)), non_blocking=True)
        barrier(world)
        e0.record()
        for _ in range(args.steps):
            u8_out.copy_(net.stylize_uint8(u8_host.to(dev, non_blocking=True)), non_blocking=True)
        e1.record()
        barrier(world)
        ms_u8 = max_over_ranks(e0.elapsed_time(e1), world, dev)

    # ---- roofline of the dominant kernel: the 3x3 256->256 gather-GEMM (ten launches per forward) -----
    k_ms, flops, k_launches = time_dominant_kernel(net, per_rank, wl["h"], wl["w"], dev)
    step_share = timer.mean_ms() * 10 / (ms / args.steps) if timer.count() else (k_ms * 10) / (ms / args.steps)
    achieved = flops / (k_ms * 1e-3) / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_ncu_conv_tc_full.json")
    if os.path.exists(tpath) and args.workload == "infer256" and per_rank == 256:
        with open(tpath) as f:
            traffic = json.load(f)["dram_bytes_per_launch"]
    roofline = {"bound": "tensor", "kernel": "conv_tc_kernel (3x3 256->256 residual conv)", "achieved": achieved,
                "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "frac": achieved / peaks["tf_sustained"],
                "traffic": traffic, "peak_source": peaks["src"] + " (sustained bf16/fp16)", "launches_timed": k_launches,
                "kernel_ms": k_ms, "kernel_share_of_step": step_share,
                "method": "back-to-back launches from a CUDA graph over rotating buffers > L2, CUDA events on the launching stream"}
    if rank != 0:
        return
    line = {"metric": wl["metric"], "value": value, "unit": wl["unit"], "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None,
            "dtype": {"fp16": "f16", "bf16": "bf16", "fp32": "f32", "fp16x3": "f16x3 (hi,lo split, fp32-class)"}[args.precision], "data": "synthetic",
            "config": {"workload": args.workload, "desc": wl["desc"], "per_gpu_batch": per_rank, "image": [wl["h"], wl["w"]],
                       "l2": ("L2 flushed (256 MB write) between timed iterations; each iteration timed with its own event pair" if small
                              else "inputs+activations per step exceed the 126 MB L2 (no flush needed)"),
                       "weights": "random init (seed 0)"},
            "whole_step_tflops": value * NET_GFLOP_256 * (wl["h"] * wl["w"]) / 65536.0 / 1e3 / world,
            "roofline": roofline,
            "e2e": {"value": total_images * args.steps / (ms_e2e / 1e3), "unit": wl["unit"],
                    "h2d_bytes_per_step": x_host.numel() * 4, "d2h_bytes_per_step": y_host.numel() * 4},
            "e2e_uint8": {"value": total_images * args.steps / (ms_u8 / 1e3), "unit": wl["unit"],
                          "h2d_bytes_per_step": u8_host.numel(), "d2h_bytes_per_step": u8_out.numel(),
                          "note": "extension beyond the reference API: StyleTransferNet.stylize_uint8 (uint8 HWC in/out)"},
            "gpu_launches": launches, "clocks": clocks}
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args.workload, wl)
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
