"""Training-step leg of bench.py: the reference's step body (train.py:168-206) on the drop-in modules.

One step = content batch (4 per GPU, 256x256) -> StyleTransferNet (train mode, Dropout2d live) -> clamp
-> VGG-19 twice -> content/style/TV losses -> NaN check -> zero_grad -> backward -> [DP: gradient
all-reduce] -> clip_grad_norm_(1.0) -> Adam(lr 1e-3, wd 1e-5) -> CosineAnnealingLR step.
"""
import json
import os
import sys

import torch

import bench as B

ROOT = os.path.dirname(os.path.abspath(__file__))


def run(args, wl, net, rank, world, dev, peaks):
    import bench_data
    from fast_neural_style_transfer_b200 import ops, parallel
    from fast_neural_style_transfer_b200 import optim as fnst_optim
    from models.vgg19_net import VGG19
    from losses import losses as L

    bsz, h, w = wl["batch"], wl["h"], wl["w"]
    vgg = VGG19()
    vgg.load_state_dict(bench_data.vgg_state_dict(seed=1))
    vgg = vgg.to(dev).eval()
    vgg.precision = {"fp32": "fp32", "fp16x3": "fp16"}.get(args.precision, "bf16")
    for p in vgg.parameters():
        p.requires_grad = False
    net.train()
    style = bench_data.image_batch(1, h, w, seed=4321, normalized=True).to(dev)
    with torch.no_grad():
        targets = [L.gram_matrix(f).squeeze(0).detach() for f in vgg(style)]           # train.py:25-37
    # train.py:135-139 / :203: same names and arguments, libfnst multi-tensor kernels unless --optimizer torch
    adam_cls, clip_fn = ((fnst_optim.Adam, fnst_optim.clip_grad_norm_) if args.optimizer == "fnst"
                         else (torch.optim.Adam, torch.nn.utils.clip_grad_norm_))
    opt = adam_cls(net.parameters(), lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-5)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=80000, eta_min=1e-7)
    dp = parallel.GradientAllReduce(net, world, overlap=os.environ.get("FNST_DP_OVERLAP", "0") != "0") if world > 1 else None
    torch.manual_seed(1000 + rank)                                                      # per-rank dropout streams
    g = torch.Generator().manual_seed(1234 + rank)
    n_host = 4
    host_batches = [bench_data.image_batch(bsz, h, w, seed=1234 + 17 * rank + i, normalized=True).pin_memory() for i in range(n_host)]
    dev_batches = [b.to(dev) for b in host_batches]
    tv_scale = 1.0 / world          # TV is a batch mean, content/style are batch sums: SUM all-reduce

    def step(content, read_losses):
        stylized = torch.clamp(net(content), -3, 3)
        with torch.no_grad():
            cf = vgg(content)
        sf = vgg(stylized)
        c_loss = L.content_loss(sf, cf)
        s_loss = L.style_loss(sf, targets)
        tv_loss = L.total_variation_loss(stylized)
        total = 1000.0 * c_loss + 1 * s_loss + 10 * tv_scale * tv_loss
        # train.py:193 (host sync); with several ranks the decision is taken on a MIN-reduced flag so that all ranks skip together
        ok = parallel.all_finite(total, world) if world > 1 else not (torch.isnan(total) or torch.isinf(total))
        if not ok:
            raise RuntimeError("invalid loss in benchmark step")
        opt.zero_grad()
        total.backward()
        if dp is not None:
            dp.all_reduce()
        clip_fn(net.parameters(), max_norm=1.0)
        opt.step()
        sched.step()
        if read_losses:
            return total.item(), c_loss.item(), s_loss.item(), tv_loss.item()            # train.py:209-212
        return None

    for i in range(args.warmup):
        step(dev_batches[i % n_host], False)
    B.barrier(world)
    sampler = B.ClockSampler(dev.index) if rank == 0 else None
    timer = ops.KernelTimer(tag_prefix="res")
    ops.kernel_timer = timer
    l0 = ops.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(dev_batches[i % n_host], False)
    e1.record()
    B.barrier(world)
    ops.kernel_timer = None
    launches = ops.launch_count - l0
    clocks = sampler.stop() if sampler else None
    ms = B.max_over_ranks(e0.elapsed_time(e1), world, dev)

    # end to end: host batch -> H2D -> step -> four loss scalars read back
    for i in range(2):
        step(host_batches[i % n_host].to(dev, non_blocking=True), True)
    B.barrier(world)
    e0.record()
    for i in range(args.steps):
        losses = step(host_batches[i % n_host].to(dev, non_blocking=True), True)
    e1.record()
    B.barrier(world)
    ms_e2e = B.max_over_ranks(e0.elapsed_time(e1), world, dev)

    k_ms, flops, k_launches = B.time_dominant_kernel(net, bsz, h, w, dev)
    tail_us, tail_bytes = B.time_optimizer_tail(net, dev) if args.optimizer == "fnst" else (float("nan"), 0.0)
    achieved = flops / (k_ms * 1e-3) / 1e12
    norm = B.time_norm_kernels(dev, peaks)
    if rank != 0:
        return None
    # DRAM traffic of the dominant kernel per launch: from the ncu --set full capture of THIS command kept under profiles/
    # (ncu cannot run inside a timed bench run); null when the capture is not there
    traffic, traffic_note = None, "not measured in this run (needs ncu); profiles/ holds the ncu --set full capture of this kernel"
    try:
        for row in json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_conv_tc.json"))):
            if "conv_tc_kernel<256, 0>" in row["kernel"] and row["grid"].startswith("(128"):
                traffic = row["dram_read_bytes"] + row["dram_write_bytes"]
                traffic_note = ("dram__bytes_read.sum + dram__bytes_write.sum per launch from profiles/r02_ncu_conv_tc.json (ncu --set full "
                                "--clock-control none of `bench.py --workload train`, cold cache per replay); algorithmic bytes = 8.9 MB "
                                "halo activations + 1.2 MB weights read, 8.4 MB output written (stays in L2)")
    except Exception:
        pass
    value = world * args.steps / (ms / 1e3)
    line = {"metric": wl["metric"], "value": value, "unit": wl["unit"], "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None,
            "dtype": {"fp16": "f16 fwd / bf16 grads", "bf16": "bf16", "fp32": "f32",
                      "fp16x3": "f16x3 fwd (hi,lo split, fp32-class) / bf16 grads"}[args.precision], "data": "synthetic",
            "config": B.train_config(world, bsz, h, w),
            "optimizer_impl": "libfnst multi-tensor kernels" if args.optimizer == "fnst" else "torch foreach",
            "precision_note": {"fp16": "fp16 activations / bf16 gradients on tcgen05 (outputs and losses within the 1e-2 gate); VGG-19 in bf16",
                               "fp16x3": "fp32-class forward on tcgen05 (fp16 hi,lo pairs, 3 MMAs per product) + bf16 backward: losses 1e-4, "
                                         "every gradient tensor within 2e-2 of the fp32 oracle; VGG-19 fp16 inside / bf16 interface"}.get(args.precision, ""),
            "images_per_s": value * bsz,
            "whole_step_tflops": value * bsz * B.TRAIN_GFLOP_IMG / 1e3 / world,
            "roofline": {"bound": "tensor", "kernel": "conv_tc_kernel (3x3 256->256 residual conv; forward launch, batch 4)",
                         "achieved": achieved, "peak": peaks["tf_burst"], "unit": "TFLOP/s",
                         "frac": achieved / peaks["tf_burst"], "frac_of_sustained_peak": achieved / peaks["tf_sustained"], "traffic": traffic,
                         "traffic_note": traffic_note,
                         "peak_source": peaks["src"] + " (burst bf16/fp16: the kernel is timed alone, ~20 ms of back-to-back launches)",
                         "launches_timed": k_launches, "kernel_ms": k_ms,
                         "kernel_share_of_step": (k_ms * 30) / (ms / args.steps),
                         "share_note": "30 launches of this shape per step: 10 forward + 10 data-gradient (same kernel) + 10 weight-gradient (wgrad_tc_kernel, same FLOPs)",
                         "method": "back-to-back launches from a CUDA graph over rotating buffers > L2, CUDA events on the launching stream"},
            "e2e": {"value": world * args.steps / (ms_e2e / 1e3), "unit": wl["unit"],
                    "h2d_bytes_per_step": host_batches[0].numel() * 4, "d2h_bytes_per_step": 16},
            "gpu_launches": launches, "clocks": clocks, "last_losses": losses}
    if norm:
        line["roofline_norm"] = norm
    if tail_bytes:
        gbs = tail_bytes / (tail_us * 1e-6) / 1e9
        line["roofline_optimizer_tail"] = {
            "bound": "hbm", "kernel": "mt_sqnorm_kernel + mt_scale_kernel + mt_adam_kernel (clip_grad_norm_ + Adam.step, 58 tensors)",
            "achieved": gbs, "peak": peaks["hbm"], "unit": "GB/s", "frac": gbs / peaks["hbm"], "traffic": None,
            "peak_source": peaks["src"] + " (copy bandwidth)", "us_per_tail": tail_us, "tail_share_of_step": tail_us / 1e3 / (ms / args.steps),
            "method": "three launches captured as one CUDA graph, replayed 20x with a 256 MB L2 flush in between, CUDA events"}
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = B.cpu_baseline("train", wl)
    return line
