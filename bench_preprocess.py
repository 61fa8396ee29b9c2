"""Input-transform leg of bench.py: Resize((256, 256)) + ToTensor + Normalize of a batch of decoded
1080 x 1920 uint8 RGB images (train.py:92-102 applied per image by data/dataset.py:21-27), as ONE launch per batch.

`value`: images resident in HBM, the batch launch replayed over rotating image sets larger than L2.  `e2e`: pinned host
uint8 images -> H2D -> transform -> float batch read back to pinned host memory.  Roofline: HBM, algorithmic bytes =
input image once + float output once.  `cpu_baseline`: the C restatement of Pillow's resize (oracle/pil_resize.c, one
thread) + ToTensor / Normalize on a bounded sample.
"""
import time

import numpy as np
import torch

import bench as B

IN_H, IN_W, OUT = 1080, 1920, 256
MEAN, STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


def run(args, rank, world, dev, peaks, with_cpu=True):
    from fast_neural_style_transfer_b200 import ops, preprocess as P
    per_rank, sets = 32, 4                       # 32 images x 6.2 MB = 199 MB per set: every set exceeds the 126 MB L2
    g = torch.Generator().manual_seed(99 + rank)
    host = [torch.randint(0, 256, (per_rank, IN_H, IN_W, 3), dtype=torch.uint8, generator=g).pin_memory() for _ in range(2)]
    dev_sets = [host[i % 2].to(dev) for i in range(sets)]
    outs = [torch.empty((per_rank, 3, OUT, OUT), dtype=torch.float32, device=dev) for _ in range(sets)]
    lists = [[s[i] for i in range(per_rank)] for s in dev_sets]
    for i in range(max(args.warmup, 3)):
        P.resize_to_tensor(lists[i % sets], (OUT, OUT), MEAN, STD, out=outs[i % sets])
    B.barrier(world)
    sampler = B.ClockSampler(dev.index) if rank == 0 else None
    steps = max(args.steps, 20)
    l0 = ops.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        P.resize_to_tensor(lists[i % sets], (OUT, OUT), MEAN, STD, out=outs[i % sets])
    e1.record()
    B.barrier(world)
    launches = ops.launch_count - l0
    clocks = sampler.stop() if sampler else None
    ms = B.max_over_ranks(e0.elapsed_time(e1), world, dev)

    # end to end: pinned host uint8 -> device -> transform -> pinned host float batch
    out_host = torch.empty((per_rank, 3, OUT, OUT), dtype=torch.float32).pin_memory()
    e2e_steps = 6
    def e2e_step(i):
        d = host[i % 2].to(dev, non_blocking=True)
        y = P.resize_to_tensor([d[k] for k in range(per_rank)], (OUT, OUT), MEAN, STD)
        out_host.copy_(y, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
    e2e_step(0)
    B.barrier(world)
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step(i + 1)
    B.barrier(world)
    ms_e2e = B.max_over_ranks((time.perf_counter() - t0) * 1e3, world, dev)
    if rank != 0:
        return None
    in_bytes, out_bytes = IN_H * IN_W * 3, 3 * OUT * OUT * 4
    per_launch_ms = ms / steps
    gbs = per_rank * (in_bytes + out_bytes) / (per_launch_ms * 1e-3) / 1e9
    line = {"metric": "input-transform images/sec (1080x1920 uint8 -> 3x256x256 float, Resize + ToTensor + Normalize)",
            "value": world * per_rank * steps / (ms / 1e3), "unit": "images/s", "n_gpus": world, "steps": steps, "warmup": max(args.warmup, 3),
            "ms_per_step": per_launch_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8 (22-bit fixed-point weights)",
            "data": "synthetic",
            "config": {"workload": "preprocess", "desc": "SURVEY 8f N3: the reference's per-image transform (train.py:92-102) on a batch of decoded frames",
                       "per_gpu_batch": per_rank, "image_in": [IN_H, IN_W], "image_out": [OUT, OUT],
                       "l2": f"{sets} rotating image sets of {per_rank * in_bytes / 1e6:.0f} MB each (> 126 MB L2)"},
            "roofline": {"bound": "hbm", "kernel": "resize_to_tensor_kernel (one launch per batch, grid z = image)", "achieved": gbs,
                         "peak": peaks["hbm"], "unit": "GB/s", "frac": gbs / peaks["hbm"], "traffic": None,
                         "algorithmic_bytes_per_image": in_bytes + out_bytes, "peak_source": peaks["src"] + " (copy bandwidth)",
                         "note": "bit-exact with Pillow: every tap is a byte load and the two passes clip to 8 bits in between; "
                                 "the kernel is issue-bound on byte loads, not bandwidth-bound"},
            "e2e": {"value": world * per_rank * e2e_steps / (ms_e2e / 1e3), "unit": "images/s",
                    "h2d_bytes_per_step": per_rank * in_bytes, "d2h_bytes_per_step": per_rank * out_bytes},
            "gpu_launches": launches, "clocks": clocks}
    if with_cpu:
        from oracle import pil_resize as R
        img = np.random.default_rng(5).integers(0, 256, (IN_H, IN_W, 3), dtype=np.uint8)
        R.to_tensor(R.resize_bilinear_u8(img, OUT, OUT), MEAN, STD)
        n, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < 3.0:
            R.to_tensor(R.resize_bilinear_u8(img, OUT, OUT), MEAN, STD)
            n += 1
        line["cpu_baseline"] = {"value": n / (time.perf_counter() - t0), "unit": "images/s", "cores": 1, "kind": "port",
                                "sample": f"{n} images in ~3 s on one thread (oracle/pil_resize.c: Pillow's two-pass bilinear resize + ToTensor + Normalize)"}
    return line
