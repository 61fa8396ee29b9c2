#!/usr/bin/env python
"""Real (concurrent) GPU timeline of the training step via torch.profiler (CUPTI): per-kernel time, busy time per stream,
union busy time, idle gaps.  Unlike an ncu launch list (serialised, cold caches) this shows overlap and bubbles.  B200 only.

    python tools/prof_train_timeline.py [steps] > gpurun_out/timeline.json
"""
import json, os, sys, collections
os.environ.setdefault("FNST_VGG19_RANDOM_INIT", "1")
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fast_neural_style_transfer_b200", "dropin"))
import bench_data
from fast_neural_style_transfer_b200 import optim as fo
from models.model import StyleTransferNet
from models.vgg19_net import VGG19
from losses import losses as L
from torch.profiler import profile, ProfilerActivity

dev = torch.device("cuda", 0)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
net = StyleTransferNet(); net.load_state_dict(bench_data.net_state_dict(0)); net = net.to(dev).train(); net.precision = "fp16"
vgg = VGG19(); vgg.load_state_dict(bench_data.vgg_state_dict(1)); vgg = vgg.to(dev).eval(); vgg.precision = "bf16"
with torch.no_grad():
    targets = [L.gram_matrix(f).squeeze(0).detach() for f in vgg(bench_data.image_batch(1, 256, 256, 4321, True).to(dev))]
opt = fo.Adam(net.parameters(), lr=1e-3, weight_decay=1e-5)
xs = [bench_data.image_batch(4, 256, 256, 1234 + i, True).to(dev) for i in range(4)]

def step(x):
    y = torch.clamp(net(x), -3, 3)
    with torch.no_grad():
        cf = vgg(x)
    sf = vgg(y)
    total = 1000.0 * L.content_loss(sf, cf) + L.style_loss(sf, targets) + 10 * L.total_variation_loss(y)
    if torch.isnan(total) or torch.isinf(total):
        raise RuntimeError("bad loss")
    opt.zero_grad(); total.backward()
    fo.clip_grad_norm_(net.parameters(), 1.0); opt.step()

for i in range(6):
    step(xs[i % 4])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(steps):
        step(xs[i % 4])
    torch.cuda.synchronize()
path = os.path.join(ROOT, "gpurun_out", "train_trace.json")
prof.export_chrome_trace(path)
ev = json.load(open(path))["traceEvents"]
ks = [e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
ks.sort(key=lambda e: e["ts"])
t0, t1 = ks[0]["ts"], max(e["ts"] + e["dur"] for e in ks)
per = collections.defaultdict(lambda: [0, 0.0])
streams = collections.defaultdict(float)
for e in ks:
    name = e["name"].split("(")[0][:60]
    per[name][0] += 1; per[name][1] += e["dur"]
    streams[e["args"].get("stream", 0)] += e["dur"]
# union busy time
iv = sorted((e["ts"], e["ts"] + e["dur"], e["name"].split("(")[0][:48]) for e in ks)
busy, cur_s, cur_e, gaps, last_name, gap_sites = 0.0, iv[0][0], iv[0][1], [], iv[0][2], []
for s, e_, nm in iv[1:]:
    if s > cur_e:
        busy += cur_e - cur_s
        gaps.append((s - cur_e, cur_e - t0))
        if s - cur_e > 15:
            gap_sites.append((round(s - cur_e, 1), last_name, nm))
        cur_s, cur_e, last_name = s, e_, nm
    else:
        if e_ >= cur_e:
            last_name = nm
        cur_e = max(cur_e, e_)
busy += cur_e - cur_s
span = t1 - t0
out = {"steps": steps, "span_us_per_step": span / steps, "union_busy_us_per_step": busy / steps, "idle_us_per_step": (span - busy) / steps,
       "sum_kernel_us_per_step": sum(v[1] for v in per.values()) / steps, "kernels_per_step": len(ks) / steps,
       "streams_busy_us_per_step": {str(k): v / steps for k, v in streams.items()},
       "gaps_over_5us_per_step": sum(1 for g, _ in gaps if g > 5) / steps, "gap_time_over_5us_per_step": sum(g for g, _ in gaps if g > 5) / steps,
       "largest_gaps_us": sorted((round(g, 1) for g, _ in gaps), reverse=True)[:12],
       "gaps_over_15us_of_the_last_step": [g for g in gap_sites[-(len(gap_sites) // steps):]],
       "top_kernels_us_per_step": [(k, v[0] / steps, round(v[1] / steps, 1)) for k, v in sorted(per.items(), key=lambda kv: -kv[1][1])[:45]]}
print(json.dumps(out, indent=1))
