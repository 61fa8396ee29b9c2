"""Where does the graphed training step spend its time?  CUDA-event timing of the step's phases."""
import sys, os, time, torch
os.environ.setdefault('FNST_VGG19_RANDOM_INIT', '1')
sys.path.insert(0, '.'); sys.path.insert(0, 'fast_neural_style_transfer_b200/dropin')
from oracle import stylenet_oracle as O
from models.model import StyleTransferNet
from models.vgg19_net import VGG19
from losses import losses as L
dev = torch.device('cuda', 0)
net = StyleTransferNet(); net.load_state_dict(O.make_net_params(seed=0)); net = net.to(dev).train()
vgg = VGG19(); vgg.load_state_dict(O.make_vgg_params(seed=1)); vgg = vgg.to(dev).eval(); vgg.precision = 'bf16'
for p in vgg.parameters(): p.requires_grad = False
with torch.no_grad():
    targets = [L.gram_matrix(f).squeeze(0).detach() for f in vgg(O.make_image(1, 256, 256, seed=4321, normalized=True).to(dev))]
opt = torch.optim.Adam(net.parameters(), lr=1e-3, weight_decay=1e-5)
x = O.make_image(4, 256, 256, seed=1, normalized=True).to(dev)
marks = {}
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e
def step(rec):
    t = [ev()]; h = [time.perf_counter()]
    def mark(): t.append(ev()); h.append(time.perf_counter())
    y = torch.clamp(net(x), -3, 3); mark()
    with torch.no_grad(): cf = vgg(x)
    mark()
    sf = vgg(y); mark()
    total = 1000.0 * L.content_loss(sf, cf) + L.style_loss(sf, targets) + 10 * L.total_variation_loss(y); mark()
    bad = torch.isnan(total) or torch.isinf(total); mark()
    opt.zero_grad(); total.backward(); mark()
    torch.nn.utils.clip_grad_norm_(net.parameters(), 1.0); opt.step(); mark()
    if rec:
        torch.cuda.synchronize()
        names = ["net fwd", "vgg(content)", "vgg(stylized)", "losses", "nan check", "backward", "clip+adam"]
        for i, n in enumerate(names):
            marks.setdefault(n, [0.0, 0.0]); marks[n][0] += t[i].elapsed_time(t[i + 1]); marks[n][1] += (h[i + 1] - h[i]) * 1e3
for _ in range(6): step(False)
torch.cuda.synchronize()
N = 20
for _ in range(N): step(True)
tot_g = sum(v[0] for v in marks.values()) / N; tot_h = sum(v[1] for v in marks.values()) / N
for n, v in marks.items(): print(f"{n:16s} gpu-timeline {v[0]/N:7.3f} ms   host {v[1]/N:7.3f} ms")
print(f"{'total':16s} gpu-timeline {tot_g:7.3f} ms   host {tot_h:7.3f} ms")
# pure replay cost of the captured graphs
st = list(net._train_graphs.values())[0]
for name, g in (("net fwd graph", st.fwd), ("net bwd graph", st.bwd)):
    torch.cuda.synchronize(); a = ev()
    for _ in range(20): g.graph.replay()
    b = ev(); torch.cuda.synchronize(); print(f"{name}: {a.elapsed_time(b)/20:.3f} ms per replay, {g.launches} launches")
for key, vg in vgg._graphs.items():
    torch.cuda.synchronize(); a = ev()
    for _ in range(20): vg.fwd.graph.replay()
    b = ev(); torch.cuda.synchronize(); print(f"vgg fwd graph {key[1]} grad={key[3]}: {a.elapsed_time(b)/20:.3f} ms, {vg.fwd.launches} launches")
    for pat, bg in vg.bwd.items():
        torch.cuda.synchronize(); a = ev()
        for _ in range(20): bg.graph.replay()
        b = ev(); torch.cuda.synchronize(); print(f"vgg bwd graph {pat}: {a.elapsed_time(b)/20:.3f} ms, {bg.launches} launches")
