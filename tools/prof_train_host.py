"""Host-side profile of the training step (cProfile) + batch-size sensitivity (host-bound check)."""
import cProfile, pstats, sys, os, io, time, json
os.environ.setdefault('FNST_VGG19_RANDOM_INIT', '1')
import torch
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
from fast_neural_style_transfer_b200 import optim as fo
opt = fo.Adam(net.parameters(), lr=1e-3, weight_decay=1e-5)
def step(x):
    y = torch.clamp(net(x), -3, 3)
    with torch.no_grad(): cf = vgg(x)
    sf = vgg(y)
    total = 1000.0 * L.content_loss(sf, cf) + L.style_loss(sf, targets) + 10 * L.total_variation_loss(y)
    if torch.isnan(total) or torch.isinf(total): raise RuntimeError
    opt.zero_grad(); total.backward()
    fo.clip_grad_norm_(net.parameters(), 1.0); opt.step()
for b in (4,):
    x = O.make_image(b, 256, 256, seed=1, normalized=True).to(dev)
    for _ in range(5): step(x)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20): step(x)
    torch.cuda.synchronize(); print(f"batch {b}: {(time.perf_counter()-t0)/20*1e3:.2f} ms/step")
x = O.make_image(4, 256, 256, seed=1, normalized=True).to(dev)
pr = cProfile.Profile(); pr.enable()
for _ in range(10): step(x)
torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(40); print(s.getvalue()[:9000])
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(45); print(s.getvalue()[:11000])
