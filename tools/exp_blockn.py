import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'fast_neural_style_transfer_b200/dropin')
import bench
from oracle import stylenet_oracle as O
from models.model import StyleTransferNet
dev = torch.device('cuda', 0)
net = StyleTransferNet(); net.load_state_dict(O.make_net_params(seed=0)); net = net.to(dev).eval()
for b in (1, 4, 8, 16):
    ms, fl, n = bench.time_dominant_kernel(net, b, 256, 256, dev)
    print(f"batch {b}: {ms*1e3:.1f} us  {fl/ms/1e9:.0f} TFLOP/s")
