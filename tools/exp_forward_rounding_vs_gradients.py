"""CPU experiment (oracle only): gradient error caused purely by rounding FORWARD activations to 16 bits, with an
exact fp32 backward (straight-through rounding).  Shows that ReLU / max-pool mask flips, not the gradient dtype,
dominate the tensor-core path's per-tensor gradient gap (precision).  Run: python tools/exp_forward_rounding_vs_gradients.py"""
import torch, sys
sys.path.insert(0, '.')
from oracle import stylenet_oracle as O
import torch.nn.functional as F
torch.set_num_threads(8)
class RoundSTE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, dt): return x.to(dt).to(x.dtype)
    @staticmethod
    def backward(ctx, g): return g, None
def rnd(x, dt): return RoundSTE.apply(x, dt) if dt is not None else x
def run(act_dt, vgg_dt):
    p = O.make_net_params(seed=0, random_affine=True); vp = O.make_vgg_params(seed=1)
    content = O.make_image(2,64,64,seed=5,normalized=True); sty = O.make_image(1,64,64,seed=6,normalized=True)
    targets = O.style_targets(vp, sty)
    leaf = {k: v.clone().requires_grad_(True) for k,v in p.items()}
    # forward with activations rounded (straight-through) after every norm/relu stage -- exact fp32 backward
    def inorm(x,g,b): return O.instance_norm(x,g,b)
    h = rnd(F.relu(inorm(O.conv_layer(content, leaf["conv1.conv.weight"], leaf["conv1.conv.bias"],2), leaf["norm1.weight"], leaf["norm1.bias"])), act_dt)
    h = rnd(F.relu(inorm(O.conv_layer(h, leaf["conv2.conv.weight"], leaf["conv2.conv.bias"],2), leaf["norm2.weight"], leaf["norm2.bias"])), act_dt)
    for i in range(5):
        pre=f"res_blocks.{i}"
        y = rnd(F.relu(inorm(O.conv_layer(h, leaf[pre+".conv1.conv.weight"], leaf[pre+".conv1.conv.bias"],1), leaf[pre+".in1.weight"], leaf[pre+".in1.bias"])), act_dt)
        y = inorm(O.conv_layer(y, leaf[pre+".conv2.conv.weight"], leaf[pre+".conv2.conv.bias"],1), leaf[pre+".in2.weight"], leaf[pre+".in2.bias"])
        h = rnd(h + y, act_dt)
    h = rnd(F.relu(inorm(O.upsample_conv(h, leaf["up1.upsample_conv.weight"], leaf["up1.upsample_conv.bias"]), leaf["norm3.weight"], leaf["norm3.bias"])), act_dt)
    h = rnd(F.relu(inorm(O.upsample_conv(h, leaf["up2.upsample_conv.weight"], leaf["up2.upsample_conv.bias"]), leaf["norm4.weight"], leaf["norm4.bias"])), act_dt)
    y = torch.clamp(O.conv_layer(h, leaf["final_conv.conv.weight"], leaf["final_conv.conv.bias"],1), -3, 3)
    def vgg(x):
        def cr(t,name): return rnd(F.relu(F.conv2d(t, vp[name+".weight"], vp[name+".bias"], padding=1)), vgg_dt)
        h = cr(cr(x,"slice1.0"),"slice1.2"); f0=h
        h = cr(cr(F.max_pool2d(h,2,2),"slice2.5"),"slice2.7"); f1=h
        h = cr(cr(cr(F.max_pool2d(h,2,2),"slice3.10"),"slice3.12"),"slice3.14"); f2=h
        h = cr(h,"slice4.16"); h = cr(cr(F.max_pool2d(h,2,2),"slice4.19"),"slice4.21"); f3=h
        return [f0,f1,f2,f3,cr(h,"slice5.23")]
    with torch.no_grad(): cf = vgg(content)
    sf = vgg(y)
    total = 1000*O.content_loss(sf,cf) + O.style_loss(sf,targets) + 10*O.total_variation_loss(y)
    total.backward()
    return float(total), {k:v.grad for k,v in leaf.items()}
l0,g0 = run(None,None)
for name,(a,v) in {"fp16 net acts only":(torch.float16,None), "bf16 vgg acts only":(None,torch.bfloat16), "fp16 net + bf16 vgg":(torch.float16,torch.bfloat16)}.items():
    l,g = run(a,v)
    gn = float(torch.sqrt(sum((x.double()**2).sum() for x in g0.values())))
    errs = {k: float((g[k].double()-g0[k].double()).norm())/max(float(g0[k].double().norm()),1e-4*gn) for k in g0}
    top = sorted(errs,key=errs.get,reverse=True)[:3]
    print(f"{name}: loss rel {abs(l/l0-1):.2e}; worst per-tensor grad err {max(errs.values()):.3e}; top {[(k,round(errs[k],4)) for k in top]}")
