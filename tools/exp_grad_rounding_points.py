"""CPU experiment (oracle arithmetic only): which 16-bit rounding points of the tensor-core training path cost gradient
accuracy?  Every tensor boundary of the GPU pipeline is modelled by a straight-through rounding op with separate
forward / backward element types, and every convolution by a function whose data-gradient and weight-gradient GEMMs see
separately rounded operands -- so a single CPU autograd pass reproduces the precision plan of backward.py and variants of
it can be compared against the exact fp32 step without a GPU.

    python tools/exp_grad_rounding_points.py [batch] [size]
"""
import sys
import torch
import torch.nn.functional as F

sys.path.insert(0, '.')
from oracle import stylenet_oracle as O

torch.set_num_threads(8)
F16, BF16 = torch.float16, torch.bfloat16


def q(x, dt, scale=1.0):
    if dt is None:
        return x
    if scale != 1.0:
        return (x * scale).to(dt).to(x.dtype) / scale
    return x.to(dt).to(x.dtype)


class Round(torch.autograd.Function):
    """y = round_fwd(x); dx = round_bwd(dy * s) / s."""
    @staticmethod
    def forward(ctx, x, fwd, bwd, scale):
        ctx.bwd, ctx.scale = bwd, scale
        return q(x, fwd)

    @staticmethod
    def backward(ctx, g):
        return q(g, ctx.bwd, ctx.scale), None, None, None


class ConvSim(torch.autograd.Function):
    """op(x, w) with operands rounded per GEMM: forward w -> cfg['w_fwd']; data gradient w -> cfg['w_bwd'];
    weight gradient activation -> cfg['a_wgrad'] (the incoming gradient is rounded by the surrounding Round ops)."""
    @staticmethod
    def forward(ctx, x, w, op, cfg):
        ctx.op, ctx.cfg = op, cfg
        ctx.save_for_backward(x, w)
        return op(x, q(w, cfg.get('w_fwd')))

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        cfg, op = ctx.cfg, ctx.op
        with torch.enable_grad():
            x1 = x.detach().requires_grad_(True)
            gx, = torch.autograd.grad(op(x1, q(w.detach(), cfg.get('w_bwd'))), x1, g)
            w1 = w.detach().requires_grad_(True)
            gw, = torch.autograd.grad(op(q(x.detach(), cfg.get('a_wgrad')), w1), w1, g)
        return gx, gw, None, None


def run(cfg, b, s):
    """cfg keys: act (fwd dtype of activations), raw (fwd dtype of raw conv outputs), g_act / g_y / g_raw (dtypes of the
    gradient at the activation, at the InstanceNorm output and at the raw conv output), gscale (scale applied before
    rounding gradients), w_fwd / w_bwd / a_wgrad (GEMM operand dtypes), vgg_act, vgg_g, vgg_w."""
    p = O.make_net_params(seed=0, random_affine=True)
    vp = O.make_vgg_params(seed=1)
    content = O.make_image(b, s, s, seed=5, normalized=True)
    sty = O.make_image(1, s, s, seed=6, normalized=True)
    targets = O.style_targets(vp, sty)
    leaf = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    gs = cfg.get('gscale', 1.0)
    R = lambda x, f, bk: Round.apply(x, cfg.get(f), cfg.get(bk), gs)

    def conv(x, name, stride):
        pad = leaf[name + ".weight"].shape[-1] // 2
        xp = F.pad(x, (pad,) * 4, mode="reflect")
        y = ConvSim.apply(xp, leaf[name + ".weight"], lambda a, w: F.conv2d(a, w, None, stride=stride), cfg)
        return y + leaf[name + ".bias"].view(1, -1, 1, 1)

    def convT(x, name):
        y = ConvSim.apply(x, leaf[name + ".weight"], lambda a, w: F.conv_transpose2d(a, w, None, stride=2, padding=1, output_padding=1), cfg)
        return y + leaf[name + ".bias"].view(1, -1, 1, 1)

    def block(x, conv_fn, norm, relu=True):
        raw = R(conv_fn(x), 'raw', 'g_raw')
        y = R(O.instance_norm(raw, leaf[norm + ".weight"], leaf[norm + ".bias"]), None, 'g_y')
        return F.relu(y) if relu else y

    h = R(block(content, lambda t: conv(t, "conv1.conv", 2), "norm1"), 'act', 'g_act')
    h = R(block(h, lambda t: conv(t, "conv2.conv", 2), "norm2"), 'act', 'g_act')
    for i in range(5):
        pre = f"res_blocks.{i}"
        y = R(block(h, lambda t: conv(t, pre + ".conv1.conv", 1), pre + ".in1"), 'act', 'g_act')
        y = block(y, lambda t: conv(t, pre + ".conv2.conv", 1), pre + ".in2", relu=False)
        h = R(h + y, 'act', 'g_act')
    h = R(block(h, lambda t: convT(t, "up1.upsample_conv"), "norm3"), 'act', 'g_act')
    h = R(block(h, lambda t: convT(t, "up2.upsample_conv"), "norm4"), 'act', 'g_act')
    y = torch.clamp(conv(h, "final_conv.conv", 1), -3, 3)

    vcfg = dict(w_fwd=cfg.get('vgg_w'), w_bwd=cfg.get('vgg_w'))

    def vgg(x):
        def cr(t, name):
            o = ConvSim.apply(t, vp[name + ".weight"], lambda a, w: F.conv2d(a, w, None, padding=1), vcfg) + vp[name + ".bias"].view(1, -1, 1, 1)
            return Round.apply(F.relu(o), cfg.get('vgg_act'), cfg.get('vgg_g'), gs)
        h = cr(cr(x, "slice1.0"), "slice1.2"); f0 = h
        h = cr(cr(F.max_pool2d(h, 2, 2), "slice2.5"), "slice2.7"); f1 = h
        h = cr(cr(cr(F.max_pool2d(h, 2, 2), "slice3.10"), "slice3.12"), "slice3.14"); f2 = h
        h = cr(h, "slice4.16"); h = cr(cr(F.max_pool2d(h, 2, 2), "slice4.19"), "slice4.21"); f3 = h
        return [f0, f1, f2, f3, cr(h, "slice5.23")]

    with torch.no_grad():
        cf = vgg(content)
    sf = vgg(y)
    total = 1000 * O.content_loss(sf, cf) + O.style_loss(sf, targets) + 10 * O.total_variation_loss(y)
    total.backward()
    return float(total), {k: v.grad for k, v in leaf.items()}


def main():
    b = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    s = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    l0, g0 = run({}, b, s)
    gn = float(torch.sqrt(sum((x.double() ** 2).sum() for x in g0.values())))
    fwd = dict(act=F16, raw=F16, w_fwd=F16, vgg_act=BF16, vgg_w=BF16)
    cur = dict(fwd, g_act=BF16, g_y=BF16, g_raw=BF16, w_bwd=BF16, a_wgrad=BF16, vgg_g=BF16)
    cases = {
        "forward 16-bit only (exact backward)": fwd,
        "backward bf16 only (exact forward)": dict(g_act=BF16, g_y=BF16, g_raw=BF16, w_bwd=BF16, a_wgrad=BF16, vgg_g=BF16),
        "CURRENT plan (fp16 fwd, bf16 grads everywhere)": cur,
        "current, gy fp32 (one-pass IN backward)": dict(cur, g_y=None),
        "current, gy + d_act fp32": dict(cur, g_y=None, g_act=None),
        "current, gy + d_act fp32, wgrad A fp16": dict(cur, g_y=None, g_act=None, a_wgrad=F16),
        "current, net grads fp16 x 2^-14 (VGG bf16)": dict(cur, g_act=F16, g_y=F16, g_raw=F16, w_bwd=F16, a_wgrad=F16, gscale=2.0 ** -14),
        "all grads fp16 x 2^-14 incl. VGG": dict(cur, g_act=F16, g_y=F16, g_raw=F16, w_bwd=F16, a_wgrad=F16, vgg_g=F16, vgg_w=F16, gscale=2.0 ** -14),
        "all fp16 x 2^-14, VGG acts fp16": dict(cur, g_act=F16, g_y=F16, g_raw=F16, w_bwd=F16, a_wgrad=F16, vgg_g=F16, vgg_w=F16, vgg_act=F16, gscale=2.0 ** -14),
    }
    bwd = dict(g_act=BF16, g_y=BF16, g_raw=BF16, w_bwd=BF16, a_wgrad=BF16, vgg_g=BF16)
    cases.update({
        "net forward exact (fp16x3 class), VGG bf16, bf16 backward": dict(bwd, vgg_act=BF16, vgg_w=BF16),
        "net forward exact (fp16x3 class), VGG fp16, bf16 backward": dict(bwd, vgg_act=F16, vgg_w=F16),
        "raw fp32 + act fp16 + w fp16, VGG bf16, bf16 backward": dict(bwd, act=F16, w_fwd=F16, vgg_act=BF16, vgg_w=BF16),
        "raw fp16 only (act, w exact), VGG bf16, bf16 backward": dict(bwd, raw=F16, vgg_act=BF16, vgg_w=BF16),
        "w fp16 only, VGG bf16, bf16 backward": dict(bwd, w_fwd=F16, vgg_act=BF16, vgg_w=BF16),
    })
    if len(sys.argv) > 3:
        cases = {k: v for k, v in cases.items() if sys.argv[3] in k}
    for name, cfg in cases.items():
        l, g = run(cfg, b, s)
        errs = {k: float((g[k].double() - g0[k].double()).norm()) / max(float(g0[k].double().norm()), 1e-4 * gn) for k in g0}
        flat = torch.cat([g[k].double().flatten() for k in g0]); flat0 = torch.cat([g0[k].double().flatten() for k in g0])
        glob = float((flat - flat0).norm() / flat0.norm())
        top = sorted(errs, key=errs.get, reverse=True)[:2]
        prof = [round(errs[k], 4) for k in ("final_conv.conv.weight", "up2.upsample_conv.weight", "up1.upsample_conv.weight", "res_blocks.4.conv2.conv.weight",
                                            "res_blocks.2.conv1.conv.weight", "res_blocks.0.conv1.conv.weight", "conv2.conv.weight", "conv1.conv.weight")]
        print(f"{name}: loss {abs(l / l0 - 1):.1e} | worst {max(errs.values()):.3e} {top[0]} | global {glob:.2e} | final..conv1 {prof}", flush=True)


if __name__ == "__main__":
    main()
