"""GPU parity at the sizes BASELINE.json names (not at reduced test sizes):

  configs[1]  perceptual-loss training step, batch 4 x 3 x 256 x 256: losses, all 58 gradients (relative L2 and
              cosine), global gradient norm -- tensor-core path and fp32 path -- against the oracle's autograd
              (train.py:168-206), arbitrated by the float64 oracle
  configs[2]  1 x 3 x 1080 x 1920 forward: fp16 tensor-core path (1e-2 / 1.0 px) and fp16x3 (1e-4)
  configs[3]  256 x 3 x 256 x 256 forward in one call, checked image by image against the oracle run in shards
  configs[0]  1 x 3 x 256 x 256 in the stated fp32 class: CUDA-core fp32 and tensor-core fp16x3, both 1e-4

Every number asserted here is also written to gpurun_out/parity_baseline_configs.json (evidence for profiles/).
"""
import json
import os

import pytest
import torch

from oracle import stylenet_oracle as O

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from fast_neural_style_transfer_b200 import engine

import test_gpu_train as T          # shared step body (drop-in modules, reference loop order)
from test_gpu_train import dropin    # noqa: F401  (pytest fixture)

DEV = "cuda"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_RESULTS = {}


def _record(key, value):
    _RESULTS[key] = value
    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_baseline_configs.json"), "w") as f:
            json.dump(_RESULTS, f, indent=1, sort_keys=True)
    except OSError:
        pass


def _grad_report(grads, ref_grads):
    gn_ref = float(torch.sqrt(sum((g.double() ** 2).sum() for g in ref_grads.values())))
    gn = float(torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())))
    rows = {}
    for k, r in ref_grads.items():
        g, r = grads[k].double().flatten(), r.double().flatten()
        denom = max(float(r.norm()), 1e-4 * gn_ref)     # conv biases under InstanceNorm have ~0 gradient (SURVEY 8c i)
        cos = float(torch.dot(g, r) / (g.norm() * r.norm() + 1e-300))
        rows[k] = (float((g - r).norm()) / denom, cos, float(r.norm()) >= 1e-4 * gn_ref)
    flat_g = torch.cat([grads[k].double().flatten() for k in ref_grads])
    flat_r = torch.cat([ref_grads[k].double().flatten() for k in ref_grads])
    glob = float((flat_g - flat_r).norm() / flat_r.norm())
    return gn, gn_ref, rows, glob


# tolerance table for configs[1]: (losses, per-tensor gradient rel-L2, per-tensor cosine, global gradient rel-L2, norm)
TRAIN_TOL = {
    # tensor-core path: fp16 activations / bf16 gradients.  north_star bounds outputs and losses (1e-2); the gradient
    # bounds are this repository's own, stated as measured at the BASELINE size.  The per-tensor bound is NOT a backward-precision
    # effect: ReLU / InstanceNorm make the gradient a discontinuous function of the forward values, and rounding ANY forward
    # operand to an 11-bit mantissa (fp16 here; TF32, the reference's own GPU default, has the same 10+1 bits) moves the
    # early layers' gradients by 5-7e-2 even with an exact backward (tools/exp_grad_rounding_points.py, profiles/r02_*).
    # The eager reference on the same GPU in its default TF32 mode shows the same gap (test_reference_tf32_gradient_gap).
    "tc": dict(loss=1e-2, grad=8e-2, cos=0.997, glob=5e-3, norm=2e-3),
    # fp32 path: 1e-4 on outputs/losses; gradients differ from the fp32 CPU oracle by isolated ReLU / max-pool mask
    # flips (the oracle itself sits that far from its float64 twin, printed by the step helper)
    "fp32": dict(loss=1e-4, grad=5e-3, cos=0.99995, glob=2e-3, norm=5e-4),
    # fp16x3 forward (fp32-class, tensor cores) + bf16 backward: what remains is the bf16 rounding of the backward GEMMs
    "tc_x3": dict(loss=1e-3, grad=2e-2, cos=0.9997, glob=3e-3, norm=1e-3),
}


def test_reference_tf32_gradient_gap():
    """Yardstick for the tensor-core gradient bound: the reference's own arithmetic (oracle port, stock torch ops) on this GPU in
    PyTorch's default mode (cuDNN TF32 convolutions) against the same code on the CPU in fp32 -- the gap any 10-bit-mantissa
    forward shows on this network at random init.  Recorded, and asserted only to be of the same class as ours."""
    p = O.make_net_params(seed=0, random_affine=True)
    vp = O.make_vgg_params(seed=1)
    content = O.make_image(4, 256, 256, seed=5, normalized=True)
    sty = O.make_image(1, 256, 256, seed=6, normalized=True)
    torch.manual_seed(99)
    ones = torch.ones((4, 256, 1, 1), device=DEV)
    drop = [torch.nn.functional.dropout2d(ones, 0.1, True).view(4, 256).cpu() for _ in range(5)]
    _, ref = O.loss_and_grads(p, vp, content, O.style_targets(vp, sty), drop)
    assert torch.backends.cudnn.allow_tf32                                   # PyTorch default
    pc, vc = {k: v.to(DEV) for k, v in p.items()}, {k: v.to(DEV) for k, v in vp.items()}
    _, got = O.loss_and_grads(pc, vc, content.to(DEV), O.style_targets(vc, sty.to(DEV)), [d.to(DEV) for d in drop])
    gn, gn_ref, rows, glob = _grad_report({k: v.cpu() for k, v in got.items()}, ref)
    worst = max(rows.items(), key=lambda kv: kv[1][0])
    print(f"reference, eager CUDA TF32 vs CPU fp32 at 4x256x256: worst tensor {worst[0]} rel_l2 {worst[1][0]:.3e}, global {glob:.3e}")
    _record("reference_tf32_gpu_vs_fp32_cpu", {"grad_worst_rel_l2": worst[1][0], "grad_worst_name": worst[0], "grad_global_rel_l2": glob})
    assert worst[1][0] > 5e-3          # a 10-bit-mantissa forward is visibly off the fp32 gradients on this network...
    assert worst[1][0] < 0.3           # ...in the same class as the fp16 tensor-core path (6.8e-2), not a bug-sized gap


@pytest.mark.parametrize("path", ["tc", "tc_x3", "fp32"])
def test_config1_training_step_4x256x256(dropin, path):
    precision, vgg_precision = {"tc": ("fp16", "bf16"), "tc_x3": ("fp16x3", "fp16"), "fp32": ("fp32", "fp32")}[path]
    got, grads, ref_losses, ref_grads = T._step(dropin, precision, vgg_precision, 4, 256, 256, seed=0)
    tol = TRAIN_TOL[path]
    out_err = T.rel_l2(got["stylized"], ref_losses["stylized"])
    rec = {"stylized_rel_l2": out_err}
    assert out_err < (1e-2 if path == "tc" else 1e-4)          # north_star: 1e-2 tensor-core path, 1e-4 fp32 class
    for k in ("content", "style", "tv", "total"):
        err = abs(got[k] / float(ref_losses[k]) - 1)
        rec["loss_" + k] = err
        print(f"[{path} 4x256x256] loss {k}: got {got[k]:.6g} ref {float(ref_losses[k]):.6g} rel {err:.2e}")
        assert err < tol["loss"], k
    gn, gn_ref, rows, glob = _grad_report(grads, ref_grads)
    worst = sorted(rows.items(), key=lambda kv: -kv[1][0])[:8]
    for k, (e, c, big) in worst:
        print(f"[{path} 4x256x256]   {k}: rel_l2 {e:.3e} cos {c:.6f}")
    rec.update(grad_norm=gn, grad_norm_ref=gn_ref, grad_global_rel_l2=glob,
               grad_worst_rel_l2=worst[0][1][0], grad_worst_name=worst[0][0],
               grad_min_cos=min(c for _, (e, c, big) in rows.items() if big))
    _record(f"config1_train_{path}", rec)
    print(f"[{path} 4x256x256] grad norm {gn:.6g} vs {gn_ref:.6g}; global rel_l2 {glob:.3e}; worst tensor {worst[0][1][0]:.3e}")
    for k, (e, c, big) in rows.items():
        assert e < tol["grad"], (k, e)
        if big:
            assert c > tol["cos"], (k, c)
    assert glob < tol["glob"]
    assert abs(gn / gn_ref - 1) < tol["norm"]


@pytest.mark.parametrize("precision,tol_l2,tol_px", [("fp16", 1e-2, 1.0), ("fp16x3", 1e-4, 0.05)])
def test_config2_forward_1080p(precision, tol_l2, tol_px):
    p = O.make_net_params(seed=0)
    x = O.make_image(1, 1080, 1920, seed=1234)
    with torch.no_grad():
        ref = O.stylenet_forward(p, x)
    y = engine.StyleNetPlan(precision).pack({k: v.to(DEV) for k, v in p.items()}).forward(x.to(DEV)).cpu()
    assert y.shape == ref.shape == (1, 3, 1080, 1920)
    err = T.rel_l2(y, ref)
    px = float((O.to_pixels(y) - O.to_pixels(ref)).abs().max())
    print(f"1x3x1080x1920 [{precision}] rel_l2 {err:.3e}  max |pixel diff| {px:.4f}")
    _record(f"config2_1080p_{precision}", {"rel_l2": err, "max_px": px})
    assert err < tol_l2 and px <= tol_px


def test_config3_batch_256_forward_matches_oracle_shards():
    """The 256-image batch in ONE forward call (what bench.py's infer256 workload times) against the oracle run in
    shards of 32 images; every image is checked on its own (a batch-level norm could hide one broken image)."""
    p = O.make_net_params(seed=0)
    x = O.make_image(256, 256, 256, seed=1234)
    y = engine.StyleNetPlan("fp16").pack({k: v.to(DEV) for k, v in p.items()}).forward(x.to(DEV)).cpu()
    worst, worst_px = 0.0, 0.0
    with torch.no_grad():
        for lo in range(0, 256, 32):
            ref = O.stylenet_forward(p, x[lo:lo + 32])
            for i in range(32):
                worst = max(worst, T.rel_l2(y[lo + i], ref[i]))
            worst_px = max(worst_px, float((O.to_pixels(y[lo:lo + 32]) - O.to_pixels(ref)).abs().max()))
    print(f"256x3x256x256 [fp16] worst per-image rel_l2 {worst:.3e}  max |pixel diff| {worst_px:.4f}")
    _record("config3_batch256_fp16", {"worst_image_rel_l2": worst, "max_px": worst_px})
    assert worst < 1e-2 and worst_px <= 1.0


@pytest.mark.parametrize("precision", ["fp32", "fp16x3"])
def test_config0_single_256_in_the_stated_fp32_class(precision):
    p = O.make_net_params(seed=0)
    x = O.make_image(1, 256, 256, seed=1234)
    with torch.no_grad():
        ref = O.stylenet_forward(p, x)
    y = engine.StyleNetPlan(precision).pack({k: v.to(DEV) for k, v in p.items()}).forward(x.to(DEV)).cpu()
    err = T.rel_l2(y, ref)
    px = float((O.to_pixels(y) - O.to_pixels(ref)).abs().max())
    print(f"1x3x256x256 [{precision}] rel_l2 {err:.3e}  max |pixel diff| {px:.5f}")
    _record(f"config0_256_{precision}", {"rel_l2": err, "max_px": px})
    assert err < 1e-4
