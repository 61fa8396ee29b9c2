"""The drop-in claim end to end: the reference's OWN scripts, unmodified, running on the drop-in modules on the GPU.

  * inference.test_inference (inference.py:27-61) on dancing.jpg with a checkpoint file: output image against the same script
    on the reference's own modules (run on the CPU in fp32);
  * train.train_style_transfer (train.py:68-302): DataLoader over a JPEG folder, style targets, 100 iterations of the loop,
    final state-dict + whole-module pickle, then a resume from a checkpoint dictionary (train.py:39-66);
  * the reference's module classes vs the drop-in's sub-module forwards (ConvLayer, UpsampleConv, ResidualBlock).

Needs the snapshot of the unmodified reference under baseline/_ref (tools/install_reference.sh; git-ignored, shipped to the GPU
box by gpurun).  PYTHONPATH = <drop-in>:<reference snapshot>: `models`, `losses` resolve to the drop-in, everything else
(train.py, inference.py, config.py, data/, utils/) to the reference."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import stylenet_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
DROPIN = os.path.join(ROOT, "fast_neural_style_transfer_b200", "dropin")
needs_ref = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "train.py")), reason="baseline/_ref missing (tools/install_reference.sh)")


def _run(code, pythonpath, cwd, extra_env=None, timeout=900):
    env = dict(os.environ, PYTHONPATH=os.pathsep.join(pythonpath), FNST_VGG19_RANDOM_INIT="1")
    env.update(extra_env or {})
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, cwd=cwd, timeout=timeout)
    assert out.returncode == 0, (out.stdout[-1500:], out.stderr[-3000:])
    return out.stdout


@needs_ref
def test_reference_inference_script_on_the_dropin(tmp_path):
    from PIL import Image
    p = O.make_net_params(seed=0)
    ckpt = tmp_path / "checkpoint.pth"
    torch.save({"model_state_dict": p}, ckpt)
    content = os.path.join(REF, "dancing.jpg")
    outs = {}
    for name, path, env in (("dropin", [DROPIN, REF], {}), ("reference_cpu", [REF], {"CUDA_VISIBLE_DEVICES": ""})):
        d = tmp_path / name
        d.mkdir()
        code = ("import inspect, inference, models.model\n"
                f"inference.test_inference({str(ckpt)!r}, {content!r}, {str(d)!r})\n"
                "print('MODEL_FILE', inspect.getfile(models.model))\n")
        log = _run(code, path, REF, env)
        assert ("fast_neural_style_transfer_b200" in log.split("MODEL_FILE")[1]) == (name == "dropin")
        outs[name] = np.asarray(Image.open(d / "noraml_output.jpg").convert("RGB")).astype(np.int32)      # file name as in inference.py:61
    diff = np.abs(outs["dropin"] - outs["reference_cpu"])
    print(f"inference.py on dancing.jpg: drop-in (GPU) vs reference modules (CPU fp32): mean |diff| {diff.mean():.4f}, max {diff.max()} of 255 (JPEG files)")
    assert outs["dropin"].shape == (256, 256, 3)
    assert diff.mean() < 0.25 and diff.max() <= 8            # two JPEG encodings of images that differ by < 1e-4 before quantisation


@needs_ref
def test_reference_training_script_on_the_dropin(tmp_path):
    """100 iterations of the reference's train_style_transfer on the drop-in (the reference's own VGG19 class cannot be
    constructed -- undefined `slice5`, weight download -- so there is no reference-module twin of this run)."""
    from PIL import Image
    rng = np.random.default_rng(0)
    data = tmp_path / "coco" / "train"
    data.mkdir(parents=True)
    for i in range(12):                                   # a small JPEG folder for data/dataset.py:7-18
        h, w = int(rng.integers(200, 400)), int(rng.integers(200, 400))
        base = rng.integers(0, 256, (h // 16 + 1, w // 16 + 1, 3), dtype=np.uint8)
        Image.fromarray(base).resize((w, h), Image.BILINEAR).save(data / f"img_{i:02d}.jpg", quality=90)
    out_dir = tmp_path / "out"
    style, monitor = os.path.join(REF, "picasso.jpg"), os.path.join(REF, "dancing.jpg")
    common = (f"style_image={style!r}, training_monitor_content_image={monitor!r}, dataset_dir={str(tmp_path / 'coco')!r}, "
              f"output_dir={str(out_dir)!r}, content_weight=1000.0, style_weight=1, tv_weight=10, num_epochs=1, batch_size=4, lr=1e-3")
    code = ("import torch, inspect\n"
            "torch.manual_seed(0)\n"
            "import train, models.model, models.vgg19_net, losses.losses\n"
            "assert all('fast_neural_style_transfer_b200' in inspect.getfile(m) for m in (models.model, models.vgg19_net, losses.losses))\n"
            "assert '_ref' in inspect.getfile(train)\n"
            f"train.train_style_transfer({common}, total_steps=100)\n"
            "print('RUN1_DONE')\n")
    log = _run(code, [DROPIN, REF], REF)
    assert "RUN1_DONE" in log and "Training completed!" in log and "Invalid loss" not in log
    line = [l for l in log.splitlines() if l.startswith("Iter [100/100]")][0]           # train.py:220-226
    total = float(line.split("Total:")[1].split("|")[0])
    print(line)
    assert np.isfinite(total) and total > 0
    final = torch.load(out_dir / "style_transfer_final.pth", map_location="cpu")          # train.py:296
    assert len(final) == 58 and all(torch.isfinite(v).all() for v in final.values())
    p0 = None
    # resume (train.py:39-66, :124-133) from a checkpoint dictionary in the reference's format, 10 more steps
    code2 = ("import torch\n"
             "torch.manual_seed(1)\n"
             "import train\n"
             "from models.model import StyleTransferNet\n"
             f"sd = torch.load({str(out_dir / 'style_transfer_final.pth')!r}, map_location='cpu')\n"
             "net = StyleTransferNet(); net.load_state_dict(sd)\n"
             "opt = torch.optim.Adam(net.parameters(), lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-5)\n"
             f"ck = {str(tmp_path / 'checkpoint_100.pth')!r}\n"
             "torch.save({'model_state_dict': sd, 'optimizer_state_dict': opt.state_dict(), 'iteration': 100}, ck)\n"
             f"train.train_style_transfer({common}, total_steps=110, checkpoint_path=ck)\n"
             "whole = torch.load(" + repr(str(out_dir / 'style_transfer.bin')) + ", weights_only=False)\n"                   # train.py:297 whole-module pickle
             "assert type(whole).__name__ == 'StyleTransferNet' and len(whole.state_dict()) == 58\n"
             "print('RUN2_DONE')\n")
    log2 = _run(code2, [DROPIN, REF], REF)
    assert "Resuming training from iteration 100" in log2 and "RUN2_DONE" in log2 and "Invalid loss" not in log2
    final2 = torch.load(out_dir / "style_transfer_final.pth", map_location="cpu")
    moved = sum(float((final2[k] - final[k]).abs().sum()) for k in final)
    assert moved > 0 and all(torch.isfinite(v).all() for v in final2.values())


@needs_ref
def test_submodule_forwards_match_the_reference_modules():
    """ConvLayer / UpsampleConv / ResidualBlock forward (models/model.py:21-22, :74-75, :86-90): the drop-in's unfused operator
    paths on the GPU against the reference's own module classes on the CPU with the same parameters."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_reference_model", os.path.join(REF, "models", "model.py"))
    ref_mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_mod)
    sys.path.insert(0, DROPIN)
    for m in [k for k in sys.modules if k.split(".")[0] in ("models", "losses", "config")]:
        del sys.modules[m]
    import models.model as mm
    g = torch.Generator().manual_seed(5)
    cases = [("ConvLayer", (3, 64, 9, 2), (2, 3, 37, 45)), ("ConvLayer", (64, 256, 3, 2), (1, 64, 20, 22)), ("ConvLayer", (32, 3, 9, 1), (1, 32, 24, 20)),
             ("UpsampleConv", (256, 64, 3, 2), (1, 256, 9, 11)), ("UpsampleConv", (64, 32, 3, 2), (2, 64, 12, 10)), ("ResidualBlock", (256,), (2, 256, 12, 14))]
    for cls, args, shape in cases:
        torch.manual_seed(11)
        ref = getattr(ref_mod, cls)(*args).eval()
        mine = getattr(mm, cls)(*args).eval()
        mine.load_state_dict(ref.state_dict())
        if cls == "ResidualBlock":                              # non-trivial affine parameters
            with torch.no_grad():
                for m in (ref, mine):
                    m.in1.weight.copy_(torch.linspace(0.5, 1.5, 256)); m.in2.bias.copy_(torch.linspace(-0.3, 0.3, 256))
        x = torch.randn(shape, generator=g)
        with torch.no_grad():
            want = ref(x)
            got = mine.cuda()(x.cuda()).cpu()
        err = float((got - want).norm() / want.norm())
        print(f"{cls}{args} on {shape}: rel_l2 {err:.2e}")
        assert got.shape == want.shape and err < 1e-5, (cls, args, err)
        with pytest.raises(RuntimeError, match="inference-time operator"):
            mine(x.cuda().requires_grad_(True))
    sys.path.remove(DROPIN)
