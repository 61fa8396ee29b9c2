"""Pins of the pre-processing oracle (oracle/pil_resize.c) and of the product's host twin of the device arithmetic:

* the oracle against the third-party libraries the reference's transform really calls -- Pillow's Image.resize(BILINEAR)
  (what torchvision's transforms.Resize does to a PIL image) and torchvision's ToTensor / Normalize -- bit for bit;
* `fnst_resize_window_host` (the SAME __host__ __device__ functions the CUDA kernel runs, compiled for the host) against the
  oracle's coefficient windows, and the shared-memory strip bound the launch relies on;
* the Python layer of fast_neural_style_transfer_b200.preprocess with the C-ABI call replaced by the oracle on host memory.
CPU only."""
import ctypes as C
import math

import numpy as np
import pytest
import torch
from PIL import Image

from oracle import pil_resize as R

MEAN, STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
SIZES = [(444, 444), (650, 650), (609, 800), (256, 256), (256, 300), (300, 256), (100, 120), (17, 23), (1080, 1920), (255, 257),
         (1, 1), (2, 3), (1000, 37), (2000, 3000)]


@pytest.mark.parametrize("hw", SIZES)
def test_resize_matches_pillow_bit_for_bit(hw):
    rng = np.random.default_rng(hw[0] * 7919 + hw[1])
    img = rng.integers(0, 256, (hw[0], hw[1], 3), dtype=np.uint8)
    for oh, ow in ((256, 256), (224, 320), (31, 500)):
        ref = np.asarray(Image.fromarray(img).resize((ow, oh), Image.BILINEAR))
        assert np.array_equal(R.resize_bilinear_u8(img, oh, ow), ref), (hw, oh, ow)


def test_smooth_images_and_strided_rows():
    yy, xx = np.mgrid[0:480, 0:640]
    img = np.stack([(yy * 255 // 479), (xx * 255 // 639), ((yy + xx) % 256)], -1).astype(np.uint8)
    ref = np.asarray(Image.fromarray(img).resize((256, 256), Image.BILINEAR))
    assert np.array_equal(R.resize_bilinear_u8(img, 256, 256), ref)
    crop = img[:, 100:400]                                            # row-strided view
    ref = np.asarray(Image.fromarray(np.ascontiguousarray(crop)).resize((256, 256), Image.BILINEAR))
    assert np.array_equal(R.resize_bilinear_u8(crop, 256, 256), ref)


def test_whole_transform_matches_torchvision():
    """transforms.Compose([Resize((256,256)), ToTensor(), Normalize]) on a PIL image (train.py:92-102) and the un-normalised
    variant of inference.py:28-31."""
    from torchvision import transforms
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (444, 517, 3), dtype=np.uint8)
    pil = Image.fromarray(img)
    t_train = transforms.Compose([transforms.Resize((256, 256)), transforms.ToTensor(), transforms.Normalize(mean=list(MEAN), std=list(STD))])
    t_inf = transforms.Compose([transforms.Resize((256, 256)), transforms.ToTensor()])
    small = R.resize_bilinear_u8(img, 256, 256)
    assert torch.equal(torch.from_numpy(R.to_tensor(small, MEAN, STD)), t_train(pil))
    assert torch.equal(torch.from_numpy(R.to_tensor(small)), t_inf(pil))


def test_host_twin_of_the_device_windows_and_strip_bound():
    from fast_neural_style_transfer_b200 import _lib
    lib = _lib.lib
    first, ln, kk = C.c_int(), C.c_int(), (C.c_int * 72)()
    pairs = [(s, o) for o in (256, 7, 300) for s in list(range(1, 40)) + [100, 255, 257, 444, 609, 800, 1080, 1920, 3000, 4000, 8960]]
    checked = 0
    for s, o in pairs:
        if math.ceil(max(s / o, 1.0)) * 2 + 1 > 72 or s == o:
            continue
        ks, ref = R.windows(s, o)
        starts, ends = [], []
        for i, (f, l, w) in enumerate(ref):
            assert lib.fnst_resize_window_host(s, o, i, C.byref(first), C.byref(ln), kk, 72) == ks
            assert (first.value, ln.value, list(kk[:ln.value])) == (f, l, w), (s, o, i)
            starts.append(f); ends.append(f + l)
        assert all(a <= b for a, b in zip(starts, starts[1:]))                          # the kernel takes lane 0's start as the strip start
        bound = min(s, math.ceil(15 * (s / o) + 2 * max(s / o, 1.0) + 3.0))              # csrc/resize.cu strip_rows_bound (it adds 1 more)
        assert all(max(ends[y:y + 16]) - starts[y] <= bound for y in range(0, o, 16)), (s, o)
        checked += 1
    assert checked > 100
    assert lib.fnst_resize_window_host(256 * 40, 256, 0, C.byref(first), C.byref(ln), kk, 72) < 0      # factor 40 > 35: refused


def test_preprocess_python_layer_with_emulated_abi(monkeypatch):
    """fast_neural_style_transfer_b200.preprocess on CPU tensors, the C-ABI entry replaced by the oracle working on the
    same host pointers: checks slot offsets, pitch, argument order, mixed image sizes, `out=`; and that without the
    emulation CPU tensors are refused (no fallback)."""
    from fast_neural_style_transfer_b200 import preprocess as P
    with pytest.raises(RuntimeError, match="CUDA"):
        P.resize_to_tensor(torch.zeros((8, 8, 3), dtype=torch.uint8))

    def emu(img, ih, iw, pitch, oh, ow, out_f, out_u8, mean, std, dev, stream):
        src = np.ctypeslib.as_array((C.c_uint8 * (ih * pitch)).from_address(img.value)).reshape(ih, pitch)[:, :iw * 3].reshape(ih, iw, 3)
        small = R.resize_bilinear_u8(src, oh, ow)
        if out_u8 is not None and getattr(out_u8, "value", None):
            np.ctypeslib.as_array((C.c_uint8 * (oh * ow * 3)).from_address(out_u8.value))[:] = small.reshape(-1)
        if out_f is not None and getattr(out_f, "value", None):
            m = None if mean is None else list(mean)
            s = None if std is None else list(std)
            np.ctypeslib.as_array((C.c_float * (3 * oh * ow)).from_address(out_f.value))[:] = R.to_tensor(small, m, s).reshape(-1)
        return 0

    def emu_batch(table, n, max_h, max_w, oh, ow, out_f, out_u8, mean, std, dev, stream):
        # fnst_image_desc[n] = {void* data; int32 h, w; int64 pitch_bytes}: the batch entry point walks the descriptor table
        desc = np.ctypeslib.as_array((C.c_int64 * (3 * n)).from_address(table.value)).reshape(n, 3)
        for i in range(n):
            ih, iw = int(desc[i, 1] & 0xFFFFFFFF), int(desc[i, 1] >> 32)
            assert ih <= max_h and iw <= max_w
            slot = lambda base, nbytes: None if base is None or not getattr(base, "value", None) else C.c_void_p(base.value + i * nbytes)
            emu(C.c_void_p(int(desc[i, 0])), ih, iw, int(desc[i, 2]), oh, ow, slot(out_f, 3 * oh * ow * 4), slot(out_u8, 3 * oh * ow), mean, std, dev, stream)
        return 0

    class Lib:
        fnst_resize_to_tensor = staticmethod(emu)
        fnst_resize_batch_to_tensor = staticmethod(emu_batch)
        fnst_last_error = staticmethod(lambda: b"emulated")

    monkeypatch.setattr(P, "lib", Lib)
    monkeypatch.setattr(P.ops, "_ctx", lambda t: (0, None))
    from torchvision import transforms
    t_train = transforms.Compose([transforms.Resize((256, 256)), transforms.ToTensor(), transforms.Normalize(mean=list(MEAN), std=list(STD))])
    rng = np.random.default_rng(9)
    imgs = [rng.integers(0, 256, s + (3,), dtype=np.uint8) for s in ((300, 400), (256, 256), (77, 91))]
    batch = P.Transform().batch([torch.from_numpy(i) for i in imgs])
    assert batch.shape == (3, 3, 256, 256)
    for got, img in zip(batch, imgs):
        assert torch.equal(got, t_train(Image.fromarray(img)))
    crop = torch.from_numpy(imgs[0])[:, 50:350]                                          # row-strided view: pitch != 3 * width
    ref = transforms.Compose([transforms.Resize((256, 256)), transforms.ToTensor()])(Image.fromarray(np.ascontiguousarray(crop.numpy())))
    assert torch.equal(P.Transform(normalize=False)(crop), ref)
    out = torch.empty((1, 3, 64, 48))
    assert P.resize_to_tensor(torch.from_numpy(imgs[2]), (64, 48), out=out) is out
    u8 = P.resize_u8(torch.from_numpy(imgs[0]), (100, 120))
    assert np.array_equal(u8.numpy(), np.asarray(Image.fromarray(imgs[0]).resize((120, 100), Image.BILINEAR)))
    with pytest.raises(RuntimeError):
        P.resize_to_tensor(torch.zeros((8, 8, 4), dtype=torch.uint8))
    with pytest.raises(ValueError):
        P.resize_to_tensor(torch.zeros((8, 8, 3), dtype=torch.uint8), mean=MEAN)


def test_oracle_matches_committed_pillow_digests(golden_dir):
    """Same pin without needing Pillow at test time: digests generated by tests/golden/make_resize_golden.py."""
    import hashlib
    import json
    import os
    import sys
    sys.path.insert(0, golden_dir)
    from make_resize_golden import image
    with open(os.path.join(golden_dir, "resize_pillow.json")) as f:
        gold = json.load(f)
    assert len(gold["cases"]) >= 8
    for case in gold["cases"]:
        (h, w), (oh, ow) = case["in"], case["out"]
        small = R.resize_bilinear_u8(image(h, w), oh, ow)
        assert hashlib.sha256(small.tobytes()).hexdigest() == case["u8_sha256"], case
        assert hashlib.sha256(R.to_tensor(small, MEAN, STD).tobytes()).hexdigest() == case["tensor_sha256"], case
