"""GPU parity of the input pre-processing (device side): fast_neural_style_transfer_b200.preprocess against
what the reference's pipeline computes per image -- torchvision's Compose([Resize((256,256)), ToTensor(), Normalize]) on a
PIL image (train.py:92-102; inference.py:28-31 without Normalize).  Bit-exact: integer resampling, IEEE float32 afterwards."""
import os
import subprocess

import numpy as np
import pytest
import torch
from PIL import Image
from torchvision import transforms

from fast_neural_style_transfer_b200 import preprocess as P

pytestmark = pytest.mark.gpu
DEV = "cuda"
MEAN, STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
T_TRAIN = transforms.Compose([transforms.Resize((256, 256)), transforms.ToTensor(), transforms.Normalize(mean=MEAN, std=STD)])
T_INFER = transforms.Compose([transforms.Resize((256, 256)), transforms.ToTensor()])


def _img(h, w, seed):
    return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)


@pytest.mark.parametrize("hw", [(444, 444), (609, 800), (1080, 1920), (256, 256), (256, 300), (100, 120), (17, 23), (2000, 3000)])
def test_transform_is_bit_identical_to_the_reference_pipeline(hw):
    img = _img(hw[0], hw[1], hw[0] + hw[1])
    pil = Image.fromarray(img)
    dev_img = torch.from_numpy(img).to(DEV)
    assert torch.equal(P.Transform()(dev_img).cpu(), T_TRAIN(pil))
    assert torch.equal(P.Transform(normalize=False)(dev_img).cpu(), T_INFER(pil))
    assert np.array_equal(P.resize_u8(dev_img, (256, 256)).cpu().numpy(), np.asarray(pil.resize((256, 256), Image.BILINEAR)))


def test_batch_of_mixed_sizes_strided_rows_and_other_output_sizes():
    imgs = [_img(300, 400, 1), _img(256, 256, 2), _img(77, 91, 3), _img(720, 1280, 4)]
    batch = P.Transform().batch([torch.from_numpy(i).to(DEV) for i in imgs])
    assert batch.shape == (4, 3, 256, 256) and batch.is_cuda
    for got, img in zip(batch.cpu(), imgs):
        assert torch.equal(got, T_TRAIN(Image.fromarray(img)))
    crop = torch.from_numpy(imgs[3]).to(DEV)[:, 200:900]                 # row-strided view (pitch != 3 * width)
    ref = T_INFER(Image.fromarray(np.ascontiguousarray(imgs[3][:, 200:900])))
    assert torch.equal(P.Transform(normalize=False)(crop).cpu(), ref)
    up = P.resize_u8(torch.from_numpy(imgs[2]).to(DEV), (300, 500))       # up-scaling, non-square
    assert np.array_equal(up.cpu().numpy(), np.asarray(Image.fromarray(imgs[2]).resize((500, 300), Image.BILINEAR)))


def test_batch_launch_equals_per_image_launches(monkeypatch):
    """One launch for the whole batch (fnst_resize_batch_to_tensor: grid z = image, descriptor table on the device) against one
    launch per image -- mixed sizes, a row-strided crop, up- and down-scaling in one batch: bit-identical."""
    imgs = [torch.from_numpy(_img(h, w, s)).to(DEV) for s, (h, w) in enumerate([(1080, 1920), (300, 401), (64, 48), (256, 256), (900, 333)])]
    imgs.append(imgs[0][100:700, 200:1500])                               # pitch != 3 * width
    for mean, std in ((MEAN, STD), (None, None)):
        batched = P.resize_to_tensor(imgs, (256, 256), mean, std)
        monkeypatch.setattr(P, "BATCH_LAUNCH_MIN", 10 ** 9)
        single = P.resize_to_tensor(imgs, (256, 256), mean, std)
        monkeypatch.undo()
        assert torch.equal(batched, single)
    other = P.resize_to_tensor(imgs[:3], (120, 200))
    for got, img in zip(other, imgs[:3]):
        assert torch.equal(got, P.resize_to_tensor(img, (120, 200))[0])


def test_errors_are_loud():
    with pytest.raises(RuntimeError):
        P.resize_to_tensor(torch.zeros((64, 64, 3), dtype=torch.uint8))                          # CPU tensor: no fallback
    with pytest.raises(RuntimeError, match="down-scaling"):
        P.resize_u8(torch.zeros((400, 40, 3), dtype=torch.uint8, device=DEV), (8, 8))           # factor 50 > 35
    with pytest.raises(RuntimeError):
        P.resize_to_tensor(torch.zeros((64, 64, 4), dtype=torch.uint8, device=DEV))


def test_native_selftest_binary():
    """The stand-alone C++ check (tests/native/resize_selftest.cu): kernel vs the C oracle through the C ABI, no Python."""
    exe = os.path.join(os.path.dirname(os.path.abspath(__file__)), "native", "bin", "resize_selftest")
    if not os.path.exists(exe):
        pytest.skip("tests/native/bin/resize_selftest not built (python -c 'import __graft_entry__ as g; g.build()')")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "ALL OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
