"""Generates tests/golden/resize_pillow.json: SHA-256 digests of what Pillow's Image.resize(BILINEAR) and torchvision's
Compose([Resize((256,256)), ToTensor(), Normalize]) -- the third-party code behind the reference's transform (train.py:92-102) --
produce for seeded random images.  Run in the build container: python tests/golden/make_resize_golden.py
(Pillow / torchvision versions are recorded in the file.)"""
import hashlib
import json
import os

import numpy as np

CASES = [(444, 444, 256, 256), (609, 800, 256, 256), (1080, 1920, 256, 256), (256, 300, 256, 256), (100, 120, 256, 256),
         (17, 23, 256, 256), (500, 333, 224, 320), (64, 64, 300, 500)]
MEAN, STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]


def image(h, w):
    return np.random.default_rng(h * 100003 + w).integers(0, 256, (h, w, 3), dtype=np.uint8)


def main():
    import PIL
    import torch
    import torchvision
    from PIL import Image
    from torchvision import transforms
    out = {"pillow": PIL.__version__, "torchvision": torchvision.__version__, "torch": torch.__version__, "cases": []}
    for h, w, oh, ow in CASES:
        img = image(h, w)
        pil = Image.fromarray(img)
        small = np.asarray(pil.resize((ow, oh), Image.BILINEAR))
        t = transforms.Compose([transforms.Resize((oh, ow)), transforms.ToTensor(), transforms.Normalize(mean=MEAN, std=STD)])(pil)
        out["cases"].append({"in": [h, w], "out": [oh, ow], "u8_sha256": hashlib.sha256(small.tobytes()).hexdigest(),
                             "tensor_sha256": hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()})
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "resize_pillow.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
