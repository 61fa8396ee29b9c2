"""Device side of the reference's input pipeline: the per-image transform

    transforms.Compose([transforms.Resize((256, 256)), transforms.ToTensor(), normalize])      # train.py:92-102
    transforms.Compose([transforms.Resize((256, 256)), transforms.ToTensor()])                  # inference.py:28-31

applied by data/dataset.py:21-27 to every decoded PIL image, as ONE libfnst launch per BATCH of decoded uint8 images that
are already on the GPU (grid z = image: a single 1080p image is 256 small blocks, latency-bound at 0.45 TB/s) (bit-identical to Pillow's bilinear resize + torchvision's ToTensor / Normalize; tests compare with
both).  Images of different sizes go straight into their slot of one (N, 3, H, W) batch tensor.  JPEG decoding itself is
not part of this module (the reference decodes with PIL on DataLoader workers).  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple, Union

import torch

from . import ops
from ._lib import check, lib

IMAGENET_MEAN = (0.485, 0.456, 0.406)        # train.py:93-96
IMAGENET_STD = (0.229, 0.224, 0.225)


BATCH_LAUNCH_MIN = 2                          # image count from which the batch goes out as ONE launch (fnst_resize_batch_to_tensor)


def _f3(v) -> "C.Array":
    if len(v) != 3:
        raise ValueError("mean / std need three values (RGB)")
    return (C.c_float * 3)(*[float(x) for x in v])


def resize_to_tensor(images: Union[torch.Tensor, Sequence[torch.Tensor]], size: Tuple[int, int] = (256, 256),
                     mean: Optional[Sequence[float]] = None, std: Optional[Sequence[float]] = None,
                     out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """images: one decoded RGB image (H, W, 3) uint8 on the GPU, or a sequence of them (sizes may differ).
    Returns float32 (N, 3, size[0], size[1]) = Normalize(mean, std)(ToTensor(Resize(size)(image))) per image
    (no Normalize when mean/std are None).  `out`: optional preallocated result."""
    single = isinstance(images, torch.Tensor)
    imgs = [images] if single else list(images)
    if not imgs:
        raise ValueError("resize_to_tensor: no images")
    if (mean is None) != (std is None):
        raise ValueError("resize_to_tensor: pass both mean and std, or neither")
    oh, ow = int(size[0]), int(size[1])
    dev_t = imgs[0]
    dev, stream = ops._ctx(dev_t)                       # raises for CPU tensors
    if out is None:
        out = torch.empty((len(imgs), 3, oh, ow), dtype=torch.float32, device=dev_t.device)
    elif out.shape != (len(imgs), 3, oh, ow) or out.dtype != torch.float32 or not out.is_contiguous() or out.device != dev_t.device:
        raise RuntimeError("resize_to_tensor: `out` must be a contiguous float32 (N, 3, H, W) tensor on the images' device")
    m3, s3 = (None, None) if mean is None else (_f3(mean), _f3(std))
    plane_bytes = 3 * oh * ow * 4
    packed = []
    for img in imgs:
        if img.device != dev_t.device or img.dtype != torch.uint8 or img.dim() != 3 or img.shape[2] != 3:
            raise RuntimeError("resize_to_tensor: every image must be a uint8 (H, W, 3) tensor on one CUDA device")
        if img.stride(2) != 1 or img.stride(1) != 3:
            img = img.contiguous()                      # rows may be strided (crops); pixels must be packed RGB
        packed.append(img)
    if len(packed) >= BATCH_LAUNCH_MIN:
        # one launch for the whole batch: the grid's z index walks a device table of (pointer, height, width, pitch)
        table = torch.tensor([[img.data_ptr(), img.shape[0] | (img.shape[1] << 32), img.stride(0)] for img in packed], dtype=torch.int64)
        if dev_t.is_cuda:
            table = table.pin_memory().to(dev_t.device, non_blocking=True)
        check(lib.fnst_resize_batch_to_tensor(C.c_void_p(table.data_ptr()), len(packed), max(i.shape[0] for i in packed),
                                              max(i.shape[1] for i in packed), oh, ow, C.c_void_p(out.data_ptr()), None, m3, s3,
                                              dev, stream), "resize_batch_to_tensor")
        ops._count()
        if dev_t.is_cuda:
            table.record_stream(torch.cuda.current_stream(dev_t.device))
        return out
    for i, img in enumerate(packed):
        check(lib.fnst_resize_to_tensor(C.c_void_p(img.data_ptr()), img.shape[0], img.shape[1], img.stride(0), oh, ow,
                                        C.c_void_p(out.data_ptr() + i * plane_bytes), None, m3, s3, dev, stream), "resize_to_tensor")
        ops._count()
    return out


def resize_u8(image: torch.Tensor, size: Tuple[int, int]) -> torch.Tensor:
    """Pillow's `Image.resize(size[::-1], Image.BILINEAR)` on a uint8 (H, W, 3) CUDA tensor -> uint8 (size[0], size[1], 3)."""
    dev, stream = ops._ctx(image)
    if image.dtype != torch.uint8 or image.dim() != 3 or image.shape[2] != 3:
        raise RuntimeError("resize_u8: uint8 (H, W, 3) tensor required")
    if image.stride(2) != 1 or image.stride(1) != 3:
        image = image.contiguous()
    out = torch.empty((int(size[0]), int(size[1]), 3), dtype=torch.uint8, device=image.device)
    check(lib.fnst_resize_to_tensor(C.c_void_p(image.data_ptr()), image.shape[0], image.shape[1], image.stride(0), out.shape[0],
                                    out.shape[1], None, C.c_void_p(out.data_ptr()), None, None, dev, stream), "resize_to_tensor")
    ops._count()
    return out


class Transform:
    """Callable with the meaning of the reference's `transform` object (train.py:98-102): decoded uint8 (H, W, 3) CUDA image ->
    float32 (3, 256, 256) tensor.  `normalize=False` gives inference.py:28-31's variant."""

    def __init__(self, size: Tuple[int, int] = (256, 256), normalize: bool = True, mean: Sequence[float] = IMAGENET_MEAN,
                 std: Sequence[float] = IMAGENET_STD):
        self.size, self.normalize, self.mean, self.std = tuple(size), normalize, tuple(mean), tuple(std)

    def __call__(self, image: torch.Tensor) -> torch.Tensor:
        return resize_to_tensor(image, self.size, self.mean if self.normalize else None, self.std if self.normalize else None)[0]

    def batch(self, images: Sequence[torch.Tensor]) -> torch.Tensor:
        """What the DataLoader's default collate builds from per-image transforms (train.py:105-107): (N, 3, H, W)."""
        return resize_to_tensor(images, self.size, self.mean if self.normalize else None, self.std if self.normalize else None)
