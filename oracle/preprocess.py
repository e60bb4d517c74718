"""Oracle for the image preprocessing of src/model.py:347-357 (test infrastructure only, like the rest of oracle/).

    image_transforms         = Resize(resize_dim, BICUBIC) -> CenterCrop(crop_dim) -> ToTensor -> Normalize     (:347-352)
    image_transforms_no_crop = Resize((resize_dim, resize_dim), BICUBIC) -> ToTensor -> Normalize                (:353-357)

The arithmetic lives in third-party code: torchvision.transforms (0.26 here; the reference's environment.yml pins its own) and
Pillow's libImaging/Resample.c (ImagingResampleHorizontal_8bpc / Vertical_8bpc, precompute_coeffs, normalize_coeffs_8bpc).
``pil_bicubic_resize`` restates that algorithm in numpy; tests pin it bit-for-bit against the Pillow installed in the image, and
``reference_transform`` simply runs the real torchvision pipeline, which is what the CUDA path is compared with.
"""
from __future__ import annotations

import math

import numpy as np
import torch

PRECISION_BITS = 32 - 8 - 2


def _bicubic(x: float) -> float:  # Resample.c: bicubic_filter, a = -0.5
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def _coeffs(in_size: int, out_size: int):
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    kk = np.zeros((out_size, ksize))
    bounds = np.zeros((out_size, 2), dtype=np.int64)
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = np.array([_bicubic((x + xmin - center + 0.5) / filterscale) for x in range(xmax)])
        ww = 0.0
        for v in w:  # same summation order as the C loop
            ww += v
        kk[xx, :xmax] = w / ww if ww != 0.0 else w
        bounds[xx] = (xmin, xmax)
    fixed = np.trunc(np.where(kk < 0, -0.5 + kk * (1 << PRECISION_BITS), 0.5 + kk * (1 << PRECISION_BITS))).astype(np.int64)
    return fixed, bounds


def _clip8(v: np.ndarray) -> np.ndarray:
    return np.clip(v >> PRECISION_BITS, 0, 255).astype(np.uint8)


def pil_bicubic_resize(img: np.ndarray, out_w: int, out_h: int) -> np.ndarray:
    """uint8 [H,W,3] -> [out_h,out_w,3], bit-identical to PIL.Image.resize((out_w, out_h), BICUBIC)."""
    a = img.astype(np.int64)
    H, W, _ = a.shape
    if out_w != W:
        kx, bx = _coeffs(W, out_w)
        tmp = np.empty((H, out_w, 3), dtype=np.uint8)
        for xx in range(out_w):
            x0, n = bx[xx]
            tmp[:, xx] = _clip8((1 << (PRECISION_BITS - 1)) + np.tensordot(a[:, x0:x0 + n], kx[xx, :n], axes=([1], [0])))
        a = tmp.astype(np.int64)
    if out_h != H:
        ky, by = _coeffs(H, out_h)
        out = np.empty((out_h, a.shape[1], 3), dtype=np.uint8)
        for yy in range(out_h):
            y0, n = by[yy]
            out[yy] = _clip8((1 << (PRECISION_BITS - 1)) + np.tensordot(ky[yy, :n], a[y0:y0 + n], axes=([0], [0])))
        return out
    return a.astype(np.uint8)


def reference_transform(pil_images, resize_dim: int, crop_dim: int, keep_img_ratio: bool) -> torch.Tensor:
    """The reference's own pipeline (src/model.py:347-357), run with the real torchvision / Pillow."""
    import torchvision.transforms as T

    norm = T.Normalize(mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225))
    if keep_img_ratio:
        tf = T.Compose([T.Resize(resize_dim, interpolation=T.InterpolationMode.BICUBIC), T.CenterCrop(crop_dim), T.ToTensor(), norm])
    else:
        tf = T.Compose([T.Resize((resize_dim, resize_dim), interpolation=T.InterpolationMode.BICUBIC), T.ToTensor(), norm])
    return torch.stack([tf(im) for im in pil_images])
