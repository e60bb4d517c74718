"""Device-side image preprocessing: the reference's torchvision pipeline (src/model.py:347-357)

    T.Resize(resize_dim, BICUBIC) -> T.CenterCrop(crop_dim) -> T.ToTensor() -> T.Normalize(IMAGENET mean, std)      keep_img_ratio
    T.Resize((resize_dim, resize_dim), BICUBIC)            -> T.ToTensor() -> T.Normalize(...)                       squash

run by ``pio_preprocess`` on uint8 HWC images.  This module is the host half: Pillow's resampling coefficient tables
(``precompute_coeffs`` + ``normalize_coeffs_8bpc`` of libImaging/Resample.c, restated with the same double arithmetic so the
22-bit fixed-point weights are identical), torchvision's output-size and centre-crop rules, batching by image size.
"""
from __future__ import annotations

import ctypes as C
import math
from functools import lru_cache
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L
from .ops import _stream, workspace

PRECISION_BITS = 32 - 8 - 2
IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def _bicubic(x: float) -> float:
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


@lru_cache(maxsize=256)
def resample_table(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray, int]:
    """Pillow's bicubic coefficients for one axis: (k int32 [out, ksize], bounds int32 [out, 2] = (first index, taps), ksize)."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    kk = np.zeros((out_size, ksize), dtype=np.float64)
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        ww = 0.0
        for x in range(xmax):
            w = _bicubic((x + xmin - center + 0.5) * ss)
            kk[xx, x] = w
            ww += w
        if ww != 0.0:
            kk[xx, :xmax] /= ww
        bounds[xx] = (xmin, xmax)
    fixed = np.trunc(np.where(kk < 0, -0.5 + kk * (1 << PRECISION_BITS), 0.5 + kk * (1 << PRECISION_BITS))).astype(np.int32)
    return fixed, bounds, ksize


def resized_size(w: int, h: int, resize_dim: int, keep_img_ratio: bool) -> Tuple[int, int]:
    """(new_w, new_h) of T.Resize(resize_dim) (short side, long side = int(size * long / short)) or T.Resize((s, s))."""
    if not keep_img_ratio:
        return resize_dim, resize_dim
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = resize_dim, int(resize_dim * long / short)
    return (new_short, new_long) if w <= h else (new_long, new_short)


_dev_tables: Dict[tuple, tuple] = {}


def _device_table(in_size: int, out_size: int, device) -> tuple:
    key = (in_size, out_size, torch.device(device).index or 0)
    t = _dev_tables.get(key)
    if t is None:
        k, b, ksize = resample_table(in_size, out_size)
        t = _dev_tables[key] = (torch.from_numpy(k).to(device), torch.from_numpy(b).to(device), ksize, b)
    return t


def preprocess_batch(imgs_u8: torch.Tensor, resize_dim: int, crop_dim: int, keep_img_ratio: bool = True,
                     mean: Sequence[float] = IMAGENET_MEAN, std: Sequence[float] = IMAGENET_STD) -> torch.Tensor:
    """uint8 images of ONE size [B,H,W,3] on the device -> normalised fp32 [B,3,S,S] (S = crop_dim, or resize_dim when squashing)."""
    if not imgs_u8.is_cuda or imgs_u8.dtype != torch.uint8 or imgs_u8.dim() != 4 or imgs_u8.shape[-1] != 3:
        raise L.PioError("preprocess_batch: uint8 CUDA tensor [B,H,W,3] expected (there is no CPU path)")
    imgs_u8 = imgs_u8.contiguous()
    B, H, W, _ = imgs_u8.shape
    new_w, new_h = resized_size(W, H, resize_dim, keep_img_ratio)
    if keep_img_ratio:
        if new_w < crop_dim or new_h < crop_dim:
            raise ValueError("CenterCrop larger than the resized image pads with zeros in torchvision: not supported")
        cw = ch = crop_dim
        top, left = int(round((new_h - ch) / 2.0)), int(round((new_w - cw) / 2.0))  # torchvision.transforms.functional.center_crop
    else:
        cw, ch, top, left = new_w, new_h, 0, 0
    dev = imgs_u8.device
    kx, bx, ksx, _ = _device_table(W, new_w, dev)
    ky, by, ksy, by_host = _device_table(H, new_h, dev)
    row_first = int(by_host[top, 0])
    row_last = int((by_host[top:top + ch, 0] + by_host[top:top + ch, 1]).max())
    rows = row_last - row_first
    out = torch.empty(B, 3, ch, cw, dtype=torch.float32, device=dev)
    nbytes = L.lib().pio_preprocess_workspace_bytes(B, rows, cw)
    ws = workspace(nbytes, dev, "preprocess")
    m3, s3 = (C.c_float * 3)(*mean), (C.c_float * 3)(*std)
    L.check(L.lib().pio_preprocess(imgs_u8.data_ptr(), B, H, W, kx.data_ptr(), bx.data_ptr(), ksx, ky.data_ptr(), by.data_ptr(), ksy,
                                   left, top, cw, ch, row_first, rows, m3, s3, out.data_ptr(), ws.data_ptr(), nbytes, _stream()))
    return out


def preprocess_images(images: Sequence, device, resize_dim: int, crop_dim: int, keep_img_ratio: bool = True) -> torch.Tensor:
    """PIL images (RGB) or uint8 [H,W,3] arrays / tensors of any sizes -> [B,3,S,S] on ``device``.  Images of equal size share one
    launch; the raw bytes are what travels to the device (a 640 x 480 JPEG frame is 0.9 MB, its 518 x 518 fp32 crop 3.2 MB)."""
    arrs: List[torch.Tensor] = []
    for im in images:
        if torch.is_tensor(im):
            t = im
        elif isinstance(im, np.ndarray):
            t = torch.from_numpy(im)
        else:
            t = torch.from_numpy(np.asarray(im.convert("RGB") if getattr(im, "mode", "RGB") != "RGB" else im).copy())
        if t.dtype != torch.uint8 or t.dim() != 3 or t.shape[-1] != 3:
            raise ValueError("images must be RGB uint8 [H,W,3]")
        arrs.append(t)
    groups: Dict[tuple, List[int]] = {}
    for i, t in enumerate(arrs):
        groups.setdefault((t.shape[0], t.shape[1]), []).append(i)
    outs: List[torch.Tensor] = [None] * len(arrs)  # type: ignore
    for idxs in groups.values():
        batch = torch.stack([arrs[i] for i in idxs]).to(device, non_blocking=True)
        res = preprocess_batch(batch, resize_dim, crop_dim, keep_img_ratio)
        for j, i in enumerate(idxs):
            outs[i] = res[j]
    return torch.stack(outs)
