"""Vectorised host-side box transforms of the eval drivers: the reference adjusts every box of every image
with a Python call (``adjust_bbox_for_transform`` / ``..._no_scale``, src/bbox_utils.py:170-250; called per box at
eval_densecap.py:313-342).  Same double arithmetic, operation for operation, over whole [B,R,4] arrays."""
from __future__ import annotations

from typing import Sequence, Tuple

import numpy as np
import torch


def adjust_bboxes(sizes: Sequence[Tuple[int, int]], bboxes, resize_dim: int, crop_dim: int, keep_img_ratio: bool = True) -> torch.Tensor:
    """``sizes``: (width, height) of every original image; ``bboxes`` [B,R,4] xywh in original pixels -> float64 [B,R,4] in the
    coordinates of the preprocessed image (``preprocess(..., keep_img_ratio)``).  keep_img_ratio=True is
    adjust_bbox_for_transform(image, box, resize_dim, crop_dim); False is adjust_bbox_for_transform_no_scale(image, box,
    resize_dim, resize_dim)."""
    b = np.asarray(torch.as_tensor(bboxes).to(torch.float64).cpu().numpy(), dtype=np.float64)
    if b.ndim != 3 or b.shape[-1] != 4 or b.shape[0] != len(sizes):
        raise ValueError("bboxes must be [B,R,4] with one (width, height) per image")
    ow = np.array([s[0] for s in sizes], dtype=np.float64)[:, None]
    oh = np.array([s[1] for s in sizes], dtype=np.float64)[:, None]
    x1, y1, w, h = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    if not keep_img_ratio:
        sw, sh = resize_dim / ow, resize_dim / oh
        return torch.from_numpy(np.stack([x1 * sw, y1 * sh, w * sw, h * sh], axis=-1))
    portrait = ow < oh
    # the reference's exact expressions, evaluated for both branches and selected (double arithmetic, same operation order)
    sw = np.where(portrait, resize_dim / ow, (resize_dim * ow) / oh / ow)
    sh = np.where(portrait, (resize_dim * oh) / ow / oh, resize_dim / oh)
    new_w = np.trunc(ow * sw)  # int(): truncation of a positive double
    new_h = np.trunc(oh * sh)
    x1, y1, w, h = x1 * sw, y1 * sh, w * sw, h * sh
    x1 = x1 - np.maximum(0, np.floor_divide(new_w - crop_dim, 2))
    y1 = y1 - np.maximum(0, np.floor_divide(new_h - crop_dim, 2))
    x1 = np.maximum(0, np.minimum(x1, crop_dim - 1))
    y1 = np.maximum(0, np.minimum(y1, crop_dim - 1))
    w = np.maximum(0, np.minimum(w, crop_dim - x1))
    h = np.maximum(0, np.minimum(h, crop_dim - y1))
    return torch.from_numpy(np.stack([x1, y1, w, h], axis=-1))
