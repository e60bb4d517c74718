"""Oracle (test infrastructure): region aggregation of patch tokens.

Restates, on CPU, what the reference does in
  * ``extract_bboxes_feats``      Patch-ioner/src/bbox_utils.py:8-109
  * ``map_traces_to_grid``        Patch-ioner/src/bbox_utils.py:158-168
  * trace pooling                 Patch-ioner/src/model.py:1049-1054
  * ``compute_region_means``      Patch-ioner/src/model.py:45-94
  * ``process_self_attention``    Patch-ioner/src/dino_extraction.py:24-34
  * ``avg_self_attn_token``       Patch-ioner/src/model.py:869

The integer part (slice bounds, trace bins) is written in plain Python so that
it can be compared bit-for-bit; the floating-point part uses torch fp32 on CPU.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import random

import torch

# ----------------------------------------------------------------------------
# integer part: box -> inclusive patch slice
# ----------------------------------------------------------------------------


def floor_div_f32(a: float, b: float) -> float:
    """``tensor //= b`` for float32 tensors (bbox_utils.py:19).

    torch's float floor_divide is *not* floor(a/b): it is
    ``(a - fmod(a, b)) / b`` followed by a sign fix and a round-to-nearest-integer
    guard (c10::div_floor_floating).  Restated here in float32 steps.
    """
    f32 = lambda v: torch.tensor(v, dtype=torch.float32).item()  # noqa: E731
    a = f32(a)
    b = f32(b)
    if b == 0:
        return f32(a / b) if a != 0 else float("nan")
    mod = f32(math.fmod(a, b))
    div = f32(f32(a - mod) / b)
    if mod != 0 and ((b < 0) != (mod < 0)):
        div = f32(div - 1.0)
    if div != 0:
        fl = float(math.floor(div))
        if f32(div - fl) > 0.5:
            fl = f32(fl + 1.0)
        return fl
    return math.copysign(0.0, f32(a / b))


def boxes_to_patch_units(bboxes: torch.Tensor, patch_size: int) -> torch.Tensor:
    """``bboxes //= patch_size; bboxes = bboxes.int()`` (bbox_utils.py:19-20).

    Float boxes use torch's float floor-division, integer boxes Python floor
    division; ``.int()`` then truncates toward zero.  Returns int32 [B,R,4]
    (x1, y1, w, h) in patch units.  Does not mutate its argument.
    """
    flat = bboxes.reshape(-1).tolist()
    out = []
    if bboxes.dtype.is_floating_point:
        for v in flat:
            q = floor_div_f32(v, float(patch_size))
            out.append(int(q) if math.isfinite(q) else 0)
    else:
        for v in flat:
            out.append(int(v) // int(patch_size))
    return torch.tensor(out, dtype=torch.int32).reshape(bboxes.shape)


def py_slice(start: int, stop: int, size: int) -> Tuple[int, int]:
    """Python ``a[start:stop]`` on a dimension of ``size`` -> [lo, hi) clamped.

    Negative indices wrap once (``+size``) and are then clamped, exactly as
    ``slice.indices`` does (bbox_utils.py:44 relies on it).
    """
    lo, hi, _ = slice(start, stop).indices(size)
    if hi < lo:
        hi = lo
    return lo, hi


def box_bounds(box_pu: Sequence[int], grid: int) -> Tuple[int, int, int, int]:
    """(y_lo, y_hi, x_lo, x_hi), hi exclusive, for one box in patch units.

    bbox_utils.py:30-34,44: ``x2 = x1 + w``, ``y2 = y1 + h`` and the region is
    ``[y1:y2+1, x1:x2+1]`` -- the end index is floor(x/ps)+floor(w/ps),
    inclusive.
    """
    x1, y1, w, h = (int(v) for v in box_pu)
    y_lo, y_hi = py_slice(y1, y1 + h + 1, grid)
    x_lo, x_hi = py_slice(x1, x1 + w + 1, grid)
    return y_lo, y_hi, x_lo, x_hi


def all_box_bounds(bboxes: torch.Tensor, patch_size: int, grid: int) -> torch.Tensor:
    """int32 [B,R,4] (y_lo,y_hi,x_lo,x_hi) -- the 'pooling indices' that must be bit-exact."""
    pu = boxes_to_patch_units(bboxes, patch_size)
    B, R, _ = pu.shape
    out = torch.zeros(B, R, 4, dtype=torch.int32)
    for i in range(B):
        for j in range(R):
            out[i, j] = torch.tensor(box_bounds(pu[i, j].tolist(), grid), dtype=torch.int32)
    return out


# ----------------------------------------------------------------------------
# floating-point part: weights and pooled embeddings
# ----------------------------------------------------------------------------


def gaussian_weights(h_span: int, w_span: int, variance: float) -> torch.Tensor:
    """bbox_utils.py:54-75 (variance != 0): exp(-(x^2+y^2)/var) on linspace(-1,1,span), sum 1."""
    y, x = torch.meshgrid(torch.linspace(-1, 1, h_span), torch.linspace(-1, 1, w_span), indexing="ij")
    w = torch.exp(-(x ** 2 + y ** 2) / variance)
    return w / w.sum()


def extract_bboxes_feats(
    patch_embeddings: torch.Tensor,
    bboxes: torch.Tensor,
    gaussian_avg: bool = False,
    gaussian_bbox_variance: float = 0.5,
    get_single_embedding_per_image: bool = False,
    patch_size: int = 14,
    attention_map: Optional[torch.Tensor] = None,
    return_debug: bool = False,
):
    """Restatement of bbox_utils.py:8-109.

    patch_embeddings [B,P,D] fp32, bboxes [B,R,4] xywh pixels.  Returns [B,R,D]
    or, in box-set mode, [B,D].  ``attention_map`` [B,P] is COPIED here (the
    reference mutates the caller's CPU copy in place, bbox_utils.py:46-48 -- the
    sequential, order-dependent rescaling *within* one call is reproduced).
    ``gaussian_bbox_variance == 0`` (:62-71): one-hot on the central patch; for an
    even span python's ``random.choice`` picks one of the two central indices --
    restated with the same calls in the same order (y then x, box by box), so a
    caller that seeds ``random`` gets the reference's picks.
    """
    B, P, D = patch_embeddings.shape
    R = bboxes.shape[1]
    g = int(P ** 0.5)
    bounds = all_box_bounds(bboxes, patch_size, g)
    pu = boxes_to_patch_units(bboxes, patch_size)
    tok = patch_embeddings.reshape(B, g, g, D).float()
    amap = attention_map.clone().reshape(B, g, g).float() if attention_map is not None else None
    total = torch.zeros(B, g, g)
    weights_dbg = torch.zeros(B, R, g, g)
    means = []
    for i in range(B):
        img = []
        for j in range(R):
            if get_single_embedding_per_image and int(pu[i, j].sum()) < 0:
                continue  # dummy [-1,-1,-1,-1] pad (bbox_utils.py:40-42)
            y0, y1, x0, x1 = bounds[i, j].tolist()
            region = tok[i, y0:y1, x0:x1]
            hs, ws = region.shape[:2]
            if amap is not None:
                w = amap[i, y0:y1, x0:x1]
                w /= w.sum()  # in place: later boxes see the rescaled map
                total[i, y0:y1, x0:x1] += w
                mean = (region * w.unsqueeze(-1)).sum(dim=(0, 1))
                weights_dbg[i, j, y0:y1, x0:x1] = w
            elif gaussian_avg:
                if gaussian_bbox_variance == 0:
                    w = torch.zeros(hs, ws)
                    cy = random.choice([hs // 2] if hs % 2 == 1 else [hs // 2 - 1, hs // 2])
                    cx = random.choice([ws // 2] if ws % 2 == 1 else [ws // 2 - 1, ws // 2])
                    w[cy, cx] = 1.0
                else:
                    w = gaussian_weights(hs, ws, gaussian_bbox_variance)
                mean = (region * w.unsqueeze(-1)).sum(dim=(0, 1))
                total[i, y0:y1, x0:x1] += w
                weights_dbg[i, j, y0:y1, x0:x1] = w
            else:
                w = torch.ones(hs, ws) / (hs * ws)
                total[i, y0:y1, x0:x1] += w
                mean = region.mean(dim=(0, 1))
                weights_dbg[i, j, y0:y1, x0:x1] = w
            img.append(mean)
        if not get_single_embedding_per_image:
            means.append(torch.stack(img))
    total = total / total.sum(dim=(1, 2), keepdim=True)
    if not get_single_embedding_per_image:
        out = torch.stack(means)
    else:
        out = (total.unsqueeze(-1) * tok).sum(dim=(1, 2))
    if return_debug:
        return out, bounds, (total if get_single_embedding_per_image else weights_dbg)
    return out


def map_traces_to_grid(traces: Sequence[dict], n_patch: int) -> torch.Tensor:
    """bbox_utils.py:158-168: histogram of trace points in Python double arithmetic."""
    grid = [[0] * n_patch for _ in range(n_patch)]
    patch_size = 1.0 / n_patch
    for t in traces:
        x, y = t["x"], t["y"]
        if 0 <= x <= 1 and 0 <= y <= 1:
            gx, gy = int(x / patch_size), int(y / patch_size)
            grid[min(gy, n_patch - 1)][min(gx, n_patch - 1)] += 1
    return torch.tensor(grid, dtype=torch.float32)


def trace_pool(
    patch_tokens: torch.Tensor,
    traces: Sequence[Sequence[dict]],
    self_attn: Optional[torch.Tensor] = None,
    return_grid: bool = False,
):
    """model.py:1049-1054: emb = sum_p w_p x_p / g^2 (a plain ``.mean``, NOT divided by sum w)."""
    B, P, D = patch_tokens.shape
    g = int(P ** 0.5)
    grid = torch.stack([map_traces_to_grid(t, g) for t in traces], dim=0)
    w = grid
    if self_attn is not None:
        w = self_attn.reshape(grid.shape) * grid
    emb = (w.unsqueeze(-1) * patch_tokens.reshape(B, g, g, D)).mean(dim=(1, 2))
    return (emb, grid) if return_grid else emb


def grid_pool(patch_tokens: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
    """The ``masks=`` generalisation of the trace branch: weights [B,R,g,g]
    -> [B,R,D] with the same  sum w x / g^2  formula.  Parity for this argument is pinned
    only through ``trace_pool`` (a trace whose histogram equals the mask)."""
    B, P, D = patch_tokens.shape
    g = int(P ** 0.5)
    w = weights.reshape(B, -1, P).float()
    return torch.einsum("brp,bpd->brd", w, patch_tokens.float()) / float(g * g)


def compute_region_means(patch_embeddings: torch.Tensor, variance: float) -> torch.Tensor:
    """model.py:45-94 (variance 0: one-hot at a python-``random`` central patch per image, :71-79, same call order)."""
    B, P, D = patch_embeddings.shape
    g = int(P ** 0.5)
    tok = patch_embeddings.reshape(B, g, g, D)
    if variance == 0:
        opts = [g // 2] if g % 2 == 1 else [g // 2 - 1, g // 2]
        w = torch.zeros(B, g, g)
        for i in range(B):
            cy = random.choice(opts)
            cx = random.choice(opts)
            w[i, cy, cx] = 1.0
        return (tok * w.unsqueeze(-1)).sum(dim=(1, 2))
    if variance >= 100:
        w = torch.full((g, g), 1 / (g * g))
    else:
        y = torch.linspace(-1, 1, g)
        x = torch.linspace(-1, 1, g)
        yy, xx = torch.meshgrid(y, x, indexing="ij")
        w = torch.exp(-(xx ** 2 + yy ** 2) / variance)
        w = w / w.sum()
    return (tok * w.unsqueeze(0).unsqueeze(-1)).sum(dim=(1, 2))


def cls_attention_map(qkv: torch.Tensor, num_global_tokens: int = 5) -> torch.Tensor:
    """dino_extraction.py:24-34 without the N x N product.

    The reference splits the last block's qkv output into 16 'heads' of 48, scales q by
    0.125, takes row 0 of q k^T, averages over the heads and soft-maxes over the patches.
    Mean-before-softmax makes that   softmax_j( <q_cls, k_j>_768 / 128 )   (SURVEY.md 8a Q1).
    qkv [B,N,3*D] -> [B,P].
    """
    B, N, C3 = qkv.shape
    D = C3 // 3
    q_cls = qkv[:, 0, :D]
    k = qkv[:, num_global_tokens:, D:2 * D]
    logits = torch.einsum("bd,bpd->bp", q_cls, k) * (0.125 / 16.0)
    return logits.softmax(dim=-1)


def process_self_attention_literal(qkv, num_attn_heads=16, scale=0.125, num_global_tokens=5):
    """Literal dino_extraction.py:24-34 (N x N), kept to pin ``cls_attention_map`` on small N."""
    B, N, C3 = qkv.shape
    D = C3 // 3
    t = qkv.reshape(B, N, 3, num_attn_heads, D // num_attn_heads).permute(2, 0, 3, 1, 4)
    q, k = t[0] * scale, t[1]
    attn = q @ k.transpose(-2, -1)
    maps = attn[:, :, 0, num_global_tokens:]
    return maps.mean(dim=1).softmax(dim=-1), maps


def avg_self_attn_token(self_attn: torch.Tensor, patch_tokens: torch.Tensor) -> torch.Tensor:
    """model.py:869: (self_attn[...,None] * patch).mean(1)."""
    return (self_attn.unsqueeze(-1) * patch_tokens).mean(dim=1)


def ctx_cleaner(dirty_embeds: torch.Tensor, ctx_embed: torch.Tensor, cleaning_type: str = "orthogonal_projection",
                alpha: float = 1.0, epsilon: float = 1e-6) -> torch.Tensor:
    """Patchioner.ctx_cleaner, src/model.py:1425-1436.  dirty_embeds [B,P,D], ctx_embed [B,D]."""
    ctx = ctx_embed.unsqueeze(1)
    if cleaning_type == "orthogonal_projection":
        projection = (dirty_embeds @ ctx.transpose(-1, -2)) / (torch.norm(ctx, dim=-1, keepdim=True) ** 2)
        return dirty_embeds - alpha * projection * ctx
    if cleaning_type == "contrastive_mask":
        ctx_norm = torch.norm(ctx, p=2, dim=2, keepdim=True) + epsilon
        return dirty_embeds * (1 - (ctx / ctx_norm))
    raise ValueError(cleaning_type)


def adjust_bbox_for_transform(orig_width: int, orig_height: int, bbox, resize_dim: int, crop_dim: int):
    """src/bbox_utils.py:170-218: box [x, y, w, h] of the original image -> the resized (short side = resize_dim) and
    centre-cropped image.  Python doubles, operation for operation."""
    x1, y1, w, h = bbox
    if orig_width < orig_height:
        scale_w = resize_dim / orig_width
        scale_h = (resize_dim * orig_height) / orig_width / orig_height
    else:
        scale_h = resize_dim / orig_height
        scale_w = (resize_dim * orig_width) / orig_height / orig_width
    new_width = int(orig_width * scale_w)
    new_height = int(orig_height * scale_h)
    x1, y1, w, h = x1 * scale_w, y1 * scale_h, w * scale_w, h * scale_h
    x1 -= max(0, (new_width - crop_dim) // 2)
    y1 -= max(0, (new_height - crop_dim) // 2)
    x1 = max(0, min(x1, crop_dim - 1))
    y1 = max(0, min(y1, crop_dim - 1))
    w = max(0, min(w, crop_dim - x1))
    h = max(0, min(h, crop_dim - y1))
    return [x1, y1, w, h]


def adjust_bbox_for_transform_no_scale(orig_width: int, orig_height: int, bbox, target_width: int, target_height: int):
    """src/bbox_utils.py:222-250: box of the original image -> the image squashed to target_width x target_height."""
    x1, y1, w, h = bbox
    scale_w = target_width / orig_width
    scale_h = target_height / orig_height
    return [x1 * scale_w, y1 * scale_h, w * scale_w, h * scale_h]


def extract_bboxes_feats_double_dino(block_fn, patch_embeddings, bboxes, cls_token, registers_tokens, patch_size=14,
                                     return_type="cls", gaussian_bbox_variance=0.5):
    """src/bbox_utils.py:300-403.  ``block_fn(x [1,L,D]) -> [1,L,D]`` stands for ``dino_model.blocks[-1]``.
    Kept as written, including the slice [bb[1] : bb[3] + 1, bb[0] : bb[2] + 1] that uses w, h as END indices (:329)."""
    N, n_boxes = patch_embeddings.shape[0], bboxes.shape[1]
    grid = int(patch_embeddings.shape[1] ** 0.5)
    D = patch_embeddings.shape[-1]
    bb = bboxes.clone()
    bb //= patch_size
    bb = bb.int()
    pe = patch_embeddings.view(N, grid, grid, D)
    if cls_token is not None:
        offset = 5 if registers_tokens is not None else 1
    else:
        assert return_type != "cls"
        offset = 0
    means = []
    for i in range(N):
        image_means = []
        for j in range(n_boxes):
            region_xy = pe[i, bb[i, j, 1]:bb[i, j, 3] + 1, bb[i, j, 0]:bb[i, j, 2] + 1, :]
            region = region_xy.reshape(1, -1, D)
            if cls_token is not None:
                parts = [cls_token[i].reshape(1, 1, D)]
                if registers_tokens is not None:
                    parts.append(registers_tokens[i].reshape(1, 4, D))
                inputs = torch.cat(parts + [region], dim=1)
            else:
                inputs = region
            outputs = block_fn(inputs)
            out_region = outputs[0, offset:]
            if return_type == "gaussian_avg":
                h_span, w_span = region_xy.shape[:2]
                y, x = torch.meshgrid(torch.linspace(-1, 1, h_span), torch.linspace(-1, 1, w_span), indexing="ij")
                wgt = torch.exp(-(x ** 2 + y ** 2) / gaussian_bbox_variance)
                wgt = wgt / wgt.sum()
                region_mean = (region_xy * wgt.unsqueeze(-1)).sum(dim=(0, 1))
            elif return_type == "avg":
                region_mean = out_region.mean(dim=0)
            else:
                region_mean = outputs[0, 0]
            image_means.append(region_mean)
        means.append(torch.stack(image_means))
    return torch.stack(means)
