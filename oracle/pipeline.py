"""Oracle (test infrastructure): the whole ``Patchioner.forward`` hot path on CPU.

Restates the glue of Patch-ioner/src/model.py:718-1058 for the ``talk2dino_decap`` /
``talk2dino_capdec`` models: ViT -> CLS attention map (:868) -> region pooling (:980-1054)
-> ``caption_tokens`` (:1392-1423: memory projection iff a bank exists, optional Talk2DINO
inversion, greedy decode).  Returns token ids (what ``decoding_method`` would receive,
decap.py:166-167) instead of detokenised strings.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch

from . import decap, dinov2, memory, pooling


class OracleModel:
    def __init__(self, vit_w, dec_w, bank: Optional[torch.Tensor], normalize: bool = True,
                 talk2dino_A_pinv=None, talk2dino_b=None, patch_size: int = 14):
        self.vit_w = vit_w
        self.dec_w = dec_w
        self.bank = memory.drop_zero_rows(bank) if bank is not None else None
        self.normalize = normalize
        self.A_pinv = talk2dino_A_pinv
        self.b = talk2dino_b
        self.patch_size = patch_size

    # model.py:1392-1423
    def caption_tokens(self, feats: torch.Tensor, project: bool = True, use_cache: bool = True,
                       compute_scores: bool = False):
        x = feats
        if self.bank is not None and project:
            x = memory.project(x, self.bank, normalize=self.normalize)
        if self.A_pinv is not None:
            x = memory.revert_transformation(x, self.A_pinv, self.b)
        return decap.decode_greedy(self.dec_w, x, compute_scores=compute_scores, use_cache=use_cache)

    @torch.no_grad()
    def forward(self, imgs, get_cls_capt=True, get_avg_self_attn_capt=False, bboxes=None, traces=None,
                get_controllable_capts=False, bs_factor=4, gaussian_avg=False, gaussian_bbox_variance=0.5,
                get_avg_patch_capt=False, gaussian_img_variance=1, use_attn_map_for_bboxes=False,
                use_attention_tracing=False, use_cache=True, return_embeds=False) -> Dict[str, object]:
        outs: Dict[str, object] = {}
        emb: Dict[str, torch.Tensor] = {}
        bs = imgs.shape[0]
        d = dinov2.forward(self.vit_w, imgs)
        patch = d["x_norm_patchtokens"]
        self_attn = pooling.cls_attention_map(d["qkv"])
        D = patch.shape[-1]
        if get_cls_capt:
            emb["cls_capt"] = d["x_norm_clstoken"]
            outs["cls_capt"] = self.caption_tokens(d["x_norm_clstoken"].clone(), use_cache=use_cache)
        if get_avg_self_attn_capt:
            e = pooling.avg_self_attn_token(self_attn, patch)
            emb["avg_self_attn_capt"] = e
            outs["avg_self_attn_capt"] = self.caption_tokens(e, use_cache=use_cache)
        if get_avg_patch_capt:
            e = pooling.compute_region_means(patch, gaussian_img_variance)
            emb["avg_patch_capt"] = e
            outs["avg_patch_capt"] = self.caption_tokens(e, use_cache=use_cache)
        if bboxes is not None and not get_controllable_capts:
            amap = self_attn if use_attn_map_for_bboxes else None
            feats = pooling.extract_bboxes_feats(patch, bboxes, gaussian_avg, gaussian_bbox_variance,
                                                 patch_size=self.patch_size, attention_map=amap)
            n_boxes = bboxes.shape[1]
            feats = feats.reshape(-1, D)
            emb["bbox_capts"] = feats
            bbox_bs = bs * bs_factor
            n_batch = math.ceil(feats.shape[0] / bbox_bs)
            ids = []
            for i in range(n_batch):  # model.py:1008-1035: chunks of bs*bs_factor regions
                s = i * bbox_bs
                e_ = s + bbox_bs if i < n_batch - 1 else feats.shape[0]
                ids.append(self.caption_tokens(feats[s:e_].clone(), use_cache=use_cache))
            ids = torch.cat(ids, dim=0)
            outs["bbox_capts"] = ids.reshape(bs, n_boxes, -1)
        elif bboxes is not None and get_controllable_capts:
            amap = self_attn if use_attn_map_for_bboxes else None
            feats = pooling.extract_bboxes_feats(patch, bboxes, gaussian_avg, gaussian_bbox_variance,
                                                 get_single_embedding_per_image=True,
                                                 patch_size=self.patch_size, attention_map=amap)
            emb["set_controllable_capts"] = feats
            outs["set_controllable_capts"] = self.caption_tokens(feats.clone(), use_cache=use_cache)
        if traces is not None:
            e = pooling.trace_pool(patch, traces, self_attn if use_attention_tracing else None)
            emb["trace_capts"] = e
            outs["trace_capts"] = self.caption_tokens(e.clone(), use_cache=use_cache)
        if return_embeds:
            outs["_embeds"] = emb
            outs["_self_attn"] = self_attn
            outs["_vit"] = d
        return outs


# ----------------------------------------------------------------------------
# synthetic inputs -- shared by tests and bench so both arms see the same data
# ----------------------------------------------------------------------------


def synth_images(B: int, S: int, seed: int = 1) -> torch.Tensor:
    return torch.randn(B, 3, S, S, generator=torch.Generator().manual_seed(seed))


def synth_boxes(B: int, R: int, S: int, seed: int = 1, degenerate_frac: float = 0.05,
                pad: Optional[str] = None) -> torch.Tensor:
    """xywh float32 in crop pixels: x,y ~ U{0..S-15}, w ~ U{14..S-x}, h ~ U{14..S-y}; 5 % degenerate
    w,h in {1..13}; ``pad='dense'`` makes the last box of each image [0,0,1,1]
    (eval_densecap.py:332), ``pad='set'`` makes it [-1,-1,-1,-1] (eval_region_set_captioning.py:268)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randint(0, S - 14, (B, R), generator=g)
    y = torch.randint(0, S - 14, (B, R), generator=g)
    u = torch.rand(B, R, generator=g)
    v = torch.rand(B, R, generator=g)
    w = 14 + torch.floor(u * (S - x - 14 + 1).float()).long()
    h = 14 + torch.floor(v * (S - y - 14 + 1).float()).long()
    deg = torch.rand(B, R, generator=g) < degenerate_frac
    w = torch.where(deg, torch.randint(1, 14, (B, R), generator=g), w)
    h = torch.where(deg, torch.randint(1, 14, (B, R), generator=g), h)
    boxes = torch.stack([x, y, w, h], dim=-1).float()
    if pad == "dense" and R > 1:
        boxes[:, -1] = torch.tensor([0.0, 0.0, 1.0, 1.0])
    if pad == "set" and R > 1:
        boxes[:, -1] = -1.0
    return boxes


def synth_traces(B: int, seed: int = 1, n_min: int = 64, n_max: int = 256, outside_frac: float = 0.03) -> List[List[dict]]:
    """Random-walk mouse traces, 3 % of points pushed outside [0,1] (bbox_utils.py:164 filter)."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(B):
        n = int(torch.randint(n_min, n_max + 1, (1,), generator=g))
        p = torch.rand(2, generator=g, dtype=torch.float64)
        steps = torch.randn(n, 2, generator=g, dtype=torch.float64) * 0.03
        pts = p + torch.cumsum(steps, dim=0)
        pts = pts - torch.floor(pts / 2.0) * 2.0      # fold into [0,2)
        pts = torch.where(pts > 1.0, 2.0 - pts, pts)  # reflect into [0,1]
        outside = torch.rand(n, generator=g) < outside_frac
        pts[outside] = pts[outside] + 1.5
        out.append([{"x": float(a), "y": float(b), "t": float(k)} for k, (a, b) in enumerate(pts.tolist())])
    return out


def synth_bank(M: int, D: int = 768, seed: int = 7, zero_frac: float = 0.001) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    bank = torch.randn(M, D, generator=g)
    nz = max(1, int(M * zero_frac)) if zero_frac > 0 else 0
    if nz:
        idx = torch.randperm(M, generator=g)[:nz]
        bank[idx] = 0.0
    return bank
