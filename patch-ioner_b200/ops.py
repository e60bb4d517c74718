"""Thin torch-tensor wrappers over the C ABI.  torch is plumbing here: device memory, streams."""
from __future__ import annotations

import ctypes as C
import math
import random
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

from . import _lib as L

MODES = {"fp32": L.PIO_FP32, "bf16": L.PIO_BF16}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise L.PioError("libpio_sm100 works on CUDA tensors only (there is no CPU fallback)")


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return L.DT_F32
    if t.dtype == torch.bfloat16:
        return L.DT_BF16
    raise L.PioError(f"unsupported dtype {t.dtype}")


_ws_cache: Dict[Tuple[int, str, int], torch.Tensor] = {}


def workspace(nbytes: int, device, tag: str = "") -> torch.Tensor:
    """A cached, 1024-byte aligned scratch buffer per (device, tag, current stream); grows monotonically.  Keyed by the
    stream so that two forwards in flight on two streams (Patchioner.forward_pipelined) never share scratch memory."""
    key = (torch.device(device).index or 0, tag, torch.cuda.current_stream(device).cuda_stream)
    buf = _ws_cache.get(key)
    need = nbytes + 1024
    if buf is None or buf.numel() < need:
        buf = torch.empty(need, dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    off = (-buf.data_ptr()) % 1024
    return buf[off:off + nbytes]


def launch_count() -> int:
    return int(L.lib().pio_launch_count())


def reset_launch_count() -> None:
    L.lib().pio_reset_launch_count()


# ------------------------------------------------------------------------------------------ dense layer
def linear(A: torch.Tensor, W: torch.Tensor, mode: str = "fp32", bias=None, act: int = L.ACT_NONE, gamma=None,
           residual=None, out: Optional[torch.Tensor] = None, out_dtype=None, alpha: float = 1.0, colscale=None,
           res_rowscale=None, w_static: bool = False) -> torch.Tensor:
    """out = residual * res_rowscale[:,None] + gamma * act(alpha * colscale * (A @ W.T) + bias).

    ``w_static=True`` promises that ``W`` is not being written by a kernel still in flight on the stream (model
    weights): the tcgen05 kernels then fetch their first weight tiles ahead of the dependent-launch wait."""
    _need_cuda(A, W)
    assert A.dim() == 2 and W.dim() == 2 and A.shape[1] == W.shape[1]
    assert A.stride(1) == 1 and W.stride(1) == 1
    M, K = A.shape
    N = W.shape[0]
    if out is None:
        out = torch.empty(M, N, dtype=out_dtype or torch.float32, device=A.device)
    p = L.PioLinear()
    p.A, p.W, p.C = A.data_ptr(), W.data_ptr(), out.data_ptr()
    p.M, p.N, p.K = M, N, K
    p.lda, p.ldw, p.ldc = A.stride(0), W.stride(0), out.stride(0)
    p.a_dt, p.c_dt = _dt(A), _dt(out)
    p.bias, p.colscale, p.gamma = _ptr(bias), _ptr(colscale), _ptr(gamma)
    p.residual, p.res_rowscale = _ptr(residual), _ptr(res_rowscale)
    p.ldres = residual.stride(0) if residual is not None else 0
    p.alpha, p.act = alpha, act
    p.w_static = 1 if w_static else 0
    L.check(L.lib().pio_linear(C.byref(p), MODES[mode], _stream()))
    return out


def linear_argmax(A: torch.Tensor, W: torch.Tensor, bias=None, with_logprob: bool = False):
    """argmax_n (A @ W.T + bias)[m, n] with the arg-max fused into the tcgen05 GEMM epilogue (bf16 operands):
    returns int32 ids [M] (first index on ties) and, optionally, log softmax at the arg-max [M]."""
    _need_cuda(A, W)
    M, K = A.shape
    N = W.shape[0]
    slabs = L.lib().pio_argmax_slabs(M, N)
    val = torch.empty(M, slabs, dtype=torch.float32, device=A.device)
    idx = torch.empty(M, slabs, dtype=torch.int32, device=A.device)
    se = torch.empty(M, slabs, dtype=torch.float32, device=A.device)
    p = L.PioLinear()
    p.A, p.W, p.C = A.data_ptr(), W.data_ptr(), None
    p.M, p.N, p.K = M, N, K
    p.lda, p.ldw, p.ldc = A.stride(0), W.stride(0), N
    p.a_dt, p.c_dt = _dt(A), L.DT_F32
    p.bias = _ptr(bias)
    p.alpha = 1.0
    p.argmax_val, p.argmax_idx, p.argmax_sumexp, p.argmax_ld = val.data_ptr(), idx.data_ptr(), se.data_ptr(), slabs
    L.check(L.lib().pio_linear(C.byref(p), L.PIO_BF16, _stream()))
    ids = torch.empty(M, 1, dtype=torch.int32, device=A.device)
    lp = torch.zeros(M, dtype=torch.float32, device=A.device) if with_logprob else None
    L.check(L.lib().pio_argmax_finish(val.data_ptr(), idx.data_ptr(), se.data_ptr(), slabs, slabs, M, ids.data_ptr(), 1, 0,
                                      _ptr(lp), _stream()))
    return (ids[:, 0], lp) if with_logprob else ids[:, 0]


def layernorm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float, out_dtype=torch.float32) -> torch.Tensor:
    _need_cuda(x, w, b)
    rows, dim = x.shape
    out = torch.empty(rows, dim, dtype=out_dtype, device=x.device)
    L.check(L.lib().pio_layernorm(x.data_ptr(), x.stride(0), w.data_ptr(), b.data_ptr(), out.data_ptr(), _dt(out),
                                  out.stride(0), rows, dim, eps, _stream()))
    return out


def l2_normalize_(x: torch.Tensor) -> torch.Tensor:
    _need_cuda(x)
    assert x.is_contiguous() and x.dtype == torch.float32
    L.check(L.lib().pio_l2_normalize(x.data_ptr(), x.shape[0], x.shape[1], _stream()))
    return x


# ------------------------------------------------------------------------------------------ ViT
def interpolate_pos_embed(pos_embed: torch.Tensor, grid: int) -> torch.Tensor:
    """DINOv2 ``interpolate_pos_encoding`` (offset 0.0): size-based bicubic, antialias, fp32; done once per
    grid on the host side of the boundary (init-time, cached) -- SURVEY.md section 7 'hard parts'."""
    pe = pos_embed.float().reshape(1, -1, pos_embed.shape[-1])
    n = pe.shape[1] - 1
    if grid * grid == n:
        return pe[0].contiguous()
    m = int(math.sqrt(n))
    patch = pe[:, 1:].reshape(1, m, m, -1).permute(0, 3, 1, 2)
    patch = F.interpolate(patch.cpu(), size=(grid, grid), mode="bicubic", antialias=True, align_corners=False)
    patch = patch.permute(0, 2, 3, 1).reshape(1, grid * grid, -1).to(pe.device)
    return torch.cat([pe[:, :1], patch], dim=1)[0].contiguous()


class Vit:
    """DINOv2 ViT-B/14-reg4 on the device; weights by torch.hub state-dict key names."""

    D, NG, PATCH = 768, 5, 14

    def __init__(self, state_dict: Dict[str, torch.Tensor], device, mode: str = "fp32"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise L.PioError("the ViT engine needs a CUDA device (there is no CPU fallback)")
        self.mode = mode
        sd = {k: v.detach().to(self.device, torch.float32).contiguous() for k, v in state_dict.items()}
        self._keep = sd
        w = L.PioVitWeights()
        w.cls_token = sd["cls_token"].data_ptr()
        w.register_tokens = sd["register_tokens"].data_ptr()
        self._patch_w = sd["patch_embed.proj.weight"].reshape(768, -1).contiguous()
        w.patch_w = self._patch_w.data_ptr()
        w.patch_b = sd["patch_embed.proj.bias"].data_ptr()
        names = {"ln1_w": "norm1.weight", "ln1_b": "norm1.bias", "qkv_w": "attn.qkv.weight", "qkv_b": "attn.qkv.bias",
                 "proj_w": "attn.proj.weight", "proj_b": "attn.proj.bias", "ls1": "ls1.gamma", "ln2_w": "norm2.weight",
                 "ln2_b": "norm2.bias", "fc1_w": "mlp.fc1.weight", "fc1_b": "mlp.fc1.bias", "fc2_w": "mlp.fc2.weight",
                 "fc2_b": "mlp.fc2.bias", "ls2": "ls2.gamma"}
        for i in range(12):
            for f, n in names.items():
                setattr(w.blk[i], f, sd[f"blocks.{i}.{n}"].data_ptr())
        w.norm_w, w.norm_b = sd["norm.weight"].data_ptr(), sd["norm.bias"].data_ptr()
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            L.check(L.lib().pio_vit_create(C.byref(h), C.byref(w), MODES[mode], _stream()))
            torch.cuda.current_stream().synchronize()
        self._h = h
        self._pos_src = sd["pos_embed"]
        self._pos: Dict[int, torch.Tensor] = {}

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and L is not None and getattr(L, "_lib", None) is not None:
            L._lib.pio_vit_destroy(h)
            self._h = None

    def pos_embed(self, grid: int) -> torch.Tensor:
        if grid not in self._pos:
            self._pos[grid] = interpolate_pos_embed(self._pos_src, grid)
        return self._pos[grid]

    def forward(self, imgs: torch.Tensor, want_attn: bool = True, want_qkv: bool = False):
        """imgs fp32 [B,3,S,S] -> (tokens [B,N,768] = final LN of [cls | 4 reg | patches], attn [B,P] | None, qkv | None)."""
        _need_cuda(imgs)
        imgs = imgs.contiguous().float()
        B, _, S, S2 = imgs.shape
        assert S == S2, "square crops only (the reference center-crops / resizes to crop_dim x crop_dim)"
        g = S // self.PATCH
        N = self.NG + g * g
        tokens = torch.empty(B, N, self.D, dtype=torch.float32, device=imgs.device)
        attn = torch.empty(B, g * g, dtype=torch.float32, device=imgs.device) if want_attn else None
        qkv = torch.empty(B, N, 3 * self.D, dtype=torch.float32, device=imgs.device) if want_qkv else None
        nbytes = L.lib().pio_vit_workspace_bytes(self._h, B, S)
        ws = workspace(nbytes, imgs.device, "vit")
        L.check(L.lib().pio_vit_forward(self._h, imgs.data_ptr(), B, S, self.pos_embed(g).data_ptr(), tokens.data_ptr(),
                                        _ptr(attn), _ptr(qkv), ws.data_ptr(), nbytes, _stream()))
        return tokens, attn, qkv

    def block_rows(self, x: torch.Tensor, bucket_nseq: Sequence[int], bucket_len: Sequence[int], layer: int = -1) -> torch.Tensor:
        """Run block ``layer`` on packed sequences, IN PLACE: ``x`` fp32 [T,768], rows grouped in buckets of equal-length
        sequences (see pio_vit_block_rows).  Returns ``x`` (the block's output, before any final norm)."""
        _need_cuda(x)
        assert x.is_contiguous() and x.dtype == torch.float32 and x.shape[1] == self.D
        n = len(bucket_nseq)
        a, b = (C.c_int * n)(*[int(v) for v in bucket_nseq]), (C.c_int * n)(*[int(v) for v in bucket_len])
        nbytes = L.lib().pio_vit_block_workspace_bytes(self._h, x.shape[0])
        ws = workspace(nbytes, x.device, "vit_block")
        L.check(L.lib().pio_vit_block_rows(self._h, int(layer), x.data_ptr(), x.shape[0], a, b, n, ws.data_ptr(), nbytes, _stream()))
        return x


def gather_rows(src: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """out[t] = src[idx[t]] for a 2-D fp32 ``src`` (rows may be strided) and int32 ``idx`` on the same device."""
    _need_cuda(src, idx)
    assert src.dim() == 2 and src.dtype == torch.float32 and src.stride(1) == 1 and idx.dtype == torch.int32
    out = torch.empty(idx.numel(), src.shape[1], dtype=torch.float32, device=src.device)
    L.check(L.lib().pio_gather_rows(src.data_ptr(), src.stride(0), idx.contiguous().data_ptr(), idx.numel(), src.shape[1], out.data_ptr(),
                                    _stream()))
    return out


def segment_mean(x: torch.Tensor, seg_start: torch.Tensor, seg_len: torch.Tensor) -> torch.Tensor:
    """Row means of contiguous segments of ``x`` [T,D] (empty segment -> NaN, like tensor.mean())."""
    _need_cuda(x, seg_start, seg_len)
    assert x.is_contiguous() and x.dtype == torch.float32 and seg_start.dtype == torch.int32 and seg_len.dtype == torch.int32
    out = torch.empty(seg_start.numel(), x.shape[1], dtype=torch.float32, device=x.device)
    L.check(L.lib().pio_segment_mean(x.data_ptr(), seg_start.contiguous().data_ptr(), seg_len.contiguous().data_ptr(), seg_start.numel(),
                                     x.shape[1], out.data_ptr(), _stream()))
    return out


def vit_attention(qkv: torch.Tensor, heads: int = 12) -> torch.Tensor:
    """One block's multi-head self-attention: qkv [B,N,3*heads*64] (fp32 or bf16) -> [B,N,heads*64]."""
    _need_cuda(qkv)
    qkv = qkv.contiguous()
    B, N, C3 = qkv.shape
    out = torch.empty(B, N, C3 // 3, dtype=qkv.dtype, device=qkv.device)
    nbytes = L.lib().pio_attention_workspace_bytes(_dt(qkv), B, N, heads)
    ws = workspace(max(nbytes, 16), qkv.device, "attn")
    L.check(L.lib().pio_vit_attention(qkv.data_ptr(), out.data_ptr(), _dt(qkv), B, N, heads, ws.data_ptr(), nbytes, _stream()))
    return out


def cls_attention(qkv: torch.Tensor, num_global: int = 5) -> torch.Tensor:
    _need_cuda(qkv)
    B, N, C3 = qkv.shape
    out = torch.empty(B, N - num_global, dtype=torch.float32, device=qkv.device)
    L.check(L.lib().pio_cls_attention(qkv.contiguous().data_ptr(), _dt(qkv), B, N, C3 // 3, num_global, out.data_ptr(), _stream()))
    return out


def ctx_clean(rows: torch.Tensor, ctx: torch.Tensor, cleaning_type: str = "orthogonal_projection", alpha: float = 1.0,
              epsilon: float = 1e-6, prenorm: bool = False) -> torch.Tensor:
    """ctx_cleaner (model.py:1425-1436) for rows [B,P,D] against ctx [B,D] -> [B,P,D]."""
    _need_cuda(rows, ctx)
    assert rows.dtype == torch.float32 and ctx.dtype == torch.float32 and rows.stride(2) == 1 and ctx.stride(1) == 1
    B, P, D = rows.shape
    out = torch.empty(B, P, D, dtype=torch.float32, device=rows.device)
    mode = {"orthogonal_projection": 0, "contrastive_mask": 1}[cleaning_type]
    L.check(L.lib().pio_ctx_clean(rows.data_ptr(), rows.stride(0), rows.stride(1), ctx.data_ptr(), ctx.stride(0), B, P, D, mode,
                                  float(alpha), float(epsilon), int(prenorm), out.data_ptr(), _stream()))
    return out


def cls_head_attention(qkv: torch.Tensor, num_global: int = 5, heads: int = 16, scale: float = 0.125) -> torch.Tensor:
    """Per-"head" CLS attention maps, softmaxed over the patches (dino_extraction.py:24-34 + model.py:871): [B, heads, P]."""
    _need_cuda(qkv)
    B, N, C3 = qkv.shape
    out = torch.empty(B, heads, N - num_global, dtype=torch.float32, device=qkv.device)
    L.check(L.lib().pio_cls_head_attention(qkv.contiguous().data_ptr(), _dt(qkv), B, N, C3 // 3, num_global, heads, float(scale),
                                           out.data_ptr(), _stream()))
    return out


# ------------------------------------------------------------------------------------------ pooling
def _token_view(patch_tokens: torch.Tensor):
    """patch tokens [B,P,D] (possibly a view into [B,N,D]) -> (ptr, img_stride, row_stride)."""
    assert patch_tokens.dtype == torch.float32 and patch_tokens.stride(2) == 1
    return patch_tokens.data_ptr(), patch_tokens.stride(0), patch_tokens.stride(1)


def pool_boxes(patch_tokens: torch.Tensor, bboxes: torch.Tensor, patch_size: int = 14, gaussian_avg: bool = False,
               gaussian_bbox_variance: float = 0.5, attention_map: Optional[torch.Tensor] = None,
               get_single_embedding_per_image: bool = False, return_bounds: bool = False):
    """extract_bboxes_feats (bbox_utils.py:8-109) on the device.  ``bboxes`` is not modified."""
    _need_cuda(patch_tokens)
    B, P, D = patch_tokens.shape
    g = int(P ** 0.5)
    R = bboxes.shape[1]
    if gaussian_avg and gaussian_bbox_variance == 0 and attention_map is None:
        return _pool_boxes_centre(patch_tokens, bboxes, patch_size, get_single_embedding_per_image, return_bounds)
    if bboxes.dtype.is_floating_point:
        bx, bdt = bboxes.to(patch_tokens.device, torch.float32).contiguous(), L.DT_F32
    else:
        bx, bdt = bboxes.to(patch_tokens.device, torch.int32).contiguous(), L.DT_I32
    mode = L.POOL_ATTN if attention_map is not None else (L.POOL_GAUSS if gaussian_avg else L.POOL_MEAN)
    amap = attention_map.to(patch_tokens.device, torch.float32).contiguous() if attention_map is not None else None
    out = torch.empty((B, D) if get_single_embedding_per_image else (B, R, D), dtype=torch.float32, device=patch_tokens.device)
    bounds = torch.empty(B, R, 4, dtype=torch.int32, device=patch_tokens.device) if return_bounds else None
    nbytes = L.lib().pio_pool_workspace_bytes(B, R, g)
    ws = workspace(nbytes, patch_tokens.device, "pool")
    ptr, istr, rstr = _token_view(patch_tokens)
    L.check(L.lib().pio_pool_boxes(ptr, istr, rstr, B, g, D, bx.data_ptr(), bdt, R, patch_size, mode,
                                   float(gaussian_bbox_variance), _ptr(amap), int(get_single_embedding_per_image),
                                   out.data_ptr(), _ptr(bounds), ws.data_ptr(), nbytes, _stream()))
    return (out, bounds) if return_bounds else out


def _centre_pick(span: int) -> int:
    """bbox_utils.py:65-69 / model.py:74-79: the central index; python ``random`` decides between the two of an even span."""
    return random.choice([span // 2] if span % 2 == 1 else [span // 2 - 1, span // 2])


def _row_index_base(patch_tokens: torch.Tensor):
    """The [B,P,D] token view as one strided 2-D matrix (row b * rows_per_image + p), when the strides allow it."""
    if patch_tokens.stride(0) % patch_tokens.stride(1) != 0:
        return None, 0
    rpi = patch_tokens.stride(0) // patch_tokens.stride(1)
    flat = torch.as_strided(patch_tokens, ((patch_tokens.shape[0] - 1) * rpi + patch_tokens.shape[1], patch_tokens.shape[2]),
                            (patch_tokens.stride(1), 1))
    return flat, rpi


def _pool_boxes_centre(patch_tokens: torch.Tensor, bboxes: torch.Tensor, patch_size: int, set_mode: bool, return_bounds: bool):
    """``gaussian_avg`` with ``gaussian_bbox_variance == 0`` (bbox_utils.py:62-71): the weight is a one-hot on the central patch of
    the box; for an even span the reference lets python's ``random.choice`` pick one of the two central indices.  The picks are made
    here on the host with the same calls in the same order (image by image, box by box, y then x; dummy boxes of the box-set mode
    are skipped before any call), so a caller that seeds ``random`` like the reference gets the reference's picks.  The slice bounds
    are the device's (bit-exact, ``box_bounds_kernel``); the pooled row is then a plain gather (one-hot x tokens is exact)."""
    B, P, D = patch_tokens.shape
    g = int(P ** 0.5)
    R = bboxes.shape[1]
    _, bounds = pool_boxes(patch_tokens, bboxes, patch_size, False, 0.5, None, False, return_bounds=True)
    bh = bounds.cpu().tolist()
    dummy = ((bboxes.detach().cpu() // patch_size).int().sum(-1) < 0).tolist() if set_mode else None  # bbox_utils.py:19-20, 40
    idx: List[int] = []
    owner: List[int] = []
    for i in range(B):
        for j in range(R):
            if set_mode and dummy[i][j]:
                continue
            y0, y1, x0, x1 = bh[i][j]
            hs, ws = y1 - y0, x1 - x0
            if hs <= 0 or ws <= 0:
                raise IndexError(f"pool_boxes: box {j} of image {i} selects no patch (the reference indexes an empty weight map here)")
            cy = _centre_pick(hs)
            cx = _centre_pick(ws)
            idx.append((y0 + cy) * g + (x0 + cx))
            owner.append(i)
    dev = patch_tokens.device
    if set_mode:
        total = torch.zeros(B, P, dtype=torch.float32)
        if idx:
            total.index_put_((torch.tensor(owner), torch.tensor(idx)), torch.ones(len(idx)), accumulate=True)
        total /= total.sum(dim=1, keepdim=True)                 # bbox_utils.py:100 (an image with only dummy boxes: 0/0 = NaN, as there)
        out = pool_grid(patch_tokens, total.to(dev).reshape(B, 1, P), 1.0)[:, 0]
    else:
        flat, rpi = _row_index_base(patch_tokens)
        if flat is not None:
            rows = torch.tensor([o * rpi + p for o, p in zip(owner, idx)], dtype=torch.int32, device=dev)
            out = gather_rows(flat, rows).reshape(B, R, D)
        else:
            w = torch.zeros(B * R, P, dtype=torch.float32)
            w[torch.arange(B * R), torch.tensor(idx)] = 1.0
            out = pool_grid(patch_tokens, w.to(dev).reshape(B, R, P), 1.0)
    return (out, bounds) if return_bounds else out


def region_centre_rows(patch_tokens: torch.Tensor) -> torch.Tensor:
    """compute_region_means with ``variance == 0`` (model.py:71-79): per image, the token of a central patch (python ``random`` picks
    among the two central indices of an even grid, y then x, image by image) -> [B,D]."""
    _need_cuda(patch_tokens)
    B, P, D = patch_tokens.shape
    g = int(P ** 0.5)
    idx = []
    for _ in range(B):
        cy = _centre_pick(g)
        cx = _centre_pick(g)
        idx.append(cy * g + cx)
    flat, rpi = _row_index_base(patch_tokens)
    if flat is not None:
        rows = torch.tensor([i * rpi + p for i, p in enumerate(idx)], dtype=torch.int32, device=patch_tokens.device)
        return gather_rows(flat, rows)
    w = torch.zeros(B, P, dtype=torch.float32)
    w[torch.arange(B), torch.tensor(idx)] = 1.0
    return pool_grid(patch_tokens, w.to(patch_tokens.device).reshape(B, 1, P), 1.0)[:, 0]


def pool_grid(patch_tokens: torch.Tensor, weights: torch.Tensor, scale: float) -> torch.Tensor:
    """out[b,r] = scale * sum_p weights[b,r,p] * x[b,p]   (traces, masks, avg_self_attn, whole-image means)."""
    _need_cuda(patch_tokens, weights)
    B, P, D = patch_tokens.shape
    g = int(P ** 0.5)
    w = weights.reshape(B, -1, P).float().contiguous()
    R = w.shape[1]
    out = torch.empty(B, R, D, dtype=torch.float32, device=patch_tokens.device)
    ptr, istr, rstr = _token_view(patch_tokens)
    L.check(L.lib().pio_pool_grid(ptr, istr, rstr, B, g, D, w.data_ptr(), R, float(scale), out.data_ptr(), _stream()))
    return out


def pack_traces(traces: Sequence[Sequence[dict]]) -> Tuple[torch.Tensor, torch.Tensor]:
    """list (per image) of lists of {'x','y','t'} dicts -> (points float64 [n,2], offsets int32 [T+1]) on the host.
    (One itemgetter call per point: the dict format itself is the cost -- 41 k points take ~25 ms of host time; callers that
    can keep traces as arrays pass the packed tuple to forward(traces=...) directly.)"""
    import numpy as np
    from operator import itemgetter

    xy = itemgetter("x", "y")
    lens = np.fromiter((len(tr) for tr in traces), dtype=np.int64, count=len(traces))
    off = np.zeros(len(traces) + 1, dtype=np.int32)
    np.cumsum(lens, out=off[1:])
    pts = np.array([xy(p) for tr in traces for p in tr], dtype=np.float64).reshape(int(off[-1]), 2)
    return torch.from_numpy(pts), torch.from_numpy(off)


def trace_bins(traces, grid: int, device, attn: Optional[torch.Tensor] = None) -> torch.Tensor:
    """map_traces_to_grid (bbox_utils.py:158-168) for a batch of traces -> counts fp32 [T,grid,grid] (x attn)."""
    pts, off = traces if isinstance(traces, tuple) else pack_traces(traces)
    T = off.numel() - 1
    pts_d = pts.to(device, non_blocking=True)
    off_d = off.to(device, non_blocking=True)
    if pts_d.numel() == 0:
        pts_d = torch.zeros(1, 2, dtype=torch.float64, device=device)
    counts = torch.empty(T, grid, grid, dtype=torch.float32, device=device)
    a = attn.contiguous() if attn is not None else None
    L.check(L.lib().pio_trace_bins(pts_d.data_ptr(), off_d.data_ptr(), T, grid, _ptr(a), counts.data_ptr(), _stream()))
    return counts


def region_mean_weights(grid: int, variance: float, device) -> torch.Tensor:
    w = torch.empty(grid * grid, dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        L.check(L.lib().pio_region_mean_weights(grid, float(variance), w.data_ptr(), _stream()))
    return w


# ------------------------------------------------------------------------------------------ memory bank
class Bank:
    """Caption memory on the device (zero rows are dropped like im2txtprojection.py:345)."""

    def __init__(self, bank: torch.Tensor, device, mode: str = "fp32"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise L.PioError("the memory bank needs a CUDA device (there is no CPU fallback)")
        self.mode = mode
        bank = bank.detach().to(torch.float32)
        bank = bank[bank.norm(dim=-1) != 0]
        b = bank.to(self.device).contiguous()
        self.M, self.D = b.shape
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            L.check(L.lib().pio_bank_create(C.byref(h), b.data_ptr(), self.M, self.D, MODES[mode], _stream()))
            torch.cuda.current_stream().synchronize()
        self._h = h

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and L is not None and getattr(L, "_lib", None) is not None:
            L._lib.pio_bank_destroy(h)
            self._h = None

    def project(self, q: torch.Tensor, temperature: float = 0.01, normalize: bool = False, partial: bool = False):
        """Im2TxtProjector.project (im2txtprojection.py:353-385).  ``q`` is not modified.
        partial=True returns the un-normalised shard partial (m [R], l [R], O [R,D])."""
        _need_cuda(q)
        q = q.float().contiguous()
        R = q.shape[0]
        out = torch.empty(R, self.D, dtype=torch.float32, device=q.device)
        m = torch.empty(R, dtype=torch.float32, device=q.device) if partial else None
        l = torch.empty(R, dtype=torch.float32, device=q.device) if partial else None
        nbytes = L.lib().pio_project_workspace_bytes(self._h, R)
        ws = workspace(nbytes, q.device, "project")
        L.check(L.lib().pio_project(self._h, q.data_ptr(), R, float(temperature), int(normalize), out.data_ptr(), _ptr(m),
                                    _ptr(l), ws.data_ptr(), nbytes, _stream()))
        return (m, l, out) if partial else out


    def best_sims(self, q: torch.Tensor, n: int, with_rows: bool = False):
        """The n largest cosine similarities of each query against the bank, descending (``return_n_best_sims``,
        im2txtprojection.py:382-383) -> [R, n] fp32 (and the bank rows [R, n] int32)."""
        _need_cuda(q)
        q = q.float().contiguous()
        R = q.shape[0]
        sims = torch.empty(R, n, dtype=torch.float32, device=q.device)
        rows = torch.empty(R, n, dtype=torch.int32, device=q.device)
        nbytes = L.lib().pio_project_workspace_bytes(self._h, R)
        ws = workspace(nbytes, q.device, "project")
        L.check(L.lib().pio_best_sims(self._h, q.data_ptr(), R, int(n), sims.data_ptr(), rows.data_ptr(), ws.data_ptr(), nbytes,
                                      _stream()))
        return (sims, rows) if with_rows else sims


def project_rescale_(O, l, m_local, m_global):
    L.check(L.lib().pio_project_rescale(O.data_ptr(), l.data_ptr(), m_local.data_ptr(), m_global.data_ptr(), O.shape[0],
                                        O.shape[1], _stream()))


def project_finish_(O, l, normalize: bool):
    L.check(L.lib().pio_project_finish(O.data_ptr(), l.data_ptr(), O.shape[0], O.shape[1], int(normalize), _stream()))
    return O


# ------------------------------------------------------------------------------------------ decoder
class Decoder:
    """DeCap prefix decoder (GPT-2 4x4x768 + Linear prefix); weights by the reference's state-dict key names."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device, mode: str = "fp32"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise L.PioError("the decoder needs a CUDA device (there is no CPU fallback)")
        self.mode = mode
        T = "decoder.transformer."
        sd = {k: v.detach().to(self.device, torch.float32).contiguous() for k, v in state_dict.items()
              if k.startswith(T) or k.startswith("clip_project.")}
        self._keep = sd
        w = L.PioDecoderWeights()
        w.wte, w.wpe = sd[T + "wte.weight"].data_ptr(), sd[T + "wpe.weight"].data_ptr()
        names = {"ln1_w": "ln_1.weight", "ln1_b": "ln_1.bias", "attn_w": "attn.c_attn.weight", "attn_b": "attn.c_attn.bias",
                 "proj_w": "attn.c_proj.weight", "proj_b": "attn.c_proj.bias", "ln2_w": "ln_2.weight", "ln2_b": "ln_2.bias",
                 "fc_w": "mlp.c_fc.weight", "fc_b": "mlp.c_fc.bias", "fc2_w": "mlp.c_proj.weight", "fc2_b": "mlp.c_proj.bias"}
        for i in range(4):
            for f, n in names.items():
                setattr(w.blk[i], f, sd[f"{T}h.{i}.{n}"].data_ptr())
        w.lnf_w, w.lnf_b = sd[T + "ln_f.weight"].data_ptr(), sd[T + "ln_f.bias"].data_ptr()
        w.prefix_w = sd["clip_project.model.0.weight"].data_ptr()
        w.prefix_b = sd["clip_project.model.0.bias"].data_ptr()
        self.prefix_size = w.prefix_size = sd["clip_project.model.0.weight"].shape[1]
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            L.check(L.lib().pio_decoder_create(C.byref(h), C.byref(w), MODES[mode], _stream()))
            torch.cuda.current_stream().synchronize()
        self._h = h
        self._keep = None  # the library owns repacked copies

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and L is not None and getattr(L, "_lib", None) is not None:
            L._lib.pio_decoder_destroy(h)
            self._h = None

    def decode(self, prefix: torch.Tensor, steps: int = 30, compute_scores: bool = False):
        """decoding_batched (decap.py:116-160) up to the ids: int32 [R,steps] (+ sum of log-probs [R])."""
        _need_cuda(prefix)
        prefix = prefix.float().contiguous()
        R = prefix.shape[0]
        assert prefix.shape[1] == self.prefix_size
        ids = torch.empty(R, steps, dtype=torch.int32, device=prefix.device)
        lp = torch.empty(R, dtype=torch.float32, device=prefix.device) if compute_scores else None
        nbytes = L.lib().pio_decode_workspace_bytes(self._h, R, steps)
        ws = workspace(nbytes, prefix.device, "decode")
        L.check(L.lib().pio_decode_greedy(self._h, prefix.data_ptr(), R, steps, ids.data_ptr(), _ptr(lp), ws.data_ptr(),
                                          nbytes, _stream()))
        return (ids, lp) if compute_scores else ids


_GPT_NAMES = {"ln1_w": "ln_1.weight", "ln1_b": "ln_1.bias", "attn_w": "attn.c_attn.weight", "attn_b": "attn.c_attn.bias",
              "proj_w": "attn.c_proj.weight", "proj_b": "attn.c_proj.bias", "ln2_w": "ln_2.weight", "ln2_b": "ln_2.bias",
              "fc_w": "mlp.c_fc.weight", "fc_b": "mlp.c_fc.bias", "fc2_w": "mlp.c_proj.weight", "fc2_b": "mlp.c_proj.bias"}


class Gpt2Decoder:
    """GPT-2 (12 heads x 64, any depth) that continues a prompt of input embeddings greedily: the language model of the
    ViECap captioner (viecap/ClipCap.py:157, search.py:108-191).  Weights by `GPT2LMHeadModel` state-dict key names
    under `prefix` (ViECap checkpoints: 'gpt.transformer.')."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device, mode: str = "fp32", prefix: str = "gpt.transformer."):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise L.PioError("the decoder needs a CUDA device (there is no CPU fallback)")
        self.mode = mode
        T = prefix
        sd = {k: v.detach().to(self.device, torch.float32).contiguous() for k, v in state_dict.items() if k.startswith(T)}
        n_layer = 1 + max(int(k[len(T) + 2:].split(".")[0]) for k in sd if k.startswith(T + "h."))
        blocks = (L.PioGptBlock * n_layer)()
        for i in range(n_layer):
            for f, n in _GPT_NAMES.items():
                setattr(blocks[i], f, sd[f"{T}h.{i}.{n}"].data_ptr())
        w = L.PioGpt2Weights()
        w.wte, w.wpe = sd[T + "wte.weight"].data_ptr(), sd[T + "wpe.weight"].data_ptr()
        w.lnf_w, w.lnf_b = sd[T + "ln_f.weight"].data_ptr(), sd[T + "ln_f.bias"].data_ptr()
        w.blk, w.n_layer, w.n_head = blocks, n_layer, 12
        self.wte = sd[T + "wte.weight"]  # fp32 [50257,768]: the hard-prompt tokens are embedded with it (word_embed)
        self.n_layer = n_layer
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            L.check(L.lib().pio_decoder_create_gpt2(C.byref(h), C.byref(w), MODES[mode], _stream()))
            torch.cuda.current_stream().synchronize()
        self._h = h

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and L is not None and getattr(L, "_lib", None) is not None:
            L._lib.pio_decoder_destroy(h)
            self._h = None

    def decode(self, prompt: torch.Tensor, steps: int = 64, compute_scores: bool = False, eos: Optional[Sequence[int]] = None):
        """prompt fp32 [R,P,768] input embeddings -> int32 ids [R,steps] (+ sum of log-probs [R]).  With ``eos`` (two token ids) the
        search stops once every row has emitted one of them (the columns after that are filled with ``eos[0]``; the tokens up to each
        row's first end of sentence -- all the reference keeps, search.py:184-190 -- are unchanged); ``self.steps_run`` = steps executed."""
        _need_cuda(prompt)
        prompt = prompt.float().contiguous()
        R, P, D = prompt.shape
        assert D == 768
        ids = torch.empty(R, steps, dtype=torch.int32, device=prompt.device)
        lp = torch.empty(R, dtype=torch.float32, device=prompt.device) if compute_scores else None
        nbytes = L.lib().pio_decode_prompt_workspace_bytes(self._h, R, P, steps)
        ws = workspace(nbytes, prompt.device, "decode_prompt")
        if eos is None:
            L.check(L.lib().pio_decode_greedy_prompt(self._h, prompt.data_ptr(), R, P, steps, ids.data_ptr(), _ptr(lp),
                                                     ws.data_ptr(), nbytes, _stream()))
            self.steps_run = steps
        else:
            ran = C.c_int(0)
            L.check(L.lib().pio_decode_greedy_prompt_eos(self._h, prompt.data_ptr(), R, P, steps, int(eos[0]), int(eos[1]), ids.data_ptr(),
                                                         _ptr(lp), C.byref(ran), ws.data_ptr(), nbytes, _stream()))
            self.steps_run = ran.value
        return (ids, lp) if compute_scores else ids

    def beam_search(self, prompt: torch.Tensor, eos: Sequence[int], beam_width: int = 5, steps: int = 64, temperature: float = 1.0):
        """beam_search (viecap/search.py:193-285) for all prompts at once: prompt fp32 [R,P,768] -> (ids int32 [R,W,steps],
        lengths int32 [R,W], length-normalised scores fp32 [R,W]), beams best first; ``self.beam_steps_run`` = steps executed."""
        _need_cuda(prompt)
        prompt = prompt.float().contiguous()
        R, P, D = prompt.shape
        assert D == 768 and len(eos) == 2  # search.py:273 ("hack"): exactly two end-of-sentence tokens
        W = int(beam_width)
        ids = torch.empty(R, W, steps, dtype=torch.int32, device=prompt.device)
        lens = torch.empty(R, W, dtype=torch.int32, device=prompt.device)
        score = torch.empty(R, W, dtype=torch.float32, device=prompt.device)
        nbytes = L.lib().pio_decode_beam_workspace_bytes(self._h, R, P, steps, W)
        ws = workspace(nbytes, prompt.device, "decode_beam")
        ran = C.c_int(0)
        L.check(L.lib().pio_decode_beam_prompt(self._h, prompt.data_ptr(), R, P, steps, W, int(eos[0]), int(eos[1]), float(temperature),
                                               ids.data_ptr(), lens.data_ptr(), score.data_ptr(), C.byref(ran), ws.data_ptr(), nbytes,
                                               _stream()))
        self.beam_steps_run = ran.value
        return ids, lens, score

    def score_tokens(self, ids: torch.Tensor, lens: torch.Tensor) -> torch.Tensor:
        """Mean negative log-likelihood per right-padded token row (int32 ids [R,n], int32 lens [R]) -> fp32 [R]."""
        _need_cuda(ids, lens)
        assert ids.dtype == torch.int32 and lens.dtype == torch.int32 and ids.dim() == 2
        ids, lens = ids.contiguous(), lens.contiguous()
        R, n = ids.shape
        out = torch.empty(R, dtype=torch.float32, device=ids.device)
        nbytes = L.lib().pio_gpt2_score_workspace_bytes(self._h, R, n)
        ws = workspace(nbytes, ids.device, "gpt2_score")
        L.check(L.lib().pio_gpt2_score_tokens(self._h, ids.data_ptr(), lens.data_ptr(), R, n, out.data_ptr(), ws.data_ptr(), nbytes,
                                              _stream()))
        return out


class Mapper:
    """ViECap mapping network (viecap/ClipCap.py:122-153); weights by the reference's key names under `prefix`."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device, mode: str = "fp32", prefix: str = "mapping_network.",
                 n_head: int = 8):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise L.PioError("the mapping network needs a CUDA device (there is no CPU fallback)")
        T = prefix
        sd = {k: v.detach().to(self.device, torch.float32).contiguous() for k, v in state_dict.items() if k.startswith(T)}
        LT = T + "transformer.layers."
        n_layer = 1 + max(int(k[len(LT):].split(".")[0]) for k in sd if k.startswith(LT))
        names = {"norm1_w": "norm1.weight", "norm1_b": "norm1.bias", "q_w": "attn.to_queries.weight",
                 "kv_w": "attn.to_keys_values.weight", "proj_w": "attn.project.weight", "proj_b": "attn.project.bias",
                 "norm2_w": "norm2.weight", "norm2_b": "norm2.bias", "fc1_w": "mlp.fc1.weight", "fc1_b": "mlp.fc1.bias",
                 "fc2_w": "mlp.fc2.weight", "fc2_b": "mlp.fc2.bias"}
        for i in range(n_layer):
            if f"{LT}{i}.attn.to_queries.bias" in sd:
                raise NotImplementedError("mapping network with biased q / kv projections (the reference builds bias=False)")
        layers = (L.PioMapperLayer * n_layer)()
        for i in range(n_layer):
            for f, n in names.items():
                setattr(layers[i], f, sd[f"{LT}{i}.{n}"].data_ptr())
        w = L.PioMapperWeights()
        lin_w, pc = sd[T + "linear.weight"], sd[T + "prefix_const"]
        assert pc.shape[1] == 768 and lin_w.shape[0] % 768 == 0
        self.clip_size = w.clip_size = lin_w.shape[1]
        w.project_len = lin_w.shape[0] // 768
        self.prefix_len = w.prefix_len = pc.shape[0]
        w.n_layer, w.n_head, w.hidden = n_layer, n_head, sd[f"{LT}0.mlp.fc1.weight"].shape[0]
        w.linear_w, w.linear_b, w.prefix_const, w.layers = lin_w.data_ptr(), sd[T + "linear.bias"].data_ptr(), pc.data_ptr(), layers
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            L.check(L.lib().pio_mapper_create(C.byref(h), C.byref(w), MODES[mode], _stream()))
            torch.cuda.current_stream().synchronize()
        self._h = h

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and L is not None and getattr(L, "_lib", None) is not None:
            L._lib.pio_mapper_destroy(h)
            self._h = None

    def forward(self, feats: torch.Tensor) -> torch.Tensor:
        """feats fp32 [R,clip_size] (unit rows) -> fp32 [R,prefix_len,768]."""
        _need_cuda(feats)
        feats = feats.float().contiguous()
        R = feats.shape[0]
        assert feats.shape[1] == self.clip_size
        out = torch.empty(R, self.prefix_len, 768, dtype=torch.float32, device=feats.device)
        nbytes = L.lib().pio_mapper_workspace_bytes(self._h, R)
        ws = workspace(nbytes, feats.device, "mapper")
        L.check(L.lib().pio_mapper_forward(self._h, feats.data_ptr(), R, out.data_ptr(), ws.data_ptr(), nbytes, _stream()))
        return out


def entity_topk(q: torch.Tensor, entities: torch.Tensor, temperature: float, k: int):
    """softmax(q . E^T / temperature) and its k best entries per row (retrieval_categories.py:87-115); q, E unit rows."""
    _need_cuda(q, entities)
    q, entities = q.float().contiguous(), entities.float().contiguous()
    R, D = q.shape
    prob = torch.empty(R, k, dtype=torch.float32, device=q.device)
    idx = torch.empty(R, k, dtype=torch.int32, device=q.device)
    L.check(L.lib().pio_entity_topk(q.data_ptr(), entities.data_ptr(), R, entities.shape[0], D, float(temperature), k,
                                    prob.data_ptr(), idx.data_ptr(), _stream()))
    return prob, idx
