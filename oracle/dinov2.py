"""Oracle (test infrastructure): DINOv2 ViT-B/14 with 4 registers, forward only.

The reference obtains this network from a THIRD-PARTY dependency that is absent from
``/root/reference``:  ``torch.hub.load('facebookresearch/dinov2', 'dinov2_vitb14_reg')``
(Patch-ioner/src/model.py:342-343; unpinned default branch, not listed in
requirements.txt, no hub cache and no network in this image).  What follows restates the
published architecture (``dinov2/models/vision_transformer.py``, ``dinov2/layers/*``,
``dinov2/hub/backbones.py``: img 518, patch 14, dim 768, depth 12, heads 12, mlp x4,
LayerScale, 4 register tokens, ``interpolate_antialias=True``, ``interpolate_offset=0.0``)
and is anchored on the reference's own call sites:

  * ``self.dino(imgs, is_training=True)`` -> dict with ``x_norm_clstoken``,
    ``x_norm_regtokens``, ``x_norm_patchtokens``, ``x_prenorm``   (model.py:783, dino_extraction.py:14-22)
  * a forward hook on ``blocks[-1].attn.qkv`` whose OUTPUT ``[B,N,3*768]`` in channel order
    ``[q | k | v]`` feeds ``process_self_attention``              (model.py:589-590, dino_extraction.py:8,24-26)
  * ``self.dino.patch_size``                                       (model.py:595)

Parity: unpinned against upstream (not available offline); cross-checked in
``tests/test_oracle_golden.py`` against the independent
``transformers.Dinov2WithRegistersModel`` port installed in the image.

State-dict key names are the hub checkpoint's, so real weights load unchanged.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

EMBED = 768
DEPTH = 12
HEADS = 12
MLP = 3072
PATCH = 14
NREG = 4
BASE_GRID = 37  # 518 / 14: pos_embed is [1, 1 + 37*37, 768]


def make_weights(seed: int = 1234, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Seed-fixed random-init weights with the hub key names.

    Scales are chosen so that activations stay O(1) through 12 blocks and attention is not
    degenerate (fan-in scaled linears, LayerScale in [0.05, 0.3]) -- random init, but a
    meaningful numerical test.
    """
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s, std=1.0: (torch.randn(*s, generator=g) * std).to(dtype)  # noqa: E731
    ru = lambda *s, lo=0.0, hi=1.0: (torch.rand(*s, generator=g) * (hi - lo) + lo).to(dtype)  # noqa: E731
    w: Dict[str, torch.Tensor] = {}
    w["cls_token"] = rn(1, 1, EMBED, std=0.5)
    w["pos_embed"] = rn(1, 1 + BASE_GRID * BASE_GRID, EMBED, std=0.2)
    w["register_tokens"] = rn(1, NREG, EMBED, std=0.5)
    w["mask_token"] = torch.zeros(1, EMBED, dtype=dtype)
    w["patch_embed.proj.weight"] = rn(EMBED, 3, PATCH, PATCH, std=1.0 / math.sqrt(3 * PATCH * PATCH))
    w["patch_embed.proj.bias"] = rn(EMBED, std=0.1)
    for i in range(DEPTH):
        p = f"blocks.{i}."
        w[p + "norm1.weight"] = ru(EMBED, lo=0.8, hi=1.2)
        w[p + "norm1.bias"] = rn(EMBED, std=0.05)
        w[p + "attn.qkv.weight"] = rn(3 * EMBED, EMBED, std=1.0 / math.sqrt(EMBED))
        w[p + "attn.qkv.bias"] = rn(3 * EMBED, std=0.1)
        w[p + "attn.proj.weight"] = rn(EMBED, EMBED, std=1.0 / math.sqrt(EMBED))
        w[p + "attn.proj.bias"] = rn(EMBED, std=0.1)
        w[p + "ls1.gamma"] = ru(EMBED, lo=0.05, hi=0.3)
        w[p + "norm2.weight"] = ru(EMBED, lo=0.8, hi=1.2)
        w[p + "norm2.bias"] = rn(EMBED, std=0.05)
        w[p + "mlp.fc1.weight"] = rn(MLP, EMBED, std=1.0 / math.sqrt(EMBED))
        w[p + "mlp.fc1.bias"] = rn(MLP, std=0.1)
        w[p + "mlp.fc2.weight"] = rn(EMBED, MLP, std=1.0 / math.sqrt(MLP))
        w[p + "mlp.fc2.bias"] = rn(EMBED, std=0.1)
        w[p + "ls2.gamma"] = ru(EMBED, lo=0.05, hi=0.3)
    w["norm.weight"] = ru(EMBED, lo=0.8, hi=1.2)
    w["norm.bias"] = rn(EMBED, std=0.05)
    return w


def interpolate_pos_embed(pos_embed: torch.Tensor, grid: int) -> torch.Tensor:
    """Upstream ``interpolate_pos_encoding`` with ``interpolate_offset=0.0``: size-based bicubic,
    ``align_corners=False``, ``antialias=True``, in fp32; identity when grid == 37.
    Returns [1, 1 + grid*grid, D].  Done once on the host by the product too."""
    n = pos_embed.shape[1] - 1
    if grid * grid == n:
        return pos_embed.float()
    pe = pos_embed.float()
    cls_pe, patch_pe = pe[:, :1], pe[:, 1:]
    m = int(math.sqrt(n))
    patch_pe = patch_pe.reshape(1, m, m, -1).permute(0, 3, 1, 2)
    patch_pe = F.interpolate(patch_pe, size=(grid, grid), mode="bicubic", antialias=True, align_corners=False)
    patch_pe = patch_pe.permute(0, 2, 3, 1).reshape(1, grid * grid, -1)
    return torch.cat([cls_pe, patch_pe], dim=1)


def prepare_tokens(w: Dict[str, torch.Tensor], imgs: torch.Tensor) -> torch.Tensor:
    """patch-embed conv 14/14 -> [cls | patches] + pos_embed -> insert 4 registers after cls."""
    B, _, H, W = imgs.shape
    g = H // PATCH
    x = F.conv2d(imgs, w["patch_embed.proj.weight"], w["patch_embed.proj.bias"], stride=PATCH)
    x = x.flatten(2).transpose(1, 2)  # [B, g*g, D], row-major (y then x)
    x = torch.cat([w["cls_token"].expand(B, -1, -1), x], dim=1)
    x = x + interpolate_pos_embed(w["pos_embed"], g)
    x = torch.cat([x[:, :1], w["register_tokens"].expand(B, -1, -1), x[:, 1:]], dim=1)
    return x


def block_forward(w, i: int, x: torch.Tensor, capture: dict | None = None) -> torch.Tensor:
    """x += ls1 * proj(MHA(LN(x)));  x += ls2 * fc2(GELU_erf(fc1(LN(x)))).  LN eps 1e-6."""
    p = f"blocks.{i}."
    B, N, D = x.shape
    h = F.layer_norm(x, (D,), w[p + "norm1.weight"], w[p + "norm1.bias"], eps=1e-6)
    qkv = F.linear(h, w[p + "attn.qkv.weight"], w[p + "attn.qkv.bias"])  # [B,N,3D] = [q|k|v]
    if capture is not None:
        capture["qkv"] = qkv
    t = qkv.reshape(B, N, 3, HEADS, D // HEADS).permute(2, 0, 3, 1, 4)
    q, k, v = t[0] * (D // HEADS) ** -0.5, t[1], t[2]
    a = (q @ k.transpose(-2, -1)).softmax(dim=-1)
    o = (a @ v).transpose(1, 2).reshape(B, N, D)
    o = F.linear(o, w[p + "attn.proj.weight"], w[p + "attn.proj.bias"])
    x = x + w[p + "ls1.gamma"] * o
    h = F.layer_norm(x, (D,), w[p + "norm2.weight"], w[p + "norm2.bias"], eps=1e-6)
    h = F.linear(h, w[p + "mlp.fc1.weight"], w[p + "mlp.fc1.bias"])
    h = F.gelu(h)  # exact erf GELU
    h = F.linear(h, w[p + "mlp.fc2.weight"], w[p + "mlp.fc2.bias"])
    return x + w[p + "ls2.gamma"] * h


@torch.no_grad()
def forward(w: Dict[str, torch.Tensor], imgs: torch.Tensor, depth: int = DEPTH) -> Dict[str, torch.Tensor]:
    """``dino(imgs, is_training=True)`` plus the hooked last-block qkv (key ``'qkv'``)."""
    x = prepare_tokens(w, imgs.float())
    cap: dict = {}
    for i in range(depth):
        x = block_forward(w, i, x, cap if i == depth - 1 else None)
    xn = F.layer_norm(x, (x.shape[-1],), w["norm.weight"], w["norm.bias"], eps=1e-6)
    return {
        "x_norm_clstoken": xn[:, 0],
        "x_norm_regtokens": xn[:, 1:1 + NREG],
        "x_norm_patchtokens": xn[:, 1 + NREG:],
        "x_prenorm": x,
        "qkv": cap["qkv"],
    }


def flops_per_image(size: int) -> float:
    """12*(24*N*D^2 + 4*N^2*D) + 2*P*588*D."""
    P = (size // PATCH) ** 2
    N = P + 1 + NREG
    return 12 * (24 * N * EMBED ** 2 + 4 * N * N * EMBED) + 2 * P * 588 * EMBED
