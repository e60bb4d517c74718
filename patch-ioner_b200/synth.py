"""Seeded synthetic weights and inputs for bench.py and examples.

Random-init weights with the reference's state-dict key names, synthetic images / boxes / traces / caption
banks.  Generators only -- no algorithm lives here.  ``tests/test_synth_cpu.py`` checks that these produce
exactly the tensors of the oracle's generators, so that the product arm and the CPU arm of the benchmark see
the same data without the product importing ``oracle/``.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch

EMBED, DEPTH, HEADS, MLP, PATCH, NREG, BASE_GRID = 768, 12, 12, 3072, 14, 4, 37
N_LAYER, N_HEAD, N_EMBD, VOCAB, N_POS = 4, 4, 768, 50257, 1024


def make_vit_weights(seed: int = 1234, dtype=torch.float32) -> Dict[str, torch.Tensor]:
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


def make_decoder_weights(seed: int = 1234, prefix_size: int = 768, init_std: float = 0.02) -> Dict[str, torch.Tensor]:
    """Seed-fixed random init following HF GPT-2 ``_init_weights`` (normal(0, 0.02), LN = (1, 0),
    residual projections scaled by 1/sqrt(2*n_layer)) and nn.Linear's default for ``clip_project``."""
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s, std=init_std: torch.randn(*s, generator=g) * std  # noqa: E731
    w: Dict[str, torch.Tensor] = {}
    T = "decoder.transformer."
    w[T + "wte.weight"] = rn(VOCAB, N_EMBD)
    w[T + "wpe.weight"] = rn(N_POS, N_EMBD)
    for i in range(N_LAYER):
        p = f"{T}h.{i}."
        w[p + "ln_1.weight"] = torch.ones(N_EMBD) + rn(N_EMBD, std=0.05)
        w[p + "ln_1.bias"] = rn(N_EMBD, std=0.02)
        w[p + "attn.c_attn.weight"] = rn(N_EMBD, 3 * N_EMBD)
        w[p + "attn.c_attn.bias"] = rn(3 * N_EMBD, std=0.01)
        w[p + "attn.c_proj.weight"] = rn(N_EMBD, N_EMBD, std=init_std / math.sqrt(2 * N_LAYER))
        w[p + "attn.c_proj.bias"] = rn(N_EMBD, std=0.01)
        w[p + "ln_2.weight"] = torch.ones(N_EMBD) + rn(N_EMBD, std=0.05)
        w[p + "ln_2.bias"] = rn(N_EMBD, std=0.02)
        w[p + "mlp.c_fc.weight"] = rn(N_EMBD, 4 * N_EMBD)
        w[p + "mlp.c_fc.bias"] = rn(4 * N_EMBD, std=0.01)
        w[p + "mlp.c_proj.weight"] = rn(4 * N_EMBD, N_EMBD, std=init_std / math.sqrt(2 * N_LAYER))
        w[p + "mlp.c_proj.bias"] = rn(N_EMBD, std=0.01)
    w[T + "ln_f.weight"] = torch.ones(N_EMBD) + rn(N_EMBD, std=0.05)
    w[T + "ln_f.bias"] = rn(N_EMBD, std=0.02)
    w["decoder.lm_head.weight"] = w[T + "wte.weight"]  # tied
    bound = 1.0 / math.sqrt(prefix_size)
    w["clip_project.model.0.weight"] = (torch.rand(N_EMBD, prefix_size, generator=g) * 2 - 1) * bound
    w["clip_project.model.0.bias"] = (torch.rand(N_EMBD, generator=g) * 2 - 1) * bound
    return w


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


# ------------------------------------------------------------------------------------------ ViECap (BASELINE configs[3])
def make_viecap_weights(seed: int = 4321, n_layer_gpt: int = 12, n_layer_map: int = 8, clip_size: int = 768, project_len: int = 10,
                        prefix_len: int = 10, std: float = 0.02) -> Dict[str, torch.Tensor]:
    """Random-init ViECap checkpoint with the reference's key names (`mapping_network.*`, `gpt.transformer.*`): 8-layer
    mapping network, GPT-2 small.  Same generator order as oracle/viecap.py::make_weights."""
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s, std=std: torch.randn(*s, generator=g) * std  # noqa: E731
    D = N_EMBD
    w: Dict[str, torch.Tensor] = {}
    M = "mapping_network."
    w[M + "linear.weight"] = rn(project_len * D, clip_size, std=0.05)
    w[M + "linear.bias"] = rn(project_len * D, std=0.02)
    w[M + "prefix_const"] = rn(prefix_len, D, std=0.5)
    for i in range(n_layer_map):
        p = f"{M}transformer.layers.{i}."
        w[p + "norm1.weight"] = torch.ones(D) + rn(D, std=0.05)
        w[p + "norm1.bias"] = rn(D)
        w[p + "attn.to_queries.weight"] = rn(D, D, std=0.04)
        w[p + "attn.to_keys_values.weight"] = rn(2 * D, D, std=0.04)
        w[p + "attn.project.weight"] = rn(D, D, std=0.03)
        w[p + "attn.project.bias"] = rn(D)
        w[p + "norm2.weight"] = torch.ones(D) + rn(D, std=0.05)
        w[p + "norm2.bias"] = rn(D)
        w[p + "mlp.fc1.weight"] = rn(2 * D, D, std=0.03)
        w[p + "mlp.fc1.bias"] = rn(2 * D)
        w[p + "mlp.fc2.weight"] = rn(D, 2 * D, std=0.03)
        w[p + "mlp.fc2.bias"] = rn(D)
    T = "gpt.transformer."
    w[T + "wte.weight"] = rn(VOCAB, D)
    w[T + "wpe.weight"] = rn(N_POS, D)
    for i in range(n_layer_gpt):
        p = f"{T}h.{i}."
        w[p + "ln_1.weight"] = torch.ones(D) + rn(D, std=0.05)
        w[p + "ln_1.bias"] = rn(D)
        w[p + "attn.c_attn.weight"] = rn(D, 3 * D)
        w[p + "attn.c_attn.bias"] = rn(3 * D, std=0.01)
        w[p + "attn.c_proj.weight"] = rn(D, D, std=std / math.sqrt(2 * n_layer_gpt))
        w[p + "attn.c_proj.bias"] = rn(D, std=0.01)
        w[p + "ln_2.weight"] = torch.ones(D) + rn(D, std=0.05)
        w[p + "ln_2.bias"] = rn(D)
        w[p + "mlp.c_fc.weight"] = rn(D, 4 * D)
        w[p + "mlp.c_fc.bias"] = rn(4 * D, std=0.01)
        w[p + "mlp.c_proj.weight"] = rn(4 * D, D, std=std / math.sqrt(2 * n_layer_gpt))
        w[p + "mlp.c_proj.bias"] = rn(D, std=0.01)
    w[T + "ln_f.weight"] = torch.ones(D) + rn(D, std=0.05)
    w[T + "ln_f.bias"] = rn(D)
    w["gpt.lm_head.weight"] = w[T + "wte.weight"]
    return w


class WordTokenizer:
    """Offline stand-in for the GPT-2 BPE tokenizer (its vocabulary files need the network): GPT-2's word / punctuation
    pre-tokenisation, one id per piece.  Only for synthetic runs (bench.py); real use passes a real tokenizer."""
    pad_token_id = None

    def __init__(self):
        import re
        self._pat = re.compile(r" ?[A-Za-z]+| ?[0-9]+| ?[^\sA-Za-z0-9]+|\s+")
        self.names: Dict[int, str] = {}

    def encode(self, text: str) -> List[int]:
        import zlib
        out = []
        for piece in self._pat.findall(text):
            i = zlib.crc32(piece.encode()) % 50000
            self.names.setdefault(i, piece)
            out.append(i)
        return out

    def decode(self, ids) -> str:
        return "".join(self.names.get(int(i), f"<{int(i)}>") for i in ids)


def synth_entities(n: int = 80, D: int = 768, seed: int = 5):
    """n entity names (the COCO vocabulary has 80) and random unit embeddings."""
    g = torch.Generator().manual_seed(seed)
    e = torch.randn(n, D, generator=g)
    return sorted(f"thing{i:02d}" for i in range(n)), e / e.norm(dim=-1, keepdim=True)
