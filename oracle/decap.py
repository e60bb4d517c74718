"""Oracle (test infrastructure): DeCap prefix decoder and its batched greedy decode.

Restates
  * ``DeCap`` / ``MLP``          Patch-ioner/src/decap/decap.py:46-79  (+ src/decap/decoder_config.pkl:
                                 GPT-2, 4 layers x 4 heads, n_embd 768, vocab 50257, gelu_new, LN eps 1e-5,
                                 tied wte / lm_head)
  * ``decoding_batched``         Patch-ioner/src/decap/decap.py:116-183
GPT-2 arithmetic itself lives in the third-party ``transformers`` package (pinned 4.46.3 in
requirements.txt:2; 5.5.0 in this image): ``modeling_gpt2.py`` -- h0 = inputs_embeds + wpe[0..T),
pre-LN blocks, causal attention scaled by head_dim^-0.5, ``Conv1D`` weights stored [in, out],
``gelu_new``.  Restated here functionally on the HF state-dict key names
(``decoder.transformer.h.{i}.attn.c_attn.weight`` ..., ``clip_project.model.0.*``) and pinned in
``tests/test_oracle_golden.py`` against ``transformers.GPT2LMHeadModel`` and against the reference's
own ``decoding_batched`` output (tests/golden).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

N_LAYER = 4
N_HEAD = 4
N_EMBD = 768
VOCAB = 50257
N_POS = 1024
ENTRY_LENGTH = 30  # decap.py:125
EOT_CLIP = 49407   # CLIP-BPE <|endoftext|>, where the detokenised caption is cut (decap.py:171)


def make_weights(seed: int = 1234, prefix_size: int = 768, init_std: float = 0.02) -> Dict[str, torch.Tensor]:
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


def gelu_new(x: torch.Tensor) -> torch.Tensor:
    """transformers/activations.py NewGELUActivation."""
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * torch.pow(x, 3.0))))


def gpt2_hidden(w, emb: torch.Tensor, kv: Optional[List[Tuple[torch.Tensor, torch.Tensor]]] = None, pos0: int = 0):
    """GPT2Model on ``inputs_embeds`` [R,T,768] at positions pos0..pos0+T; optional KV cache
    (list of per-layer (k,v) [R,H,t,hd], extended in place).  Returns ln_f(hidden) [R,T,768]."""
    Tp = "decoder.transformer."
    R, T, D = emb.shape
    hd = D // N_HEAD
    x = emb + w[Tp + "wpe.weight"][pos0:pos0 + T]
    for i in range(N_LAYER):
        p = f"{Tp}h.{i}."
        h = F.layer_norm(x, (D,), w[p + "ln_1.weight"], w[p + "ln_1.bias"], eps=1e-5)
        qkv = h @ w[p + "attn.c_attn.weight"] + w[p + "attn.c_attn.bias"]
        q, k, v = qkv.split(D, dim=-1)
        q = q.reshape(R, T, N_HEAD, hd).transpose(1, 2)
        k = k.reshape(R, T, N_HEAD, hd).transpose(1, 2)
        v = v.reshape(R, T, N_HEAD, hd).transpose(1, 2)
        if kv is not None:
            if kv[i] is not None:
                k = torch.cat([kv[i][0], k], dim=2)
                v = torch.cat([kv[i][1], v], dim=2)
            kv[i] = (k, v)
        S = k.shape[2]
        att = (q @ k.transpose(-2, -1)) * hd ** -0.5
        # causal: query at absolute position pos0+a sees keys 0..pos0+a
        qpos = torch.arange(pos0, pos0 + T)[:, None]
        kpos = torch.arange(S)[None, :]
        att = att.masked_fill(kpos > qpos, float("-inf")).softmax(dim=-1)
        o = (att @ v).transpose(1, 2).reshape(R, T, D)
        x = x + (o @ w[p + "attn.c_proj.weight"] + w[p + "attn.c_proj.bias"])
        h = F.layer_norm(x, (D,), w[p + "ln_2.weight"], w[p + "ln_2.bias"], eps=1e-5)
        h = gelu_new(h @ w[p + "mlp.c_fc.weight"] + w[p + "mlp.c_fc.bias"])
        x = x + (h @ w[p + "mlp.c_proj.weight"] + w[p + "mlp.c_proj.bias"])
    return F.layer_norm(x, (D,), w[Tp + "ln_f.weight"], w[Tp + "ln_f.bias"], eps=1e-5)


def prefix_embed(w, feats: torch.Tensor) -> torch.Tensor:
    """decap.py:124: clip_project = one Linear (MLP with two sizes has no activation, :51-58)."""
    return F.linear(feats.float(), w["clip_project.model.0.weight"], w["clip_project.model.0.bias"])


@torch.no_grad()
def decode_greedy(w, feats: torch.Tensor, compute_scores: bool = False, use_cache: bool = True,
                  steps: int = ENTRY_LENGTH, return_margin: bool = False):
    """decap.py:116-160.  30 fixed steps, no EOS early exit; next token = argmax of the softmax
    PROBABILITIES of the last position (first index wins ties, :136,141); its wte row is appended.

    ``use_cache=False`` re-runs the whole growing sequence every step exactly like the reference
    (465 token-positions per region, lm-head applied to the last one only here -- the other
    positions' logits are discarded by the reference, :133); ``use_cache=True`` is the same
    arithmetic with a KV cache.  Returns tokens [R,30] int64 (and scores [R] = exp(sum log p)).
    ``return_margin`` (instead of scores) also returns, per step, the gap between the best and the second-best LOGIT [R,steps]
    and the standard deviation of the logits over the vocabulary [R,steps] -- tests use them to tell a near-tie that reduced
    precision may flip from a real mismatch."""
    wte = w["decoder.transformer.wte.weight"]
    R = feats.shape[0]
    emb = prefix_embed(w, feats).reshape(R, 1, -1)
    tokens = []
    logps = []
    margins, spreads = [], []
    kv = [None] * N_LAYER if use_cache else None
    seq = emb
    for t in range(steps):
        if use_cache:
            h = gpt2_hidden(w, seq[:, -1:], kv, pos0=t)[:, -1]
        else:
            h = gpt2_hidden(w, seq)[:, -1]
        logits = h @ wte.T
        probs = F.softmax(logits, -1)
        nxt = torch.argmax(probs, -1)
        if return_margin:
            top2 = logits.topk(2, dim=-1).values
            margins.append(top2[:, 0] - top2[:, 1])
            spreads.append(logits.std(dim=-1))
        if compute_scores:
            logps.append(torch.log(probs).gather(1, nxt[:, None])[:, 0])
        tokens.append(nxt)
        seq = torch.cat([seq, wte[nxt][:, None, :]], dim=1)
    tokens = torch.stack(tokens, dim=1)
    if return_margin:
        return tokens, torch.stack(margins, dim=1), torch.stack(spreads, dim=1)
    if compute_scores:
        return tokens, torch.exp(torch.stack(logps, dim=1).sum(dim=-1))
    return tokens


def cut_at_eot(ids: List[int]) -> List[int]:
    """decap.py:171: the detokenised string is cut at the first CLIP <|endoftext|> (id 49407)."""
    return ids[: ids.index(EOT_CLIP)] if EOT_CLIP in ids else list(ids)


FLOPS_PER_TOKEN = 2 * (N_LAYER * 12 * N_EMBD * N_EMBD + VOCAB * N_EMBD)  # 133.8 MFLOP
