"""Oracle (test infrastructure): the ViECap captioner that BASELINE config 4 puts behind the pooled embeddings.

Restates, functionally on the reference's state-dict key names,
  * ``VieCap.forward``            Patch-ioner/src/viecap/entrypoint.py:98-162 (in-place L2 normalise :108, soft prompt,
                                  hard prompt, right padding with ``pad_token_id or 0`` :105,126, no attention mask)
  * ``MappingNetwork`` and its    Patch-ioner/src/viecap/ClipCap.py:7-153 (Linear(clip -> project_len x 768) tokens +
    ``Transformer``               learnt ``prefix_const``; pre-LN layers, q / kv projections without bias, softmax over
                                  keys, ``project`` with bias, ReLU MLP with mlp_ratio 2; the last prefix_len tokens)
  * ``image_text_simiarlity`` /   Patch-ioner/src/viecap/retrieval_categories.py:60-115 (softmax(q.E^T / T) on unit rows,
    ``top_k_categories``          top-k, stop at the first probability below the threshold)
  * ``compose_discrete_prompts``  Patch-ioner/src/viecap/utils.py:55-74
  * ``greedy_search``             Patch-ioner/src/viecap/search.py:108-191 (arg-max on the LOGITS, 64 tokens, no early exit
                                  for batches > 1, sentence cut after the first '.' / ' .' token)
  * ``beam_search``               Patch-ioner/src/viecap/search.py:193-285 (the reference's default, entrypoint.py:77,139-143:
                                  one call per region, width 5, length-normalised scores, stopped beams frozen)
GPT-2 itself is third-party (``transformers``, 4.46.3 pinned by the reference, 5.5.0 here): the same arithmetic as
``oracle/decap.py`` with 12 heads.  Pinned by ``tests/golden/make_golden_viecap.py``, which loads the seeded weights
below into the UNMODIFIED reference classes (and ``transformers.GPT2LMHeadModel``) and stores their outputs in
``tests/golden/viecap.pt``; ``tests/test_viecap_cpu.py`` checks this file against them.
"""
from __future__ import annotations

import math
import re
import zlib
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

N_EMBD = 768
VOCAB = 50257
N_POS = 1024
MAX_LEN = 64  # search.py:113


class ToyTokenizer:
    """Stand-in for the GPT-2 BPE tokenizer (its vocabulary files are not in this image): GPT-2's pre-tokenisation regex
    on words / punctuation, one id per piece (crc32 mod 50000, remembered for ``decode``).  Like the real one, a word is
    tokenised together with its leading blank, so a sentence tokenises to the concatenation of its words' tokens."""
    pad_token_id = None
    _pat = re.compile(r" ?[A-Za-z]+| ?[0-9]+| ?[^\sA-Za-z0-9]+|\s+")

    def __init__(self):
        self.names: Dict[int, str] = {}

    def encode(self, text: str) -> List[int]:
        out = []
        for piece in self._pat.findall(text):
            i = zlib.crc32(piece.encode()) % 50000
            self.names.setdefault(i, piece)
            out.append(i)
        return out

    def decode(self, ids: Sequence[int]) -> str:
        return "".join(self.names.get(int(i), f"<{int(i)}>") for i in ids)


def make_weights(seed: int = 4321, n_layer_gpt: int = 2, n_layer_map: int = 2, clip_size: int = 768, project_len: int = 10,
                 prefix_len: int = 10, std: float = 0.02) -> Dict[str, torch.Tensor]:
    """Seed-fixed ViECap checkpoint (``ClipCaptionModel`` key names: ``mapping_network.*``, ``gpt.transformer.*``)."""
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
    w["gpt.lm_head.weight"] = w[T + "wte.weight"]  # tied
    return w


def _layers(w, prefix: str) -> int:
    return 1 + max(int(k[len(prefix):].split(".")[0]) for k in w if k.startswith(prefix))


def mapping_network(w, feats: torch.Tensor, n_head: int = 8) -> torch.Tensor:
    """ClipCap.py:122-153.  feats [R,clip] -> [R,prefix_len,768]."""
    M = "mapping_network."
    D = N_EMBD
    R = feats.shape[0]
    pc = w[M + "prefix_const"]
    x = F.linear(feats, w[M + "linear.weight"], w[M + "linear.bias"]).view(R, -1, D)  # :147
    clen = x.shape[1]
    x = torch.cat([x, pc.unsqueeze(0).expand(R, *pc.shape)], dim=1)                   # :148-149
    n, hd = x.shape[1], D // n_head
    LT = M + "transformer.layers."
    for i in range(_layers(w, LT)):
        p = f"{LT}{i}."
        h = F.layer_norm(x, (D,), w[p + "norm1.weight"], w[p + "norm1.bias"], eps=1e-5)
        q = F.linear(h, w[p + "attn.to_queries.weight"]).reshape(R, n, n_head, hd)            # :55 (bias=False, :104)
        kv = F.linear(h, w[p + "attn.to_keys_values.weight"]).reshape(R, n, 2, n_head, hd)    # :56
        k, v = kv[:, :, 0], kv[:, :, 1]
        att = torch.einsum("bnhd,bmhd->bnmh", q, k) * hd ** -0.5                              # :58
        att = att.softmax(dim=2)                                                              # :65
        o = torch.einsum("bnmh,bmhd->bnhd", att, v).reshape(R, n, D)
        x = x + F.linear(o, w[p + "attn.project.weight"], w[p + "attn.project.bias"])        # :67, :92-93
        h = F.layer_norm(x, (D,), w[p + "norm2.weight"], w[p + "norm2.bias"], eps=1e-5)
        h = F.relu(F.linear(h, w[p + "mlp.fc1.weight"], w[p + "mlp.fc1.bias"]))               # :24-26
        x = x + F.linear(h, w[p + "mlp.fc2.weight"], w[p + "mlp.fc2.bias"])                   # :94
    return x[:, clen:, :]                                                                     # :151


def entity_probs(feats: torch.Tensor, texts_embeddings: torch.Tensor, temperature: float) -> torch.Tensor:
    """retrieval_categories.py:87-94 (both sides re-normalised, fp32 on the CPU)."""
    q = feats.float() / feats.float().norm(dim=-1, keepdim=True)
    e = texts_embeddings.float() / texts_embeddings.float().norm(dim=-1, keepdim=True)
    return F.softmax(q @ e.T / temperature, dim=-1)


def pick_entities(entities_text: Sequence[str], probs: torch.Tensor, top_k: int, threshold: float) -> List[List[str]]:
    """retrieval_categories.py:97-115: top-k in descending order, stop at the first one below the threshold."""
    p, idx = torch.topk(probs, k=top_k, dim=-1)
    out = []
    for i in range(probs.shape[0]):
        cur = []
        for j in range(top_k):
            if p[i, j] < threshold:
                break
            cur.append(entities_text[int(idx[i, j])])
        out.append(cur)
    return out


def compose_prompt(entities: Sequence[str]) -> str:
    """utils.py:55-74."""
    if len(entities) == 0:
        return "There are something in image."
    return "There are" + ",".join(" " + e for e in entities) + " in image."


def gpt2_hidden(w, emb: torch.Tensor, n_head: int = 12, kv: Optional[list] = None, pos0: int = 0) -> torch.Tensor:
    """GPT2Model on inputs_embeds [R,T,768] at positions pos0..pos0+T-1 -> ln_f(hidden).  ``kv`` (a list with one
    entry per layer, extended in place) is the KV cache the reference's greedy_search uses (search.py:150-163)."""
    Tp = "gpt.transformer."
    R, T, D = emb.shape
    hd = D // n_head
    x = emb + w[Tp + "wpe.weight"][pos0:pos0 + T]
    mask = torch.arange(pos0 + T)[None, :] > torch.arange(pos0, pos0 + T)[:, None]  # key position > query position
    for i in range(_layers(w, Tp + "h.")):
        p = f"{Tp}h.{i}."
        h = F.layer_norm(x, (D,), w[p + "ln_1.weight"], w[p + "ln_1.bias"], eps=1e-5)
        qkv = h @ w[p + "attn.c_attn.weight"] + w[p + "attn.c_attn.bias"]
        q, k, v = (t.reshape(R, T, n_head, hd).transpose(1, 2) for t in qkv.split(D, dim=-1))
        if kv is not None:
            if kv[i] is not None:
                k, v = torch.cat([kv[i][0], k], dim=2), torch.cat([kv[i][1], v], dim=2)
            kv[i] = (k, v)
        att = ((q @ k.transpose(-2, -1)) * hd ** -0.5).masked_fill(mask, float("-inf")).softmax(dim=-1)
        o = (att @ v).transpose(1, 2).reshape(R, T, D)
        x = x + (o @ w[p + "attn.c_proj.weight"] + w[p + "attn.c_proj.bias"])
        h = F.layer_norm(x, (D,), w[p + "ln_2.weight"], w[p + "ln_2.bias"], eps=1e-5)
        h = h @ w[p + "mlp.c_fc.weight"] + w[p + "mlp.c_fc.bias"]
        h = 0.5 * h * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (h + 0.044715 * torch.pow(h, 3.0))))
        x = x + (h @ w[p + "mlp.c_proj.weight"] + w[p + "mlp.c_proj.bias"])
    return F.layer_norm(x, (D,), w[Tp + "ln_f.weight"], w[Tp + "ln_f.bias"], eps=1e-5)


@torch.no_grad()
def greedy_ids(w, prompt: torch.Tensor, steps: int = MAX_LEN, return_margin: bool = False, use_cache: bool = False):
    """search.py:146-171 up to the token ids: [R,steps] int64.  (The reference's step loop runs the model once more
    after the last token; that run has no effect on the tokens.)  ``return_margin`` also gives, per step, the gap
    between the best and second-best logit -- tests use it to tell a real mismatch from a near-tie."""
    wte = w["gpt.transformer.wte.weight"]
    seq = prompt.float()
    toks, margins = [], []
    kv = [None] * _layers(w, "gpt.transformer.h.") if use_cache else None
    for t in range(steps):
        if use_cache:  # prefill once, then one position per step like the reference (search.py:150-163)
            new = seq if t == 0 else seq[:, -1:]
            logits = gpt2_hidden(w, new, kv=kv, pos0=seq.shape[1] - new.shape[1])[:, -1] @ wte.T
        else:
            logits = gpt2_hidden(w, seq)[:, -1] @ wte.T
        top2 = logits.topk(2, dim=-1).values
        margins.append(top2[:, 0] - top2[:, 1])
        nxt = torch.argmax(logits, dim=-1)
        toks.append(nxt)
        seq = torch.cat([seq, wte[nxt][:, None, :]], dim=1)
    ids = torch.stack(toks, dim=1)
    return (ids, torch.stack(margins, dim=1)) if return_margin else ids


def stopping_weights(w: Dict[str, torch.Tensor], eos: Sequence[int], seed: int = 77, start: int = 24, slope: float = 0.35,
                     gain: float = 3.0) -> Dict[str, torch.Tensor]:
    """A copy of ``w`` whose language model drifts towards the end-of-sentence tokens as the position grows, so that beam search
    on seeded random weights actually STOPS (different beams at different steps): the position embeddings gain a component
    along a fixed unit direction u that grows linearly from position ``start`` on, and the tied embedding rows of the
    end-of-sentence tokens are moved along u.  Test scaffolding only (used by tests/golden/make_golden_viecap_beam.py and the
    tests that rebuild its inputs)."""
    g = torch.Generator().manual_seed(seed)
    u = torch.randn(N_EMBD, generator=g)
    u = u / u.norm()
    out = dict(w)
    wpe = w["gpt.transformer.wpe.weight"].clone()
    pos = torch.arange(N_POS, dtype=torch.float32)
    wpe += (slope * (pos - start).clamp(min=0))[:, None] * u[None, :]
    out["gpt.transformer.wpe.weight"] = wpe
    wte = w["gpt.transformer.wte.weight"].clone()
    for k, e in enumerate(eos):
        wte[int(e)] += (gain - 0.4 * k) * u
    out["gpt.transformer.wte.weight"] = wte
    if "gpt.lm_head.weight" in out:  # tied in GPT2LMHeadModel: a state dict that carries both must carry the same tensor
        out["gpt.lm_head.weight"] = wte
    return out


@torch.no_grad()
def beam_search_ids(w, prompt: torch.Tensor, eos: Sequence[int], beam_width: int = 5, steps: int = MAX_LEN,
                    temperature: float = 1.0):
    """search.py:193-285 for ONE prompt [1,P,768] (entrypoint.py:139-143 calls it region by region): returns
    (tokens [W, n] int64, seq_lengths [W] float, scores / seq_lengths [W]) in the reference's final beam order BEFORE its
    ``argsort`` -- the caller sorts like :281-283.  No KV cache (the reference re-runs the whole sequence per step); stopped
    beams keep their score (all of their mass on token 0, :250-251) and their length; the length-normalised scores pick the
    top ``beam_width`` of the ``W x vocab`` continuations (:252-257); the loop ends when every beam has emitted '.' / ' .'."""
    wte = w["gpt.transformer.wte.weight"]
    W = beam_width
    generated = prompt.float()
    scores = None
    tokens = None
    seq_lengths = torch.ones(W)
    is_stopped = torch.zeros(W, dtype=torch.bool)
    for _ in range(steps):
        logits = gpt2_hidden(w, generated)[:, -1] @ wte.T
        logits = logits / (temperature if temperature > 0 else 1.0)
        logits = logits.softmax(-1).log()
        if scores is None:
            scores, next_tokens = logits.topk(W, -1)
            generated = generated.expand(W, *generated.shape[1:])
            next_tokens, scores = next_tokens.permute(1, 0), scores.squeeze(0)
            tokens = next_tokens
        else:
            logits[is_stopped] = -float("inf")
            logits[is_stopped, 0] = 0
            scores_sum = scores[:, None] + logits
            seq_lengths[~is_stopped] += 1
            avg = scores_sum / seq_lengths[:, None]
            avg, next_tokens = avg.view(-1).topk(W, -1)
            src = torch.div(next_tokens, scores_sum.shape[1], rounding_mode="trunc")
            seq_lengths = seq_lengths[src]
            next_tokens = (next_tokens % scores_sum.shape[1]).unsqueeze(1)
            tokens = torch.cat((tokens[src], next_tokens), dim=1)
            generated = generated[src]
            scores = avg * seq_lengths
            is_stopped = is_stopped[src]
        generated = torch.cat((generated, wte[next_tokens.squeeze()].view(W, 1, -1)), dim=1)
        is_stopped = is_stopped + (next_tokens.eq(eos[0]) | next_tokens.eq(eos[1])).squeeze()
        if is_stopped.all():
            break
    return tokens, seq_lengths, scores / seq_lengths


def beam_sentences(tokens: torch.Tensor, seq_lengths: torch.Tensor, avg_scores: torch.Tensor) -> List[List[int]]:
    """search.py:279-283: every beam cut at its length, best length-normalised score first."""
    outs = [[int(t) for t in row[:int(n)]] for row, n in zip(tokens.tolist(), seq_lengths.tolist())]
    return [outs[i] for i in avg_scores.argsort(descending=True).tolist()]


@torch.no_grad()
def perplexity(w, token_ids: Sequence[int]) -> float:
    """entrypoint.py:164-177 for one sentence: exp of the causal-LM loss of ``GPT2LMHeadModel(input_ids, labels=input_ids)``
    (mean cross-entropy of token i+1 given tokens 0..i; NaN for fewer than two tokens)."""
    ids = torch.tensor([int(i) for i in token_ids], dtype=torch.long)
    wte = w["gpt.transformer.wte.weight"]
    logits = gpt2_hidden(w, wte[ids][None])[0] @ wte.T
    if ids.numel() < 2:
        return float("nan")
    return float(torch.exp(F.cross_entropy(logits[:-1], ids[1:])))


def cut_sentence(ids: Sequence[int], eos: Sequence[int]) -> List[int]:
    """search.py:184-190: keep tokens up to and including the first end-of-sentence token (all of them if none)."""
    ids = [int(i) for i in ids]
    for i, t in enumerate(ids):
        if t in eos:
            return ids[:i + 1]
    return ids


@torch.no_grad()
def viecap_forward(w, feats: torch.Tensor, entities_text: Sequence[str], texts_embeddings: torch.Tensor, tokenizer,
                   temperature: float = 0.01, top_k: int = 3, threshold: float = 0.4, using_hard_prompt: bool = True,
                   soft_prompt_first: bool = True, only_hard_prompt: bool = False, steps: int = MAX_LEN,
                   use_cache: bool = False):
    """entrypoint.py:98-147 with greedy search.  Returns (sentences, ids [R,steps], prompt embeddings [R,P,768],
    hard-prompt tokens [R,Lmax] or None).  ``feats`` is normalised IN PLACE like the reference (:108)."""
    pad_id = tokenizer.pad_token_id if tokenizer.pad_token_id is not None else 0
    feats /= feats.norm(2, dim=-1, keepdim=True)
    cont = mapping_network(w, feats)
    hard = None
    if using_hard_prompt:
        ents = pick_entities(entities_text, entity_probs(feats, texts_embeddings, temperature), top_k, threshold)
        toks = [torch.tensor(tokenizer.encode(compose_prompt(e))) for e in ents]
        hard = torch.nn.utils.rnn.pad_sequence(toks, batch_first=True, padding_value=pad_id)
        disc = w["gpt.transformer.wte.weight"][hard]
        emb = disc if only_hard_prompt else (torch.cat([cont, disc], 1) if soft_prompt_first else torch.cat([disc, cont], 1))
    else:
        emb = cont
    ids = greedy_ids(w, emb, steps, use_cache=use_cache)
    eos = [tokenizer.encode(e)[-1] for e in (".", " .")]
    sentences = [tokenizer.decode(cut_sentence(r, eos)) for r in ids.tolist()]
    return sentences, ids, emb, hard
