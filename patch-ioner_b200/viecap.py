"""ViECap captioner on the pooled region embeddings (BASELINE config 4's second model).

Mirror of ``VieCap`` (Patch-ioner/src/viecap/entrypoint.py:15-162) for the configuration the reference ships
(``configs/mlp.viecap.k.yaml``: GPT-2 language model, greedy search; beam search -- the reference's default -- when
``using_greedy_search`` is false):

    feats --L2 normalise in place (:108)--> mapping network (ClipCap.py:122-153)     -> 10 soft-prompt embeddings
          --softmax(q.E^T / T), top-k >= threshold (retrieval_categories.py:87-115)  -> entity names
          --"There are a, b in image." (utils.py:55-74), GPT-2 BPE, right padded     -> hard-prompt embeddings (wte)
    [soft | hard] prompt --GPT-2, KV-cached greedy, 64 tokens (search.py:108-191)    -> ids -> cut after the first '.'

Everything numeric runs in libpio_sm100 (`pio_mapper_forward`, `pio_entity_topk`, `pio_gather_rows`,
`pio_decode_greedy_prompt`); the host composes the prompt strings and detokenises.  The entity retrieval that the
reference forces onto the CPU (retrieval_categories.py:87-88) stays on the device.  There is no CPU fallback.

No network in this image, so besides the reference's path-based config (`weight_path`, `files_path`, `language_model`)
the config may carry the objects directly: `state_dict`, `entities_text`, `texts_embeddings`, `tokenizer`.
"""
from __future__ import annotations

import json
import os
import pickle
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib as L
from . import ops

DEFAULTS = {  # entrypoint.py:60-79
    "language_model": "gpt2", "continuous_prompt_length": 10, "clip_project_length": 10, "temperature": 0.01, "top_k": 3,
    "threshold": 0.2, "disable_all_entities": False, "name_of_entities_text": "vinvl_vgoi_entities", "prompt_ensemble": False,
    "weight_path": None, "files_path": None, "using_hard_prompt": False, "soft_prompt_first": False, "only_hard_prompt": False,
    "using_greedy_search": False, "beam_width": 5, "text_prompt": None,
}
MAX_LEN = 64                       # search.py:113
END_OF_SENTENCES = (".", " .")     # search.py:114

_ENTITY_FILES = {  # entrypoint.py:188-216: (vocabulary file, embedding file stem)
    "coco_entities": ("coco_categories.json", "coco_embeddings"),
    "vinvl_vgoi_entities": ("vgcocooiobjects_v1_class2ind.json", "vgoi_embeddings"),
    "vinvl_vg_entities": ("VG-SGG-dicts-vgoi6-clipped.json", "vg_embeddings"),
}


def compose_discrete_prompt(entities: Sequence[str]) -> str:
    """utils.py:55-74."""
    if len(entities) == 0:
        return "There are something in image."
    return "There are" + ",".join(" " + e for e in entities) + " in image."


class VieCap:
    def __init__(self, args: Dict, device, clip_name: Optional[str] = None, precision: str = "fp32"):
        cfg = dict(DEFAULTS)
        cfg.update(args)
        self.args = cfg
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise L.PioError("patchioner_b200 runs on CUDA (sm_100a) only: there is no CPU fallback")
        if "gpt" not in cfg["language_model"]:
            raise NotImplementedError("only the GPT-2 language model is built (opt_search, search.py:16-105, is not)")
        self.beam_width = int(cfg["beam_width"])
        if not cfg["using_greedy_search"] and not 1 <= self.beam_width <= 8:
            raise ValueError(f"beam_width {self.beam_width}: the device beam search is built for 1..8 beams")
        self.clip_hidden_size = cfg.get("clip_hidden_size") or (640 if "RN" in (clip_name or "") else 512)  # entrypoint.py:24-29

        sd = cfg.get("state_dict")
        if sd is None:
            if not cfg["weight_path"]:
                raise ValueError("viecap: weight_path (or state_dict) is required")
            sd = torch.load(cfg["weight_path"], map_location="cpu", weights_only=False)
        self.mapper = ops.Mapper(sd, self.device, precision)
        if self.mapper.clip_size != self.clip_hidden_size:
            raise ValueError(f"clip_hidden_size {self.clip_hidden_size} != mapping network input {self.mapper.clip_size}")
        if self.mapper.prefix_len != cfg["continuous_prompt_length"]:
            raise ValueError("continuous_prompt_length does not match the checkpoint's prefix_const")
        if not any(k.startswith("gpt.transformer.") for k in sd):
            # ClipCaptionPrefix checkpoints may omit the frozen GPT-2: take it from `gpt2_state_dict`
            gsd = cfg.get("gpt2_state_dict")
            if gsd is None:
                raise ValueError("the checkpoint holds no 'gpt.transformer.*' weights: pass gpt2_state_dict "
                                 "(GPT2LMHeadModel.from_pretrained needs the network)")
            sd = {("gpt." + k if not k.startswith("gpt.") else k): v for k, v in gsd.items()}
        self.gpt = ops.Gpt2Decoder(sd, self.device, precision)

        self.tokenizer = cfg.get("tokenizer")
        if self.tokenizer is None:
            from transformers import AutoTokenizer  # entrypoint.py:39 (needs the files of `language_model` on disk)
            self.tokenizer = AutoTokenizer.from_pretrained(cfg["language_model"])
        pad = getattr(self.tokenizer, "pad_token_id", None)
        self.pad_id = pad if pad is not None else 0                                    # entrypoint.py:105
        self.eos = [self.tokenizer.encode(e)[-1] for e in END_OF_SENTENCES]            # search.py:135

        self.entities_text: Optional[List[str]] = None
        self.texts_embeddings = None
        if cfg["using_hard_prompt"]:
            ents, emb = cfg.get("entities_text"), cfg.get("texts_embeddings")
            if ents is None or emb is None:
                ents, emb = self._load_entities(cfg, cfg.get("suffix") or clip_name or "")
            self.entities_text = list(ents)
            emb = torch.as_tensor(emb).to(self.device, torch.float32)
            self.texts_embeddings = (emb / emb.norm(dim=-1, keepdim=True)).contiguous()  # retrieval_categories.py:90
            # A word is tokenised with its leading blank, so the prompt's tokens are the concatenation of fixed pieces:
            # tokenise every entity once instead of every prompt of every region.
            enc = self.tokenizer.encode
            self._head, self._tail, self._comma = enc("There are"), enc(" in image."), enc(",")
            self._something = enc("There are something in image.")
            self._entity_tokens = [enc(" " + e) for e in self.entities_text]
            probe = self.entities_text[:2]
            if enc(compose_discrete_prompt(probe)) != self._compose_tokens(list(range(len(probe)))):
                self._entity_tokens = None  # a tokenizer that merges across word boundaries: tokenise whole prompts

    @staticmethod
    def _load_entities(cfg, suffix: str):
        """entrypoint.py:178-221 for the JSON-list vocabularies (coco / vinvl)."""
        name = cfg["name_of_entities_text"]
        if name not in _ENTITY_FILES:
            raise NotImplementedError(f"name_of_entities_text={name!r}: pass entities_text / texts_embeddings in the config")
        vocab, stem = _ENTITY_FILES[name]
        directory = os.path.join(cfg["files_path"] or "", "annotations/vocabulary")
        with open(os.path.join(directory, vocab)) as f:
            ents = json.load(f)
        if isinstance(ents, dict):
            ents = list(ents.keys())
        keep_all = not cfg["disable_all_entities"]
        ents = sorted(e.lower().strip() for e in ents if keep_all or len(e.split()) == 1)  # load_annotations.py:91-103
        suffix = suffix.replace("/", "")
        fname = f"{stem}_{suffix}{'_with_ensemble' if cfg['prompt_ensemble'] else ''}.pickle"
        with open(os.path.join(directory, fname), "rb") as f:
            emb = pickle.load(f)
        return ents, emb

    def _compose_tokens(self, entity_rows: Sequence[int]) -> List[int]:
        if len(entity_rows) == 0:
            return list(self._something)
        out = list(self._head)
        for j, e in enumerate(entity_rows):
            if j:
                out += self._comma
            out += self._entity_tokens[e]
        return out + self._tail

    def detect_entities(self, feats: torch.Tensor) -> List[List[int]]:
        """Rows of ``entities_text`` kept per region: top-k by probability, stop at the first one below the threshold
        (retrieval_categories.py:97-115).  ``feats`` must already be unit rows."""
        k = min(int(self.args["top_k"]), len(self.entities_text))
        prob, idx = ops.entity_topk(feats, self.texts_embeddings, self.args["temperature"], k)
        prob, idx = prob.cpu().tolist(), idx.cpu().tolist()  # R x k numbers: the strings are composed on the host
        thr = float(self.args["threshold"])
        out = []
        for p, i in zip(prob, idx):
            cur = []
            for pj, ij in zip(p, i):
                if pj < thr:
                    break
                cur.append(ij)
            out.append(cur)
        return out

    def hard_prompt_token_lists(self, feats: torch.Tensor) -> List[List[int]]:
        """Token ids of every region's hard prompt, unpadded (entrypoint.py:117-121).  ``feats`` must be unit rows."""
        rows = self.detect_entities(feats)
        if self._entity_tokens is not None:
            return [self._compose_tokens(r) for r in rows]
        return [self.tokenizer.encode(compose_discrete_prompt([self.entities_text[i] for i in r])) for r in rows]

    def _pad(self, toks: Sequence[Sequence[int]], length: Optional[int] = None) -> torch.Tensor:
        lmax = length if length is not None else max(len(t) for t in toks)
        flat = torch.full((len(toks), lmax), self.pad_id, dtype=torch.int32)
        for i, t in enumerate(toks):
            flat[i, :len(t)] = torch.tensor(t, dtype=torch.int32)
        return flat.to(self.device, non_blocking=True)

    def hard_prompt_tokens(self, feats: torch.Tensor) -> torch.Tensor:
        """int32 [R,Lmax] on the device, right padded with pad_id over the whole call (entrypoint.py:117-126)."""
        return self._pad(self.hard_prompt_token_lists(feats))

    def _normalised(self, image_features: torch.Tensor) -> torch.Tensor:
        """fp32 unit rows on the device; normalises IN PLACE when the argument already is such a tensor (entrypoint.py:108)."""
        if image_features.device != self.device or image_features.dtype != torch.float32 or not image_features.is_contiguous():
            image_features = image_features.to(self.device, torch.float32).contiguous()
        return ops.l2_normalize_(image_features)

    def _join(self, cont: torch.Tensor, hard: Optional[torch.Tensor]) -> torch.Tensor:
        if hard is None:
            return cont
        R, Lh = hard.shape
        disc = ops.gather_rows(self.gpt.wte, hard.reshape(-1)).reshape(R, Lh, 768)       # word_embed (ClipCap.py:196-201)
        if self.args["only_hard_prompt"]:
            return disc
        return torch.cat((cont, disc), 1) if self.args["soft_prompt_first"] else torch.cat((disc, cont), 1)

    @torch.no_grad()
    def prompt_embeddings(self, image_features: torch.Tensor) -> torch.Tensor:
        """[R,P,768] fp32 input embeddings of the language model for ONE call of the reference (entrypoint.py:108-136)."""
        feats = self._normalised(image_features)
        cont = self.mapper.forward(feats)
        return self._join(cont, self.hard_prompt_tokens(feats) if self.args["using_hard_prompt"] else None)

    @torch.no_grad()
    def forward_ids(self, image_features: torch.Tensor, chunk: int = 4096, pad_group: Optional[int] = None) -> torch.Tensor:
        """int32 [R,64] generated ids on the device (before the sentence cut).

        The reference right-pads the hard prompts of ONE ``forward`` call to their longest and attends to the padding
        (entrypoint.py:126, no attention mask), so a caption depends on which regions share its call.  ``pad_group`` = rows
        per reference call (``Patchioner.forward`` captions boxes in calls of ``bs * bs_factor`` regions, model.py:981-1013;
        None = one call for everything).  Regions are decoded together across calls whenever their calls pad to the same
        length, in chunks of ``chunk`` rows that only bound the workspace (3.5 MB of KV cache per region, bf16)."""
        feats = image_features.reshape(-1, image_features.shape[-1])
        R = feats.shape[0]
        out = torch.empty(R, MAX_LEN, dtype=torch.int32, device=self.device)
        if R == 0:
            return out
        feats = self._normalised(feats)
        cont = torch.cat([self.mapper.forward(feats[s:s + chunk]) for s in range(0, R, chunk)], 0)
        if not self.args["using_greedy_search"]:
            chunk = max(1, chunk // self.beam_width)  # the cache holds beam_width rows per region
        if not self.args["using_hard_prompt"]:
            for s in range(0, R, chunk):
                out[s:s + chunk] = self._search(cont[s:s + chunk])
            return out
        toks = self.hard_prompt_token_lists(feats)
        G = R if not pad_group else int(pad_group)
        buckets: Dict[int, List[int]] = {}
        for g0 in range(0, R, G):
            rows = range(g0, min(g0 + G, R))
            buckets.setdefault(max(len(toks[i]) for i in rows), []).extend(rows)
        for length, rows in sorted(buckets.items()):
            for s in range(0, len(rows), chunk):
                part = rows[s:s + chunk]
                hard = self._pad([toks[i] for i in part], length)
                whole = len(part) == R
                idx = None if whole else torch.tensor(part, dtype=torch.long, device=self.device)
                ids = self._search(self._join(cont if whole else cont[idx], hard))
                if whole:
                    out = ids
                else:
                    out[idx] = ids
        return out

    def _search(self, prompt: torch.Tensor) -> torch.Tensor:
        """int32 [R,64] ids of the chosen sentence per prompt.  Greedy search (search.py:108-191) when the config says so, else the
        reference's default (entrypoint.py:77,139-143): beam search of ``beam_width`` beams, best beam returned.  The reference
        keeps ``tokens[:length]`` of that beam (search.py:280); here the positions from ``length`` on are overwritten with the
        end-of-sentence id, so that ``cut`` (the same rule greedy search uses) yields exactly those ``length`` tokens."""
        if self.args["using_greedy_search"]:
            return self.gpt.decode(prompt, MAX_LEN, eos=self.eos)  # stops once every row has ended; `cut` keeps the same tokens
        ids, lens, _ = self.gpt.beam_search(prompt, self.eos, self.beam_width, MAX_LEN)
        best, n = ids[:, 0], lens[:, 0:1]
        pos = torch.arange(MAX_LEN, device=best.device, dtype=torch.int32)[None, :]
        return torch.where(pos < n, best, torch.full_like(best, int(self.eos[0])))

    def cut(self, ids: Sequence[int]) -> List[int]:
        """search.py:184-190: keep up to and including the first end-of-sentence token."""
        for i, t in enumerate(ids):
            if t in self.eos:
                return list(ids[:i + 1])
        return list(ids)

    @torch.no_grad()
    def forward(self, image_features: torch.Tensor, compute_scores: bool = False, pad_group: Optional[int] = None):
        """entrypoint.py:98-162: list of sentences.  Note: a batch of ONE region stops at the first '.', which gives the
        same sentence as cutting afterwards (search.py:173-176 vs :184-190).  ``pad_group``: see ``forward_ids``."""
        ids = self.forward_ids(image_features, pad_group=pad_group).cpu().tolist()
        sentences = [self.tokenizer.decode(self.cut(r)) for r in ids]
        if compute_scores:
            return sentences, self.compute_perplexity(sentences)
        return sentences

    @torch.no_grad()
    def compute_perplexity(self, sentences: Sequence[str], chunk: int = 1024) -> List[float]:
        """entrypoint.py:164-177: exp of the language model's causal loss of every (re-tokenised) sentence.  The reference
        runs one sentence at a time in a Python loop; here all sentences go through one batched pass (right padded: with
        causal attention the padding cannot influence the scored positions)."""
        toks = [self.tokenizer.encode(s) for s in sentences]
        out: List[float] = []
        for s in range(0, len(toks), chunk):
            part = toks[s:s + chunk]
            n = max(1, max(len(t) for t in part))
            if n > 128:
                raise ValueError(f"a sentence of {n} tokens exceeds the 128 positions the scorer is built for")
            ids = torch.zeros(len(part), n, dtype=torch.int32)
            for i, t in enumerate(part):
                ids[i, :len(t)] = torch.tensor(t, dtype=torch.int32)
            lens = torch.tensor([len(t) for t in part], dtype=torch.int32)
            nll = self.gpt.score_tokens(ids.to(self.device), lens.to(self.device))
            out += torch.exp(nll).cpu().tolist()
        return out

    __call__ = forward
