"""Host-side detokenisation hook (decap.py:162-181).

The reference decodes ids with the CLIP BPE vocabulary (src/clip/simple_tokenizer.py, 49408 entries,
vocabulary file ``bpe_simple_vocab_16e6.txt.gz`` -- a third-party asset that is not shipped here).  Set
``PIO_CLIP_BPE`` to that file to get real text; otherwise ids are rendered as ``"<id> <id> ..."`` with the
CLIP end-of-text id mapped to ``<|endoftext|>`` so that the caller's cut-at-EOT logic still applies.
Detokenisation is a 'next' row, not part of the measured path.
"""
from __future__ import annotations

import gzip
import os
from functools import lru_cache
from typing import Callable, List

EOT_ID, SOT_ID = 49407, 49406


def _bytes_to_unicode():
    bs = list(range(ord("!"), ord("~") + 1)) + list(range(ord("¡"), ord("¬") + 1)) + list(range(ord("®"), ord("ÿ") + 1))
    cs = bs[:]
    n = 0
    for b in range(2 ** 8):
        if b not in bs:
            bs.append(b)
            cs.append(2 ** 8 + n)
            n += 1
    return dict(zip(bs, [chr(c) for c in cs]))


class ClipBpeDecoder:
    """Decode-only CLIP BPE: rebuilds the id -> token table from the merges file."""

    def __init__(self, bpe_path: str):
        b2u = _bytes_to_unicode()
        self.byte_decoder = {v: k for k, v in b2u.items()}
        merges = gzip.open(bpe_path).read().decode("utf-8").split("\n")[1:49152 - 256 - 2 + 1]
        vocab = list(b2u.values())
        vocab = vocab + [v + "</w>" for v in vocab]
        vocab += ["".join(m.split()) for m in merges]
        vocab += ["<|startoftext|>", "<|endoftext|>"]
        self.decoder = dict(enumerate(vocab))

    def __call__(self, ids: List[int]) -> str:
        text = "".join(self.decoder[int(t)] for t in ids)  # KeyError for ids >= 49408, like the reference
        return bytearray(self.byte_decoder[c] for c in text).decode("utf-8", errors="replace").replace("</w>", " ")


def _id_renderer(ids: List[int]) -> str:
    return " ".join("<|endoftext|>" if int(t) == EOT_ID else ("<|startoftext|>" if int(t) == SOT_ID else str(int(t)))
                    for t in ids)


@lru_cache(maxsize=1)
def default_detokenizer() -> Callable[[List[int]], str]:
    path = os.environ.get("PIO_CLIP_BPE")
    if path and os.path.exists(path):
        dec = ClipBpeDecoder(path)

        def safe(ids):
            try:
                return dec(ids)
            except KeyError:  # random-init weights emit ids outside the CLIP vocabulary
                return _id_renderer(ids)

        return safe
    return _id_renderer
