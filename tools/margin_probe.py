"""Where do bf16 decoder captions leave the fp32 oracle's?  Top-2 logit gap (in units of the logits' standard deviation) at every
first diverging position, against the distribution of that gap over all positions.  python tools/margin_probe.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import decap as o_decap
from patchioner_b200 import ops
dev = torch.device('cuda:0')
w = o_decap.make_weights(seed=1234)
R = 256
feats = torch.randn(R, 768, generator=torch.Generator().manual_seed(32))
feats = feats / feats.norm(dim=-1, keepdim=True)
ref, margin, spread = o_decap.decode_greedy(w, feats, use_cache=True, return_margin=True)
for mode in ("bf16", "fp32"):
    ids = ops.Decoder(w, dev, mode).decode(feats.to(dev), 30).cpu().long()
    div = []
    for r in range(R):
        ne = (ids[r] != ref[r]).nonzero()
        if len(ne):
            t = int(ne[0])
            div.append((margin[r, t] / spread[r, t]).item())
    rel = (margin / spread).flatten()
    print(mode, "rows differing", len(div), "of", R, "| margin/spread at first divergence: max %.4g median %.4g" % (max(div) if div else 0, sorted(div)[len(div)//2] if div else 0),
          "| all positions: median %.4g, 5%% quantile %.4g, 1%% quantile %.4g" % (rel.median().item(), rel.quantile(0.05).item(), rel.quantile(0.01).item()))
