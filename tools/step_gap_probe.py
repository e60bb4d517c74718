"""Is the steady-state bench step (model(...) in a loop) longer than the sum of its stages?  Times, without any host sync inside the
loop, (a) N calls of model(...) and (b) N iterations of the four stage calls of bench.stage_breakdown, each with per-stage events."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from patchioner_b200 import Patchioner, ops, synth  # noqa: E402

dev = torch.device("cuda:0")
B, R, S, N = 64, 64, 518, 12
cfg = {"decap_weights": synth.make_decoder_weights(1234), "prefix_size": 768, "support_memory_size": 591753,
       "dino_model": "dinov2_vitb14_reg", "normalize": True, "resize_dim": S, "crop_dim": S, "dino_weights": synth.make_vit_weights(1234),
       "memory_bank": synth.synth_bank(591753, 768, seed=7), "precision": "bf16"}
model = Patchioner.from_config(cfg, device=dev)
imgs = synth.synth_images(B, S, seed=1).to(dev)
boxes = synth.synth_boxes(B, R, S, seed=1).to(dev)
kw = dict(get_cls_capt=False, bboxes=boxes, gaussian_avg=True, gaussian_bbox_variance=1.0, return_ids=True)
ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
for _ in range(3):
    model(imgs, **kw)
torch.cuda.synchronize()
marks = [ev() for _ in range(N + 1)]
t0 = time.perf_counter()
marks[0].record()
for i in range(N):
    model(imgs, **kw)
    marks[i + 1].record()
t_issue = time.perf_counter() - t0
torch.cuda.synchronize()
print("model(...) per step ms:", " ".join(f"{marks[i].elapsed_time(marks[i + 1]):.1f}" for i in range(N)), f"| host issue time per step {t_issue / N * 1e3:.1f} ms")
P = (S // 14) ** 2
m = [[ev() for _ in range(5)] for _ in range(N)]
torch.cuda.synchronize()
for i in range(N):
    m[i][0].record()
    tokens, attn, _ = model.dino.forward(imgs, want_attn=True)
    m[i][1].record()
    feats = ops.pool_boxes(tokens[:, 5:], boxes, 14, True, 1.0, None)
    m[i][2].record()
    pre = model.embed_tokens(feats.reshape(-1, 768))
    m[i][3].record()
    model.decoder.decode(pre, 30)
    m[i][4].record()
torch.cuda.synchronize()
for i in range(N):
    st = [m[i][k].elapsed_time(m[i][k + 1]) for k in range(4)]
    gap = m[i - 1][4].elapsed_time(m[i][0]) if i else 0.0
    print(f"stages it {i}: vit {st[0]:.2f} pool {st[1]:.2f} project {st[2]:.2f} decode {st[3]:.2f} sum {sum(st):.2f} gap-before {gap:.2f}")
print(f"stage loop per iteration: {m[0][0].elapsed_time(m[N - 1][4]) / N:.2f} ms")
