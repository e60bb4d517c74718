"""Run-to-run reproducibility of the bf16 ViT forward (the probe that found the programmatic-dependent-launch race of round 2:
profiles/r02y_*, r02z_*, r02ab_*, r02ac_*).  TAG=<label> [PIO_PDL=0 | PIO_PDL_OFF=<kind mask>] python tools/determinism_probe.py"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import dinov2 as o_vit, pipeline as o_pipe
from patchioner_b200 import ops
dev = torch.device("cuda:0")
vit = ops.Vit(o_vit.make_weights(seed=1234), dev, "bf16")
for B, S in ((4, 224), (2, 224), (1, 518)):
    imgs = o_pipe.synth_images(B, S, seed=31).to(dev)
    ref = None
    bad = 0
    worst = 0.0
    for it in range(30):
        t, a, _ = vit.forward(imgs)
        t = t.clone()
        if ref is None:
            ref = t
        elif not torch.equal(ref, t):
            bad += 1
            worst = max(worst, (ref - t).abs().max().item())
    print(os.environ.get("TAG", ""), f"B={B} S={S}: {bad}/29 runs differ from the first, max |diff| {worst:.3g}")
