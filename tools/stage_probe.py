"""Run one stage of the hot path alone (for ncu launch lists).  python tools/stage_probe.py vit|text|all [B] [S] [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from patchioner_b200 import ops, synth  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "vit"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
S = int(sys.argv[3]) if len(sys.argv) > 3 else 518
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
prec = os.environ.get("PIO_PRECISION", "bf16")
dev = torch.device("cuda:0")
ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
if what in ("vit", "all"):
    vit = ops.Vit(synth.make_vit_weights(1234), dev, prec)
    imgs = synth.synth_images(B, S, seed=1).to(dev)
    for i in range(reps):
        e0, e1 = ev(), ev()
        e0.record()
        tokens, attn, _ = vit.forward(imgs)
        e1.record()
        torch.cuda.synchronize()
        print(f"vit forward B={B} S={S}: {e0.elapsed_time(e1):.2f} ms", flush=True)
if what in ("text", "all"):
    R = B * 64
    dec = ops.Decoder(synth.make_decoder_weights(1234), dev, prec)
    bank = ops.Bank(synth.synth_bank(int(os.environ.get("PIO_BANK", "591753")), 768, seed=7), dev, prec)
    q = torch.randn(R, 768, device=dev)
    for i in range(reps):
        e0, e1, e2 = ev(), ev(), ev()
        e0.record()
        pre = bank.project(q, normalize=True)
        e1.record()
        ids = dec.decode(pre, int(os.environ.get("PIO_STEPS", "30")))
        e2.record()
        torch.cuda.synchronize()
        print(f"project R={R}: {e0.elapsed_time(e1):.2f} ms   decode: {e1.elapsed_time(e2):.2f} ms", flush=True)
