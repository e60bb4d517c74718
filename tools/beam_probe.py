"""Time ViECap's searches on the device: greedy (pio_decode_greedy_prompt) vs beam search (pio_decode_beam_prompt) of R prompts.
python tools/beam_probe.py [R] [P] [layers] [steps] [W]   (seeded random GPT-2 weights biased to stop: oracle.viecap.stopping_weights)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import viecap as ov  # noqa: E402  (weights only: the searches run in libpio_sm100)
from patchioner_b200 import ops  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 64
P = int(sys.argv[2]) if len(sys.argv) > 2 else 20
layers = int(sys.argv[3]) if len(sys.argv) > 3 else 12
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 64
W = int(sys.argv[5]) if len(sys.argv) > 5 else 5
dev = torch.device("cuda:0")
tok = ov.ToyTokenizer()
eos = [tok.encode(e)[-1] for e in (".", " .")]
w = ov.stopping_weights(ov.make_weights(n_layer_gpt=layers), eos, start=P + 8)
g = torch.Generator().manual_seed(1)
prompts = (torch.randn(R, P, 768, generator=g) * 0.3).to(dev)
ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
for mode in ("bf16", "fp32"):
    dec = ops.Gpt2Decoder(w, dev, mode)
    for name, fn in (("greedy", lambda: dec.decode(prompts, steps)), ("beam", lambda: dec.beam_search(prompts, eos, W, steps))):
        fn()
        torch.cuda.synchronize()
        e0, e1 = ev(), ev()
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        extra = ""
        if name == "beam":
            extra = f"  steps run {dec.beam_steps_run}, best-beam lengths {out[1][:, 0].float().mean().item():.1f} mean / {int(out[1][:, 0].max())} max"
        print(f"{mode} {name:6s} R={R} P={P} L={layers} steps={steps} W={W}: {e0.elapsed_time(e1):8.2f} ms{extra}", flush=True)
