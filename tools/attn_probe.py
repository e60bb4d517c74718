"""Time ViT attention alone.  python tools/attn_probe.py [B] [N] [reps] [dtype]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from patchioner_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1374
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
dt = torch.bfloat16 if (len(sys.argv) <= 4 or sys.argv[4] == "bf16") else torch.float32
dev = torch.device("cuda:0")
qkv = torch.randn(B, N, 2304, device=dev).to(dt)
for _ in range(2):
    ops.vit_attention(qkv)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for _ in range(reps):
    ops.vit_attention(qkv)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"attention B={B} N={N} {dt}: {ms:.3f} ms  {4.0 * B * 12 * N * N * 64 / ms / 1e9:.1f} TFLOP/s", flush=True)
