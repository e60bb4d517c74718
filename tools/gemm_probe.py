"""Time the dense layer (pio_linear) alone for the hot shapes.  Usage: python tools/gemm_probe.py [mode] [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from patchioner_b200 import _lib as L, ops  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "bf16"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dt = torch.bfloat16 if mode == "bf16" else torch.float32
dev = torch.device("cuda:0")
shapes = [("vit qkv", 87936, 2304, 768, "bias", dt), ("vit proj+res", 87936, 768, 768, "res", torch.float32),
          ("vit fc1+gelu", 87936, 3072, 768, "gelu", dt), ("vit fc2+res", 87936, 768, 3072, "res", torch.float32),
          ("plain bf16 out", 87936, 3072, 768, "none", dt), ("plain f32 out", 87936, 3072, 768, "none", torch.float32),
          ("lm_head", 4096, 50257, 768, "none", torch.float32), ("dec qkv", 4096, 2304, 768, "bias", dt),
          ("square 8192", 8192, 8192, 8192, "none", dt),
          ("dec proj+res", 4096, 768, 768, "res", torch.float32), ("dec fc+gelu", 4096, 3072, 768, "gelu", dt),
          ("dec fc2+res", 4096, 768, 3072, "res", torch.float32), ("proj O", 4096, 768, 16384, "res", torch.float32)]
if os.environ.get("PIO_PROBE_ONLY"):
    shapes = [s for s in shapes if s[0].startswith(os.environ["PIO_PROBE_ONLY"])]
for name, M, N, K, epi, odt in shapes:
    A = torch.randn(M, K, device=dev).to(dt)
    W = (torch.randn(N, K, device=dev) / K ** 0.5).to(dt)
    bias = torch.randn(N, device=dev)
    ld = (N + 7) // 8 * 8
    C = torch.zeros(M, ld, device=dev, dtype=odt)[:, :N]
    kw = {}
    if epi in ("bias", "gelu", "res"):
        kw["bias"] = bias
    if epi == "gelu":
        kw["act"] = L.ACT_GELU_ERF
    if epi == "res":
        kw["gamma"] = bias
        kw["residual"] = C
    for _ in range(2):
        ops.linear(A, W, mode, out=C, **kw)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        ops.linear(A, W, mode, out=C, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:16s} M={M:6d} N={N:6d} K={K:5d} epi={epi:5s} out={str(odt)[6:]:9s} {ms:8.3f} ms  {2.0 * M * N * K / ms / 1e9:8.1f} TFLOP/s", flush=True)
