"""One GEMM shape, a few launches (for ncu).  python tools/gemm_one.py M N K [epi]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from patchioner_b200 import _lib as L, ops  # noqa: E402

M, N, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
epi = sys.argv[4] if len(sys.argv) > 4 else "bias"
dev = torch.device("cuda:0")
A = torch.randn(M, K, device=dev).bfloat16()
W = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
bias = torch.randn(N, device=dev)
C = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
for _ in range(4):
    ops.linear(A, W, "bf16", bias=bias, out=C, act=L.ACT_GELU_ERF if epi == "gelu" else L.ACT_NONE)
torch.cuda.synchronize()
