"""Which columns beyond N does the TMA-store epilogue touch?  (debug aid)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from patchioner_b200 import ops
dev = torch.device("cuda:0")
for N in (5003, 5004, 5000, 4999):
    M, K = 513, 768
    A = torch.randn(M, K, device=dev).bfloat16(); W = torch.randn(N, K, device=dev).bfloat16()
    for dt in (torch.bfloat16, torch.float32):
        ld = (N + 63) // 64 * 64 + 64
        buf = torch.full((M, ld), 7.0, dtype=dt, device=dev)
        ops.linear(A, W, "bf16", out=buf[:, :N])
        torch.cuda.synchronize()
        bad = (buf[:, N:] != 7.0).any(dim=0).nonzero().flatten().tolist()
        badrows = (buf[:, N:] != 7.0).any(dim=1).sum().item()
        print(N, dt, "touched columns beyond N:", [N + b for b in bad][:20], "rows:", badrows, flush=True)
