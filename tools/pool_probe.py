"""Time the pooling kernels at BASELINE sizes with an L2 flush between runs."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from patchioner_b200 import ops, synth  # noqa: E402

dev = torch.device("cuda:0")
B, g, R, D = 64, 37, 64, 768
tokens = torch.randn(B, 5 + g * g, D, device=dev)
patch = tokens[:, 5:]
boxes = synth.synth_boxes(B, R, 518, seed=3, pad="dense").to(dev)
bytes_ = B * (g * g * D * 4 + R * D * 4 + R * 16)
big = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731


def timeit(fn, n=5):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        big.zero_()  # flush L2
        e0, e1 = ev(), ev()
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


for name, kw in (("mean", {}), ("gauss", {"gaussian_avg": True, "gaussian_bbox_variance": 1.0})):
    ms = timeit(lambda: ops.pool_boxes(patch, boxes, **kw))
    print(f"pool_boxes {name} B=64 R=64: {ms * 1e3:.1f} us  {bytes_ / ms / 1e6:.0f} GB/s  ({bytes_ / ms / 1e6 / 6551:.2%} of measured HBM peak)")
tr = synth.synth_traces(256, seed=1)
tok2 = torch.randn(256, 5 + g * g, D, device=dev)
w = ops.trace_bins(tr, g, dev)
ms = timeit(lambda: ops.pool_grid(tok2[:, 5:], w.reshape(256, 1, -1), 1.0 / (g * g)))
by = 256 * (g * g * D * 4 + D * 4 + g * g * 4)
print(f"trace pool_grid B=256: {ms * 1e3:.1f} us {by / ms / 1e6:.0f} GB/s ({by / ms / 1e6 / 6551:.2%})")
