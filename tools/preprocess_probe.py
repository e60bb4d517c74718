"""Device preprocessing vs the reference's host pipeline: python tools/preprocess_probe.py [B] [H] [W]"""
import os
import sys
import time

import numpy as np
import torch
from PIL import Image

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import preprocess as o_pre  # noqa: E402  (only as the timed host baseline)
from patchioner_b200 import preprocess as pre  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
H = int(sys.argv[2]) if len(sys.argv) > 2 else 480
W = int(sys.argv[3]) if len(sys.argv) > 3 else 640
dev = torch.device("cuda:0")
rng = np.random.RandomState(0)
arr = rng.randint(0, 256, (B, H, W, 3), dtype=np.uint8)
pil = [Image.fromarray(a) for a in arr]
host = torch.from_numpy(arr).pin_memory()
for _ in range(3):
    out = pre.preprocess_batch(host.to(dev, non_blocking=True), 518, 518, True)
torch.cuda.synchronize()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
reps = 10
e0.record()
for _ in range(reps):
    d = host.to(dev, non_blocking=True)
e1.record()
for _ in range(reps):
    out = pre.preprocess_batch(d, 518, 518, True)
e2.record()
torch.cuda.synchronize()
t_copy, t_k = e0.elapsed_time(e1) / reps, e1.elapsed_time(e2) / reps
t0 = time.perf_counter()
ref = o_pre.reference_transform(pil[:8], 518, 518, True)
t_host = (time.perf_counter() - t0) / 8 * B * 1e3
in_b, out_b = B * H * W * 3, B * 3 * 518 * 518 * 4
print(f"{B} x {H}x{W} -> 518: H2D of the raw bytes {t_copy:.3f} ms, pio_preprocess {t_k:.3f} ms ({B / t_k * 1e3:.0f} img/s, "
      f"{(in_b + out_b) / t_k / 1e6:.0f} GB/s of algorithmic bytes); torchvision on one host core {t_host:.0f} ms ({B / t_host * 1e3:.0f} img/s); "
      f"identical: {torch.equal(out[:8].cpu(), ref)}")
