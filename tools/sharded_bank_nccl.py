"""BASELINE config 5 on real GPUs: caption memory bank row-sharded over the ranks, NCCL all-reduce of the partial softmax.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/sharded_bank_nccl.py [bank_rows] [regions] [reps]

Every rank holds rows [lo, hi) of the same seeded bank and all the queries; pio_project(partial) -> all_reduce(MAX) on m ->
pio_project_rescale -> all_reduce(SUM) on [O | l] -> pio_project_finish (patch-ioner_b200/dist.py).  Rank 0 checks the
result against the unsharded projection on its own GPU and prints one JSON line (device-timed, max over ranks).
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

M = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
R = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
if rank == 0:
    import __graft_entry__ as ge
    ge.build()
dist.barrier()
from patchioner_b200 import dist as pd, ops, synth  # noqa: E402

bank = synth.synth_bank(M, 768, seed=7)
bank = bank[bank.abs().sum(dim=1) > 0]  # zero rows are dropped at load (im2txtprojection.py:345)
lo, hi = pd.shard_range(bank.shape[0], rank, world)
shard = ops.Bank(bank[lo:hi].contiguous(), dev, "bf16")
q = torch.randn(R, 768, generator=torch.Generator().manual_seed(11)).to(dev)
out = None
for _ in range(2):
    out = pd.project_sharded(shard, q.clone(), 0.01, True)
torch.cuda.synchronize()
dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    out = pd.project_sharded(shard, q.clone(), 0.01, True)
e1.record()
torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    del shard
    full = ops.Bank(bank, dev, "bf16")
    ref = full.project(q.clone(), normalize=True)
    cos = torch.nn.functional.cosine_similarity(out.double(), ref.double(), dim=-1).min().item()
    ms = float(t.item())
    print(json.dumps({"what": "sharded caption-memory projection (BASELINE configs[4])", "n_gpus": world, "bank_rows": int(bank.shape[0]),
                      "regions": R, "ms": ms, "regions_per_s": R / ms * 1e3, "tflops_aggregate": 4.0 * bank.shape[0] * 768 * R / ms / 1e9,
                      "min_cosine_vs_unsharded": cos, "allreduce_bytes_per_rank": R * 4 + R * 769 * 4}), flush=True)
    assert cos >= 0.999, cos
dist.barrier()
dist.destroy_process_group()
