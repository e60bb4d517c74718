"""BASELINE configs[4] on real GPUs: caption memory bank row-sharded over the ranks, NCCL all-reduce of the partial softmax.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/sharded_bank_nccl.py [bank_rows] [regions] [reps]

Every rank holds rows [lo, hi) of the same seeded bank and all the queries; pio_project(partial) -> all_reduce(MAX) on m ->
pio_project_rescale -> all_reduce(SUM) on [O | l] -> pio_project_finish (patch-ioner_b200/dist.py).  Rank 0 checks the
result against the unsharded projection on its own GPU and prints one JSON line (device-timed, max over ranks, with the
per-phase split).  ``run_sharded_bank`` is also what bench.py calls for its ``sharded_bank`` sub-record when WORLD_SIZE > 1.
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

BLOCK = 125_000  # the synthetic bank is defined block-wise so that a rank only generates the rows of its own shard


def bank_rows(lo: int, hi: int, dev) -> torch.Tensor:
    """rows [lo, hi) of the seeded synthetic bank (block b = randn(BLOCK, 768) with seed 700 + b, generated on the device)"""
    out = []
    for b in range(lo // BLOCK, (hi - 1) // BLOCK + 1):
        g = torch.Generator(device=dev).manual_seed(700 + b)
        blk = torch.randn(BLOCK, 768, device=dev, generator=g)
        out.append(blk[max(lo - b * BLOCK, 0):min(hi - b * BLOCK, BLOCK)])
    return torch.cat(out, 0).contiguous()


def run_sharded_bank(dev, rank: int, world: int, M: int = 1_000_000, R: int = 4096, reps: int = 5, check: bool = True) -> dict:
    from patchioner_b200 import dist as pd, ops

    lo, hi = pd.shard_range(M, rank, world)
    shard = ops.Bank(bank_rows(lo, hi, dev), dev, "bf16")
    q = torch.randn(R, 768, generator=torch.Generator().manual_seed(11)).to(dev)
    marks = {}

    def ev(name):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        marks.setdefault(name, []).append(e)

    def once(timed: bool):
        """patch-ioner_b200/dist.py::project_sharded, unrolled so that the phases can be bracketed by events"""
        qq = q.clone()
        if timed:
            ev("t0")
        m, l, O = shard.project(qq, temperature=0.01, partial=True)
        if timed:
            ev("t1")
        m_glob = m.clone()
        dist.all_reduce(m_glob, op=dist.ReduceOp.MAX)
        if timed:
            ev("t2")
        ops.project_rescale_(O, l, m, m_glob)
        packed = torch.cat([O, l[:, None]], dim=1).contiguous()
        if timed:
            ev("t3")
        dist.all_reduce(packed, op=dist.ReduceOp.SUM)
        if timed:
            ev("t4")
        out = ops.project_finish_(packed[:, :-1].contiguous(), packed[:, -1].contiguous(), True)
        if timed:
            ev("t5")
        return out

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
    dist.barrier()
    for _ in range(reps):  # a second, instrumented pass for the per-phase split (events between the phases)
        once(True)
    torch.cuda.synchronize()
    names = ["local_gemms", "allreduce_max", "rescale_pack", "allreduce_sum", "finish"]
    ph = torch.tensor([sum(marks[f"t{i}"][k].elapsed_time(marks[f"t{i + 1}"][k]) for k in range(reps)) / reps for i in range(5)], device=dev)
    dist.all_reduce(ph, op=dist.ReduceOp.MAX)
    rec = None
    if rank == 0:
        ms = float(t.item())
        rec = {"what": "sharded caption-memory projection (BASELINE configs[4])", "n_gpus": world, "bank_rows": M, "regions": R,
               "ms": ms, "regions_per_s": R / ms * 1e3, "tflops_aggregate": 4.0 * M * 768 * R / ms / 1e9,
               "phases_ms_max_over_ranks": {n: round(float(v), 4) for n, v in zip(names, ph.tolist())},
               "allreduce_bytes_per_rank": R * 4 + R * 769 * 4, "scaling": "strong (fixed bank and queries)"}
        if check:
            del shard
            torch.cuda.empty_cache()
            full = ops.Bank(bank_rows(0, M, dev), dev, "bf16")
            ref = full.project(q.clone(), normalize=True)
            rec["min_cosine_vs_unsharded"] = torch.nn.functional.cosine_similarity(out.double(), ref.double(), dim=-1).min().item()
            del full
    dist.barrier()
    return rec


def main():
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
    rec = run_sharded_bank(dev, rank, world, M, R, reps)
    if rank == 0:
        print(json.dumps(rec), flush=True)
        assert rec["min_cosine_vs_unsharded"] >= 0.999, rec
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
