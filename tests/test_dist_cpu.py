"""CPU, world_size 2, gloo: the N>1 host logic -- image sharding and the sharded-bank merge choreography."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import memory as o_mem
from oracle import pipeline as o_pipe


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import importlib.util

    spec = importlib.util.spec_from_file_location("pio_dist", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                                          "patch-ioner_b200", "dist.py"))
    pd = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(pd)

    bank = o_mem.drop_zero_rows(o_pipe.synth_bank(4001, 768, seed=3))
    q = torch.randn(37, 768, generator=torch.Generator().manual_seed(4))
    lo, hi = pd.shard_range(bank.shape[0], rank, world)
    m, l, O = o_mem.project_partial(q, bank[lo:hi])

    def rescale_(O_, l_, ml, mg):  # the CPU stand-in for pio_project_rescale (test infrastructure)
        f = torch.exp(ml - mg)
        O_ *= f[:, None]
        l_ *= f

    def finish_(O_, l_, normalize):  # stand-in for pio_project_finish
        out = O_ / l_[:, None]
        return out / out.norm(dim=-1, keepdim=True) if normalize else out

    merged = pd.merge_partials(m, l, O, rescale_, finish_, True)
    ids = torch.full((3, 30), rank, dtype=torch.int32)
    allids = pd.gather_ids(ids)
    if rank == 0:
        torch.save({"merged": merged, "ids": allids}, os.path.join(out_dir, "r0.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_bank_merge_world2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = torch.load(os.path.join(tmp_path, "r0.pt"))
    bank = o_mem.drop_zero_rows(o_pipe.synth_bank(4001, 768, seed=3))
    q = torch.randn(37, 768, generator=torch.Generator().manual_seed(4))
    ref = o_mem.project(q, bank, normalize=True)
    torch.testing.assert_close(got["merged"], ref, rtol=1e-4, atol=1e-6)
    assert got["ids"].shape == (6, 30) and got["ids"][:3].eq(0).all() and got["ids"][3:].eq(1).all()


def test_shard_range_partitions():
    import importlib.util

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("pio_dist", os.path.join(root, "patch-ioner_b200", "dist.py"))
    pd = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(pd)
    for n in (0, 1, 7, 64, 591753):
        for w in (1, 2, 3, 8):
            rs = [pd.shard_range(n, r, w) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in rs]
            assert max(sizes) - min(sizes) <= 1
