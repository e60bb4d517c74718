"""Multi-GPU plumbing: one process per GPU over torch.distributed.

* images and their regions are independent units -> contiguous image shards per rank, no data-path collective;
* only a ROW-SHARDED caption bank needs an exchange: all_reduce(MAX) on the running max m, local rescale by
  exp(m_i - m), all_reduce(SUM) on [O | l], then O / l and the L2 normalisation.
The numeric steps are the library's (pio_project / pio_project_rescale / pio_project_finish); this module only
orders the collectives, so the same choreography is testable on CPU with gloo.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of n units owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def merge_partials(m: torch.Tensor, l: torch.Tensor, O: torch.Tensor, rescale_: Callable, finish_: Callable,
                   normalize: bool = True, group=None) -> torch.Tensor:
    """Combine per-rank flash-style partials (m [R], l [R], O [R,D]) of a row-sharded bank.

    ``rescale_(O, l, m_local, m_global)`` multiplies O and l in place by exp(m_local - m_global);
    ``finish_(O, l, normalize)`` divides by l and L2-normalises.  R*4 bytes (MAX) + R*(D+1)*4 bytes (SUM) per rank.
    """
    m_glob = m.clone()
    dist.all_reduce(m_glob, op=dist.ReduceOp.MAX, group=group)
    rescale_(O, l, m, m_glob)
    packed = torch.cat([O, l[:, None]], dim=1).contiguous()  # one SUM all-reduce for [O | l]
    dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    O_sum, l_sum = packed[:, :-1].contiguous(), packed[:, -1].contiguous()
    return finish_(O_sum, l_sum, normalize)


def project_sharded(bank_shard, q: torch.Tensor, temperature: float = 0.01, normalize: bool = True, group=None) -> torch.Tensor:
    """Im2TxtProjector.project over a bank whose rows are sharded across the ranks of `group` (queries replicated)."""
    from . import ops

    m, l, O = bank_shard.project(q, temperature=temperature, partial=True)
    return merge_partials(m, l, O, ops.project_rescale_, ops.project_finish_, normalize, group)


def gather_ids(ids: torch.Tensor, group=None) -> Optional[torch.Tensor]:
    """All ranks' [R_i, T] id blocks concatenated in rank order (equal R_i) -- only if a single caller wants them."""
    world = dist.get_world_size(group)
    out = [torch.empty_like(ids) for _ in range(world)]
    dist.all_gather(out, ids.contiguous(), group=group)
    return torch.cat(out, dim=0)
