"""Oracle (test infrastructure): DeCap caption-memory projection and the Talk2DINO inverse map.

Restates
  * ``Im2TxtProjector.project``   Patch-ioner/src/decap/im2txtprojection/im2txtprojection.py:353-385
  * zero-row filter at load       .../im2txtprojection.py:343-345
  * ``get_pseudo_inverse`` / ``revert_transformation``   Patch-ioner/src/embedding_utils.py:3-24
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch


def drop_zero_rows(bank: torch.Tensor) -> torch.Tensor:
    """im2txtprojection.py:345: rows with zero norm are removed when the bank is loaded."""
    return bank[bank.norm(dim=-1) != 0]


def project(
    q: torch.Tensor,
    bank: torch.Tensor,
    temperature: float = 0.01,
    normalize: bool = False,
    return_n_best_sims: Optional[int] = None,
):
    """q [R,D], bank [M,D] fp32 -> [R,D].

    K = bank/|bank| (recomputed every call, :367); q/|q| (:368, in place in the reference --
    not mutated here); sim = q K^T (:370); P = softmax(sim/T) (:376); out = P @ bank with the
    RAW rows (:377); out/|out| if ``normalize`` (:379-380).  ``return_n_best_sims=n`` also
    returns the n largest sims per row, descending (:382-383).
    """
    q = q.float()
    bank = bank.float()
    k = bank / bank.norm(dim=-1, keepdim=True)
    qn = q / q.norm(dim=-1, keepdim=True)
    sim = qn @ k.T
    p = (sim / temperature).softmax(dim=-1)
    out = p @ bank
    if normalize:
        out = out / out.norm(dim=-1, keepdim=True)
    if return_n_best_sims:
        return out, sim.sort(dim=-1, descending=True).values[:, :return_n_best_sims]
    return out


def project_partial(q: torch.Tensor, bank_shard: torch.Tensor, temperature: float = 0.01) -> Tuple[torch.Tensor, ...]:
    """One rank's share of a row-sharded bank: returns (m [R], l [R], O [R,D]) with
    m = max_j s_j/T over the shard, l = sum exp(s_j/T - m), O = sum exp(s_j/T - m) bank_j."""
    q = q.float()
    k = bank_shard / bank_shard.norm(dim=-1, keepdim=True)
    s = (q / q.norm(dim=-1, keepdim=True)) @ k.T / temperature
    m = s.max(dim=-1).values
    e = torch.exp(s - m[:, None])
    return m, e.sum(dim=-1), e @ bank_shard.float()


def merge_partials(ms, ls, Os, normalize: bool = True) -> torch.Tensor:
    """all_reduce(MAX) on m, rescale, all_reduce(SUM) on [O | l], O/l, L2-normalise."""
    m = torch.stack(ms).max(dim=0).values
    l = sum(li * torch.exp(mi - m) for mi, li in zip(ms, ls))
    O = sum(Oi * torch.exp(mi - m)[:, None] for mi, Oi in zip(ms, Os))
    out = O / l[:, None]
    if normalize:
        out = out / out.norm(dim=-1, keepdim=True)
    return out


def get_pseudo_inverse(A: torch.Tensor) -> torch.Tensor:
    """embedding_utils.py:3-15: SVD pseudo-inverse with a 1e-10 cut-off (init time, host)."""
    U, S, Vh = torch.linalg.svd(A, full_matrices=False)
    S_pinv = torch.zeros_like(S)
    nz = S > 1e-10
    S_pinv[nz] = 1.0 / S[nz]
    return Vh.T @ torch.diag(S_pinv) @ U.T


def revert_transformation(features: torch.Tensor, A_pinv: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """embedding_utils.py:24: (x - b) @ A_pinv^T   (768 -> 512 for Talk2DINO ViT-B)."""
    return (features - b) @ A_pinv.t()
