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


def talk2dino_project_clip_txt(w, textual_embedding: torch.Tensor, act: str = "tanh") -> torch.Tensor:
    """``ProjectionLayer.project_clip_txt`` (Patch-ioner/src/talk2dino/talk2dino.py:73-83), applied to the CLIP text
    features when a caption memory is built (im2txtprojection.py:519-523): Linear, then (act, Linear) per hidden layer."""
    f = {"tanh": torch.tanh, "relu": torch.relu, None: None}[act]
    x = torch.nn.functional.linear(textual_embedding.float(), w["linear_layer.weight"], w["linear_layer.bias"])
    i = 0
    while f"hidden_layers.{i}.weight" in w:
        if f is not None:
            x = f(x)
        x = torch.nn.functional.linear(x, w[f"hidden_layers.{i}.weight"], w[f"hidden_layers.{i}.bias"])
        i += 1
    return x


def make_talk2dino_weights(seed: int = 77, clip_dim: int = 512, dino_dim: int = 768, hidden_layers: int = 1):
    g = torch.Generator().manual_seed(seed)
    w = {"linear_layer.weight": torch.randn(dino_dim, clip_dim, generator=g) * clip_dim ** -0.5,
         "linear_layer.bias": torch.randn(dino_dim, generator=g) * 0.1}
    for i in range(hidden_layers):
        w[f"hidden_layers.{i}.weight"] = torch.randn(dino_dim, dino_dim, generator=g) * dino_dim ** -0.5
        w[f"hidden_layers.{i}.bias"] = torch.randn(dino_dim, generator=g) * 0.1
    return w
