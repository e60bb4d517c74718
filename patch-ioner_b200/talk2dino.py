"""Talk2DINO inverse map, init-time host maths (Patch-ioner/src/embedding_utils.py:3-15)."""
import torch


def pseudo_inverse(A: torch.Tensor) -> torch.Tensor:
    """SVD pseudo-inverse with the reference's 1e-10 cut-off on the singular values."""
    U, S, Vh = torch.linalg.svd(A, full_matrices=False)
    S_pinv = torch.zeros_like(S)
    nz = S > 1e-10
    S_pinv[nz] = 1.0 / S[nz]
    return Vh.T @ torch.diag(S_pinv) @ U.T
