#!/usr/bin/env python
"""Golden vectors for the Talk2DINO inversion from the UNMODIFIED reference file
``Patch-ioner/src/embedding_utils.py`` (loaded by path: it needs nothing but torch).

    python tests/golden/make_golden_revert.py        # build container only (needs /root/reference)

``memory.pt["revert"]`` (make_golden.py) pins the same function with the real ``vitb_mlp_infonce.pth`` weights, which
cannot travel to the GPU box; this fixture uses a seeded synthetic first layer (A [768,512], b [768]) so that the
``-m gpu`` test can rebuild the inputs, and stores the reference's reverted embeddings plus eight rows and two checksums of its pseudo-inverse.
"""
import importlib.util
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/Patch-ioner/src/embedding_utils.py"


def inputs():
    g = torch.Generator().manual_seed(2111)
    A = torch.randn(768, 512, generator=g) * 0.05          # Talk2DINO linear_layer.weight: CLIP 512 -> 768
    b = torch.randn(768, generator=g) * 0.1
    x = torch.randn(24, 768, generator=g)
    x = x / x.norm(dim=-1, keepdim=True)                    # what Im2TxtProjector.project(normalize=True) hands over
    return A, b, x


def main():
    spec = importlib.util.spec_from_file_location("ref_embedding_utils", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    A, b, x = inputs()
    A_pinv = ref.get_pseudo_inverse(A)
    out = {"in_sums": [float(t.double().sum()) for t in (A, b, x)],
           "A_pinv_head": A_pinv[:8].clone(), "A_pinv_sum": float(A_pinv.double().sum()),
           "A_pinv_abs_sum": float(A_pinv.double().abs().sum()),
           "revert": ref.revert_transformation(x, A_pinv=A_pinv, b=b),
           "revert_via_layer": ref.revert_transformation(x, linear_layer=type("L", (), {"weight": A, "bias": b})())}
    torch.save(out, os.path.join(HERE, "revert.pt"))
    print({k: (tuple(v.shape) if torch.is_tensor(v) else v) for k, v in out.items()})


if __name__ == "__main__":
    main()
