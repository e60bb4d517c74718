#!/usr/bin/env python
"""Golden vectors for the ``variance == 0`` branches (one-hot on a central patch, python ``random`` picks between the two
central indices of an even span) from the UNMODIFIED reference: ``extract_bboxes_feats`` (src/bbox_utils.py:62-71, dense and
box-set mode) and ``compute_region_means`` (src/model.py:71-79).

    python tests/golden/make_golden_centre.py        # build container only (needs /root/reference)

The inputs are the seeded ones of the pooling fixture (``_pool_inputs`` in tests/test_oracle_golden.py builds the same tensors);
``random.seed(SEED)`` is set before every reference call, and the tests do the same before calling the oracle / the device path.
"""
import importlib.util
import os
import random
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
SEED = 20251


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(HERE, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def inputs(g, B, R, D):
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import pipeline as o_pipe  # the seeded box generator the pooling fixture uses

    S = g * 14
    gen = torch.Generator().manual_seed(100 + g)
    tok = torch.randn(B, g * g, D, generator=gen)
    torch.rand(B, g * g, generator=gen)  # the attention map of the pooling fixture (keeps the generator in step)
    boxes = o_pipe.synth_boxes(B, R, S, seed=5 + g, degenerate_frac=0.2)
    boxes[0, 0] = torch.tensor([float(S - 20), float(S - 20), 100.0, 100.0])
    boxes[1, 1] = torch.tensor([3.5, 7.25, 27.9, 13.99])
    boxes_set = boxes.clone()
    boxes_set[:, -1] = -1.0
    boxes_dense = boxes.clone()
    boxes_dense[:, -1] = torch.tensor([0.0, 0.0, 1.0, 1.0])
    return tok, boxes_dense, boxes_set


def main():
    mg = _load("make_golden")
    mg.install_shims()
    sys.path.insert(0, mg.REF)
    import src.bbox_utils as ref_bbox
    import src.model as ref_model

    out = {"seed": SEED}
    for name, (B, g, R, D) in {"g16": (2, 16, 6, 64), "g37": (2, 37, 9, 64)}.items():
        tok, bd, bs = inputs(g, B, R, D)
        rec = {"shape": (B, g, R, D), "in_tok_sum": float(tok.double().sum())}
        random.seed(SEED)
        rec["dense"] = ref_bbox.extract_bboxes_feats(tok, bd.clone(), gaussian_avg=True, gaussian_bbox_variance=0)
        random.seed(SEED)
        rec["set"] = ref_bbox.extract_bboxes_feats(tok, bs.clone(), gaussian_avg=True, gaussian_bbox_variance=0,
                                                   get_single_embedding_per_image=True)
        random.seed(SEED)
        rec["region_means_0"] = ref_model.compute_region_means(tok, 0)
        out[name] = rec
    torch.save(out, os.path.join(HERE, "centre.pt"))
    print({k: ({kk: (tuple(v.shape) if torch.is_tensor(v) else v) for kk, v in r.items()} if isinstance(r, dict) else r) for k, r in out.items()})


if __name__ == "__main__":
    main()
