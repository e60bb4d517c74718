"""CPU oracle for the Patch-ioner patch -> region -> caption hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker (or as the CPU arm being timed), never as the thing shipped.  The
product path (``patch-ioner_b200``) raises when its CUDA library is missing and
never falls back to this code.

What it is: a plain torch-CPU / pure-Python restatement of the reference's
algorithm for the hot path named in BASELINE.json, each
function citing the reference ``file:line`` it follows.

Parity pinning: the reference ships no tests and no golden vectors for this
path, so the restatement is pinned against OUTPUTS OF THE
REFERENCE ITSELF: ``tests/golden/make_golden.py`` imports the unmodified
reference from ``/root/reference`` (shimmed imports only, no edits), runs its
own functions (``extract_bboxes_feats``, ``map_traces_to_grid``,
``process_self_attention``, ``Im2TxtProjector.project``, ``decoding_batched``,
``compute_region_means``, ``revert_transformation`` and the full
``Patchioner.forward``) on seeded inputs and commits the results as fixtures
under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks every oracle
function against them.  One piece is *not* pinned by reference code: the DINOv2
ViT-B/14-reg4 arithmetic lives in ``facebookresearch/dinov2`` (torch.hub,
unpinned, absent from ``/root/reference`` and from this offline image).  It is
restated from its published architecture in ``oracle/dinov2.py`` and
cross-checked against the independent ``transformers.Dinov2WithRegistersModel``
port installed in the image -> for the ViT alone: "parity unpinned against
upstream, cross-checked against the HF port".
"""

from . import dinov2, pooling, memory, decap, pipeline  # noqa: F401
