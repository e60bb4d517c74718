"""CPU: the Talk2DINO text projection of the oracle against the reference class's outputs (tests/golden/talk2dino.pt), and the
host side of patch-ioner_b200/bank_builder.py (file names, on-disk round trip, row shards)."""
import os

import pytest
import torch

from oracle import memory as om


def test_talk2dino_projection_oracle_matches_reference(golden):
    g = golden("talk2dino")
    for key, want in g["out"].items():
        hidden, act = int(key[1]), key.split("_")[1]
        w = om.make_talk2dino_weights(seed=77 + hidden, hidden_layers=hidden)
        got = om.talk2dino_project_clip_txt(w, g["feats"], act)
        torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-5)


def test_bank_file_round_trip_and_names(tmp_path):
    from patchioner_b200 import bank_builder as bb

    # im2txtprojection.py:234
    assert bb.bank_filename("coco_train_karpathy", "ViT-B/16", 591753) == "coco_train_karpathy_text_embeddings-ViT-B.16-591753.h5"
    emb = torch.randn(7, 768)
    emb[3] = 0  # zero rows are stored; the loader filters them (im2txtprojection.py:342-345)
    texts = [f"caption {i} é" for i in range(7)]
    path = bb.write_bank(str(tmp_path / "bank.h5"), emb, texts, name="coco")
    e2, t2 = bb.read_bank(path)
    assert torch.equal(e2, emb) and t2 == texts
    with pytest.raises(ValueError):
        bb.write_bank(str(tmp_path / "bad.pt"), emb, texts[:-1])
    out, written = bb.build_bank(emb, texts, None, out_path=str(tmp_path / "raw.pt"))
    assert torch.equal(out, emb) and os.path.exists(written)
    # row shards cover the bank exactly once
    for M, world in ((10, 3), (591753, 8), (5, 8)):
        spans = [bb.shard_rows(M, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == M and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_builder_mlp_fails_loudly_without_gpu():
    from patchioner_b200 import PioError, bank_builder as bb

    with pytest.raises(PioError):
        bb.talk2dino_project(torch.randn(4, 512), om.make_talk2dino_weights(), "tanh", device="cpu")
