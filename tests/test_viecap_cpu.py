"""CPU: the ViECap oracle (oracle/viecap.py) against the outputs of the reference's own classes (tests/golden/viecap.pt,
made by tests/golden/make_golden_viecap.py), plus the host logic of patch-ioner_b200/viecap.py that needs no GPU."""
import pytest
import torch

from oracle import viecap as ov


def _run(golden):
    g = golden("viecap")
    w = ov.make_weights()
    tok = ov.ToyTokenizer()
    feats = g["feats"].clone()
    sentences, ids, emb, hard = ov.viecap_forward(w, feats, g["entities"], g["ent_emb"], tok)
    return g, w, tok, feats, sentences, ids, emb, hard


def test_oracle_matches_reference_outputs(golden):
    g, w, tok, feats, sentences, ids, emb, hard = _run(golden)
    # mapping network: same arithmetic as the reference module -> tight
    torch.testing.assert_close(emb[:, :10], g["cont"], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(ov.entity_probs(feats, g["ent_emb"], 0.01), g["probs"], rtol=1e-5, atol=1e-6)
    assert ov.pick_entities(g["entities"], g["probs"], 3, 0.4) == g["detected"]
    assert [] in g["detected"] and any(len(d) >= 2 for d in g["detected"])  # 'something' and a multi-entity prompt are covered
    assert torch.equal(hard, g["hard"])
    eos = [tok.encode(e)[-1] for e in (".", " .")]
    assert [ov.cut_sentence(r, eos) for r in ids.tolist()] == g["sentence_ids"]   # greedy_search of the reference
    assert sentences == g["sentences"]


def test_prompt_composition_and_cut():
    assert ov.compose_prompt([]) == "There are something in image."
    assert ov.compose_prompt(["dog"]) == "There are dog in image."
    assert ov.compose_prompt(["person", "traffic light", "kite"]) == "There are person, traffic light, kite in image."
    assert ov.cut_sentence([5, 6, 7, 6], [7]) == [5, 6, 7]
    assert ov.cut_sentence([5, 6], [7]) == [5, 6]
    from patchioner_b200.viecap import compose_discrete_prompt
    for e in ([], ["dog"], ["a b", "c"]):
        assert compose_discrete_prompt(e) == ov.compose_prompt(e)


def test_piecewise_tokenisation_equals_whole_prompt():
    """The product tokenises every entity once and concatenates; that must equal tokenising the whole prompt."""
    tok = ov.ToyTokenizer()
    ents = ["person", "traffic light", "tv", "hot dog"]
    head, tail, comma = tok.encode("There are"), tok.encode(" in image."), tok.encode(",")
    per = [tok.encode(" " + e) for e in ents]
    for rows in ([0], [1, 3], [2, 0, 1]):
        pieces = list(head)
        for j, r in enumerate(rows):
            pieces += (comma if j else []) + per[r]
        assert pieces + tail == tok.encode(ov.compose_prompt([ents[r] for r in rows]))


def test_cached_greedy_equals_full_recompute():
    w = ov.make_weights(seed=9, n_layer_gpt=2, n_layer_map=1)
    prompt = torch.randn(3, 7, 768, generator=torch.Generator().manual_seed(2)) * 0.3
    a, margin = ov.greedy_ids(w, prompt, 6, return_margin=True)
    b = ov.greedy_ids(w, prompt, 6, use_cache=True)
    assert margin.min() > 1e-4 and torch.equal(a, b)


def test_oracle_perplexity_matches_reference(golden):
    g = golden("viecap")
    w = ov.make_weights()
    got = [ov.perplexity(w, i) for i in g["score_ids"]]
    torch.testing.assert_close(torch.tensor(got), torch.tensor(g["perplexities"]), rtol=1e-4, atol=0)
    assert ov.perplexity(w, [5]) != ov.perplexity(w, [5])  # NaN for a single token, like the reference's empty loss


def test_entity_files_are_read_like_the_reference(tmp_path):
    """entrypoint.py:178-221 + load_annotations.py:91-103: vocabulary JSON lower-cased, stripped, SORTED; embeddings pickle named
    by the suffix with '/' removed and '_with_ensemble' when prompt_ensemble is set."""
    import json
    import pickle

    from patchioner_b200.viecap import DEFAULTS, VieCap

    d = tmp_path / "annotations" / "vocabulary"
    d.mkdir(parents=True)
    (d / "coco_categories.json").write_text(json.dumps(["Person ", "traffic light", "Dog", "cat"]))
    emb = torch.randn(4, 8)
    with open(d / "coco_embeddings_ViT-B16_t2d__with_ensemble.pickle", "wb") as f:
        pickle.dump(emb, f)
    cfg = dict(DEFAULTS, name_of_entities_text="coco_entities", files_path=str(tmp_path), prompt_ensemble=True)
    ents, e = VieCap._load_entities(cfg, "ViT-B/16_t2d_")
    assert ents == ["cat", "dog", "person", "traffic light"] and torch.equal(e, emb)
    cfg["disable_all_entities"] = True          # single-word entities only
    assert VieCap._load_entities(cfg, "ViT-B/16_t2d_")[0] == ["cat", "dog", "person"]
    with pytest.raises(NotImplementedError):
        VieCap._load_entities(dict(cfg, name_of_entities_text="open_image_entities"), "x")


def test_oracle_beam_search_matches_reference_golden(golden):
    """oracle/viecap.py::beam_search_ids against the sentences of the unmodified reference beam_search
    (tests/golden/make_golden_viecap_beam.py): all five beams of every prompt, in the reference's order."""
    import importlib.util
    import os

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_golden_viecap_beam.py")
    spec = importlib.util.spec_from_file_location("make_golden_viecap_beam", path)
    bm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bm)
    rec = golden("viecap_beam")
    w, prompts, eos = bm.inputs()
    assert float(prompts.double().sum()) == rec["prompt_sum"]
    lens_seen = set()
    for r in (0, 2, 4):
        toks, sl, avg = ov.beam_search_ids(w, prompts[r:r + 1], eos, rec["W"], rec["steps"])
        assert ov.beam_sentences(toks, sl, avg) == rec["regions"][r]["sentences"]
        lens_seen.update(int(v) for v in sl.tolist())
    assert len(lens_seen) >= 3   # beams of different lengths: the length normalisation and the frozen stopped beams are exercised
