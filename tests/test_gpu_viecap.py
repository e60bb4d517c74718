"""GPU (-m gpu): the ViECap captioner (mapping network, entity retrieval, prompt-continuing GPT-2 greedy decode) through
the C ABI, against oracle/viecap.py and the reference outputs in tests/golden/viecap.pt."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import viecap as ov


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def ops():
    from patchioner_b200 import ops as _ops

    return _ops


@pytest.fixture(scope="module")
def weights():
    return ov.make_weights()


def _unit(x):
    return x / x.norm(dim=-1, keepdim=True)


def test_mapper_against_oracle(dev, ops, weights):
    g = torch.Generator().manual_seed(3)
    feats = _unit(torch.randn(37, 768, generator=g))
    ref = ov.mapping_network(weights, feats)
    got = ops.Mapper(weights, dev, "fp32").forward(feats.to(dev)).cpu()
    torch.testing.assert_close(got, ref, rtol=2e-4, atol=2e-4)
    got16 = ops.Mapper(weights, dev, "bf16").forward(feats.to(dev)).cpu()
    cos = torch.nn.functional.cosine_similarity(got16.reshape(-1, 768), ref.reshape(-1, 768), dim=-1)
    assert cos.min() > 0.995, cos.min()
    assert ops.Mapper(weights, dev, "fp32").forward(feats[:0].to(dev)).shape == (0, 10, 768)


def test_entity_topk_against_oracle(dev, ops):
    g = torch.Generator().manual_seed(5)
    for E, R, k in ((12, 40, 3), (80, 300, 3), (1000, 64, 5), (3, 10, 3)):
        ent = _unit(torch.randn(E, 768, generator=g))
        q = _unit(torch.randn(R, 768, generator=g) + 3 * ent[torch.randint(0, E, (R,), generator=g)])
        probs = ov.entity_probs(q, ent, 0.01)
        rp, ri = torch.topk(probs, k, dim=-1)
        p, i = ops.entity_topk(q.to(dev), ent.to(dev), 0.01, k)
        torch.testing.assert_close(p.cpu(), rp, rtol=2e-3, atol=1e-5)
        clear = (rp[:, :-1] - rp[:, 1:]).min(dim=1).values > 1e-4   # rows whose order is not a near-tie
        assert torch.equal(i.cpu().long()[clear], ri[clear])
    with pytest.raises(Exception, match="k 4"):
        ops.entity_topk(q.to(dev), ent.to(dev), 0.01, 4)


def _agree(ids, ref, margin, tol):
    """rows equal up to the first step whose top-2 logit gap is below tol (a legitimate flip point)"""
    ok = 0
    for r in range(ref.shape[0]):
        same = True
        for t in range(ref.shape[1]):
            if margin[r, t] < tol:
                break
            if ids[r, t] != ref[r, t]:
                same = False
                break
        ok += same
    return ok / ref.shape[0]


@pytest.mark.parametrize("P,steps", [(1, 5), (2, 3), (14, 12), (33, 40), (49, 6), (55, 8)])  # 2..49: batched prefill; else one position at a time
def test_prompt_decode_fp32_against_oracle(dev, ops, weights, P, steps):
    g = torch.Generator().manual_seed(P)
    R = 7
    prompt = torch.randn(R, P, 768, generator=g) * 0.3
    ref, margin = ov.greedy_ids(weights, prompt, steps, return_margin=True)
    dec = ops.Gpt2Decoder(weights, dev, "fp32")
    ids = dec.decode(prompt.to(dev), steps).cpu().long()
    assert _agree(ids, ref, margin, 1e-4) == 1.0
    assert (ids == ref).all(dim=1).float().mean() >= 0.85
    again = dec.decode(prompt.to(dev), steps).cpu().long()
    assert torch.equal(ids, again)


def test_prompt_decode_bf16_runs_and_mostly_agrees(dev, ops, weights):
    g = torch.Generator().manual_seed(11)
    R, P, steps = 64, 18, 16
    prompt = torch.randn(R, P, 768, generator=g) * 0.3
    ref, margin = ov.greedy_ids(weights, prompt, steps, return_margin=True)
    ids = ops.Gpt2Decoder(weights, dev, "bf16").decode(prompt.to(dev), steps).cpu().long()
    assert ids.min() >= 0 and ids.max() < 50257
    assert (ids[:, 0] == ref[:, 0]).float().mean() >= 0.8   # bf16 flips near-ties; reported, not a parity claim


def test_gpt2_small_depth_and_long_cache(dev, ops):
    """12 layers (GPT-2 small), prompt + 64 tokens = 91 cache positions: runs, deterministic, fp32 == oracle on row 0."""
    w = ov.make_weights(seed=77, n_layer_gpt=12, n_layer_map=1)
    g = torch.Generator().manual_seed(1)
    prompt = torch.randn(3, 28, 768, generator=g) * 0.3
    dec = ops.Gpt2Decoder(w, dev, "fp32")
    assert dec.n_layer == 12
    ids = dec.decode(prompt.to(dev), 64).cpu().long()
    ref, margin = ov.greedy_ids(w, prompt[:1], 64, return_margin=True)
    assert _agree(ids[:1], ref, margin, 1e-4) == 1.0
    ids16 = ops.Gpt2Decoder(w, dev, "bf16").decode(prompt.to(dev), 64).cpu().long()
    assert ids16.shape == (3, 64) and ids16.min() >= 0 and ids16.max() < 50257


def test_viecap_forward_against_reference_golden(dev, golden, weights):
    """VieCap.forward through the library, fp32, against the sentences the reference's greedy_search produced."""
    from patchioner_b200.viecap import VieCap

    g = golden("viecap")
    tok = ov.ToyTokenizer()
    cfg = {"state_dict": weights, "entities_text": g["entities"], "texts_embeddings": g["ent_emb"], "tokenizer": tok,
           "clip_hidden_size": 768, "temperature": 0.01, "top_k": 3, "threshold": 0.4, "using_hard_prompt": True,
           "soft_prompt_first": True, "using_greedy_search": True}
    vc = VieCap(cfg, dev, "ViT-B/16", precision="fp32")
    feats = g["feats"].clone().to(dev)
    assert [[g["entities"][i] for i in r] for r in vc.detect_entities(_unit(g["feats"]).to(dev))] == g["detected"]
    assert torch.equal(vc.hard_prompt_tokens(_unit(g["feats"]).to(dev)).cpu().long(), g["hard"])
    emb = vc.prompt_embeddings(feats.clone())
    torch.testing.assert_close(emb[:, :10].cpu(), g["cont"], rtol=2e-4, atol=2e-4)
    ids = vc.forward_ids(feats.clone()).cpu().tolist()
    assert [vc.cut(r) for r in ids] == g["sentence_ids"]
    assert vc.forward(feats.clone()) == g["sentences"]
    # the reference normalises its argument in place (entrypoint.py:108)
    f = g["feats"].clone().to(dev)
    vc.forward(f)
    torch.testing.assert_close(f.norm(dim=-1).cpu(), torch.ones(f.shape[0]), rtol=1e-5, atol=1e-5)
    # hard prompt first / soft prompt only
    for extra in ({"soft_prompt_first": False}, {"using_hard_prompt": False}, {"only_hard_prompt": True}):
        v2 = VieCap({**cfg, **extra}, dev, "ViT-B/16", precision="fp32")
        want = ov.viecap_forward(weights, g["feats"].clone(), g["entities"], g["ent_emb"], ov.ToyTokenizer(),
                                 using_hard_prompt=extra.get("using_hard_prompt", True),
                                 soft_prompt_first=extra.get("soft_prompt_first", True),
                                 only_hard_prompt=extra.get("only_hard_prompt", False), steps=10)
        got = v2.gpt.decode(v2.prompt_embeddings(g["feats"].clone().to(dev)), 10).cpu().long()
        assert (got == want[1]).all(dim=1).float().mean() >= 0.8
    with pytest.raises(ValueError, match="beam_width"):
        VieCap({**cfg, "using_greedy_search": False, "beam_width": 9}, dev, "ViT-B/16")


def test_patchioner_with_viecap_region_sets(dev, golden, weights):
    """BASELINE config 4: region-set embeddings (one per image) captioned by ViECap behind Patchioner.forward."""
    from oracle import dinov2 as o_vit
    from oracle import pipeline as o_pipe
    from patchioner_b200 import Patchioner

    g = golden("viecap")
    tok = ov.ToyTokenizer()
    vcfg = {"state_dict": weights, "entities_text": g["entities"], "texts_embeddings": g["ent_emb"], "tokenizer": tok,
            "clip_hidden_size": 768, "project_length": 10, "temperature": 0.01, "top_k": 3, "threshold": 0.4,
            "using_hard_prompt": True, "soft_prompt_first": True, "using_greedy_search": True}
    m = Patchioner.from_config({"prefix_size": 768, "support_memory_size": 0, "dino_model": "dinov2_vitb14_reg", "normalize": False,
                                "resize_dim": 224, "crop_dim": 224, "dino_weights": o_vit.make_weights(seed=1234),
                                "clip_model_name": "ViT-B/16", "viecap": vcfg, "precision": "fp32"}, device=dev)
    B, S, R = 2, 224, 4
    imgs = o_pipe.synth_images(B, S, seed=1)
    boxes = o_pipe.synth_boxes(B, R, S, seed=2, pad="set")
    out = m(imgs, get_cls_capt=False, bboxes=boxes.clone(), get_controllable_capts=True)
    assert len(out["set_controllable_capts"]) == B and all(isinstance(s, str) for s in out["set_controllable_capts"])
    emb = m.region_embeddings(imgs.to(dev), bboxes=boxes.clone(), get_controllable_capts=True)["set"]
    want = ov.viecap_forward(weights, emb.cpu().clone(), g["entities"], g["ent_emb"], tok)[0]
    assert out["set_controllable_capts"] == want
    piped = list(m.forward_pipelined([{"imgs": imgs, "bboxes": boxes.clone()}] * 3, get_cls_capt=False, get_controllable_capts=True))
    assert len(piped) == 3 and all(p["set_controllable_capts"] == want for p in piped)
    ids = m(imgs, get_cls_capt=False, bboxes=boxes.clone(), get_controllable_capts=True, return_ids=True)["set_controllable_capts"]
    assert ids.shape == (B, 64) and [tok.decode(m.viecap.cut(r)) for r in ids.cpu().tolist()] == want
    dense = m(imgs, get_cls_capt=True, bboxes=o_pipe.synth_boxes(B, R, S, seed=1, pad="dense"))
    assert len(dense["bbox_capts"]) == B and len(dense["bbox_capts"][0]) == R and isinstance(dense["cls_capt"][0], str)
    with pytest.raises(Exception, match="not supported with viecap"):
        m(imgs, get_cls_capt=False, bboxes=boxes.clone(), return_n_best_sims=2)


def test_cabi_errors(dev, ops, weights):
    from patchioner_b200 import _lib as L

    dec = ops.Gpt2Decoder(weights, dev, "fp32")
    with pytest.raises(L.PioError, match="exceed the cache"):
        dec.decode(torch.zeros(2, 70, 768, device=dev), 64)
    lib = L.lib()
    ids = torch.zeros(2, 4, dtype=torch.int32, device=dev)
    assert lib.pio_decode_greedy_prompt(dec._h, None, 2, 3, 4, ids.data_ptr(), None, None, 0, None) != 0
    assert b"null argument" in lib.pio_last_error()
    # the DeCap entry point refuses a decoder without prefix projection instead of dereferencing a null weight
    ws = torch.zeros(lib.pio_decode_prompt_workspace_bytes(dec._h, 2, 3, 4) + (64 << 20), dtype=torch.uint8, device=dev)
    rc = lib.pio_decode_greedy(dec._h, torch.zeros(2, 768, device=dev).data_ptr(), 2, 4, ids.data_ptr(), None, ws.data_ptr(), ws.numel(), None)
    assert rc != 0 and b"no prefix projection" in lib.pio_last_error()
    bad = dict(weights)
    bad["mapping_network.transformer.layers.0.attn.to_queries.bias"] = torch.zeros(768)
    with pytest.raises(NotImplementedError):
        ops.Mapper(bad, dev, "fp32")


def test_bank_builder_talk2dino_mlp_against_reference_golden(dev, golden, tmp_path):
    """Caption-memory builder: the Talk2DINO text projection on the device vs the reference class's outputs, then
    a bank built from it is loaded by Patchioner-side code and projects like the oracle."""
    from oracle import memory as om
    from patchioner_b200 import bank_builder as bb
    from patchioner_b200 import ops

    g = golden("talk2dino")
    for key, want in g["out"].items():
        hidden, act = int(key[1]), key.split("_")[1]
        w = om.make_talk2dino_weights(seed=77 + hidden, hidden_layers=hidden)
        got = bb.talk2dino_project(g["feats"], w, act, device=dev, rows_per_call=10)
        torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-4)
    with pytest.raises(ValueError):
        bb.talk2dino_project(g["feats"], w, "sigmoid", device=dev)
    # old checkpoints name the hidden layer linear_layer2 (talk2dino.py:85-91)
    w1 = om.make_talk2dino_weights(seed=78, hidden_layers=1)
    old = {"linear_layer.weight": w1["linear_layer.weight"], "linear_layer.bias": w1["linear_layer.bias"],
           "linear_layer2.weight": w1["hidden_layers.0.weight"], "linear_layer2.bias": w1["hidden_layers.0.bias"]}
    torch.testing.assert_close(bb.talk2dino_project(g["feats"], old, "tanh", device=dev), g["out"]["h1_tanh"], rtol=1e-4, atol=1e-4)
    # build -> write -> read -> project
    feats = torch.randn(500, 512, generator=torch.Generator().manual_seed(8))
    texts = [f"t{i}" for i in range(500)]
    emb, path = bb.build_bank(feats, texts, w1, "tanh", out_path=str(tmp_path / bb.bank_filename("synthetic", "ViT-B/16", 500)), device=dev)
    bank, t2 = bb.read_bank(path)
    assert t2 == texts and torch.equal(bank, emb)
    q = torch.randn(9, 768, generator=torch.Generator().manual_seed(9))
    want = om.project(q.clone(), om.drop_zero_rows(bank), normalize=True)
    got = ops.Bank(bank, dev, "fp32").project(q.to(dev), normalize=True).cpu()
    assert torch.nn.functional.cosine_similarity(got, want, dim=-1).min() > 0.9999
    from oracle import dinov2 as o_vit
    from oracle import decap as o_decap
    from oracle import pipeline as o_pipe
    from patchioner_b200 import Patchioner
    m = Patchioner.from_config({"decap_weights": o_decap.make_weights(seed=1234), "prefix_size": 768, "support_memory_size": 500,
                                "dino_model": "dinov2_vitb14_reg", "normalize": True, "resize_dim": 224, "crop_dim": 224,
                                "dino_weights": o_vit.make_weights(seed=1234), "memory_bank": path, "calculate_argmax_text": True,
                                "precision": "fp32"}, device=dev)
    caps = m(o_pipe.synth_images(2, 224, seed=1), get_cls_capt=True)["cls_capt"]
    assert len(caps) == 2 and all(c in texts for c in caps)
    from patchioner_b200 import _lib as L
    with pytest.raises(L.PioError, match="fp32 mode only"):
        ops.linear(torch.randn(128, 64, device=dev).bfloat16(), torch.randn(64, 64, device=dev).bfloat16(), "bf16", act=L.ACT_TANH)


def test_viecap_compute_scores_against_reference_golden(dev, golden, weights):
    """compute_scores = per-sentence perplexity (entrypoint.py:164-177) vs the reference's own numbers."""
    from patchioner_b200.viecap import VieCap

    g = golden("viecap")
    tok = ov.ToyTokenizer()
    cfg = {"state_dict": weights, "entities_text": g["entities"], "texts_embeddings": g["ent_emb"], "tokenizer": tok,
           "clip_hidden_size": 768, "temperature": 0.01, "top_k": 3, "threshold": 0.4, "using_hard_prompt": True,
           "soft_prompt_first": True, "using_greedy_search": True}
    vc = VieCap(cfg, dev, "ViT-B/16", precision="fp32")
    got = vc.compute_perplexity(g["score_sentences"])
    torch.testing.assert_close(torch.tensor(got), torch.tensor(g["perplexities"]), rtol=2e-3, atol=0)
    one = vc.compute_perplexity(["bench"])
    assert len(one) == 1 and one[0] != one[0]   # a single token has no next-token loss: NaN, like the reference

    class RoundTrip(ov.ToyTokenizer):
        """ids the random model invents decode to a word that encodes back to the same id (a real BPE vocabulary does that)"""

        def decode(self, ids):
            return "".join(self.names.get(int(i), " q" + "".join(chr(97 + int(c)) for c in str(int(i)))) for i in ids)

        def encode(self, text):
            out = []
            for piece in self._pat.findall(text):
                w_ = piece.strip()
                if w_.startswith("q") and len(w_) > 1 and all("a" <= c <= "j" for c in w_[1:]):
                    out.append(int("".join(str(ord(c) - 97) for c in w_[1:])))
                else:
                    out += super().encode(piece)
            return out

    rt = RoundTrip()
    v2 = VieCap({**cfg, "tokenizer": rt}, dev, "ViT-B/16", precision="fp32")
    sentences, scores = v2.forward(g["feats"].clone().to(dev), compute_scores=True)
    assert len(scores) == len(sentences) == g["feats"].shape[0]
    assert [rt.encode(s_) for s_ in sentences] == g["sentence_ids"]      # 64-token rows: the > 48 KB shared-memory path of the scorer
    want = [ov.perplexity(weights, i) for i in g["sentence_ids"]]
    torch.testing.assert_close(torch.tensor(scores), torch.tensor(want), rtol=2e-3, atol=0)
    with pytest.raises(ValueError, match="128 positions"):
        vc.compute_perplexity([" ".join(["dog"] * 130)])
    vc16 = VieCap(cfg, dev, "ViT-B/16", precision="bf16")
    got16 = vc16.compute_perplexity(g["score_sentences"])
    torch.testing.assert_close(torch.tensor(got16), torch.tensor(g["perplexities"]), rtol=0.1, atol=0)


def test_viecap_padding_follows_the_reference_calls(dev, golden, weights):
    """Hard prompts are padded per reference call and the padding is attended (entrypoint.py:126): captioning rows in calls of
    G rows must equal the oracle run call by call, also when rows of different calls are decoded together."""
    from oracle import dinov2 as o_vit
    from oracle import pipeline as o_pipe
    from oracle import pooling as o_pool
    from patchioner_b200 import Patchioner
    from patchioner_b200.viecap import VieCap

    g = golden("viecap")
    tok = ov.ToyTokenizer()
    cfg = {"state_dict": weights, "entities_text": g["entities"], "texts_embeddings": g["ent_emb"], "tokenizer": tok,
           "clip_hidden_size": 768, "temperature": 0.01, "top_k": 3, "threshold": 0.4, "using_hard_prompt": True,
           "soft_prompt_first": True, "using_greedy_search": True}
    vc = VieCap(cfg, dev, "ViT-B/16", precision="fp32")
    lens = [len(t) for t in vc.hard_prompt_token_lists(_unit(g["feats"]).to(dev))]
    assert len(set(lens)) > 1                                   # the fixture has prompts of different lengths
    for G in (2, 4, 5):
        want = []
        for s in range(0, 6, G):
            want += ov.viecap_forward(weights, g["feats"][s:s + G].clone(), g["entities"], g["ent_emb"], tok, steps=24)[1].tolist()
        got = vc.forward_ids(g["feats"].clone().to(dev), pad_group=G)[:, :24].cpu().tolist()
        assert got == want, G
    assert vc.forward_ids(g["feats"].clone().to(dev), chunk=2).cpu().tolist() == vc.forward_ids(g["feats"].clone().to(dev)).cpu().tolist()

    # through Patchioner.forward: boxes are captioned in calls of bs * bs_factor regions (model.py:981-1013)
    m = Patchioner.from_config({"prefix_size": 768, "support_memory_size": 0, "dino_model": "dinov2_vitb14_reg", "normalize": False,
                                "resize_dim": 224, "crop_dim": 224, "dino_weights": o_vit.make_weights(seed=1234),
                                "clip_model_name": "ViT-B/16", "viecap": cfg, "precision": "fp32"}, device=dev)
    B, S, R = 2, 224, 5
    imgs = o_pipe.synth_images(B, S, seed=3)
    boxes = o_pipe.synth_boxes(B, R, S, seed=3, pad="dense")
    emb = m.region_embeddings(imgs.to(dev), bboxes=boxes.clone())["bbox"].reshape(-1, 768).cpu()
    for bs_factor in (1, 2, 4):
        per_call = B * bs_factor
        want = []
        for s in range(0, B * R, per_call):
            want += ov.viecap_forward(weights, emb[s:s + per_call].clone(), g["entities"], g["ent_emb"], tok)[0]
        out = m(imgs, get_cls_capt=False, bboxes=boxes.clone(), bs_factor=bs_factor)["bbox_capts"]
        assert [c for row in out for c in row] == want, bs_factor


def _golden_script(name):
    import importlib.util
    import os

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".py")
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_beam_search_against_reference_golden(dev, ops, golden):
    """pio_decode_beam_prompt (all prompts x 5 beams as one KV-cached batch, beams re-ordered through the cache-row table) against
    the sentences the UNMODIFIED reference beam_search (search.py:193-285) produced prompt by prompt without a cache
    (tests/golden/make_golden_viecap_beam.py): every beam of every prompt, token for token, in the reference's order; lengths and
    length-normalised scores against the oracle transcript of the same loop."""
    bm = _golden_script("make_golden_viecap_beam")
    rec = golden("viecap_beam")
    w, prompts, eos = bm.inputs()
    assert float(prompts.double().sum()) == rec["prompt_sum"] and eos == rec["eos"]
    dec = ops.Gpt2Decoder(w, dev, "fp32")
    ids, lens, score = dec.beam_search(prompts.to(dev), eos, rec["W"], rec["steps"])
    ids, lens, score = ids.cpu(), lens.cpu(), score.cpu()
    assert dec.beam_steps_run < rec["steps"]                 # every beam ended: the loop stopped early (search.py:275-276)
    for r, reg in enumerate(rec["regions"]):
        got = [ids[r, k, :int(lens[r, k])].tolist() for k in range(rec["W"])]
        assert got == reg["sentences"], (r, got, reg["sentences"])
        order = reg["oracle_avg"].argsort(descending=True)
        torch.testing.assert_close(score[r], reg["oracle_avg"][order], rtol=1e-4, atol=1e-4)
        assert lens[r].tolist() == [int(v) for v in reg["oracle_lens"][order].tolist()]
    # a single prompt, and a width-1 search (= greedy until the first end-of-sentence token)
    one, l1, _ = dec.beam_search(prompts[2:3].to(dev), eos, rec["W"], rec["steps"])
    assert one[0, 0, :int(l1[0, 0])].tolist() == rec["regions"][2]["sentences"][0]
    w1, lw, _ = dec.beam_search(prompts.to(dev), eos, 1, rec["steps"])
    greedy = dec.decode(prompts.to(dev), rec["steps"]).cpu()
    for r in range(prompts.shape[0]):
        n = int(lw[r, 0])
        assert w1[r, 0, :n].tolist() == greedy[r, :n].tolist()
    # steps exhausted before the beams end: lengths stay within the budget, nothing is read past it
    short, ls, _ = dec.beam_search(prompts.to(dev), eos, rec["W"], 3)
    assert int(ls.max()) <= 3 and dec.beam_steps_run == 3
    for r, reg in enumerate(rec["regions"]):
        toks, sl, avg = ov.beam_search_ids(w, prompts[r:r + 1], eos, rec["W"], 3)
        want = ov.beam_sentences(toks, sl, avg)
        assert [short[r, k, :int(ls[r, k])].tolist() for k in range(rec["W"])] == want


def test_beam_search_bf16_mostly_agrees(dev, ops, golden):
    bm = _golden_script("make_golden_viecap_beam")
    rec = golden("viecap_beam")
    w, prompts, eos = bm.inputs()
    dec = ops.Gpt2Decoder(w, dev, "bf16")
    ids, lens, score = dec.beam_search(prompts.to(dev), eos, rec["W"], rec["steps"])
    first = [int(ids[r, 0, 0]) == reg["sentences"][0][0] for r, reg in enumerate(rec["regions"])]
    assert sum(first) >= len(first) - 1, first
    assert bool(torch.isfinite(score).all()) and int(lens.min()) >= 1


def test_viecap_forward_with_beam_search(dev, ops, weights, golden):
    """VieCap.forward with the reference's default search (using_greedy_search False): mapping network -> entities -> prompt ->
    device beam search, against the oracle's beam search on the oracle's prompt embeddings, region by region."""
    from patchioner_b200.viecap import VieCap

    g = golden("viecap")
    tok = ov.ToyTokenizer()
    eos = [tok.encode(e)[-1] for e in (".", " .")]
    ws = ov.stopping_weights(weights, eos)
    cfg = {"state_dict": ws, "tokenizer": tok, "entities_text": g["entities"], "texts_embeddings": g["ent_emb"], "using_hard_prompt": True,
           "soft_prompt_first": True, "threshold": 0.4, "clip_hidden_size": 768, "using_greedy_search": False, "beam_width": 5}
    v = VieCap(cfg, dev, "ViT-B/16", precision="fp32")
    got = v(g["feats"].clone().to(dev))
    _, _, emb, _ = ov.viecap_forward(ws, g["feats"].clone(), g["entities"], g["ent_emb"], ov.ToyTokenizer(), steps=1)
    want = []
    for r in range(emb.shape[0]):
        toks, sl, avg = ov.beam_search_ids(ws, emb[r:r + 1], eos, 5, ov.MAX_LEN)
        want.append(tok.decode(ov.beam_sentences(toks, sl, avg)[0]))
    assert got == want, (got, want)


def test_greedy_search_early_exit_keeps_the_sentences(dev, ops):
    """pio_decode_greedy_prompt_eos: the search stops once every row has emitted '.' / ' .'; the reference runs all 64 positions for a
    batch and cuts afterwards (search.py:184-190) -- the kept tokens must be identical, in both arithmetic modes, on the
    kernel-per-op path (40 rows) and on the persistent-kernel path (8 rows: runs to the end, same sentences)."""
    tok = ov.ToyTokenizer()
    eos = [tok.encode(e)[-1] for e in (".", " .")]
    w = ov.stopping_weights(ov.make_weights(), eos)
    g = torch.Generator().manual_seed(77)
    for mode in ("fp32", "bf16"):
        dec = ops.Gpt2Decoder(w, dev, mode)
        for R in (40, 8):
            prompts = (torch.randn(R, 14, 768, generator=g) * 0.3).to(dev)
            full = dec.decode(prompts, 48).cpu().tolist()
            early = dec.decode(prompts, 48, eos=eos).cpu().tolist()
            ran = dec.steps_run
            cut_full, cut_early = [ov.cut_sentence(r, eos) for r in full], [ov.cut_sentence(r, eos) for r in early]
            assert cut_full == cut_early
            assert max(len(c) for c in cut_full) < 48        # every row ends: the early exit has something to skip
            if R == 40:
                assert ran < 48 and ran % 8 == 0 and ran >= max(len(c) for c in cut_full)
                assert all(t == eos[0] for r in early for t in r[ran:])


def test_patchioner_with_viecap_default_search_is_beam_search(dev, golden, weights):
    """The reference's default (`using_greedy_search` absent -> False, entrypoint.py:77): Patchioner.forward with a ViECap block
    captions the dense boxes by beam search -- strings, ids (return_ids) and the pipelined serving loop agree with the oracle's
    beam search on the region embeddings of the same forward."""
    from oracle import dinov2 as o_vit
    from oracle import pipeline as o_pipe
    from patchioner_b200 import Patchioner

    g = golden("viecap")
    tok = ov.ToyTokenizer()
    eos = [tok.encode(e)[-1] for e in (".", " .")]
    ws = ov.stopping_weights(weights, eos)
    vcfg = {"state_dict": ws, "entities_text": g["entities"], "texts_embeddings": g["ent_emb"], "tokenizer": tok, "clip_hidden_size": 768,
            "project_length": 10, "temperature": 0.01, "top_k": 3, "threshold": 0.4, "using_hard_prompt": False}  # no search key: beam
    m = Patchioner.from_config({"prefix_size": 768, "support_memory_size": 0, "dino_model": "dinov2_vitb14_reg", "normalize": False,
                                "resize_dim": 224, "crop_dim": 224, "dino_weights": o_vit.make_weights(seed=1234),
                                "clip_model_name": "ViT-B/16", "viecap": vcfg, "precision": "fp32"}, device=dev)
    assert m.viecap.args["using_greedy_search"] is False and m.viecap.beam_width == 5
    B, S, R = 2, 224, 3
    imgs = o_pipe.synth_images(B, S, seed=1)
    boxes = o_pipe.synth_boxes(B, R, S, seed=1, pad="dense")
    out = m(imgs, get_cls_capt=False, bboxes=boxes.clone())["bbox_capts"]
    emb = m.region_embeddings(imgs.to(dev), bboxes=boxes.clone())["bbox"].reshape(-1, 768).cpu()
    emb = emb / emb.norm(dim=-1, keepdim=True)
    cont = ov.mapping_network(ws, emb)
    want = []
    for r in range(cont.shape[0]):
        toks, sl, avg = ov.beam_search_ids(ws, cont[r:r + 1], eos, 5, ov.MAX_LEN)
        want.append(tok.decode(ov.beam_sentences(toks, sl, avg)[0]))
    assert [s for img in out for s in img] == want
    ids = m(imgs, get_cls_capt=False, bboxes=boxes.clone(), return_ids=True)["bbox_capts"].reshape(-1, 64)
    assert [tok.decode(m.viecap.cut(r)) for r in ids.cpu().tolist()] == want
    piped = list(m.forward_pipelined([{"imgs": imgs, "bboxes": boxes.clone()}] * 2, get_cls_capt=False))
    assert all([s for img in p["bbox_capts"] for s in img] == want for p in piped)


def test_beam_search_edge_cases(dev, ops, golden):
    """Shapes around the edges of pio_decode_beam_prompt: no prompts, a one-position prompt (the position-by-position prefill
    path), a single step, the widest beam -- against the oracle's beam search prompt by prompt."""
    bm = _golden_script("make_golden_viecap_beam")
    w, prompts, eos = bm.inputs()
    dec = ops.Gpt2Decoder(w, dev, "fp32")
    ids, lens, score = dec.beam_search(prompts[:0].to(dev), eos, 5, 8)
    assert ids.shape == (0, 5, 8) and lens.shape == (0, 5) and score.shape == (0, 5)
    for P, W, steps in ((1, 5, 6), (14, 8, 5), (3, 2, 1), (14, 1, 9)):
        pr = prompts[:3, :P].contiguous()
        ids, lens, score = dec.beam_search(pr.to(dev), eos, W, steps)
        ids, lens, score = ids.cpu(), lens.cpu(), score.cpu()
        for r in range(pr.shape[0]):
            toks, sl, avg = ov.beam_search_ids(w, pr[r:r + 1], eos, W, steps)
            want = ov.beam_sentences(toks, sl, avg)
            got = [ids[r, k, :int(lens[r, k])].tolist() for k in range(W)]
            assert got == want, (P, W, steps, r, got, want)
            torch.testing.assert_close(score[r], avg[avg.argsort(descending=True)], rtol=1e-4, atol=1e-4)
    with pytest.raises(Exception, match="beam width"):
        dec.beam_search(prompts[:1].to(dev), eos, 9, 4)
    with pytest.raises(Exception, match="exceed the cache"):
        dec.beam_search(prompts[:1].to(dev), eos, 2, 120)
