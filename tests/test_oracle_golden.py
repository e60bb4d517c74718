"""CPU: the oracle restatement against fixtures produced by the unmodified reference
(tests/golden/make_golden.py).  Integer results bit-exact, fp32 results to fp32 round-off."""
import math
import os

import pytest
import torch

from oracle import decap as o_decap
from oracle import dinov2 as o_vit
from oracle import memory as o_mem
from oracle import pipeline as o_pipe
from oracle import pooling as o_pool


def csum(t):
    return float(t.double().sum())


def _golden_script(name):
    """import tests/golden/<name>.py (the generator scripts also define the seeded inputs of their fixtures)"""
    import importlib.util

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".py")
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _pool_inputs(g, B, R, D):
    S = g * 14
    gen = torch.Generator().manual_seed(100 + g)
    tok = torch.randn(B, g * g, D, generator=gen)
    amap = torch.rand(B, g * g, generator=gen).softmax(dim=-1)
    boxes = o_pipe.synth_boxes(B, R, S, seed=5 + g, degenerate_frac=0.2)
    boxes[0, 0] = torch.tensor([float(S - 20), float(S - 20), 100.0, 100.0])
    boxes[1, 1] = torch.tensor([3.5, 7.25, 27.9, 13.99])
    boxes_set = boxes.clone()
    boxes_set[:, -1] = -1.0
    boxes_dense = boxes.clone()
    boxes_dense[:, -1] = torch.tensor([0.0, 0.0, 1.0, 1.0])
    return tok, amap, boxes_dense, boxes_set


@pytest.mark.parametrize("name", ["g16", "g37"])
def test_pooling_matches_reference(golden, name):
    rec = golden("pooling")[name]
    B, g, R, D = rec["shape"]
    tok, amap, bd, bs = _pool_inputs(g, B, R, D)
    assert csum(tok) == rec["in_tok_sum"] and csum(amap) == rec["in_amap_sum"]
    # integer part: bit exact
    assert torch.equal(o_pool.boxes_to_patch_units(bd, 14), rec["patch_units"])
    tol = dict(rtol=2e-5, atol=2e-6)
    torch.testing.assert_close(o_pool.extract_bboxes_feats(tok, bd), rec["mean"], **tol)
    torch.testing.assert_close(o_pool.extract_bboxes_feats(tok, bd, True, 0.5), rec["gauss_0.5"], **tol)
    torch.testing.assert_close(o_pool.extract_bboxes_feats(tok, bd, True, 1.0), rec["gauss_1.0"], **tol)
    torch.testing.assert_close(o_pool.extract_bboxes_feats(tok, bd, attention_map=amap), rec["attn"], **tol)
    torch.testing.assert_close(o_pool.extract_bboxes_feats(tok, bs, get_single_embedding_per_image=True), rec["set_mean"], **tol)
    torch.testing.assert_close(o_pool.extract_bboxes_feats(tok, bs, True, 1.0, True), rec["set_gauss_1.0"], **tol)
    torch.testing.assert_close(o_pool.extract_bboxes_feats(tok, bs, get_single_embedding_per_image=True, attention_map=amap),
                               rec["set_attn"], **tol)
    torch.testing.assert_close(o_pool.extract_bboxes_feats(tok, bd.long()), rec["mean_intboxes"], **tol)
    for v, key in ((1, "region_means_1"), (100, "region_means_100"), (0.3, "region_means_0.3")):
        torch.testing.assert_close(o_pool.compute_region_means(tok, v), rec[key], **tol)


@pytest.mark.parametrize("name", ["g16", "g37"])
def test_variance_zero_centre_patch_matches_reference(golden, name):
    """variance 0: the reference's python-``random`` centre picks, reproduced call for call (tests/golden/make_golden_centre.py)."""
    import random

    centre = _golden_script("make_golden_centre")
    rec_all = golden("centre")
    rec = rec_all[name]
    B, g, R, D = rec["shape"]
    tok, bd, bs = centre.inputs(g, B, R, D)
    assert csum(tok) == rec["in_tok_sum"]
    random.seed(rec_all["seed"])
    assert torch.equal(o_pool.extract_bboxes_feats(tok, bd, True, 0), rec["dense"])
    random.seed(rec_all["seed"])
    torch.testing.assert_close(o_pool.extract_bboxes_feats(tok, bs, True, 0, True), rec["set"], rtol=2e-5, atol=2e-6)
    random.seed(rec_all["seed"])
    assert torch.equal(o_pool.compute_region_means(tok, 0), rec["region_means_0"])


def test_attention_mutation_is_order_dependent(golden):
    """Q2: the in-place rescale makes overlapping boxes see modified weights -- the oracle must differ
    from an 'independent boxes' computation exactly where the reference does."""
    rec = golden("pooling")["g16"]
    B, g, R, D = rec["shape"]
    tok, amap, bd, _ = _pool_inputs(g, B, R, D)
    indep = torch.stack([o_pool.extract_bboxes_feats(tok, bd[:, j:j + 1], attention_map=amap)[:, 0] for j in range(R)], 1)
    seq = o_pool.extract_bboxes_feats(tok, bd, attention_map=amap)
    torch.testing.assert_close(seq, rec["attn"], rtol=2e-5, atol=2e-6)
    assert (indep - seq).abs().max() > 1e-4  # the quirk is real on these inputs


def test_float_floor_division_matches_torch():
    vals = torch.tensor([0.0, 13.999, 14.0, 27.9, 3.5, -1.0, -14.0, -0.5, 517.0, 503.99, 1e-7, 41.999996, 42.0])
    t = vals.clone()
    t //= 14
    mine = torch.tensor([o_pool.floor_div_f32(float(v), 14.0) for v in vals])
    assert torch.equal(t, mine)
    gen = torch.Generator().manual_seed(3)
    r = (torch.rand(5000, generator=gen) * 560 - 20)
    t = r.clone()
    t //= 14
    assert torch.equal(t, torch.tensor([o_pool.floor_div_f32(float(v), 14.0) for v in r]))


def test_py_slice_semantics():
    x = list(range(16))
    for a in range(-20, 21):
        for b in range(-20, 21):
            lo, hi = o_pool.py_slice(a, b, 16)
            assert x[a:b] == x[lo:hi]


@pytest.mark.parametrize("g", [16, 37])
def test_traces_match_reference(golden, g):
    rec = golden("traces")[f"g{g}"]
    traces = o_pipe.synth_traces(4, seed=40 + g, n_min=20, n_max=80, outside_frac=0.1)
    traces[0] += [{"x": k / g, "y": (g - k) / g, "t": 0.0} for k in range(g + 1)]
    traces[1] += [{"x": 1.0, "y": 1.0, "t": 0}, {"x": 0.0, "y": 0.0, "t": 0}, {"x": 0.29, "y": 0.57, "t": 0}]
    assert [len(t) for t in traces] == rec["npts"]
    grids = torch.stack([o_pool.map_traces_to_grid(t, g) for t in traces])
    assert torch.equal(grids, rec["grids"])  # bins are integer-exact
    gen = torch.Generator().manual_seed(300 + g)
    tok = torch.randn(4, g * g, 64, generator=gen)
    sa = torch.rand(4, g * g, generator=gen).softmax(-1)
    assert csum(tok) == rec["in_tok_sum"]
    torch.testing.assert_close(o_pool.trace_pool(tok, traces), rec["pool"], rtol=2e-5, atol=1e-7)
    torch.testing.assert_close(o_pool.trace_pool(tok, traces, sa), rec["pool_attn"], rtol=2e-5, atol=1e-9)
    # the masks= generalisation is pinned through the trace branch
    torch.testing.assert_close(o_pool.grid_pool(tok, grids[:, None])[:, 0], rec["pool"], rtol=2e-5, atol=1e-7)


def test_cls_attention_map_matches_reference(golden):
    rec = golden("self_attn")
    gen = torch.Generator().manual_seed(77)
    B, N, D = 2, 41, 768
    qkv = torch.randn(B, N, 3 * D, generator=gen)
    patch = torch.randn(B, N - 5, D, generator=gen)
    assert csum(qkv) == rec["in_qkv_sum"]
    sa = o_pool.cls_attention_map(qkv)
    torch.testing.assert_close(sa, rec["self_attn"], rtol=1e-4, atol=1e-7)
    lit, maps = o_pool.process_self_attention_literal(qkv)
    torch.testing.assert_close(lit, rec["self_attn"], rtol=1e-5, atol=1e-8)
    torch.testing.assert_close(maps, rec["self_attn_maps"], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(o_pool.avg_self_attn_token(sa, patch), rec["avg_self_attn_token"], rtol=1e-4, atol=1e-7)


def test_memory_projection_matches_reference(golden):
    rec = golden("memory")
    bank = o_pipe.synth_bank(3000, 768, seed=7, zero_frac=0.002)
    gen = torch.Generator().manual_seed(8)
    q = torch.randn(16, 768, generator=gen)
    q[3] = bank[11] * 2.5 + 0.01 * torch.randn(768, generator=gen)
    assert csum(q) == rec["in_q_sum"] and csum(bank) == rec["in_bank_sum"]
    fb = o_mem.drop_zero_rows(bank)
    assert fb.shape[0] == rec["M_after_filter"] < 3000
    torch.testing.assert_close(o_mem.project(q, fb, normalize=True), rec["out_norm"], rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(o_mem.project(q, fb, normalize=False), rec["out_raw"], rtol=1e-4, atol=1e-5)
    _, sims = o_mem.project(q, fb, normalize=True, return_n_best_sims=5)
    torch.testing.assert_close(sims, rec["best_sims"], rtol=1e-5, atol=1e-6)
    # sharded form == monolithic form
    parts = [o_mem.project_partial(q, s) for s in fb.chunk(4)]
    merged = o_mem.merge_partials([p[0] for p in parts], [p[1] for p in parts], [p[2] for p in parts])
    torch.testing.assert_close(merged, rec["out_norm"], rtol=1e-4, atol=1e-6)


def test_revert_transformation_matches_reference(golden):
    """SURVEY 8 row a11 (embedding_utils.py:3-24): seeded synthetic first layer (travels to the GPU box) ..."""
    inputs = _golden_script("make_golden_revert").inputs
    rec = golden("revert")
    A, b, x = inputs()
    assert [csum(t) for t in (A, b, x)] == rec["in_sums"]
    A_pinv = o_mem.get_pseudo_inverse(A)
    torch.testing.assert_close(A_pinv[:8], rec["A_pinv_head"], rtol=1e-4, atol=1e-6)
    assert abs(csum(A_pinv) - rec["A_pinv_sum"]) < 1e-2 and abs(csum(A_pinv.abs()) / rec["A_pinv_abs_sum"] - 1) < 1e-5
    got = o_mem.revert_transformation(x, A_pinv, b)
    torch.testing.assert_close(got, rec["revert"], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(got, rec["revert_via_layer"], rtol=1e-4, atol=1e-5)
    # the product's init-time host maths is the same SVD
    from patchioner_b200.talk2dino import pseudo_inverse

    torch.testing.assert_close(pseudo_inverse(A), A_pinv, rtol=0, atol=0)


_T2D = "/root/reference/Patch-ioner/src/viecap/training/talk2dino/weights/vitb_mlp_infonce.pth"


@pytest.mark.skipif(not os.path.exists(_T2D), reason="the real Talk2DINO weights live in the build container's /root/reference only")
def test_revert_transformation_real_weights_golden(golden):
    """... and memory.pt['revert']: the reference's own output with the in-tree vitb_mlp_infonce.pth first layer (build box only)."""
    rec = golden("memory")
    sd = torch.load(_T2D, map_location="cpu")
    A, b = sd["linear_layer.weight"].float(), sd["linear_layer.bias"].float()
    A_pinv = o_mem.get_pseudo_inverse(A)
    assert abs(csum(A_pinv) - rec["A_pinv_sum"]) <= 1e-3 * max(1.0, abs(rec["A_pinv_sum"]))
    torch.testing.assert_close(o_mem.revert_transformation(rec["out_norm"], A_pinv, b), rec["revert"], rtol=1e-4, atol=1e-5)


def test_decoder_matches_reference(golden):
    rec = golden("decoder")
    w = o_decap.make_weights(seed=1234)
    assert csum(w["decoder.transformer.wte.weight"]) == rec["wte_sum"]
    gen = torch.Generator().manual_seed(9)
    feats = torch.randn(6, 768, generator=gen)
    feats = feats / feats.norm(dim=-1, keepdim=True)
    assert csum(feats) == rec["in_feats_sum"]
    e0 = o_decap.prefix_embed(w, feats).reshape(6, 1, -1)
    logits0 = o_decap.gpt2_hidden(w, e0)[:, -1] @ w["decoder.transformer.wte.weight"].T
    top = logits0.topk(8, dim=-1)
    assert torch.equal(top.indices, rec["logits0_top_indices"])
    torch.testing.assert_close(top.values, rec["logits0_top_values"], rtol=1e-4, atol=1e-5)
    ids_nc, sc = o_decap.decode_greedy(w, feats, compute_scores=True, use_cache=False)
    ids_c = o_decap.decode_greedy(w, feats, use_cache=True)
    assert torch.equal(ids_nc, rec["ids"])  # literal no-cache form == reference ids
    assert torch.equal(ids_c, rec["ids"])   # KV-cache form == reference ids
    torch.testing.assert_close(sc, rec["scores"].float(), rtol=1e-3, atol=0)


def test_gpt2_restatement_matches_transformers():
    from transformers import GPT2Config, GPT2LMHeadModel

    w = o_decap.make_weights(seed=4321)
    cfg = GPT2Config(vocab_size=50257, n_positions=1024, n_embd=768, n_layer=4, n_head=4, activation_function="gelu_new",
                     layer_norm_epsilon=1e-5, resid_pdrop=0.0, embd_pdrop=0.0, attn_pdrop=0.0)
    m = GPT2LMHeadModel(cfg).eval()
    sd = {k[len("decoder."):]: v for k, v in w.items() if k.startswith("decoder.")}
    m.load_state_dict(sd, strict=False)
    x = torch.randn(3, 7, 768, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        ref = m(inputs_embeds=x).logits
    mine = o_decap.gpt2_hidden(w, x) @ w["decoder.transformer.wte.weight"].T
    torch.testing.assert_close(mine, ref, rtol=1e-4, atol=1e-4)


def test_full_forward_matches_reference(golden):
    rec = golden("forward")
    vit_w = o_vit.make_weights(seed=1234)
    dec_w = o_decap.make_weights(seed=1234)
    bank = o_pipe.synth_bank(3000, 768, seed=7, zero_frac=0.002)
    B, S, R = 2, 224, 4
    imgs = o_pipe.synth_images(B, S, seed=1)
    boxes = o_pipe.synth_boxes(B, R, S, seed=1, pad="dense")
    boxes_set = o_pipe.synth_boxes(B, R, S, seed=2, pad="set")
    traces = o_pipe.synth_traces(B, seed=1)
    d = o_vit.forward(vit_w, imgs)
    torch.testing.assert_close(d["x_norm_clstoken"], rec["vit_cls"], rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(d["x_norm_patchtokens"][:, :100], rec["vit_patch_0_100"], rtol=1e-4, atol=1e-4)
    for variant, with_bank in (("decap", True), ("capdec", False)):
        g = rec[variant]
        assert csum(imgs) == g["in_img_sum"]
        m = o_pipe.OracleModel(vit_w, dec_w, bank if with_bank else None)

        def ids(out, *keys):
            return torch.cat([out[k].reshape(-1, 30) for k in keys], 0)

        o = m.forward(imgs, get_cls_capt=True, bboxes=boxes.clone())
        assert torch.equal(ids(o, "cls_capt", "bbox_capts"), g["cls+bbox_mean"])
        o = m.forward(imgs, get_cls_capt=False, bboxes=boxes.clone(), gaussian_avg=True, gaussian_bbox_variance=1.0)
        assert torch.equal(ids(o, "bbox_capts"), g["bbox_gauss1"])
        o = m.forward(imgs, get_cls_capt=False, bboxes=boxes.clone(), use_attn_map_for_bboxes=True)
        assert torch.equal(ids(o, "bbox_capts"), g["bbox_attn"])
        o = m.forward(imgs, get_cls_capt=False, bboxes=boxes_set.clone(), get_controllable_capts=True,
                      gaussian_avg=True, gaussian_bbox_variance=1.0)
        assert torch.equal(ids(o, "set_controllable_capts"), g["set_gauss1"])
        o = m.forward(imgs, get_cls_capt=False, traces=traces)
        assert torch.equal(ids(o, "trace_capts"), g["trace"])
        o = m.forward(imgs, get_cls_capt=False, traces=traces, use_attention_tracing=True)
        assert torch.equal(ids(o, "trace_capts"), g["trace_attn"])
        o = m.forward(imgs, get_cls_capt=False, get_avg_self_attn_capt=True, get_avg_patch_capt=True, gaussian_img_variance=1.0)
        assert torch.equal(ids(o, "avg_self_attn_capt", "avg_patch_capt"), g["avg_self_attn+avg_patch"])


def test_dinov2_restatement_matches_hf_port():
    """Independent second opinion on the ViT arithmetic (upstream source is not available offline)."""
    from transformers import Dinov2WithRegistersConfig, Dinov2WithRegistersModel

    depth = 2
    w = o_vit.make_weights(seed=99)
    cfg = Dinov2WithRegistersConfig(hidden_size=768, num_hidden_layers=depth, num_attention_heads=12, mlp_ratio=4,
                                    image_size=518, patch_size=14, num_register_tokens=4, layerscale_value=1.0,
                                    hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, layer_norm_eps=1e-6,
                                    qkv_bias=True, use_swiglu_ffn=False)
    m = Dinov2WithRegistersModel(cfg).eval()
    sd = m.state_dict()
    new = {}
    new["embeddings.cls_token"] = w["cls_token"]
    new["embeddings.mask_token"] = w["mask_token"]
    new["embeddings.register_tokens"] = w["register_tokens"]
    new["embeddings.position_embeddings"] = w["pos_embed"]
    new["embeddings.patch_embeddings.projection.weight"] = w["patch_embed.proj.weight"]
    new["embeddings.patch_embeddings.projection.bias"] = w["patch_embed.proj.bias"]
    for i in range(depth):
        p, q = f"blocks.{i}.", f"encoder.layer.{i}."
        qw, kw, vw = w[p + "attn.qkv.weight"].chunk(3, 0)
        qb, kb, vb = w[p + "attn.qkv.bias"].chunk(3, 0)
        new[q + "norm1.weight"], new[q + "norm1.bias"] = w[p + "norm1.weight"], w[p + "norm1.bias"]
        new[q + "norm2.weight"], new[q + "norm2.bias"] = w[p + "norm2.weight"], w[p + "norm2.bias"]
        new[q + "attention.attention.query.weight"], new[q + "attention.attention.query.bias"] = qw, qb
        new[q + "attention.attention.key.weight"], new[q + "attention.attention.key.bias"] = kw, kb
        new[q + "attention.attention.value.weight"], new[q + "attention.attention.value.bias"] = vw, vb
        new[q + "attention.output.dense.weight"], new[q + "attention.output.dense.bias"] = w[p + "attn.proj.weight"], w[p + "attn.proj.bias"]
        new[q + "layer_scale1.lambda1"], new[q + "layer_scale2.lambda1"] = w[p + "ls1.gamma"], w[p + "ls2.gamma"]
        new[q + "mlp.fc1.weight"], new[q + "mlp.fc1.bias"] = w[p + "mlp.fc1.weight"], w[p + "mlp.fc1.bias"]
        new[q + "mlp.fc2.weight"], new[q + "mlp.fc2.bias"] = w[p + "mlp.fc2.weight"], w[p + "mlp.fc2.bias"]
    new["layernorm.weight"], new["layernorm.bias"] = w["norm.weight"], w["norm.bias"]
    assert set(new) == set(sd), (set(sd) ^ set(new))
    m.load_state_dict(new)
    for S in (224, 518 if False else 266):  # 16x16 and 19x19 grids: both exercise the pos-embed resize
        imgs = o_pipe.synth_images(1, S, seed=3)
        with torch.no_grad():
            ref = m(pixel_values=imgs).last_hidden_state
        mine = o_vit.forward(w, imgs, depth=depth)
        got = torch.cat([mine["x_norm_clstoken"][:, None], mine["x_norm_regtokens"], mine["x_norm_patchtokens"]], 1)
        torch.testing.assert_close(got, ref, rtol=2e-4, atol=2e-4)


def test_flop_model():
    assert abs(o_vit.flops_per_image(224) / 1e9 - 47.08) < 0.05
    assert abs(o_vit.flops_per_image(518) / 1e9 - 304.23) < 0.05
    assert abs(o_decap.FLOPS_PER_TOKEN / 1e6 - 133.82) < 0.05
