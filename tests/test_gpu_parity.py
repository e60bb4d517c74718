"""GPU (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle and the golden fixtures.

Bars (BASELINE.json north_star): pooling indices / trace bins bit-exact; region embeddings cosine >= 0.9999
(fp32 mode) or >= 0.999 (bf16 mode); greedy token ids identical on >= 99 % of regions in fp32 mode.
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import decap as o_decap
from oracle import dinov2 as o_vit
from oracle import memory as o_mem
from oracle import pipeline as o_pipe
from oracle import pooling as o_pool


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def ops():
    from patchioner_b200 import ops as _ops

    return _ops


def cos_min(a, b):
    return torch.nn.functional.cosine_similarity(a.double().reshape(-1, a.shape[-1]), b.double().reshape(-1, b.shape[-1]), dim=-1).min().item()


# ----------------------------------------------------------------------------------------- dense layers
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (300, 200, 192), (77, 50, 588), (1030, 768, 768), (257, 2304, 768)])
def test_linear_fp32(dev, ops, M, N, K):
    g = torch.Generator().manual_seed(M * 7 + N)
    A = torch.randn(M, K, generator=g)
    W = torch.randn(N, K, generator=g) / math.sqrt(K)
    bias = torch.randn(N, generator=g)
    gamma = torch.rand(N, generator=g)
    res = torch.randn(M, N, generator=g)
    ref = res + gamma * torch.nn.functional.gelu(A @ W.T + bias)
    out = ops.linear(A.to(dev), W.to(dev), "fp32", bias=bias.to(dev), act=1, gamma=gamma.to(dev), residual=res.to(dev))
    torch.testing.assert_close(out.cpu(), ref, rtol=1e-4, atol=1e-4)
    out2 = ops.linear(A.to(dev), W.to(dev), "fp32")
    torch.testing.assert_close(out2.cpu(), A @ W.T, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (128, 64, 128), (300, 200, 192), (1030, 768, 768), (257, 2304, 768),
                                   (4096, 768, 3072), (70, 1000, 640), (20000, 768, 768), (513, 5003, 768), (2000, 5003, 192), (19000, 2304, 768)])
def test_linear_bf16_tcgen05(dev, ops, M, N, K):
    """tcgen05 GEMM: products of bf16 inputs are exact in fp32, so only the accumulation order differs."""
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g).bfloat16()
    W = (torch.randn(N, K, generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, generator=g)
    ref = A.float() @ W.float().T
    out = ops.linear(A.to(dev), W.to(dev), "bf16")
    torch.testing.assert_close(out.cpu(), ref, rtol=2e-3, atol=2e-3)
    # epilogue: bias + gelu + layerscale + residual, bf16 and fp32 outputs
    gamma = torch.rand(N, generator=g)
    res = torch.randn(M, N, generator=g)
    ref2 = res + gamma * torch.nn.functional.gelu(ref + bias)
    out2 = ops.linear(A.to(dev), W.to(dev), "bf16", bias=bias.to(dev), act=1, gamma=gamma.to(dev), residual=res.to(dev))
    torch.testing.assert_close(out2.cpu(), ref2, rtol=2e-3, atol=2e-3)
    out3 = ops.linear(A.to(dev), W.to(dev), "bf16", bias=bias.to(dev), out_dtype=torch.bfloat16)
    torch.testing.assert_close(out3.float().cpu(), ref + bias, rtol=2e-2, atol=2e-2)
    # in-place residual stream  x += gamma * (A W^T + bias)  (the copy engine's reduce-add when the layout allows it)
    X = res.clone().to(dev)
    ops.linear(A.to(dev), W.to(dev), "bf16", bias=bias.to(dev), gamma=gamma.to(dev), residual=X, out=X)
    torch.testing.assert_close(X.cpu(), res + gamma * (ref + bias), rtol=2e-3, atol=2e-3)
    # gelu to bf16 through a padded leading dimension (what the ViT's fc1 does)
    ld = (N + 63) // 64 * 64 + 64
    buf = torch.full((M, ld), 7.0, dtype=torch.bfloat16, device=dev)
    ops.linear(A.to(dev), W.to(dev), "bf16", bias=bias.to(dev), act=1, out=buf[:, :N])
    torch.testing.assert_close(buf[:, :N].float().cpu(), torch.nn.functional.gelu(ref + bias), rtol=2e-2, atol=2e-2)
    assert (buf[:, N:] == 7.0).all(), "columns beyond N must not be written"


def test_linear_inplace_residual_and_rowscale(dev, ops):
    for mode, dt, tol in (("fp32", torch.float32, 1e-4), ("bf16", torch.bfloat16, 3e-3)):
        g = torch.Generator().manual_seed(5)
        A = torch.randn(200, 256, generator=g).to(dt)
        W = (torch.randn(96, 256, generator=g) / 16).to(dt)
        O = torch.randn(200, 96, generator=g)
        rs = torch.rand(200, generator=g)
        cs = torch.rand(96, generator=g)
        ref = O * rs[:, None] + 0.5 * cs * (A.float() @ W.float().T)
        Od = O.to(dev)
        ops.linear(A.to(dev), W.to(dev), mode, residual=Od, out=Od, res_rowscale=rs.to(dev), colscale=cs.to(dev), alpha=0.5)
        torch.testing.assert_close(Od.cpu(), ref, rtol=tol, atol=tol)


@pytest.mark.parametrize("M,N,K", [(300, 5003, 256), (4096, 50257, 768), (33, 700, 64)])
def test_linear_fused_argmax(dev, ops, M, N, K):
    g = torch.Generator().manual_seed(N)
    A = torch.randn(M, K, generator=g).bfloat16()
    W = (torch.randn(N, K, generator=g) / math.sqrt(K)).bfloat16()
    W[7] = W[3]  # exact ties: the first index must win
    Ad, Wd = A.to(dev), W.to(dev)
    logits = ops.linear(Ad, Wd, "bf16")          # same kernel, materialised
    ids, lp = ops.linear_argmax(Ad, Wd, with_logprob=True)
    want = logits.argmax(dim=-1)
    assert torch.equal(ids.long(), want)
    ref_lp = torch.log_softmax(logits.double(), -1).gather(1, want[:, None])[:, 0].float()
    torch.testing.assert_close(lp, ref_lp, rtol=1e-4, atol=1e-4)


def test_layernorm(dev, ops):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1000, 768, generator=g) * 3 + 1
    w, b = torch.rand(768, generator=g) + 0.5, torch.randn(768, generator=g)
    ref = torch.nn.functional.layer_norm(x, (768,), w, b, eps=1e-6)
    out = ops.layernorm(x.to(dev), w.to(dev), b.to(dev), 1e-6)
    torch.testing.assert_close(out.cpu(), ref, rtol=1e-5, atol=1e-5)
    out16 = ops.layernorm(x.to(dev), w.to(dev), b.to(dev), 1e-6, out_dtype=torch.bfloat16)
    torch.testing.assert_close(out16.float().cpu(), ref, rtol=1e-2, atol=1e-2)


# ----------------------------------------------------------------------------------------- pooling
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
def test_pooling_against_golden(dev, ops, golden, name):
    rec = golden("pooling")[name]
    B, g, R, D = rec["shape"]
    tok, amap, bd, bs = _pool_inputs(g, B, R, D)
    t = tok.to(dev)
    tol = dict(rtol=1e-4, atol=1e-5)
    out, bounds = ops.pool_boxes(t, bd, return_bounds=True)
    assert torch.equal(bounds.cpu(), o_pool.all_box_bounds(bd, 14, g))  # pooling indices: bit exact
    torch.testing.assert_close(out.cpu(), rec["mean"], **tol)
    torch.testing.assert_close(ops.pool_boxes(t, bd, gaussian_avg=True, gaussian_bbox_variance=0.5).cpu(), rec["gauss_0.5"], **tol)
    torch.testing.assert_close(ops.pool_boxes(t, bd, gaussian_avg=True, gaussian_bbox_variance=1.0).cpu(), rec["gauss_1.0"], **tol)
    a = amap.to(dev)
    torch.testing.assert_close(ops.pool_boxes(t, bd, attention_map=a).cpu(), rec["attn"], **tol)
    assert torch.equal(a.cpu(), amap)  # the caller's map is not modified
    torch.testing.assert_close(ops.pool_boxes(t, bs, get_single_embedding_per_image=True).cpu(), rec["set_mean"], **tol)
    torch.testing.assert_close(ops.pool_boxes(t, bs, gaussian_avg=True, gaussian_bbox_variance=1.0,
                                              get_single_embedding_per_image=True).cpu(), rec["set_gauss_1.0"], **tol)
    torch.testing.assert_close(ops.pool_boxes(t, bs, attention_map=a, get_single_embedding_per_image=True).cpu(), rec["set_attn"], **tol)
    torch.testing.assert_close(ops.pool_boxes(t, bd.long()).cpu(), rec["mean_intboxes"], **tol)
    for v, key in ((1, "region_means_1"), (100, "region_means_100"), (0.3, "region_means_0.3")):
        w = ops.region_mean_weights(g, v, dev)
        got = ops.pool_grid(t, w.reshape(1, 1, -1).expand(B, 1, -1), 1.0)[:, 0]
        torch.testing.assert_close(got.cpu(), rec[key], **tol)


def test_pooling_bounds_bit_exact_random(dev, ops):
    """20k random boxes incl. negatives, fractions and off-grid extents: indices equal the oracle's bit for bit."""
    g, B, R, D = 37, 8, 2500, 32
    gen = torch.Generator().manual_seed(11)
    boxes = (torch.rand(B, R, 4, generator=gen) * 640 - 60)
    boxes[:, ::7] = torch.floor(boxes[:, ::7])
    boxes[:, ::11, 2:] = torch.randint(0, 15, (B, len(range(0, R, 11)), 2), generator=gen).float()
    tok = torch.randn(B, g * g, D, generator=gen).to(dev)
    _, bounds = ops.pool_boxes(tok, boxes, return_bounds=True)
    assert torch.equal(bounds.cpu(), o_pool.all_box_bounds(boxes, 14, g))
    _, bounds_i = ops.pool_boxes(tok, boxes.long(), return_bounds=True)
    assert torch.equal(bounds_i.cpu(), o_pool.all_box_bounds(boxes.long(), 14, g))


def test_pooling_many_boxes_per_image(dev, ops):
    """R = 600 boxes per image (a warp owns more than 32 boxes: the per-warp bounds prefetch reloads), values vs the oracle."""
    g, B, R, D = 16, 2, 600, 32
    gen = torch.Generator().manual_seed(13)
    tok = torch.randn(B, g * g, D, generator=gen)
    boxes = o_pipe.synth_boxes(B, R, 224, seed=13, degenerate_frac=0.1)
    for gauss in (False, True):
        ref = o_pool.extract_bboxes_feats(tok, boxes.clone(), gauss, 1.0)
        got = ops.pool_boxes(tok.to(dev), boxes, gaussian_avg=gauss, gaussian_bbox_variance=1.0).cpu()
        torch.testing.assert_close(got, ref, rtol=1e-4, atol=1e-5, equal_nan=True)


def test_pooling_edge_cases(dev, ops):
    g, D = 16, 64
    tok = torch.randn(1, g * g, D, generator=torch.Generator().manual_seed(2))
    boxes = torch.tensor([[[-1.0, -1.0, -1.0, -1.0], [500.0, 500.0, 10.0, 10.0], [0.0, 0.0, 223.0, 223.0], [100.0, 50.0, 0.0, 0.0]]])
    ref = o_pool.extract_bboxes_feats(tok, boxes)
    out = ops.pool_boxes(tok.to(dev), boxes).cpu()
    assert torch.isnan(ref[0, 0]).all() and torch.isnan(out[0, 0]).all()  # empty slice -> NaN like tensor.mean()
    assert torch.isnan(ref[0, 1]).all() and torch.isnan(out[0, 1]).all()
    torch.testing.assert_close(out[0, 2:], ref[0, 2:], rtol=1e-4, atol=1e-5)
    refg = o_pool.extract_bboxes_feats(tok, boxes, True, 1.0)
    outg = ops.pool_boxes(tok.to(dev), boxes, gaussian_avg=True, gaussian_bbox_variance=1.0).cpu()
    torch.testing.assert_close(outg, refg, rtol=1e-4, atol=1e-5, equal_nan=True)
    # all-dummy box set -> 0/0 map -> NaN embedding, like the reference
    sets = torch.full((1, 3, 4), -1.0)
    assert torch.isnan(ops.pool_boxes(tok.to(dev), sets, get_single_embedding_per_image=True)).all()
    assert torch.isnan(o_pool.extract_bboxes_feats(tok, sets, get_single_embedding_per_image=True)).all()


def test_pooling_full_size_properties(dev, ops):
    """BASELINE config 2 size (64 x 518 px, 64 boxes): size-independent properties instead of the slow oracle."""
    B, g, R, D = 64, 37, 64, 768
    gen = torch.Generator().manual_seed(3)
    tokens = torch.randn(B, 5 + g * g, D, generator=gen).to(dev)
    patch = tokens[:, 5:]                      # a strided view, like x_norm_patchtokens
    boxes = o_pipe.synth_boxes(B, R, 518, seed=3, pad="dense")
    out = ops.pool_boxes(patch, boxes)
    # linearity: pooling(a x + y) = a pooling(x) + pooling(y)
    y = torch.randn_like(tokens)
    lhs = ops.pool_boxes((2.5 * tokens + y)[:, 5:], boxes)
    rhs = 2.5 * out + ops.pool_boxes(y[:, 5:], boxes)
    torch.testing.assert_close(lhs, rhs, rtol=1e-4, atol=1e-4)
    # constant tokens pool to the constant for every weighting (weights sum to 1)
    ones = torch.ones_like(tokens)
    for kw in ({}, {"gaussian_avg": True, "gaussian_bbox_variance": 1.0}):
        torch.testing.assert_close(ops.pool_boxes(ones[:, 5:], boxes, **kw), torch.ones(B, R, D, device=dev), rtol=1e-5, atol=1e-5)
    # a sample of images against the oracle
    sub = [0, 31, 63]
    ref = o_pool.extract_bboxes_feats(patch[sub].cpu(), boxes[sub], True, 1.0)
    got = ops.pool_boxes(patch[sub], boxes[sub], gaussian_avg=True, gaussian_bbox_variance=1.0).cpu()
    assert cos_min(got, ref) >= 0.9999
    torch.testing.assert_close(got, ref, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("g", [16, 37])
def test_traces_against_golden(dev, ops, golden, g):
    rec = golden("traces")[f"g{g}"]
    traces = o_pipe.synth_traces(4, seed=40 + g, n_min=20, n_max=80, outside_frac=0.1)
    traces[0] += [{"x": k / g, "y": (g - k) / g, "t": 0.0} for k in range(g + 1)]
    traces[1] += [{"x": 1.0, "y": 1.0, "t": 0}, {"x": 0.0, "y": 0.0, "t": 0}, {"x": 0.29, "y": 0.57, "t": 0}]
    counts = ops.trace_bins(traces, g, dev)
    assert torch.equal(counts.cpu(), rec["grids"])  # bins: integer exact
    gen = torch.Generator().manual_seed(300 + g)
    tok = torch.randn(4, g * g, 64, generator=gen)
    sa = torch.rand(4, g * g, generator=gen).softmax(-1)
    P = g * g
    got = ops.pool_grid(tok.to(dev), counts.reshape(4, 1, P), 1.0 / P)[:, 0]
    torch.testing.assert_close(got.cpu(), rec["pool"], rtol=1e-4, atol=1e-6)
    wa = ops.trace_bins(traces, g, dev, sa.to(dev))
    got = ops.pool_grid(tok.to(dev), wa.reshape(4, 1, P), 1.0 / P)[:, 0]
    torch.testing.assert_close(got.cpu(), rec["pool_attn"], rtol=1e-4, atol=1e-8)
    # empty trace list -> zero grid
    z = ops.trace_bins([[], [{"x": 2.0, "y": 0.5, "t": 0}]], g, dev)
    assert float(z.abs().sum()) == 0.0


def test_trace_bins_large_random(dev, ops):
    g = 37
    traces = o_pipe.synth_traces(256, seed=9)
    counts = ops.trace_bins(traces, g, dev).cpu()
    ref = torch.stack([o_pool.map_traces_to_grid(t, g) for t in traces])
    assert torch.equal(counts, ref)


def test_cls_attention_against_golden(dev, ops, golden):
    rec = golden("self_attn")
    gen = torch.Generator().manual_seed(77)
    qkv = torch.randn(2, 41, 3 * 768, generator=gen)
    sa = ops.cls_attention(qkv.to(dev))
    torch.testing.assert_close(sa.cpu(), rec["self_attn"], rtol=1e-4, atol=1e-7)


# ----------------------------------------------------------------------------------------- attention
def _attn_ref(qkv, H=12):
    B, N, C3 = qkv.shape
    D = C3 // 3
    t = qkv.float().reshape(B, N, 3, H, D // H).permute(2, 0, 3, 1, 4)
    a = ((t[0] * (D // H) ** -0.5) @ t[1].transpose(-2, -1)).softmax(dim=-1)
    return (a @ t[2]).transpose(1, 2).reshape(B, N, D)


@pytest.mark.parametrize("B,N", [(2, 261), (1, 1374), (3, 128), (1, 77)])
def test_vit_attention_fp32(dev, ops, B, N):
    qkv = torch.randn(B, N, 2304, generator=torch.Generator().manual_seed(N))
    out = ops.vit_attention(qkv.to(dev)).cpu()
    torch.testing.assert_close(out, _attn_ref(qkv), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("B,N", [(2, 261), (1, 1374), (3, 128), (1, 77), (2, 300)])
def test_vit_attention_bf16_tcgen05(dev, ops, B, N):
    qkv = (torch.randn(B, N, 2304, generator=torch.Generator().manual_seed(N)) * 1.5).bfloat16()
    out = ops.vit_attention(qkv.to(dev)).float().cpu()
    ref = _attn_ref(qkv)
    torch.testing.assert_close(out, ref, rtol=2e-2, atol=2e-2)
    assert cos_min(out, ref) >= 0.9995


@pytest.mark.parametrize("late_gain", [6.0, 40.0, 400.0])
def test_vit_attention_bf16_rising_max(dev, ops, late_gain):
    """Scores that keep growing along the key axis: the lazily updated softmax reference must be raised
    (> 2^8 above it: rescale of the tensor-memory accumulator; > 2^64: the tile is redone) and stay exact."""
    B, N = 2, 700
    g = torch.Generator().manual_seed(int(late_gain))
    qkv = torch.randn(B, N, 2304, generator=g)
    ramp = torch.linspace(1.0, late_gain, N)[None, :, None]
    qkv[:, :, 768:1536] *= ramp            # keys grow with their position
    qkv[:, 350:, 768:1536] *= 3.0          # and jump in the middle of a key tile
    qkv = qkv.bfloat16()
    out = ops.vit_attention(qkv.to(dev)).float().cpu()
    ref = _attn_ref(qkv)
    assert torch.isfinite(out).all()
    torch.testing.assert_close(out, ref, rtol=3e-2, atol=3e-2)
    assert cos_min(out, ref) >= 0.999


# ----------------------------------------------------------------------------------------- ViT
@pytest.fixture(scope="module")
def vit_w():
    return o_vit.make_weights(seed=1234)


@pytest.mark.parametrize("S", [224, 518])
def test_vit_fp32_against_oracle(dev, ops, vit_w, golden, S):
    B = 2 if S == 224 else 1
    imgs = o_pipe.synth_images(B, S, seed=1)
    ref = o_vit.forward(vit_w, imgs)
    vit = ops.Vit(vit_w, dev, "fp32")
    tokens, attn, qkv = vit.forward(imgs.to(dev), want_attn=True, want_qkv=True)
    tokens, attn, qkv = tokens.cpu(), attn.cpu(), qkv.cpu()
    ref_tok = torch.cat([ref["x_norm_clstoken"][:, None], ref["x_norm_regtokens"], ref["x_norm_patchtokens"]], 1)
    assert cos_min(tokens, ref_tok) >= 0.9999
    torch.testing.assert_close(tokens, ref_tok, rtol=2e-3, atol=2e-3)
    torch.testing.assert_close(qkv, ref["qkv"], rtol=2e-3, atol=2e-3)
    torch.testing.assert_close(attn, o_pool.cls_attention_map(ref["qkv"]), rtol=5e-3, atol=1e-6)
    if S == 224:
        g = golden("forward")
        torch.testing.assert_close(tokens[:, 0], g["vit_cls"], rtol=2e-3, atol=2e-3)


def test_vit_bf16_against_oracle(dev, ops, vit_w):
    imgs = o_pipe.synth_images(2, 224, seed=1)
    ref = o_vit.forward(vit_w, imgs)
    vit = ops.Vit(vit_w, dev, "bf16")
    tokens, attn, _ = vit.forward(imgs.to(dev), want_attn=True)
    ref_tok = torch.cat([ref["x_norm_clstoken"][:, None], ref["x_norm_regtokens"], ref["x_norm_patchtokens"]], 1)
    c = cos_min(tokens.cpu(), ref_tok)
    assert c >= 0.999, c


# ----------------------------------------------------------------------------------------- memory projection
@pytest.mark.parametrize("mode,cmin", [("fp32", 0.9999), ("bf16", 0.999)])
def test_project_against_golden(dev, ops, golden, mode, cmin):
    rec = golden("memory")
    bank = o_pipe.synth_bank(3000, 768, seed=7, zero_frac=0.002)
    gen = torch.Generator().manual_seed(8)
    q = torch.randn(16, 768, generator=gen)
    q[3] = bank[11] * 2.5 + 0.01 * torch.randn(768, generator=gen)
    b = ops.Bank(bank, dev, mode)
    assert b.M == rec["M_after_filter"]
    qd = q.to(dev)
    out = b.project(qd, normalize=True).cpu()
    assert torch.equal(qd.cpu(), q)  # the query is not normalised in place
    assert cos_min(out, rec["out_norm"]) >= cmin
    if mode == "fp32":
        torch.testing.assert_close(out, rec["out_norm"], rtol=1e-3, atol=1e-4)
        torch.testing.assert_close(b.project(qd, normalize=False).cpu(), rec["out_raw"], rtol=1e-3, atol=1e-3)


def test_project_multi_chunk_and_sharded(dev, ops, monkeypatch):
    """M spans several chunks; the sharded (m, l, O) form merged like the NCCL path equals the monolithic one."""
    monkeypatch.setenv("PIO_PROJECT_MAX_CHUNK", "16384")
    bank = o_pipe.synth_bank(40000, 768, seed=21, zero_frac=0.001)
    q = torch.randn(300, 768, generator=torch.Generator().manual_seed(22))
    ref = o_mem.project(q, o_mem.drop_zero_rows(bank), normalize=True)
    full = ops.Bank(bank, dev, "fp32")
    out = full.project(q.to(dev), normalize=True).cpu()
    assert cos_min(out, ref) >= 0.9999
    torch.testing.assert_close(out, ref, rtol=1e-3, atol=1e-4)
    shards = [ops.Bank(s, dev, "fp32") for s in bank.chunk(4)]
    parts = [s.project(q.to(dev), partial=True) for s in shards]
    m = torch.stack([p[0] for p in parts]).max(dim=0).values          # all_reduce(MAX)
    for pm, pl, pO in parts:
        ops.project_rescale_(pO, pl, pm, m)
    O = sum(p[2] for p in parts)                                      # all_reduce(SUM)
    l = sum(p[1] for p in parts)
    merged = ops.project_finish_(O, l, True).cpu()
    assert cos_min(merged, ref) >= 0.9999


def test_project_bf16_fused_multi_chunk_and_sharded(dev, ops, monkeypatch):
    """bf16 fast path (exp fused in the GEMM epilogue, lagging reference max): several chunks, a near-duplicate query
    (logit ~ 1/T = 100 -> the overflow guard), opposite-direction queries, and the sharded (m, l, O) merge."""
    monkeypatch.setenv("PIO_PROJECT_MAX_CHUNK", "16384")
    bank = o_pipe.synth_bank(70000, 768, seed=23, zero_frac=0.001)
    gen = torch.Generator().manual_seed(24)
    q = torch.randn(200, 768, generator=gen)
    q[5] = bank[60001] * 3.0 + 0.02 * torch.randn(768, generator=gen)   # match late in the bank: reference max jumps
    q[6] = bank[17] * 0.5
    q[7] = -bank[100]
    ref = o_mem.project(q, o_mem.drop_zero_rows(bank), normalize=True)
    full = ops.Bank(bank, dev, "bf16")
    out = full.project(q.to(dev), normalize=True).cpu()
    assert torch.isfinite(out).all()
    assert cos_min(out, ref) >= 0.999, cos_min(out, ref)
    shards = [ops.Bank(s, dev, "bf16") for s in bank.chunk(3)]
    parts = [s.project(q.to(dev), partial=True) for s in shards]
    m = torch.stack([p[0] for p in parts]).max(dim=0).values
    for pm, pl, pO in parts:
        ops.project_rescale_(pO, pl, pm, m)
    merged = ops.project_finish_(sum(p[2] for p in parts), sum(p[1] for p in parts), True).cpu()
    assert cos_min(merged, ref) >= 0.999, cos_min(merged, ref)


@pytest.mark.parametrize("mode,cmin", [("fp32", 0.9999), ("bf16", 0.999)])
def test_project_few_queries_large_chunks(dev, ops, mode, cmin):
    """Few queries take the large-chunk path (up to 131072 bank rows per GEMM): same result as the oracle."""
    bank = o_pipe.synth_bank(300000, 768, seed=33, zero_frac=0.001)
    q = torch.randn(7, 768, generator=torch.Generator().manual_seed(34))
    q[2] = bank[299000] * 1.7 + 0.02 * torch.randn(768, generator=torch.Generator().manual_seed(35))
    ref = o_mem.project(q, o_mem.drop_zero_rows(bank), normalize=True)
    out = ops.Bank(bank, dev, mode).project(q.to(dev), normalize=True).cpu()
    assert cos_min(out, ref) >= cmin


# ----------------------------------------------------------------------------------------- decoder
def test_decode_fp32_against_golden(dev, ops, golden):
    rec = golden("decoder")
    w = o_decap.make_weights(seed=1234)
    gen = torch.Generator().manual_seed(9)
    feats = torch.randn(6, 768, generator=gen)
    feats = feats / feats.norm(dim=-1, keepdim=True)
    dec = ops.Decoder(w, dev, "fp32")
    ids, lp = dec.decode(feats.to(dev), 30, compute_scores=True)
    assert torch.equal(ids.cpu().long(), rec["ids"])
    torch.testing.assert_close(torch.exp(lp).cpu(), rec["scores"].float(), rtol=1e-2, atol=0)


def test_decode_fp32_token_parity_rate(dev, ops):
    """>= 99 % of regions decode to identical ids in fp32 (north_star)."""
    w = o_decap.make_weights(seed=1234)
    R = 128
    feats = torch.randn(R, 768, generator=torch.Generator().manual_seed(31))
    feats = feats / feats.norm(dim=-1, keepdim=True)
    ref = o_decap.decode_greedy(w, feats, use_cache=True)
    dec = ops.Decoder(w, dev, "fp32")
    ids = dec.decode(feats.to(dev), 30).cpu().long()
    same = (ids == ref).all(dim=1).float().mean().item()
    assert same >= 0.99, same


def test_decode_bf16_runs_and_mostly_agrees(dev, ops):
    w = o_decap.make_weights(seed=1234)
    R = 64
    feats = torch.randn(R, 768, generator=torch.Generator().manual_seed(32))
    feats = feats / feats.norm(dim=-1, keepdim=True)
    ref = o_decap.decode_greedy(w, feats, use_cache=True)
    dec = ops.Decoder(w, dev, "bf16")
    ids = dec.decode(feats.to(dev), 30).cpu().long()
    assert ids.min() >= 0 and ids.max() < 50257
    first = (ids[:, 0] == ref[:, 0]).float().mean().item()
    assert first >= 0.8, first  # bf16 flips near-ties; reported, not a parity claim


def test_decode_bf16_differs_from_the_oracle_only_at_near_ties(dev, ops):
    """What the bf16 mode's caption differences ARE: with random-init weights the logits are nearly flat, and
    every FIRST position at which a bf16 caption leaves the fp32 oracle's is one where the oracle's best and second-best logit are
    within a few percent of one standard deviation of the logits (measured on B200: <= 1.6 %, the lowest ~1.5 % of all positions;
    the median position has a gap of 83 %).  Rows whose 30 positions all have a clear gap decode identically."""
    w = o_decap.make_weights(seed=1234)
    R = 256
    feats = torch.randn(R, 768, generator=torch.Generator().manual_seed(32))
    feats = feats / feats.norm(dim=-1, keepdim=True)
    ref, margin, spread = o_decap.decode_greedy(w, feats, use_cache=True, return_margin=True)
    rel = margin / spread
    ids = ops.Decoder(w, dev, "bf16").decode(feats.to(dev), 30).cpu().long()
    differing = 0
    for r in range(R):
        ne = (ids[r] != ref[r]).nonzero()
        if len(ne):
            differing += 1
            t = int(ne[0])
            assert rel[r, t] < 0.04, (r, t, rel[r, t].item())      # a near-tie, never a clear arg-max
    clear = rel.min(dim=1).values >= 0.04                        # rows without any near-tie
    assert clear.sum() >= R // 4
    assert torch.equal(ids[clear], ref[clear])
    assert differing <= R // 5                                   # ~9 % on B200


# ----------------------------------------------------------------------------------------- whole forward
def _model(dev, precision, with_bank, golden_bank=True):
    from patchioner_b200 import Patchioner

    vit_w = o_vit.make_weights(seed=1234)
    dec_w = o_decap.make_weights(seed=1234)
    bank = o_pipe.synth_bank(3000, 768, seed=7, zero_frac=0.002) if with_bank else None
    return Patchioner.from_config({"decap_weights": dec_w, "prefix_size": 768, "support_memory_size": 3000 if with_bank else 0,
                                   "dino_model": "dinov2_vitb14_reg", "normalize": True, "resize_dim": 224, "crop_dim": 224,
                                   "dino_weights": vit_w, "memory_bank": bank, "precision": precision}, device=dev)


@pytest.mark.parametrize("variant,with_bank", [("decap", True), ("capdec", False)])
def test_forward_ids_against_reference_golden(dev, golden, variant, with_bank):
    """The reference's own Patchioner.forward (tests/golden/forward.pt) vs ours, fp32: identical token ids."""
    g = golden("forward")[variant]
    m = _model(dev, "fp32", with_bank)
    B, S, R = 2, 224, 4
    imgs = o_pipe.synth_images(B, S, seed=1)
    boxes = o_pipe.synth_boxes(B, R, S, seed=1, pad="dense")
    boxes_set = o_pipe.synth_boxes(B, R, S, seed=2, pad="set")
    traces = o_pipe.synth_traces(B, seed=1)

    def ids(out, *keys):
        return torch.cat([out[k].reshape(-1, 30) for k in keys], 0).cpu().long()

    total = same = 0

    def check(got, want):
        nonlocal total, same
        total += want.shape[0]
        same += int((got == want).all(dim=1).sum())

    o = m(imgs, get_cls_capt=True, bboxes=boxes.clone(), return_ids=True)
    check(ids(o, "cls_capt", "bbox_capts"), g["cls+bbox_mean"])
    o = m(imgs, get_cls_capt=False, bboxes=boxes.clone(), gaussian_avg=True, gaussian_bbox_variance=1.0, return_ids=True)
    check(ids(o, "bbox_capts"), g["bbox_gauss1"])
    o = m(imgs, get_cls_capt=False, bboxes=boxes.clone(), use_attn_map_for_bboxes=True, return_ids=True)
    check(ids(o, "bbox_capts"), g["bbox_attn"])
    o = m(imgs, get_cls_capt=False, bboxes=boxes_set.clone(), get_controllable_capts=True, gaussian_avg=True,
          gaussian_bbox_variance=1.0, return_ids=True)
    check(ids(o, "set_controllable_capts"), g["set_gauss1"])
    o = m(imgs, get_cls_capt=False, traces=traces, return_ids=True)
    check(ids(o, "trace_capts"), g["trace"])
    o = m(imgs, get_cls_capt=False, traces=traces, use_attention_tracing=True, return_ids=True)
    check(ids(o, "trace_capts"), g["trace_attn"])
    o = m(imgs, get_cls_capt=False, get_avg_self_attn_capt=True, get_avg_patch_capt=True, gaussian_img_variance=1.0, return_ids=True)
    check(ids(o, "avg_self_attn_capt", "avg_patch_capt"), g["avg_self_attn+avg_patch"])
    assert same / total >= 0.99, (same, total)


def test_forward_strings_and_keys(dev):
    m = _model(dev, "fp32", True)
    imgs = o_pipe.synth_images(2, 224, seed=1)
    boxes = o_pipe.synth_boxes(2, 3, 224, seed=1)
    bcopy = boxes.clone()
    out = m(imgs, bboxes=boxes, traces=o_pipe.synth_traces(2, seed=1), compute_scores=True)
    assert torch.equal(boxes, bcopy)
    assert set(out) == {"cls_capt", "cls_capt_scores", "bbox_capts", "bbox_scores", "trace_capts", "trace_capts_scores"}
    assert len(out["cls_capt"]) == 2 and len(out["bbox_capts"]) == 2 and len(out["bbox_capts"][0]) == 3
    assert all(isinstance(s, str) for s in out["cls_capt"])
    seen = []
    m.decoding_method = lambda ids: (seen.append(list(ids)) or "x")
    out = m(imgs, get_cls_capt=True)
    assert out["cls_capt"] == ["x", "x"] and len(seen) == 2 and len(seen[0]) == 30


def test_masks_generalise_traces(dev):
    """masks= is pinned through the trace branch: a mask equal to a trace histogram gives the same embedding."""
    m = _model(dev, "fp32", False)
    imgs = o_pipe.synth_images(2, 224, seed=4).to(dev)
    traces = o_pipe.synth_traces(2, seed=4)
    grids = torch.stack([o_pool.map_traces_to_grid(t, 16) for t in traces])
    e = m.region_embeddings(imgs, traces=traces, masks=grids[:, None])
    torch.testing.assert_close(e["mask"][:, 0], e["trace"], rtol=1e-6, atol=1e-7)


def test_forward_bf16_embeddings(dev):
    """bf16 mode: region embeddings within cosine >= 0.999 of the fp32 oracle (north_star)."""
    m = _model(dev, "bf16", False)
    vit_w = o_vit.make_weights(seed=1234)
    imgs = o_pipe.synth_images(2, 224, seed=1)
    boxes = o_pipe.synth_boxes(2, 4, 224, seed=1, pad="dense")
    d = o_vit.forward(vit_w, imgs)
    ref = o_pool.extract_bboxes_feats(d["x_norm_patchtokens"], boxes, True, 1.0)
    e = m.region_embeddings(imgs.to(dev), bboxes=boxes, gaussian_avg=True, gaussian_bbox_variance=1.0)
    c = cos_min(e["bbox"].cpu(), ref)
    assert c >= 0.999, c


# ----------------------------------------------------------------------------------------- widened forward modes
def test_cls_head_attention_maps(dev, ops):
    """Per-"head" CLS maps (16 x 48-channel re-cut) and the disentangled tokens of model.py:871-872 vs the oracle's literal
    restatement of process_self_attention (itself pinned to the reference's output in tests/golden/self_attention.pt)."""
    gen = torch.Generator().manual_seed(31)
    B, g = 2, 16
    N = 5 + g * g
    qkv = torch.randn(B, N, 2304, generator=gen)
    patch = torch.randn(B, g * g, 768, generator=gen)
    _, maps = o_pool.process_self_attention_literal(qkv)
    want_maps = maps.softmax(dim=-1)
    want_tok = (patch.unsqueeze(1) * want_maps.unsqueeze(-1)).mean(dim=2)
    got_maps = ops.cls_head_attention(qkv.to(dev))
    torch.testing.assert_close(got_maps.cpu(), want_maps, rtol=1e-4, atol=1e-7)
    got_tok = ops.pool_grid(patch.to(dev), got_maps, 1.0 / (g * g))
    torch.testing.assert_close(got_tok.cpu(), want_tok, rtol=1e-4, atol=1e-7)
    got16 = ops.cls_head_attention(qkv.to(dev).bfloat16())
    torch.testing.assert_close(got16.cpu(), want_maps, rtol=5e-2, atol=1e-4)


@pytest.mark.parametrize("mode,tol", [("fp32", 1e-5), ("bf16", 4e-3)])
def test_best_sims(dev, ops, mode, tol, monkeypatch):
    """return_n_best_sims (im2txtprojection.py:382-383): the n largest cosines per query, descending, over several chunks."""
    monkeypatch.setenv("PIO_PROJECT_MAX_CHUNK", "16384")
    bank = o_pipe.synth_bank(40000, 768, seed=41, zero_frac=0.001)
    fb = o_mem.drop_zero_rows(bank)
    gen = torch.Generator().manual_seed(42)
    q = torch.randn(50, 768, generator=gen)
    q[3] = fb[39000] * 2.0 + 0.05 * torch.randn(768, generator=gen)
    q[4] = fb[5] + fb[20000]
    sim = (q / q.norm(dim=-1, keepdim=True)) @ (fb / fb.norm(dim=-1, keepdim=True)).T
    want = sim.sort(dim=-1, descending=True)
    b = ops.Bank(bank, dev, mode)
    sims, rows = b.best_sims(q.to(dev), 5, with_rows=True)
    torch.testing.assert_close(sims.cpu(), want.values[:, :5], rtol=0, atol=tol)
    assert (sims[:, :-1] >= sims[:, 1:]).all()
    if mode == "fp32":
        assert torch.equal(rows.cpu().long(), want.indices[:, :5])
    else:
        assert torch.equal(rows[3:5, 0].cpu().long(), want.indices[3:5, 0])
    one = b.best_sims(q.to(dev), 1)
    torch.testing.assert_close(one.cpu(), want.values[:, :1], rtol=0, atol=tol)


def test_forward_argmax_text_and_attn_heads(dev):
    """calculate_argmax_text (model.py:1408-1411) with return_n_best_sims, and get_attn_heads_capt (model.py:950-960)."""
    from patchioner_b200 import Patchioner

    vit_w, dec_w = o_vit.make_weights(seed=1234), o_decap.make_weights(seed=1234)
    bank = o_pipe.synth_bank(3000, 768, seed=7, zero_frac=0.002)
    fb = o_mem.drop_zero_rows(bank)
    texts = [f"caption {i}".encode() for i in range(bank.shape[0])]
    cfg = {"decap_weights": dec_w, "prefix_size": 768, "support_memory_size": 3000, "dino_model": "dinov2_vitb14_reg",
           "normalize": True, "resize_dim": 224, "crop_dim": 224, "dino_weights": vit_w, "memory_bank": bank, "precision": "fp32"}
    imgs = o_pipe.synth_images(2, 224, seed=1)
    boxes = o_pipe.synth_boxes(2, 3, 224, seed=1, pad="dense")
    d = o_vit.forward(vit_w, imgs)
    feats = o_pool.extract_bboxes_feats(d["x_norm_patchtokens"], boxes.clone(), False, 0.5).reshape(-1, 768)
    sim = (feats / feats.norm(dim=-1, keepdim=True)) @ (fb / fb.norm(dim=-1, keepdim=True)).T
    want = sim.sort(dim=-1, descending=True)

    m = Patchioner.from_config(dict(cfg, calculate_argmax_text=True, memory_bank_texts=texts), device=dev)
    out = m(imgs, get_cls_capt=True, bboxes=boxes, return_n_best_sims=3, compute_scores=True)
    assert set(out) == {"cls_capt", "cls_capt_scores", "bbox_capts", "bbox_scores", "bbox_sims"}
    flat = [c for per_img in out["bbox_capts"] for c in per_img]
    assert flat == [f"caption {i}" for i in want.indices[:, 0].tolist()]   # texts are indexed by the FILTERED row, like the reference
    torch.testing.assert_close(torch.tensor(out["bbox_sims"]).reshape(-1, 3), want.values[:, :3], rtol=0, atol=2e-4)
    assert out["bbox_scores"] == [[1.0] * 3] * 2 and out["cls_capt_scores"] == [1.0, 1.0]

    m2 = Patchioner.from_config(cfg, device=dev)
    with pytest.raises(ValueError):
        m2(imgs, bboxes=boxes, return_n_best_sims=3)      # decoder path: the reference fails too (model.py:1033)
    o2 = m2(imgs, get_cls_capt=False, get_attn_heads_capt=True, return_ids=True)
    assert o2["attn_heads_capts"].shape == (2, 16, 30)
    _, maps = o_pool.process_self_attention_literal(d["qkv"])
    tok = (d["x_norm_patchtokens"].unsqueeze(1) * maps.softmax(dim=-1).unsqueeze(-1)).mean(dim=2).reshape(-1, 768)
    ref = o_pipe.OracleModel(vit_w, dec_w, bank).caption_tokens(tok)
    agree = (o2["attn_heads_capts"].reshape(-1, 30).cpu().long() == ref).all(dim=1).float().mean().item()
    assert agree >= 0.99, agree


def test_forward_pipelined_matches_forward(dev):
    """The serving loop (host->device copy of batch i+1 under batch i's kernels) returns what forward returns, in order."""
    m = _model(dev, "fp32", True)
    batches = [{"imgs": o_pipe.synth_images(2, 224, seed=s).pin_memory(), "bboxes": o_pipe.synth_boxes(2, 3, 224, seed=s).pin_memory()}
               for s in (1, 2, 3)]
    want = [m(b["imgs"], bboxes=b["bboxes"].clone(), get_cls_capt=False, return_ids=True)["bbox_capts"].cpu() for b in batches]
    got = [o["bbox_capts"].cpu() for o in m.forward_pipelined(iter(batches), get_cls_capt=False, return_ids=True)]
    assert len(got) == 3 and all(torch.equal(a, b) for a, b in zip(got, want))


@pytest.mark.parametrize("kind", ["orthogonal_projection", "contrastive_mask"])
def test_ctx_clean(dev, ops, kind):
    gen = torch.Generator().manual_seed(51)
    tok = torch.randn(3, 5 + 100, 768, generator=gen)
    d = tok[:, 5:]                      # strided view, like x_norm_patchtokens
    c = torch.randn(3, 768, generator=gen)
    want = o_pool.ctx_cleaner(d, c, kind, alpha=0.7)
    got = ops.ctx_clean(tok.to(dev)[:, 5:], c.to(dev), kind, alpha=0.7)
    torch.testing.assert_close(got.cpu(), want, rtol=1e-4, atol=1e-5)
    dn, cn = d / d.norm(dim=-1, keepdim=True), c / c.norm(dim=-1, keepdim=True)
    got = ops.ctx_clean(tok.to(dev)[:, 5:], c.to(dev), kind, alpha=0.7, prenorm=True)
    torch.testing.assert_close(got.cpu(), o_pool.ctx_cleaner(dn, cn, kind, alpha=0.7), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("after", [True, False])
def test_forward_cleaning_type(dev, after):
    """cleaning_type (model.py:879-922): patch tokens replaced by context-cleaned, projected tokens; box captions skip the projection."""
    m = _model(dev, "fp32", True)
    vit_w, dec_w = o_vit.make_weights(seed=1234), o_decap.make_weights(seed=1234)
    bank = o_mem.drop_zero_rows(o_pipe.synth_bank(3000, 768, seed=7, zero_frac=0.002))
    imgs = o_pipe.synth_images(2, 224, seed=1)
    boxes = o_pipe.synth_boxes(2, 3, 224, seed=1, pad="dense")
    d = o_vit.forward(vit_w, imgs)
    patch, cls = d["x_norm_patchtokens"], d["x_norm_clstoken"]
    if after:
        cleaned = o_pool.ctx_cleaner(o_mem.project(patch, bank, normalize=True), o_mem.project(cls, bank, normalize=True),
                                     "orthogonal_projection", 0.8)
    else:
        cleaned = o_mem.project(o_pool.ctx_cleaner(patch / patch.norm(dim=-1, keepdim=True), cls / cls.norm(dim=-1, keepdim=True),
                                                   "orthogonal_projection", 0.8), bank, normalize=True)
    feats = o_pool.extract_bboxes_feats(cleaned, boxes.clone(), False, 0.5).reshape(-1, 768)
    want = o_pipe.OracleModel(vit_w, dec_w, bank).caption_tokens(feats, project=False)
    out = m(imgs, get_cls_capt=False, bboxes=boxes, cleaning_type="orthogonal_projection", alpha=0.8, clean_after_projection=after,
            return_ids=True)
    got = out["bbox_capts"].reshape(-1, 30).cpu().long()
    assert (got == want).all(dim=1).float().mean().item() >= 0.99


def test_caption_surface_maps_onto_forward(dev):
    """caption(caption_from=..., use_gaussian_weighting=..., use_attention_weighting=...) == forward with the mapped flags."""
    from patchioner_b200 import AutoModel

    vit_w, dec_w = o_vit.make_weights(seed=1234), o_decap.make_weights(seed=1234)
    m = AutoModel.from_pretrained({"decap_weights": dec_w, "prefix_size": 768, "support_memory_size": 0, "dino_model": "dinov2_vitb14_reg",
                                   "resize_dim": 224, "crop_dim": 224, "dino_weights": vit_w, "precision": "fp32"}, device=dev)
    imgs = o_pipe.synth_images(2, 224, seed=3)
    boxes = o_pipe.synth_boxes(2, 3, 224, seed=3)
    traces = o_pipe.synth_traces(2, seed=3)
    a = m.caption(imgs, "patches", bboxes=boxes, use_gaussian_weighting=True, gaussian_variance=1.0, return_ids=True)
    b = m(imgs, get_cls_capt=False, bboxes=boxes, gaussian_avg=True, gaussian_bbox_variance=1.0, return_ids=True)
    assert set(a) == {"bbox_capts"} and torch.equal(a["bbox_capts"], b["bbox_capts"])
    a = m.caption(imgs, "patches", traces=traces, use_attention_weighting=True, return_ids=True)
    b = m(imgs, get_cls_capt=False, traces=traces, use_attention_tracing=True, return_ids=True)
    assert torch.equal(a["trace_capts"], b["trace_capts"])
    a = m.caption(imgs, "patches", bboxes=boxes, region_sets=True, return_ids=True)
    assert a["set_controllable_capts"].shape == (2, 30)
    assert set(m.caption(imgs, "cls", return_ids=True)) == {"cls_capt"}
    assert set(m.caption(imgs, "avg_self_attn", return_ids=True)) == {"avg_self_attn_capt"}
    assert set(m.caption(imgs, "patches", return_ids=True)) == {"avg_patch_capt"}
    from PIL import Image
    pil = [Image.new("RGB", (320, 240), (200, 30, 90)), Image.new("RGB", (100, 300), (5, 5, 5))]
    assert m.preprocess(pil, keep_img_ratio=True).shape == (2, 3, 224, 224) and m.preprocess(pil, keep_img_ratio=False).shape == (2, 3, 224, 224)


def test_variance_zero_centre_patch_against_reference_golden(dev, ops, golden):
    """gaussian variance 0 (bbox_utils.py:62-71, model.py:71-79): one-hot on a central patch, python ``random`` decides for even
    spans.  With ``random`` seeded like the generator (tests/golden/make_golden_centre.py, unmodified reference) the device path
    returns the reference's rows: dense mode is a gather (bit-exact), box-set mode a weighted sum (fp32 round-off)."""
    import random

    from oracle import pooling as o_pool

    centre = _golden_script("make_golden_centre")
    rec_all = golden("centre")
    for name in ("g16", "g37"):
        rec = rec_all[name]
        B, g, R, D = rec["shape"]
        tok, bd, bs = centre.inputs(g, B, R, D)
        assert float(tok.double().sum()) == rec["in_tok_sum"]
        t = tok.to(dev)
        random.seed(rec_all["seed"])
        got = ops.pool_boxes(t, bd.to(dev), 14, True, 0.0).cpu()
        assert torch.equal(got, rec["dense"])
        random.seed(rec_all["seed"])
        assert torch.equal(o_pool.extract_bboxes_feats(tok, bd, True, 0.0), rec["dense"])
        random.seed(rec_all["seed"])
        got = ops.pool_boxes(t, bs.to(dev), 14, True, 0.0, get_single_embedding_per_image=True).cpu()
        torch.testing.assert_close(got, rec["set"], rtol=2e-5, atol=2e-6)
        random.seed(rec_all["seed"])
        assert torch.equal(ops.region_centre_rows(t).cpu(), rec["region_means_0"])
        # a strided view (patch tokens behind cls + registers, as the forward hands them over)
        full = torch.cat([torch.zeros(B, 5, D), tok], dim=1).to(dev)
        random.seed(rec_all["seed"])
        assert torch.equal(ops.pool_boxes(full[:, 5:], bd.to(dev), 14, True, 0.0).cpu(), rec["dense"])


def test_cabi_rejects_bad_arguments(dev, ops):
    """Error behaviour of the C ABI: status code + message through PioError, nothing launched, no crash."""
    from patchioner_b200 import PioError

    x = torch.randn(4, 5 * 5, 768, device=dev)
    boxes = torch.tensor([[[0.0, 0.0, 28.0, 28.0]]] * 4, device=dev)
    with pytest.raises(PioError, match="variance 0"):
        ops.region_mean_weights(5, 0.0, dev)                         # the C entry has no weight form for the python-random centre (ops.region_centre_rows is the path)
    A = torch.randn(64, 100, device=dev).bfloat16()                   # K = 100: 200-byte rows, not a legal TMA stride
    W = torch.randn(32, 100, device=dev).bfloat16()
    with pytest.raises(PioError, match="multiples of 8"):
        ops.linear(A, W, "bf16")
    bank = ops.Bank(o_pipe.synth_bank(300, 768, seed=1), dev, "fp32")
    with pytest.raises(PioError, match="outside"):
        bank.best_sims(torch.randn(3, 768, device=dev), 33)
    w = o_decap.make_weights(seed=1234)
    dec = ops.Decoder(w, dev, "fp32")
    with pytest.raises(PioError, match="steps"):
        dec.decode(torch.randn(2, 768, device=dev), 33)
    # empty inputs are fine and launch nothing
    m = _model(dev, "fp32", True)
    o = m(o_pipe.synth_images(2, 224, seed=1), get_cls_capt=False, bboxes=torch.zeros(2, 0, 4))
    assert o["bbox_capts"] == [[], []]
    assert ops.pool_boxes(x[:0], boxes[:0], 14).shape == (0, 1, 768)
    assert bank.project(torch.randn(0, 768, device=dev)).shape == (0, 768)


def test_decode_repeatable(dev, ops):
    """Repeated decodes of the same prefixes through recycled buffers: identical ids and scores, identical launch counts
    (no state leaks from one call into the next through the KV cache or the workspace)."""
    w = o_decap.make_weights(seed=1234)
    gen = torch.Generator().manual_seed(77)
    for mode in ("bf16", "fp32"):
        dec = ops.Decoder(w, dev, mode)
        pre = torch.randn(37, 768, generator=gen).to(dev)
        first, counts = None, []
        for i in range(5):
            ops.reset_launch_count()
            r = dec.decode(pre, 30, True)
            torch.cuda.synchronize()
            counts.append(ops.launch_count())
            got = (r[0].clone(), r[1].clone())
            del r  # the result buffers go back to the allocator, so the next call sees the same pointers
            if first is None:
                first = got
            assert torch.equal(got[0], first[0]) and torch.allclose(got[1], first[1], rtol=1e-5, atol=1e-6)
        assert len(set(counts)) == 1, counts


def test_caption_bboxes_crops(dev):
    """caption_bboxes_type (model.py:1356-1390): PIL crop of every box -> transform -> caption of the crop's CLS token."""
    from PIL import Image
    import numpy as np

    m = _model(dev, "fp32", False)
    seen = []
    m.decoding_method = lambda ids: (seen.append(list(ids)) or "c%d" % len(seen))
    rng = np.random.RandomState(3)
    pil = [Image.fromarray(rng.randint(0, 255, (260, 300, 3), dtype=np.uint8)), Image.fromarray(rng.randint(0, 255, (240, 240, 3), dtype=np.uint8))]
    boxes = torch.tensor([[[10.0, 20.0, 100.0, 80.0], [50.0, 50.0, 120.0, 150.0]], [[0.0, 0.0, 240.0, 240.0], [30.0, 40.0, 60.0, 60.0]]])
    out = m(pil, bboxes=boxes, caption_bboxes_type="cls_capt")
    assert set(out) == {"bbox_capts"} and [len(r) for r in out["bbox_capts"]] == [2, 2]
    got = [list(r) for r in seen]
    crops = torch.stack([m.image_transforms_no_crop(im.crop((x, y, x + w, y + h))) for im, bb in zip(pil, boxes.tolist()) for (x, y, w, h) in bb])
    want = m(crops, get_cls_capt=True, return_ids=True)["cls_capt"].cpu().tolist()
    assert got == want


@pytest.mark.parametrize("keep", [True, False])
def test_preprocess_on_device_matches_torchvision(dev, keep):
    """pio_preprocess vs the reference's own torchvision pipeline (src/model.py:347-357) on images of mixed sizes: the resized
    bytes are Pillow's bit for bit, so the normalised floats are identical."""
    from PIL import Image
    import numpy as np
    from oracle import preprocess as o_pre
    from patchioner_b200 import preprocess as pre

    rng = np.random.RandomState(5)
    sizes = [(480, 640), (640, 480), (480, 640), (333, 1001), (600, 600), (230, 300)]
    pil = [Image.fromarray(rng.randint(0, 256, (h, w, 3), dtype=np.uint8)) for (h, w) in sizes]
    for resize_dim, crop_dim in ((518, 518), (224, 224)):
        want = o_pre.reference_transform(pil, resize_dim, crop_dim, keep)
        got = pre.preprocess_images(pil, dev, resize_dim, crop_dim, keep).cpu()
        assert got.shape == want.shape
        assert torch.equal(got, want), (got - want).abs().max()
    m = _model(dev, "fp32", False)
    assert torch.equal(m.preprocess(pil, keep_img_ratio=keep, on_device=True).cpu(), m.preprocess(pil, keep_img_ratio=keep))


@pytest.mark.parametrize("mode,tol", [("fp32", 2e-3), ("bf16", 6e-2)])
def test_double_dino_feats(dev, mode, tol):
    """double_DINO_for_bboxes (bbox_utils.py:300-403): the last block re-run on [cls | registers | box patches] per box."""
    m = _model(dev, mode, False)
    vit_w = o_vit.make_weights(seed=1234)
    imgs = o_pipe.synth_images(2, 224, seed=6)
    boxes = o_pipe.synth_boxes(2, 5, 224, seed=6)
    boxes[0, 0] = torch.tensor([20.0, 30.0, 150.0, 120.0])   # x < w, y < h: a non-empty [y : h + 1, x : w + 1] slice
    boxes[1, 1] = torch.tensor([0.0, 0.0, 223.0, 223.0])
    boxes[1, 2] = torch.tensor([200.0, 200.0, 10.0, 10.0])   # end before start: empty region
    d = o_vit.forward(vit_w, imgs)
    block = lambda x: o_vit.block_forward(vit_w, 11, x)      # noqa: E731
    tokens, _, _ = m.dino.forward(imgs.to(dev), want_attn=False)
    for rt, use_cls in (("avg", True), ("cls", True), ("avg", False), ("gaussian_avg", True)):
        want = o_pool.extract_bboxes_feats_double_dino(block, d["x_norm_patchtokens"], boxes.clone(),
                                                       d["x_norm_clstoken"] if use_cls else None,
                                                       d["x_norm_regtokens"] if use_cls else None, 14, rt, 0.5)
        got = m.double_dino_feats(tokens, boxes, rt, use_cls, 0.5).cpu()
        ok = ~torch.isnan(want).any(-1)
        assert torch.equal(torch.isnan(got).any(-1), ~ok), rt
        if mode == "fp32":
            torch.testing.assert_close(got[ok], want[ok], rtol=tol, atol=tol)
        nz = ok & (want.abs().sum(-1) > 0)  # an empty rectangle pools to exact zeros in 'gaussian_avg'
        assert (got[ok & ~nz] == 0).all()
        assert cos_min(got[nz], want[nz]) >= (0.9999 if mode == "fp32" else 0.999), (rt, use_cls)
    out = m(imgs, get_cls_capt=False, bboxes=boxes, double_DINO_for_bboxes=True, double_DINO_use_cls=True, return_ids=True)
    assert out["bbox_capts"].shape == (2, 5, 30)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_nan_embedding_decodes_like_torch_argmax(dev, ops, mode):
    """An empty box pools to NaN (like the reference); its NaN logits must decode to token 0 at every step, as torch.argmax does
    (first NaN index), instead of indexing the embedding table out of bounds."""
    dec = ops.Decoder(o_decap.make_weights(seed=1234), dev, mode)
    pre = torch.randn(5, 768, generator=torch.Generator().manual_seed(3))
    pre[2] = float("nan")
    ids = dec.decode(pre.to(dev), 30).cpu()
    assert (ids[2] == 0).all()
    clean = dec.decode(pre[[0, 1, 3, 4]].to(dev), 30).cpu()
    assert torch.equal(ids[[0, 1, 3, 4]], clean)      # the NaN row does not disturb its neighbours
    m = _model(dev, mode, True)
    out = m(o_pipe.synth_images(1, 224, seed=1), get_cls_capt=False, bboxes=torch.tensor([[[500.0, 500.0, 10.0, 10.0], [0.0, 0.0, 100.0, 100.0]]]),
            return_ids=True)
    assert (out["bbox_capts"][0, 0] == 0).all()


def test_full_size_bf16_against_fp32_mode(dev):
    """BASELINE configs[1] geometry (518 px, 64 boxes / image; 8 images to keep the fp32 SIMT arm short): the bf16 tensor-core path
    stays within the north-star tolerance of the fp32 parity mode at every stage -- tokens, region embeddings, projected prefixes."""
    m32, m16 = _model(dev, "fp32", True), _model(dev, "bf16", True)
    m32.resize_dim = m32.crop_dim = m16.resize_dim = m16.crop_dim = 518
    imgs = o_pipe.synth_images(8, 518, seed=21).to(dev)
    boxes = o_pipe.synth_boxes(8, 64, 518, seed=21, pad="dense")
    t32, a32, _ = m32.dino.forward(imgs)
    t16, a16, _ = m16.dino.forward(imgs)
    assert cos_min(t16.cpu(), t32.cpu()) >= 0.999
    torch.testing.assert_close(a16.cpu(), a32.cpu(), rtol=5e-2, atol=1e-5)
    e32 = m32.region_embeddings(imgs, bboxes=boxes, gaussian_avg=True, gaussian_bbox_variance=1.0)["bbox"].reshape(-1, 768)
    e16 = m16.region_embeddings(imgs, bboxes=boxes, gaussian_avg=True, gaussian_bbox_variance=1.0)["bbox"].reshape(-1, 768)
    ok = ~torch.isnan(e32).any(-1)
    assert cos_min(e16[ok].cpu(), e32[ok].cpu()) >= 0.999
    p32, p16 = m32.embed_tokens(e32[ok]), m16.embed_tokens(e32[ok])    # same queries: isolates the projection
    assert cos_min(p16.cpu(), p32.cpu()) >= 0.999


def test_deterministic_and_batch_invariant(dev):
    """Same inputs twice -> identical ids (bf16 and fp32); in fp32 mode the caption of an image does not depend on its batch."""
    imgs = o_pipe.synth_images(4, 224, seed=31)
    boxes = o_pipe.synth_boxes(4, 3, 224, seed=31, pad="dense")
    for mode in ("bf16", "fp32"):
        m = _model(dev, mode, True)
        a = m(imgs, get_cls_capt=True, bboxes=boxes, return_ids=True)
        b = m(imgs, get_cls_capt=True, bboxes=boxes, return_ids=True)
        assert torch.equal(a["bbox_capts"], b["bbox_capts"]) and torch.equal(a["cls_capt"], b["cls_capt"])
        if mode == "fp32":
            one = m(imgs[2:3], get_cls_capt=True, bboxes=boxes[2:3], return_ids=True)
            assert torch.equal(one["bbox_capts"][0], a["bbox_capts"][2]) and torch.equal(one["cls_capt"][0], a["cls_capt"][2])


@pytest.mark.parametrize("M,N,K", [(32, 768, 16384), (200, 768, 40000), (128, 192, 8192)])
def test_linear_split_k_accumulate(dev, ops, M, N, K):
    """Long-K accumulation with few output tiles: the GEMM splits K over the SMs and the partial tiles meet through the bulk
    reduce-add; same result as the unsplit product."""
    g = torch.Generator().manual_seed(K)
    A = (torch.randn(M, K, generator=g) / 8).bfloat16()
    W = (torch.randn(N, K, generator=g) / 8).bfloat16()
    C0 = torch.randn(M, N, generator=g)
    ref = C0.double() + A.double() @ W.double().T
    X = C0.clone().to(dev)
    ops.linear(A.to(dev), W.to(dev), "bf16", residual=X, out=X)
    torch.testing.assert_close(X.cpu().double(), ref, rtol=2e-3, atol=2e-2)


@pytest.mark.parametrize("M,N,K", [(64, 768, 3072), (256, 768, 3072), (130, 1000, 2048), (1, 768, 1536), (300, 2304, 1536), (64, 2304, 768)])
def test_linear_bf16_deterministic_split_k(dev, ops, M, N, K):
    """Long-K GEMMs with few output tiles are cut along K; the partial tiles are summed in split order by the last-arriving
    CTA, so the result is exact to accumulation order AND bit-identical from run to run (greedy decoding must be repeatable)."""
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g).bfloat16()
    W = (torch.randn(N, K, generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, generator=g)
    res = torch.randn(M, N, generator=g)
    ref = A.float() @ W.float().T
    Ad, Wd = A.to(dev), W.to(dev)
    out = ops.linear(Ad, Wd, "bf16")
    torch.testing.assert_close(out.cpu(), ref, rtol=2e-3, atol=2e-3)
    for _ in range(3):
        assert torch.equal(ops.linear(Ad, Wd, "bf16"), out)
    # the decode step's fc2: x += A W^T + b in place, fp32 residual stream
    X = res.clone().to(dev)
    ops.linear(Ad, Wd, "bf16", bias=bias.to(dev), residual=X, out=X)
    torch.testing.assert_close(X.cpu(), res + ref + bias, rtol=2e-3, atol=2e-3)
    X2 = res.clone().to(dev)
    ops.linear(Ad, Wd, "bf16", bias=bias.to(dev), residual=X2, out=X2)
    assert torch.equal(X, X2)
    # bf16 output with a column scale, and on a second stream (its own workspace)
    gamma = torch.rand(N, generator=g)
    o16 = ops.linear(Ad, Wd, "bf16", bias=bias.to(dev), gamma=gamma.to(dev), out_dtype=torch.bfloat16)
    torch.testing.assert_close(o16.float().cpu(), gamma * (ref + bias), rtol=2e-2, atol=2e-2)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        o2 = ops.linear(Ad, Wd, "bf16")
    s.synchronize()
    assert torch.equal(o2, out)


# ----------------------------------------------------------------------------------------- round-2 evidence
def _golden_script(name):
    import importlib.util
    import os

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".py")
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("mode,rtol,atol,cmin", [("fp32", 2e-4, 2e-4, 0.99999), ("bf16", 3e-2, 3e-2, 0.999)])
def test_embed_inversion_against_reference_golden(dev, golden, mode, rtol, atol, cmin):
    """Patchioner(talk2dino_weights=...).embed_tokens == revert_transformation (embedding_utils.py:17-24)
    of the unmodified reference (tests/golden/make_golden_revert.py)."""
    from patchioner_b200 import Patchioner

    rec = golden("revert")
    A, b, x = _golden_script("make_golden_revert").inputs()
    assert [float(t.double().sum()) for t in (A, b, x)] == rec["in_sums"]
    m = Patchioner.from_config({"decap_weights": o_decap.make_weights(seed=1234), "prefix_size": 768, "support_memory_size": 0,
                                "dino_model": "dinov2_vitb14_reg", "normalize": True, "resize_dim": 224, "crop_dim": 224,
                                "dino_weights": o_vit.make_weights(seed=1234), "precision": mode,
                                "talk2dino_weights": {"linear_layer.weight": A, "linear_layer.bias": b}}, device=dev)
    assert m.embed_inversion
    torch.testing.assert_close(m.talk2dino_A_pinv[:8].cpu(), rec["A_pinv_head"], rtol=1e-4, atol=1e-6)
    got = m.embed_tokens(x.to(dev)).float().cpu()
    assert got.shape == (24, 512)
    assert cos_min(got, rec["revert"]) >= cmin
    torch.testing.assert_close(got, rec["revert"], rtol=rtol, atol=atol)


def test_patch_and_register_captions_against_oracle(dev):
    """SURVEY 8f.4 / src/model.py:957-979: get_patch_capts (every patch token, grouped per image) and get_register_capts
    (the 4 register tokens) go through caption_tokens like the reference; ids vs the CPU oracle, fp32 mode."""
    m = _model(dev, "fp32", True)
    vit_w, dec_w = o_vit.make_weights(seed=1234), o_decap.make_weights(seed=1234)
    bank = o_pipe.synth_bank(3000, 768, seed=7, zero_frac=0.002)
    imgs = o_pipe.synth_images(2, 224, seed=5)
    out = m(imgs, get_cls_capt=False, get_patch_capts=True, get_register_capts=True, return_ids=True, compute_scores=True)
    assert set(out) == {"patch_tokens_capts", "patch_tokens_scores", "register_capts", "register_scores"}
    assert out["patch_tokens_capts"].shape == (2, 256, 30) and out["register_capts"].shape == (2, 4, 30)
    d = o_vit.forward(vit_w, imgs)
    om = o_pipe.OracleModel(vit_w, dec_w, bank)
    # all 8 register rows, and a spread of 96 patch rows (the oracle decodes on the CPU)
    ref_reg, ref_reg_sc = om.caption_tokens(d["x_norm_regtokens"].reshape(-1, 768), compute_scores=True)
    got_reg = out["register_capts"].reshape(-1, 30).cpu().long()
    assert (got_reg == ref_reg).all(dim=1).float().mean().item() >= 0.99
    torch.testing.assert_close(torch.as_tensor(out["register_scores"], dtype=torch.float32).reshape(-1),
                               torch.as_tensor(ref_reg_sc, dtype=torch.float32).reshape(-1), rtol=2e-3, atol=1e-30)
    sel = torch.arange(0, 512, 16).tolist() + torch.arange(5, 512, 8).tolist()
    ref_patch = om.caption_tokens(d["x_norm_patchtokens"].reshape(-1, 768)[sel])
    got_patch = out["patch_tokens_capts"].reshape(-1, 30).cpu().long()[sel]
    agree = (got_patch == ref_patch).all(dim=1).float().mean().item()
    assert agree >= 0.99, agree
    # strings: grouped [B][P] / [B][4] lists like the reference
    s = m(imgs[:1], get_cls_capt=False, get_patch_capts=True, get_register_capts=True)
    assert len(s["patch_tokens_capts"]) == 1 and len(s["patch_tokens_capts"][0]) == 256 and len(s["register_capts"][0]) == 4


def test_vit_bf16_against_oracle_518(dev, ops, vit_w):
    """bf16 tensor-core ViT at the BASELINE geometry (518 px: N = 1374, the 128-key attention tiles) against the CPU ORACLE, B = 1."""
    imgs = o_pipe.synth_images(1, 518, seed=3)
    ref = o_vit.forward(vit_w, imgs)
    vit = ops.Vit(vit_w, dev, "bf16")
    tokens, attn, _ = vit.forward(imgs.to(dev), want_attn=True)
    ref_tok = torch.cat([ref["x_norm_clstoken"][:, None], ref["x_norm_regtokens"], ref["x_norm_patchtokens"]], 1)
    c = cos_min(tokens.cpu(), ref_tok)
    assert c >= 0.999, c
    torch.testing.assert_close(attn.cpu(), o_pool.cls_attention_map(ref["qkv"]), rtol=8e-2, atol=2e-5)


def test_linear_chained_dynamic_weight(dev, ops):
    """W produced by the previous call on the same stream (w_static = 0, the default) must be waited for --
    the weight-tile prefetch ahead of the dependent-launch wait is only taken when the caller vouches for a static W."""
    g = torch.Generator().manual_seed(77)
    a = torch.randn(512, 768, generator=g).to(dev).to(torch.bfloat16)
    w1 = (torch.randn(768, 768, generator=g) / 28).to(dev).to(torch.bfloat16)
    x = torch.randn(300, 768, generator=g).to(dev).to(torch.bfloat16)
    for _ in range(5):
        y = ops.linear(a, w1, "bf16", out_dtype=torch.bfloat16)             # [512, 768], then used as W
        z = ops.linear(x, y, "bf16")                                        # [300, 512]
        ref = x.float() @ (a.float() @ w1.float().T).to(torch.bfloat16).float().T
        torch.testing.assert_close(z, ref, rtol=2e-2, atol=2e-1)
        z2 = ops.linear(x, y, "bf16", w_static=True)                        # y is complete by now: same numbers
        assert torch.equal(z, z2)


def test_forward_pipelined_strings_match_forward(dev):
    """The string-returning serving loop (ids -> pinned host copy on the batch's stream -> batched detokenisation at
    consumption time) returns exactly what forward() returns, grouped [B][R] / [B] like the reference."""
    m = _model(dev, "bf16", True)
    batches = [{"imgs": o_pipe.synth_images(2, 224, seed=s).pin_memory(), "bboxes": o_pipe.synth_boxes(2, 3, 224, seed=s).pin_memory()}
               for s in (1, 2, 3, 4)]
    want = [m(b["imgs"], bboxes=b["bboxes"].clone(), get_cls_capt=True) for b in batches]
    got = list(m.forward_pipelined(iter(batches), get_cls_capt=True))
    assert len(got) == 4
    for a, b in zip(got, want):
        assert a["bbox_capts"] == b["bbox_capts"] and a["cls_capt"] == b["cls_capt"]
        assert len(a["bbox_capts"]) == 2 and len(a["bbox_capts"][0]) == 3 and isinstance(a["bbox_capts"][0][0], str)
    m.decoding_method = lambda ids: "|".join(str(i) for i in ids[:2])      # the reference's hook (model.py:105), row by row
    hooked = list(m.forward_pipelined(iter(batches[:2]), get_cls_capt=False))
    assert all(len(s.split("|")) == 2 for s in hooked[0]["bbox_capts"][0])


def test_vit_bf16_is_reproducible_run_to_run(dev, ops, vit_w):
    """Round-2 finding: launched as a programmatic dependent of the qkv GEMM, the tcgen05 attention kernel made the bf16 ViT differ
    from run to run at exactly this shape (4 x 224 px: 12 of 29 forwards, max |diff| 0.06).  The attention kernel is now launched
    fully serialised and pdl_wait() carries an async-proxy fence: every forward must be bit-identical."""
    for B, S in ((4, 224), (3, 224), (1, 518)):
        imgs = o_pipe.synth_images(B, S, seed=31).to(dev)
        vit = ops.Vit(vit_w, dev, "bf16")
        ref = vit.forward(imgs)[0].clone()
        for _ in range(12):
            assert torch.equal(vit.forward(imgs)[0], ref), (B, S)


def test_vit_bf16_is_reproducible_under_tensor_pipe_contention(dev, ops, vit_w):
    """The attention kernel's P buffer is re-used per key tile; round 1 let the softmax overwrite it as soon as S of the next tile
    was ready, although the tensor core might still be reading it for P V -- harmless only while the tensor pipe keeps ahead, which
    MMAs of OTHER kernels on the same SMs can break.  With the half-tile commits and waits of round 2 the ViT forward must be
    bit-identical whether or not a second stream keeps the tensor pipes busy with unrelated GEMMs."""
    imgs = o_pipe.synth_images(4, 224, seed=31).to(dev)
    vit = ops.Vit(vit_w, dev, "bf16")
    ref = vit.forward(imgs)[0].clone()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    A = torch.randn(2048, 768, device=dev).bfloat16()
    W = torch.randn(2304, 768, device=dev).bfloat16()
    C = torch.empty(2048, 2304, device=dev, dtype=torch.bfloat16)
    for _ in range(6):
        with torch.cuda.stream(side):
            for _ in range(200):                         # ~20 us each: a steady stream of MMAs next to the attention CTAs
                ops.linear(A, W, "bf16", out=C)
        got = [vit.forward(imgs)[0].clone() for _ in range(4)]
        torch.cuda.synchronize()
        for t in got:
            assert torch.equal(t, ref)


def test_vit_bf16_is_reproducible_with_programmatic_overlap_back_on(dev):
    """The discriminating check for the attention kernel's P-buffer race: with programmatic dependent launch re-enabled inside the
    ViT (PIO_VIT_PDL=1, read once per process -> a fresh interpreter) every forward must still be bit-identical.  A library built
    with -DPIO_ATTN_WAIT_PV=0 (the round-1 behaviour) fails exactly this (profiles/r02bp_attention_race_root_cause.txt)."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PIO_VIT_PDL="1", TAG="pdl_on")
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "determinism_probe.py")], env=env, capture_output=True, text=True,
                         timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if "runs differ" in ln]
    assert len(lines) == 3 and all(" 0/29 runs differ" in ln for ln in lines), out.stdout


def test_project_bf16_ignores_stale_workspace_contents(dev, ops, monkeypatch):
    """The fused projection no longer zeroes its whole P buffer per call, only the K-padding columns of the last chunk: a workspace
    full of NaN bit patterns (whatever an earlier call of another size left there) must not reach the result."""
    monkeypatch.setenv("PIO_PROJECT_MAX_CHUNK", "16384")
    bank = o_pipe.synth_bank(40013, 768, seed=27, zero_frac=0.001)      # last chunk: 40013 - 32768 rows -> 59 padding columns
    q = torch.randn(800, 768, generator=torch.Generator().manual_seed(28))   # > 768 queries: the deterministic (no split-K) path
    ref = o_mem.project(q, o_mem.drop_zero_rows(bank), normalize=True)
    b16 = ops.Bank(bank, dev, "bf16")
    first = b16.project(q.to(dev), normalize=True).clone()
    from patchioner_b200 import _lib as L

    ws = ops.workspace(L.lib().pio_project_workspace_bytes(b16._h, 800), dev, "project")
    ws.view(torch.int32).fill_(0x7FC07FC0)                               # bf16 / fp32 NaNs everywhere
    again = b16.project(q.to(dev), normalize=True)
    assert bool(torch.isfinite(again).all())
    assert torch.equal(again, first)
    assert cos_min(again.cpu(), ref) >= 0.999
