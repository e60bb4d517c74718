"""GPU (-m gpu): the persistent fused decode kernel (csrc/decode_fused_sm100.cu) -- phase by phase against a torch
emulation of the same arithmetic (bf16 operands, fp32 accumulation, bf16 activations), and end to end against the
kernel-per-op decode of round 1 and the fp32 parity mode.  decoding_batched: src/decap/decap.py:116-160."""
import ctypes as C
import math
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import decap as o_decap


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
def dec_w():
    return o_decap.make_weights(seed=1234)


def bf(t):
    return t.to(torch.bfloat16).float()


def gelu_new(x):
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x ** 3)))


def emulate_step0(w, prefix):
    """expected workspace contents after each global phase of step 0 (position 0): list of (name, [(region, tensor), ...]) in the
    kernel's phase order -- block 0: QKV* ATTN PROJ FC* FC2; blocks 1..: LN1r QKV ATTN PROJ FC* FC2; then LNFr LMHEAD (PICK).
    (* LayerNorm applied while loading: no h rows in memory; fc2 leaves split-K partials that the NEXT LayerNorm phase folds into x.)"""
    T = "decoder.transformer."
    out = []
    x = bf(prefix) @ bf(w["clip_project.model.0.weight"]).T + w["clip_project.model.0.bias"] + w[T + "wpe.weight"][0]
    for i in range(4):
        p = f"{T}h.{i}."
        h = bf(F.layer_norm(x, (768,), w[p + "ln_1.weight"], w[p + "ln_1.bias"], eps=1e-5))
        if i > 0:
            out.append((f"L{i}.ln1", [("h", h), ("x", x)]))
        qkv = bf(h @ bf(w[p + "attn.c_attn.weight"]) + w[p + "attn.c_attn.bias"])
        out.append((f"L{i}.qkv", [("qkv", qkv)]))
        att = qkv[:, 1536:]                      # one key: softmax = 1, output = v
        out.append((f"L{i}.attn", [("att", att)]))
        x = x + att @ bf(w[p + "attn.c_proj.weight"]) + w[p + "attn.c_proj.bias"]
        out.append((f"L{i}.proj", [("x", x)]))
        h = bf(F.layer_norm(x, (768,), w[p + "ln_2.weight"], w[p + "ln_2.bias"], eps=1e-5))
        f = bf(gelu_new(h @ bf(w[p + "mlp.c_fc.weight"]) + w[p + "mlp.c_fc.bias"]))
        out.append((f"L{i}.fc", [("f", f)]))
        x = x + f @ bf(w[p + "mlp.c_proj.weight"]) + w[p + "mlp.c_proj.bias"]
        out.append((f"L{i}.fc2", []))
    h = bf(F.layer_norm(x, (768,), w[T + "ln_f.weight"], w[T + "ln_f.bias"], eps=1e-5))
    out.append(("lnf", [("h", h), ("x", x)]))
    logits = h @ bf(w[T + "wte.weight"]).T
    out.append(("lmhead", [("logits", logits)]))
    return out


def _regions(ops, dec, R, dev):
    from patchioner_b200 import _lib as L

    offs = (C.c_longlong * 8)()
    L.check(L.lib().pio_decode_debug_layout(dec._h, R, offs, 8))
    nbytes = L.lib().pio_decode_workspace_bytes(dec._h, R, 30)
    ws = ops.workspace(nbytes, dev, "decode")

    def view(i, dtype, shape):
        n = int(torch.tensor(shape).prod()) * torch.empty(0, dtype=dtype).element_size()
        return ws[offs[i]:offs[i] + n].view(dtype).reshape(shape)

    return {"x": view(0, torch.float32, (R, 768)), "h": view(1, torch.bfloat16, (R, 768)), "qkv": view(2, torch.bfloat16, (R, 2304)),
            "f": view(3, torch.bfloat16, (R, 3072)), "att": view(4, torch.bfloat16, (R, 768))}


@pytest.mark.parametrize("R", [32, 5])
def test_fused_decode_phase_by_phase(dev, ops, dec_w, R, monkeypatch):
    """Stop the persistent kernel after each global phase of the first position and compare the workspace region that phase
    writes with the emulation: localises a wrong phase in one run."""
    monkeypatch.setenv("PIO_FUSED_CHECK", "1")
    dec = ops.Decoder(dec_w, dev, "bf16")
    g = torch.Generator().manual_seed(5)
    prefix = torch.randn(R, 768, generator=g)
    prefix = (prefix / prefix.norm(dim=-1, keepdim=True)).to(dev)
    w = {k: v.to(dev).float() for k, v in dec_w.items()}
    exp = emulate_step0(w, prefix)
    worst = []
    for gp, (name, checks) in enumerate(exp):
        monkeypatch.setenv("PIO_FUSED_STOP_PHASE", str(gp))
        ids = dec.decode(prefix, 30)
        torch.cuda.synchronize()
        for region, want in checks:
            if region == "logits":
                continue
            got = _regions(ops, dec, R, dev)[region].float()
            err = (got - want).abs().max().item()
            scale = want.abs().max().item()
            worst.append((f"{name}:{region}", err, scale))
            print(f"phase {gp:2d} {name:10s} {region:4s} max|diff| {err:.4g} (max|ref| {scale:.4g})")
    bad = [(n, e, s) for n, e, s in worst if not (e <= 0.02 * max(s, 1.0))]
    assert not bad, bad
    # first token: arg-max of the emulated logits, where the emulated margin is clear of bf16 noise
    monkeypatch.setenv("PIO_FUSED_STOP_PHASE", str(len(exp)))
    ids = dec.decode(prefix, 30)
    torch.cuda.synchronize()
    logits = exp[-1][1][0][1]
    top2 = logits.topk(2, dim=-1)
    clear = (top2.values[:, 0] - top2.values[:, 1]) > 0.05
    assert clear.sum() >= R // 2
    assert torch.equal(ids[:, 0].long()[clear], top2.indices[:, 0][clear])


def _prefix(R, dev, seed=9):
    g = torch.Generator().manual_seed(seed)
    p = torch.randn(R, 768, generator=g)
    return (p / p.norm(dim=-1, keepdim=True)).to(dev)


@pytest.mark.parametrize("R", [1, 7, 16, 32, 33, 64])
def test_fused_decode_matches_kernel_per_op_decode(dev, ops, dec_w, R, monkeypatch):
    """Same arithmetic, different summation order inside the tensor cores / split-K: the two bf16 paths agree on nearly all
    captions (random-init logit margins are tiny), and the fused one repeats bit for bit."""
    monkeypatch.setenv("PIO_FUSED_CHECK", "1")
    monkeypatch.setenv("PIO_DECODE_FUSED_MAX_ROWS", "64")  # the default route stops at 32 rows (where the fused kernel is faster)
    dec = ops.Decoder(dec_w, dev, "bf16")
    prefix = _prefix(R, dev)
    monkeypatch.setenv("PIO_DECODE_FUSED", "1")
    a = dec.decode(prefix, 30).clone()
    b = dec.decode(prefix, 30).clone()
    assert torch.equal(a, b), "fused decode is not repeatable"
    monkeypatch.setenv("PIO_DECODE_FUSED", "0")
    c = dec.decode(prefix, 30).clone()
    same = (a == c)
    prefix_len = same.long().cumprod(dim=1).sum(dim=1).float().mean().item()
    first = same[:, 0].float().mean().item()
    print(f"R={R}: identical captions {same.all(dim=1).float().mean().item():.3f}, mean common prefix {prefix_len:.1f}/30, first token {first:.3f}")
    assert first >= 0.9 and prefix_len >= 20.0


def test_fused_decode_against_fp32_oracle(dev, ops, dec_w, monkeypatch):
    """bf16 fused decode vs the CPU oracle (fp32): agreement of the same order as the kernel-per-op bf16 path."""
    monkeypatch.setenv("PIO_FUSED_CHECK", "1")
    monkeypatch.setenv("PIO_DECODE_FUSED_MAX_ROWS", "64")
    R = 48
    dec = ops.Decoder(dec_w, dev, "bf16")
    prefix = _prefix(R, dev, seed=10)
    ref = o_decap.decode_greedy(dec_w, prefix.cpu())
    monkeypatch.setenv("PIO_DECODE_FUSED", "1")
    a = dec.decode(prefix, 30).cpu().long()
    monkeypatch.setenv("PIO_DECODE_FUSED", "0")
    c = dec.decode(prefix, 30).cpu().long()
    pa = (a == ref).long().cumprod(dim=1).sum(dim=1).float().mean().item()
    pc = (c == ref).long().cumprod(dim=1).sum(dim=1).float().mean().item()
    print(f"mean common prefix with the fp32 oracle: fused {pa:.1f}, kernel-per-op {pc:.1f}")
    assert (a[:, 0] == ref[:, 0]).float().mean().item() >= 0.8 and pa >= pc - 4.0


def test_fused_decode_nan_row_and_scores_fallback(dev, ops, dec_w, monkeypatch):
    """A NaN prefix row decodes to token 0 at every step (torch.argmax on NaN logits) without disturbing its neighbours;
    compute_scores takes the kernel-per-op path (the fused kernel does not carry the sum of exponentials)."""
    monkeypatch.setenv("PIO_FUSED_CHECK", "1")
    dec = ops.Decoder(dec_w, dev, "bf16")
    prefix = _prefix(6, dev, seed=11)
    clean = dec.decode(prefix, 30).clone()
    prefix2 = prefix.clone()
    prefix2[2] = float("nan")
    got = dec.decode(prefix2, 30)
    assert (got[2] == 0).all()
    keep = [0, 1, 3, 4, 5]
    assert torch.equal(got[keep], clean[keep])
    ids, lp = dec.decode(prefix, 30, compute_scores=True)
    assert torch.isfinite(lp).all() and (ids[:, 0] == clean[:, 0]).float().mean() >= 0.8


def test_fused_decode_speed_report(dev, ops, dec_w, monkeypatch):
    """Not a pass/fail bar: prints us per decode step of both paths for the BASELINE small-batch sizes."""
    monkeypatch.setenv("PIO_DECODE_FUSED_MAX_ROWS", "64")
    dec = ops.Decoder(dec_w, dev, "bf16")
    for R in (8, 32, 64):
        prefix = _prefix(R, dev, seed=12)
        res = {}
        for flag in ("1", "0"):
            monkeypatch.setenv("PIO_DECODE_FUSED", flag)
            for _ in range(3):
                dec.decode(prefix, 30)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                dec.decode(prefix, 30)
            e1.record()
            torch.cuda.synchronize()
            res[flag] = e0.elapsed_time(e1) / 10
        print(f"R={R}: fused {res['1']:.3f} ms ({res['1'] / 30 * 1e3:.1f} us/step), kernel-per-op {res['0']:.3f} ms ({res['0'] / 30 * 1e3:.1f} us/step)")


@pytest.mark.parametrize("R,P,steps,layers", [(64, 18, 16, 2), (5, 28, 64, 12), (33, 3, 40, 12)])
def test_fused_prompt_decode_matches_kernel_per_op(dev, ops, R, P, steps, layers, monkeypatch):
    """ViECap's language model (greedy_search, src/viecap/search.py:108-191): GPT-2 (12 heads x 64) continuing a prompt -- the
    generation loop after the batched prefill runs in the persistent kernel (start at ln_f of the last prompt position,
    head_dim-64 attention over up to 128 cached positions).  Against the kernel-per-op bf16 path and for repeatability."""
    from oracle import viecap as ov

    monkeypatch.setenv("PIO_FUSED_CHECK", "1")
    monkeypatch.setenv("PIO_DECODE_FUSED_MAX_ROWS", "64")
    w = ov.make_weights(seed=77, n_layer_gpt=layers, n_layer_map=1)
    dec = ops.Gpt2Decoder(w, dev, "bf16")
    g = torch.Generator().manual_seed(21)
    prompt = (torch.randn(R, P, 768, generator=g) * 0.3).to(dev)
    monkeypatch.setenv("PIO_DECODE_FUSED", "1")
    a = dec.decode(prompt, steps).clone()
    b = dec.decode(prompt, steps).clone()
    assert torch.equal(a, b), "fused prompt decode is not repeatable"
    monkeypatch.setenv("PIO_DECODE_FUSED", "0")
    c = dec.decode(prompt, steps).clone()
    same = (a == c)
    prefix_len = same.long().cumprod(dim=1).sum(dim=1).float().mean().item()
    print(f"R={R} P={P} steps={steps} L={layers}: first token {same[:, 0].float().mean().item():.3f}, mean common prefix {prefix_len:.1f}/{steps}")
    assert a.min() >= 0 and a.max() < 50257
    assert same[:, 0].float().mean().item() >= 0.9 and prefix_len >= 0.5 * steps


def test_fused_prompt_decode_speed_report(dev, ops, monkeypatch):
    """prints us per generated position of GPT-2 small at the ViECap bench batch (64 rows, 64 tokens)"""
    from oracle import viecap as ov

    monkeypatch.setenv("PIO_DECODE_FUSED_MAX_ROWS", "64")
    w = ov.make_weights(seed=77, n_layer_gpt=12, n_layer_map=1)
    dec = ops.Gpt2Decoder(w, dev, "bf16")
    prompt = (torch.randn(64, 24, 768, generator=torch.Generator().manual_seed(3)) * 0.3).to(dev)
    res = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("PIO_DECODE_FUSED", flag)
        for _ in range(2):
            dec.decode(prompt, 64)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            dec.decode(prompt, 64)
        e1.record()
        torch.cuda.synchronize()
        res[flag] = e0.elapsed_time(e1) / 5
    print(f"GPT-2 small, 64 rows, 24-token prompt + 64 tokens: fused {res['1']:.2f} ms, kernel-per-op {res['0']:.2f} ms")
