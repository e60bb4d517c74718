"""CPU: bench.py's reference arm (the oracle port on the host cores) prints one JSON line with the contract's keys, for the
headline workload and for the ViECap workload, and stays silent on ranks other than 0."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
        "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"}


def _run(extra, env=None):
    e = dict(os.environ, **(env or {}))
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--size", "224"] + extra, capture_output=True, text=True, timeout=600, env=e, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    return [ln for ln in p.stdout.splitlines() if ln.startswith("{")]


@pytest.mark.parametrize("workload", ["dense", "regionset-viecap"])
def test_reference_arm_prints_the_contract_line(workload):
    lines = _run(["--workload", workload])
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert KEYS <= set(d) and d["impl"] == "reference" and d["unit"] == "captions/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["gpu_launches"] == 0
    assert "workload" in d["config"] and d["vs_baseline"] is None and d["higher_is_better"] is True


def test_reference_arm_other_ranks_exit_silently():
    assert _run(["--gpus", "2"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []


def test_gpu_eager_baseline_pipeline_matches_oracle():
    """tools/gpu_eager.py (bench.py's eager-PyTorch-on-GPU baseline) computes what the oracle computes: checked here on
    the CPU (the module is device-agnostic) stage by stage, token ids included, with and without the KV cache."""
    import torch

    from oracle import decap as od
    from oracle import dinov2 as ov
    from oracle import memory as om
    from oracle import pipeline as op
    from oracle import pooling as opool
    from tools.gpu_eager import EagerPipeline

    vw, dw = ov.make_weights(1234), od.make_weights(1234)
    bank = op.synth_bank(3000, 768, seed=7, zero_frac=0.002)
    ep = EagerPipeline(vw, dw, bank, "cpu")
    imgs = op.synth_images(2, 224, seed=1)
    boxes = op.synth_boxes(2, 5, 224, seed=1, pad="dense")
    xn, attn = ep.vit(imgs)
    ref = ov.forward(vw, imgs)
    torch.testing.assert_close(xn[:, 5:], ref["x_norm_patchtokens"], rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(attn, opool.cls_attention_map(ref["qkv"]), rtol=1e-3, atol=1e-7)
    for gauss in (True, False):
        feats = ep.pool(xn[:, 5:], ep.box_weights(boxes, 16, gauss, 1.0))
        rf = opool.extract_bboxes_feats(ref["x_norm_patchtokens"], boxes.clone(), gauss, 1.0)
        torch.testing.assert_close(feats, rf, rtol=1e-4, atol=1e-4)
    pr = ep.project(feats.reshape(-1, 768))
    rp = om.project(rf.reshape(-1, 768), om.drop_zero_rows(bank), normalize=True)
    torch.testing.assert_close(pr, rp, rtol=1e-4, atol=1e-5)
    assert torch.equal(ep.decode(pr[:4], use_cache=True), od.decode_greedy(dw, rp[:4], use_cache=True))
    assert torch.equal(ep.dense_step(imgs, boxes, gaussian=False, use_cache=True)[:3], od.decode_greedy(dw, rp[:3], use_cache=True))
