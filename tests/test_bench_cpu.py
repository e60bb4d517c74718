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
