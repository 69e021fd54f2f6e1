"""bench.py command-line contract on a machine WITHOUT a GPU: the product arm refuses to run (no CPU fallback), the reference
arm prints one JSON line with the contract's keys, and under a multi-rank launch only rank 0 runs the reference arm."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")


def _run(args, env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, BENCH, *args], capture_output=True, text=True, env=env, cwd=ROOT, timeout=600)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_product_arm_fails_loudly_without_a_gpu():
    r = _run(["--steps", "1", "--warmup", "1", "--no-cpu"])
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout) and "no CPU fallback" in (r.stderr + r.stdout)


def test_reference_arm_json_line_and_rank_gating():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-batch", "1"])
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("train images/sec") and d["unit"] == "images/s"
    assert d["higher_is_better"] is True and d["steps"] == 1 and d["value"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == os.cpu_count() and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    # ranks other than 0 of a torchrun launch exit 0 without work
    r1 = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-batch", "1"], {"RANK": "1", "WORLD_SIZE": "2"})
    assert r1.returncode == 0 and not [l for l in r1.stdout.splitlines() if l.startswith("{")]
