"""bench.py's driver contract, checked on the CPU box: the reference arm prints ONE JSON line with the required keys
(metric / unit / config shared with the native arm, `impl: reference`, cpu_baseline, zero-copy e2e), and the native arm
refuses to run without a CUDA device instead of falling back to anything."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600)


def test_reference_arm_json_line():
    r = _run("--impl", "reference", "--size", "128", "--steps", "2", "--warmup", "1", "--ref-iters", "3")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "problem-iters/s" and d["value"] > 0 and d["vs_baseline"] is None
    assert "workload" in d["config"] and d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_native_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    r = _run("--steps", "1", "--warmup", "0")
    assert r.returncode != 0 and "CUDA" in (r.stderr + r.stdout)
