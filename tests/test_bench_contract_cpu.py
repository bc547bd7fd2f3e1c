"""bench.py's contract on CPU: the reference arm prints exactly one JSON line with the agreed keys
(tiny geometry here so it runs in seconds), and our own arm refuses to run without a GPU instead
of falling back to anything."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--model", "tiny", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "decode_tok_s" and j["unit"] == "tok/s" and j["higher_is_better"] is True
    assert j["n_gpus"] == 1 and j["steps"] == 1 and j["value"] > 0 and j["vs_baseline"] is None and j["data"] == "synthetic"
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == j["value"] and isinstance(cb["sample"], str)
    e = j["e2e"]
    assert (e["value"], e["unit"], e["h2d_bytes_per_step"], e["d2h_bytes_per_step"]) == (j["value"], j["unit"], 0, 0)
    assert "workload" in j["config"] and "model" not in j["config"]


def test_our_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--model", "tiny", "--steps", "1", "--warmup", "3", "--no-cpu-baseline", "--no-micro"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode != 0
    assert not [l for l in out.stdout.splitlines() if l.strip().startswith("{")]     # no number without the device
