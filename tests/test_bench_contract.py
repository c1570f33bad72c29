"""bench.py's reference arm runs on host cores only, so its JSON line can be checked here: one line,
the contract's keys, the reference-arm additions (impl, cpu_baseline, e2e with zero copy bytes), and
the same metric / config naming as the GPU arm reports."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    # torchrun hands OMP_NUM_THREADS=1 to its workers: the arm must still use every host thread it may
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="2")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--gpus", "2"], capture_output=True, text=True, timeout=900, env=env)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "queries/s" and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and d["n_gpus"] == 2 and d["steps"] == 1
    assert d["config"]["workload"] == "c3" and d["config"]["rows"] == 1_000_000 and d["config"]["dim"] == 768
    assert d["config"]["batch"] == 1024 and d["config"]["k"] == 100 and "model" not in d["config"]
    assert "1M x 768" in d["metric"] and d["value"] > 0 and d["ms_per_step"] > 0
    cb, e = d["cpu_baseline"], d["e2e"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert cb["cores"] == len(os.sched_getaffinity(0)) and d["host"]["torch_threads"] == cb["cores"]
    # ranks other than 0 exit 0 without work and without output
    env1 = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r1 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                         "--warmup", "1", "--gpus", "2"], capture_output=True, text=True, timeout=120, env=env1)
    assert r1.returncode == 0 and not r1.stdout.strip()


def test_gpu_arm_fails_loudly_without_a_device():
    """No CPU fallback: without CUDA the product arm must exit non-zero, not print a number."""
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("GPU present")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode != 0
    assert not [l for l in res.stdout.splitlines() if l.startswith("{")]
