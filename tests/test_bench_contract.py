"""bench.py's contract on a box without a GPU: the reference arm (the CPU restatement of lumo's renderer — the one place besides
the tests where oracle/ may run) prints ONE JSON line with the keys the driver reads; the repo arm refuses to run without the
CUDA library and a device instead of falling back to anything."""
import json
import os
import subprocess
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=300):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "1", "--workload", "cornell", "--cpu-seconds", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1                                    # exactly one JSON line on stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "Mrays/s" and d["unit"] == "Mrays/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"]["workload"] == "cornell" and d["config"]["integrator"] == "PathTrace" and d["vs_baseline"] is None
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["sample"] and abs(cb["value"] - d["value"]) <= 1e-9 * d["value"]
    e = d["e2e"]
    assert e["unit"] == d["unit"] and abs(e["value"] - d["value"]) <= 1e-9 * d["value"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_repo_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the repo arm runs")
    r = _run("--steps", "1", "--warmup", "0", "--other-workloads", "", "--no-cpu", timeout=600)
    assert r.returncode != 0
    assert not any(l.strip().startswith("{") and '"value"' in l for l in r.stdout.splitlines())   # no number without the CUDA path
