"""bench.py contract on the CPU: the reference arm (C/OpenMP port from oracle/ only) prints one JSON line with the keys the
driver reads, on the same `config` keys as the GPU arm; the GPU arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=600, cwd=ROOT)


def test_reference_arm_line():
    r = _run(["--impl", "reference", "--refine", "0", "--steps", "2", "--warmup", "1", "--no-rk4"])
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["metric"] == "pa_laplace_matvec_gdofs" and d["unit"] == "GDOF/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 2 and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None
    cfg = d["config"]
    assert cfg["workload"].startswith("wave-tank-big8") and cfg["hexes_per_gpu"] == 4096 and cfg["dofs_global"] == 299520 and cfg["order"] == 4
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "GDOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the process must not have mapped the product library
    assert "liblpf_b200" not in r.stdout + r.stderr


def test_gpu_arm_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present")
    r = _run(["--steps", "1", "--warmup", "1", "--no-cpu"])
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout) or "CUDA" in (r.stderr + r.stdout)
