"""CPU: the parts of bench.py's contract that do not need a GPU -- the reference arm's JSON line (it runs the CPU
oracle port on the host cores: the one place bench.py may execute oracle/) and the refusal of the GPU arm to run
without a device (no CPU fallback on the product path)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args, env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + list(args), cwd=ROOT, capture_output=True, text=True,
                          timeout=900, env=dict(os.environ, **(env or {})))


def test_reference_arm_line():
    r = run_bench("--impl", "reference", "--steps", "2", "--warmup", "0", "--batch", "64")
    assert r.returncode == 0, r.stderr[-400:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    sys.path.insert(0, ROOT)
    import bench
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == bench.UNIT and d["higher_is_better"] is True
    assert d["config"] == bench.arm_config(64, 1)                 # our arm's config, word for word
    assert d["value"] > 0 and d["steps"] == 2 and d["n_gpus"] == 1 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    r = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--batch", "64", "--gpus", "2", env={"RANK": "1"})
    assert r.returncode == 0 and not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]


def test_gpu_arm_refuses_to_run_without_a_device():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    r = run_bench("--steps", "1", "--warmup", "1", "--batch", "64")
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
