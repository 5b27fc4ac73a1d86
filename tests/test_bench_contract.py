"""bench.py contract checks that need no GPU: the reference arm's JSON line, and the product arm refusing to run
without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "3",
                          "--workload", "newcastle"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "cell-updates/s" and line["unit"] == "cell-updates/s"
    assert line["value"] > 0 and line["ms_per_step"] > 0 and line["steps"] == 2 and line["warmup"] == 3 and line["n_gpus"] == 1
    assert line["higher_is_better"] is True and line["scaling"] == "weak" and line["vs_baseline"] is None
    assert line["dtype"] == "f64" and line["data"] == "synthetic" and line["config"]["workload"] == "newcastle"
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0


def test_product_arm_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "3", "--workload", "newcastle"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode != 0
    assert "no CPU fallback" in (res.stderr + res.stdout)
    assert not any(l.startswith("{") for l in res.stdout.splitlines())       # no result line
