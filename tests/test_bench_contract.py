"""bench.py's reference arm runs on the host alone: its JSON line must carry the contract's keys and
the same `config` / metric / unit as the GPU arm prints (the driver divides one by the other)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "tailseq-import")),
                    reason="oracle/_ref is built by __graft_entry__.build() where /root/reference exists")
def test_reference_arm_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3",
                        "--ref-clusters", "3000"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["unit"] == "clusters/s" and line["higher_is_better"] is True and line["value"] > 0
    assert line["config"]["workload"].startswith("MiSeq v2 full run") and line["config"]["clusters_per_tile"] == 1070000
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["steps"] == 1 and line["warmup"] >= 3


def test_gpu_arm_refuses_without_a_gpu():
    """No CPU fallback: without a CUDA device the product arm stops with an error, it does not print a number."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "3", "--no-cli", "--no-extra"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0
    assert not any(ln.startswith("{") and '"value"' in ln for ln in r.stdout.splitlines())
