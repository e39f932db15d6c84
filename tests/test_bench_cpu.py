"""bench.py contract checks that need no GPU: the reference arm falls back to the CPU port on a bounded sample and prints
one well-formed JSON line; our arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

from conftest import HAS_GPU

REPO = Path(__file__).resolve().parents[1]


@pytest.mark.skipif(HAS_GPU, reason="GPU present: the reference arm would run the reference kernels instead")
def test_reference_arm_falls_back_to_cpu_port():
    res = subprocess.run([sys.executable, str(REPO / "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "3",
                          "--workload", "dam", "--n-side", "24"], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    line = [l for l in res.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["impl"] == "reference" and d["metric"] == "particle-iterations/sec" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["steps"] == 2 and d["n_gpus"] == 1 and d["vs_baseline"] is None


@pytest.mark.skipif(HAS_GPU, reason="GPU present")
def test_our_arm_refuses_to_run_without_a_gpu():
    res = subprocess.run([sys.executable, str(REPO / "bench.py"), "--steps", "2", "--warmup", "3"], capture_output=True,
                         text=True, timeout=600)
    assert res.returncode != 0 and "no CPU fallback" in (res.stderr + res.stdout)
