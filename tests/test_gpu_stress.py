"""Randomized parity (GPU vs oracle) over modification sets, decoy modes and counts, windows, top-k, batch sizes, stored
decoys, the expanded variable-modification mode and the split-spectrum scoring path (tools/gpu_stress.py).  The fixed
cases of test_gpu_parity.py pin the known branches; this one looks for the combinations nobody thought of (it found the
permuted-target decoys of candidates that carry variable modifications)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_random_configurations_bit_exact():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gpu_stress.py"), "24"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "mismatches: 0" in r.stdout
