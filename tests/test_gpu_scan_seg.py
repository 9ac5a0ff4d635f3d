"""The chain of long inputs (pfb_scan.cu: scan_seg_* + scan_chain_seg_kernel) is chosen by the number of tiles
(> 8192, i.e. > 1.7e7 elements), so the adversarial inputs of tests/test_gpu_kernels.py (ties, zeros, subnormals,
60-binade sweeps, degenerate weights) would never reach it: this test re-runs the scan / resampling tests of that file in a
child process with PFB_SCAN_SEG_MIN=0, which sends every single-run exact scan through the new chain."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_adversarial_scan_inputs_through_the_segmented_chain():
    env = dict(os.environ, PFB_SCAN_SEG_MIN="0")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_kernels.py"), "-m", "gpu", "-x", "-q",
                        "-k", "scan or cdf or cumsum or resample or systematic or multinomial or slice_scaled or step",
                        "-p", "no:cacheprovider"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    tail = (r.stdout or "")[-2000:] + (r.stderr or "")[-1000:]
    assert r.returncode == 0, tail
    assert " passed" in r.stdout and "no tests ran" not in r.stdout, tail
