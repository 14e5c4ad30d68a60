"""The compiled C++ harness (examples/harness.cpp): the reference's node call order through the header-only class wrappers
(LaserProcessingClass -> OdomEstimationClass -> LaserMappingClass) on the GPU; exit code 0 = tracked the ground truth."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_harness_runs():
    exe = os.path.join(ROOT, "examples", "harness")
    if not os.path.exists(exe):
        pytest.fail("examples/harness missing: run __graft_entry__.build()")
    r = subprocess.run([exe, "14"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    print(r.stdout)
    assert r.returncode == 0, r.stdout
    assert "14 frames" in r.stdout
