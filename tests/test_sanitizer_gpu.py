"""compute-sanitizer over the product path ("race detection"): memcheck, racecheck and
synccheck of a smoke-size forward + device-resident rollout (tools/sanitize.py) must report nothing.
The tensor-core kernels hand shared-memory operands between warps through named barriers and
mbarriers; a hand-off that is one phase off shows up here, not in a parity test."""

import os
import shutil
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("tool", ["memcheck", "racecheck", "synccheck"])
def test_compute_sanitizer_is_clean(tool):
    exe = shutil.which("compute-sanitizer") or "/usr/local/cuda/bin/compute-sanitizer"
    if not os.path.exists(exe):
        pytest.skip("compute-sanitizer not installed")
    res = subprocess.run([exe, "--tool", tool, "--error-exitcode", "1", sys.executable,
                          os.path.join(ROOT, "tools", "sanitize.py")], capture_output=True, text=True, timeout=900,
                         cwd=ROOT)
    tail = res.stdout[-3000:] + res.stderr[-1000:]
    assert res.returncode == 0 and "SANITIZE_WORKLOAD_OK" in res.stdout, tail
    assert "0 errors" in res.stdout or "0 hazards" in res.stdout, tail
