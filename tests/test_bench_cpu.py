"""bench.py's CPU-only legs: the JSON contract of the reference arm and the synthetic workloads."""

import json
import os
import subprocess
import sys

import numpy as np

from lagrangebench_b200 import synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tgv2d",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "particle-steps/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["config"]["workload"] == "tgv2d"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "2500 particles" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_quiet_synthetic_statistics_only_rescale_the_motion():
    """quiet=True (bench.py) changes the velocity / acceleration scales, not the cloud: same lattice
    jitter, same particle types, per-step displacement about 250 times smaller."""
    a = synthetic.make_case("ldc3d", n_future=3, seed=4, dims=(12, 10, 9))
    q = synthetic.make_case("ldc3d", n_future=3, seed=4, dims=(12, 10, 9), quiet=True)
    assert np.array_equal(a["particle_type"], q["particle_type"]) and np.array_equal(a["box"], q["box"])
    assert np.array_equal(a["positions"][:, 0], q["positions"][:, 0])
    fluid = a["particle_type"] == 0
    da = np.abs(a["positions"][fluid, 1] - a["positions"][fluid, 0]).max()
    dq = np.abs(q["positions"][fluid, 1] - q["positions"][fluid, 0]).max()
    assert 0 < dq < da / 100
    dx = q["metadata"]["dx"]
    assert q["metadata"]["vel_std"][0] == 2.0e-4 * dx and q["metadata"]["acc_std"][0] == 1.0e-7 * dx
    # 400 steps of ballistic drift stay far inside one cutoff radius
    assert 400 * 5 * q["metadata"]["vel_std"][0] < 0.5 * q["metadata"]["default_connectivity_radius"]
