"""Slab decomposition on real GPUs: a 2-rank NCCL rollout must reproduce the 1-GPU rollout
(needs >= 2 visible GPUs; the single-rank path of the same code is checked on one)."""

import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOL = os.path.join(ROOT, "tools", "dist_check.py")


def _run(cmd):
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0 and "DIST_CHECK_OK" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]
    return res.stdout


def test_single_rank_domain_path_equals_engine():
    out = _run([sys.executable, TOOL, "--case", "rpf3d_8k", "--steps", "2", "--mp", "2"])
    assert "max|dpos|=0.000e+00" in out  # world == 1: same kernels, same order -> bitwise


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_decomposition_matches_single_gpu():
    _run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
          "--master-addr", "127.0.0.1", "--master-port", "29577", TOOL, "--case", "rpf3d_8k", "--steps", "3"])
