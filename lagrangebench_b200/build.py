"""Build the in-tree CUDA library ``lagrangebench_b200/_lb200.so`` for sm_100a with nvcc.

The library is a plain C-ABI shared object (``include/lb200.h``): no torch, no pybind.
``nvcc`` cross-compiles without a GPU; the built ``.so`` is git-ignored but travels to the
GPU box with the repo snapshot.
"""

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO_PATH = os.path.join(HERE, "_lb200.so")
STAMP = SO_PATH + ".stamp"
SOURCES = ["neighbor.cu", "features.cu", "gns.cu", "gns_tc.cu", "gns_tc2.cu", "node_tc2.cu", "peer.cu", "rollout.cu"]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC", "-shared", "-cudart", "shared",
    "-Xlinker", "-rpath=/usr/local/cuda/lib64",
]


def _flags():
    """LB200_BUILD_CROSSCHECK=1 also compiles the superseded first-generation tensor-core kernels and the A/B
    variants of the message kernel (tests then compare the product kernels against them)."""
    extra = ["-DLB200_CROSSCHECK"] if os.environ.get("LB200_BUILD_CROSSCHECK") == "1" else []
    return NVCC_FLAGS + extra


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _digest():
    h = hashlib.sha256()
    h.update(" ".join(_flags()).encode())
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for name in sorted(os.listdir(root)):
            if name.endswith((".cu", ".cuh", ".h")):
                with open(os.path.join(root, name), "rb") as f:
                    h.update(name.encode())
                    h.update(f.read())
    return h.hexdigest()


def build(force=False, verbose=False):
    """Compile if sources changed since the last build.  Returns the path of the .so."""
    digest = _digest()
    if not force and os.path.exists(SO_PATH) and os.path.exists(STAMP):
        with open(STAMP) as f:
            if f.read().strip() == digest:
                return SO_PATH
    cmd = [_nvcc()] + _flags() + (["-Xptxas", "-v"] if verbose else []) + \
        ["-o", SO_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building _lb200.so")
    if verbose:
        sys.stderr.write(res.stdout + res.stderr)
    with open(STAMP, "w") as f:
        f.write(digest)
    return SO_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
