"""Phase timeline of one node-kernel CTA (cross-check build: LB200_BUILD_CROSSCHECK=1).

    LB200_BUILD_CROSSCHECK=1 python tools/node_trace.py [workload]

Runs one forward and prints, per worker of CTA 0 of the last processor node kernel, the SM-clock
time of every phase boundary (ids in csrc/node_tc2.cu: 0 entry, 1 alloc done, 2 weights in TMEM, 3 wait done,
10 operands built, 11 GEMM 1 issued, 12 GEMM 1 done, 13 hidden written, 14 residual requested, 15 GEMM 2 done,
16 tile done, 20/21 switch barrier, 22 pass-B weights in TMEM, 30 rows converted, 31 GEMMs issued, 32 done,
33 tile done, 40 end)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lagrangebench_b200 import GNS, _cabi, case_builder, synthetic  # noqa: E402
from lagrangebench_b200 import models as lbmodels  # noqa: E402

NAMES = {0: "entry", 1: "alloc+vec", 2: "weightsA", 3: "pdl wait", 5: "A0 req1", 6: "A0 req2", 7: "A0 data", 10: "A0 built", 11: "G1 issued", 12: "G1 done",
         13: "E1 hidden", 14: "resid req", 15: "G2 done", 16: "tile A end", 20: "switch>", 21: "switch<", 22: "weightsB",
         30: "B0 built", 31: "GB issued", 32: "GB done", 33: "tile B end", 40: "end"}


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "ldc3d_28k"
    c = synthetic.make_case(name, 6, 0, 0, np.float64, quiet=True)
    d = c["metadata"]["dim"]
    case = case_builder(c["box"], c["metadata"], 6, cfg_neighbors={"multiplier": c["multiplier"]},
                        external_force_fn=c["force"], dtype="float64", noise_std=0.0)
    feats, _ = case.allocate_eval((c["positions"][:, :6], c["particle_type"]))
    node_in = sum(int(np.prod(feats[k].shape[1:])) for k in ("vel_hist", "bound", "force") if k in feats)
    params = lbmodels.init_params(node_in, d, 128, 10, 16, seed=0)
    model = GNS(d, 128, 2, 10, 16)
    for _ in range(3):
        model.apply(params, {}, (feats, c["particle_type"]))
    torch.cuda.synchronize()
    lib = C.CDLL(_cabi.library_path())
    buf = (C.c_longlong * (4 * 64 * 2))()
    cnt = (C.c_int * 4)()
    rc = lib.lb200_debug_node_trace(buf, cnt)
    assert rc == 0, rc
    arr = np.array(buf[:]).reshape(4, 64, 2)
    ghz = 1.9
    t0 = min(arr[w, 0, 1] for w in range(4) if cnt[w] > 0)
    for w in range(4):
        print(f"worker {w}: " + "  ".join(f"{NAMES.get(int(arr[w, i, 0]), int(arr[w, i, 0]))}@{(arr[w, i, 1] - t0) / ghz / 1e3:.2f}"
                                        for i in range(cnt[w])))


if __name__ == "__main__":
    main()
