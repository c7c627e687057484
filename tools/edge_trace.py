"""Phase timeline of one message-kernel worker (cross-check build: LB200_BUILD_CROSSCHECK=1).

    LB200_BUILD_CROSSCHECK=1 python tools/edge_trace.py [workload]

Prints, for warp 0 (issues the GEMMs) and warp 1 of worker 0 of CTA 0 of the last message-kernel launch,
the SM-clock time of every phase boundary over pipeline iterations 8..11 (ids in csrc/gns_tc2.cu:
E1: 10 gathers issued, 11 GEMM 1 done, 12 hidden written, 13 past the hand-off (+ GEMM 2 issued);
A: 0 begin, 1 staged rows there, 2 operand written, 3 past the hand-off (+ GEMM 1 issued, next bulk copy);
E2: 20 residual requested, 21 GEMM 2 done, 22 partial sums written, 23 past LayerNorm's barrier, 30 end)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lagrangebench_b200 import GNS, _cabi, case_builder, synthetic  # noqa: E402
from lagrangebench_b200 import models as lbmodels  # noqa: E402

NAMES = {0: "A>", 1: "A.staged", 2: "A.built", 3: "A<", 10: "E1.gath", 11: "E1.g1", 12: "E1.built", 13: "E1<",
         20: "E2.req", 21: "E2.g2", 22: "E2.part", 23: "E2.ln", 30: "E2<"}


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
    buf = (C.c_longlong * (4 * 96 * 2))()
    cnt = (C.c_int * 4)()
    assert lib.lb200_debug_edge_trace(buf, cnt) == 0
    arr = np.array(buf[:]).reshape(4, 96, 2)
    ghz = 1.9
    t0 = min(arr[w, 0, 1] for w in range(4) if cnt[w] > 0)
    for w in range(4):
        print(f"warp {w}:")
        line = []
        for i in range(cnt[w]):
            line.append(f"{NAMES.get(int(arr[w, i, 0]), int(arr[w, i, 0]))}@{(arr[w, i, 1] - t0) / ghz / 1e3:.2f}")
            if int(arr[w, i, 0]) == 30:
                print("   " + "  ".join(line))
                line = []
        if line:
            print("   " + "  ".join(line))


if __name__ == "__main__":
    main()
