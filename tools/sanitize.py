"""Smoke-size workload for compute-sanitizer ("race detection"):

    compute-sanitizer --tool memcheck  --error-exitcode 1 python tools/sanitize.py
    compute-sanitizer --tool racecheck --error-exitcode 1 python tools/sanitize.py
    compute-sanitizer --tool synccheck --error-exitcode 1 python tools/sanitize.py

One forward through the per-step API and a three-step device-resident rollout (eager + graph replay)
of a small periodic cloud and a small walled one: every kernel of the product path runs.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lagrangebench_b200 import GNS, RolloutEngine, case_builder, synthetic  # noqa: E402
from lagrangebench_b200 import models as lbmodels  # noqa: E402


def main():
    mp = int(os.environ.get("SANITIZE_MP", "2"))
    for name, dims in (("tgv2d", (20, 20)), ("ldc3d", (12, 10, 9))):
        c = synthetic.make_case(name, 6, 3, 0, np.float32, dims=dims, quiet=True)
        d = c["metadata"]["dim"]
        case = case_builder(c["box"], c["metadata"], 6, cfg_neighbors={"multiplier": c["multiplier"]},
                            external_force_fn=c["force"], dtype="float32", noise_std=0.0)
        feats, nbrs = case.allocate_eval((c["positions"][:, :6], c["particle_type"]))
        node_in = sum(int(np.prod(feats[k].shape[1:])) for k in ("vel_hist", "bound", "force") if k in feats)
        params = lbmodels.init_params(node_in, d, 128, mp, 16, seed=0)
        model = GNS(d, 128, 2, mp, 16)
        out, _ = model.apply(params, {}, (feats, c["particle_type"]))
        assert torch.isfinite(out["acc"]).all()
        engine = RolloutEngine(case, model, params)
        window = torch.as_tensor(c["positions"][:, :6]).cuda().contiguous()
        targets = torch.as_tensor(c["positions"][:, 6:9]).permute(1, 0, 2).cuda().contiguous()
        preds, _ = engine.run(window, c["particle_type"], targets, 3)
        torch.cuda.synchronize()
        assert torch.isfinite(preds).all()
        print(f"{name} {dims}: N={window.shape[0]} ok", flush=True)
    print("SANITIZE_WORKLOAD_OK", flush=True)


if __name__ == "__main__":
    main()
