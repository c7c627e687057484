"""Long-horizon sanity check of the device-resident rollout on a quiet synthetic cloud: prints the
edge count, the largest displacement from the start and the number of re-allocations every few steps.

    python tools/long_rollout_check.py [workload] [steps] [chunk]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lagrangebench_b200 import GNS, RolloutEngine, case_builder, synthetic  # noqa: E402
from lagrangebench_b200 import models as lbmodels  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "ldc3d_28k"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 10
spec = synthetic.make_case(name, 6, steps, 0, np.float64, quiet=True)
d = spec["metadata"]["dim"]
dx = spec["metadata"]["dx"]
case = case_builder(spec["box"], spec["metadata"], 6, cfg_neighbors={"multiplier": spec["multiplier"]},
                    external_force_fn=spec["force"], dtype="float64", noise_std=0.0)
node_in = 5 * d + (0 if any(spec["metadata"]["periodic_boundary_conditions"]) else 2 * d) + (d if spec["force"] is not None else 0)
params = lbmodels.init_params(node_in, d, 128, 10, 16, seed=0)
model = GNS(d, 128, 2, 10, 16)
engine = RolloutEngine(case, model, params, steps_per_sync=chunk)
dev = torch.device("cuda")
window = torch.as_tensor(spec["positions"][:, :6]).to(dev).contiguous()
targets = torch.as_tensor(spec["positions"][:, 6:6 + steps]).permute(1, 0, 2).to(dev).contiguous()
ptype = torch.as_tensor(spec["particle_type"]).to(dev)
p0 = window[:, -1].clone()
fluid = ptype == 0
nbrs = None
for s0 in range(0, steps, chunk):
    preds, nbrs = engine.run(window, ptype, targets[s0:s0 + chunk], chunk, nbrs)
    torch.cuda.synchronize()
    disp = (window[:, -1] - p0).abs()
    print(f"step {s0 + chunk:4d}  E {nbrs.n_edges:8d}  e_cap {nbrs.max_occupancy:8d}  cell_cap {nbrs.cell_list_capacity}  "
          f"max disp/dx all {disp.max().item() / dx:.4f} fluid {disp[fluid].max().item() / dx:.4f}  "
          f"reallocs {engine.n_reallocations}  finite {bool(torch.isfinite(window).all())}", flush=True)
