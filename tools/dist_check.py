"""Multi-GPU check: a slab-decomposed rollout must reproduce the single-GPU rollout.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tools/dist_check.py --case rpf3d_8k --steps 3
"""
import argparse
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lagrangebench_b200 import GNS, RolloutEngine, case_builder, synthetic  # noqa: E402
from lagrangebench_b200 import models as lbmodels  # noqa: E402
from lagrangebench_b200.domain import DistributedRollout  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default="rpf3d_8k")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--mp", type=int, default=10)
    ap.add_argument("--same-gpu", action="store_true",
                    help="all ranks on cuda:0 (time-sliced contexts), gloo for the host plumbing: the peer-memory "
                         "path on a one-GPU box")
    ap.add_argument("--spread", type=float, default=0.0,
                    help="velocity scale in units of dx per step (0: the case's own statistics): makes particles "
                         "cross slab faces, so ghost re-selection and migration run")
    ap.add_argument("--sync", type=int, default=32, help="steps per host synchronisation")
    ap.add_argument("--margin", type=float, default=0.25)
    ap.add_argument("--dtype", default="float64", choices=["float32", "float64"],
                    help="position dtype (reference default float64).  In float32 the open-axis displacement of a "
                         "slab (a - b) and the periodic one of the whole box round differently, and a pair within "
                         "1e-7 of the cutoff can be an edge in one cloud and not in the other")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = 0 if args.same_gpu else int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        if args.same_gpu:
            dist.init_process_group("gloo")
        else:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    npd = np.float64 if args.dtype == "float64" else np.float32
    tdt = torch.float64 if args.dtype == "float64" else torch.float32
    c = synthetic.make_case(args.case, 6, args.steps, 0, npd)  # future frames: where kinematic particles are put
    if args.spread > 0:  # a coherent drift along every axis plus the case's jitter
        dx = c["metadata"]["dx"]
        drift = args.spread * dx * np.arange(6, dtype=npd)[None, :, None]
        c["positions"] = np.mod(c["positions"][:, :6] + drift, c["box"].astype(npd)).astype(npd)  # drops the future frames
        c["metadata"]["vel_mean"] = [args.spread * dx] * c["metadata"]["dim"]
    d = c["metadata"]["dim"]
    n = c["positions"].shape[0]
    node_in = 5 * d + (2 * d if not any(c["metadata"]["periodic_boundary_conditions"]) else 0) + \
        (d if c["force"] is not None else 0)
    params = lbmodels.init_params(node_in, d, 128, args.mp, 16, seed=0)
    dr = DistributedRollout(c["box"], c["metadata"], params, args.mp, force=c["force"], dtype=tdt,
                            multiplier=c["multiplier"], halo_margin=args.margin, steps_per_sync=args.sync)
    dr.scatter(c["positions"], c["particle_type"])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dr.run(args.steps)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    pos = dr.gather_positions(n)
    owned = torch.tensor([dr.window.shape[0]], device="cpu" if args.same_gpu else "cuda")
    if world > 1:
        dist.all_reduce(owned)
    if rank == 0:
        case = case_builder(c["box"], c["metadata"], 6, cfg_neighbors={"multiplier": c["multiplier"]},
                            external_force_fn=c["force"], dtype=args.dtype)
        model = GNS(d, 128, 2, args.mp, 16)
        eng = RolloutEngine(case, model, params)
        window = torch.as_tensor(c["positions"][:, :6]).cuda().contiguous()
        fut = c["positions"][:, 6:6 + args.steps]
        targets = torch.as_tensor(fut).permute(1, 0, 2).cuda().contiguous() if fut.shape[1] == args.steps else None
        ref, _ = eng.run(window, c["particle_type"], targets, args.steps)
        diff = case.displacement(pos, ref[-1]).abs().max().item()
        dx = c["metadata"]["dx"]
        print(f"world={world} N={n} steps={args.steps} owned_total={int(owned)} edges_rank0={dr.edges_last} "
              f"ghosts_rank0={dr.n_ghost_left + dr.n_ghost_right} max|dpos|={diff:.3e} ({diff / dx:.2e} dx) "
              f"{1e3 * dt / args.steps:.2f} ms/step reallocs={dr.n_reallocations} selections={dr.n_selections} "
              f"migrations={dr.n_migrations}", flush=True)
        assert int(owned) == n, "particles lost or duplicated in migration"
        assert diff <= 1e-5 * dx + 8 * np.finfo(npd).eps * float(np.max(c["box"])), "decomposed rollout diverged"
        print("DIST_CHECK_OK", flush=True)
    dr.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
