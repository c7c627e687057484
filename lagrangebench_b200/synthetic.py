"""Synthetic particle clouds of the shapes BASELINE.json names (no datasets are reachable).

Jittered lattices with a short position history, metadata in the reference's
``metadata.json`` schema, particle types per ``lagrangebench/utils.py:17-25`` and the force
fields of the reference datasets' ``force.py`` as :class:`PiecewiseForce`.  SURVEY.md §8(d)
describes the recipe; sizes follow the table at the top of SURVEY.md §8.
"""

import numpy as np

from .case_setup import PiecewiseForce
from .utils import NodeType

# name -> lattice dims, dx, periodic, features; N = prod(dims) unless a mask removes sites
CASES = {
    # TGV 2D: 2 500 particles, periodic unit box, r = 0.029, no force
    "tgv2d": dict(dims=(50, 50), dx=0.02, periodic=True, force=None, multiplier=1.25),
    # RPF 2D: 3 200 particles, 1 x 2 periodic box, dx = 0.025, r = 0.036, force +-x by y
    "rpf2d": dict(dims=(40, 80), dx=0.025, periodic=True, force="rpf", multiplier=1.25),
    # DAM 2D: ~5 740 particles, walls, gravity, free surface (variable edge count)
    "dam2d": dict(dims=(274, 106), dx=0.02, periodic=False, force="gravity", multiplier=2.0, mask="dam"),
    # LDC 3D (reference dataset size 8 160): walls + moving lid
    "ldc3d": dict(dims=(24, 20, 17), dx=0.05, periodic=False, force=None, multiplier=2.0, walls="ldc"),
    # LDC 3D at the size BASELINE.json quotes (~28k)
    "ldc3d_28k": dict(dims=(40, 28, 25), dx=0.03, periodic=False, force=None, multiplier=2.0, walls="ldc"),
    # RPF 3D 1 M particles: 100 x 200 x 50 lattice at dx = 0.05, periodic, force +-x by y
    "rpf3d_1m": dict(dims=(100, 200, 50), dx=0.05, periodic=True, force="rpf", multiplier=1.25),
    "rpf3d_8k": dict(dims=(20, 40, 10), dx=0.05, periodic=True, force="rpf", multiplier=1.25),
    "rpf3d_125k": dict(dims=(50, 100, 25), dx=0.05, periodic=True, force="rpf", multiplier=1.25),
}


def _radius(dx):
    """1.45 dx rounded to 2 significant digits (data_gen/lagrangebench_data/gen_dataset.py:193-197)."""
    r = 1.45 * dx
    return float(f"{r:.2g}")


def make_case(name, input_seq_length=6, n_future=0, seed=0, dtype=np.float32, dims=None, quiet=False):
    """-> dict(metadata, box, positions (N, T, d), particle_type (N,), force, multiplier, name)

    ``T = input_seq_length + n_future``; the future frames continue the same motion and are
    what kinematic particles are overridden with during a rollout.

    ``quiet=True`` (bench.py): velocity / acceleration statistics four to five orders of magnitude
    below ``dx`` per step, so that a RANDOM-INIT network (whose normalised output is O(1)) keeps the
    cloud next to its lattice for hundreds of steps.  With the default statistics an untrained
    model drives fluid particles through the walls within ~50 steps; they pile up in the border
    cells and the edge count explodes, which measures the blow-up, not the path.  The work per
    step (full neighbor rebuild, same edge count) does not depend on the choice."""
    spec = dict(CASES[name])
    if dims is not None:
        spec["dims"] = tuple(dims)
    dims_, dx = spec["dims"], spec["dx"]
    d = len(dims_)
    rng = np.random.default_rng(seed)
    box = np.array(dims_, dtype=np.float64) * dx
    grid = np.stack(np.meshgrid(*[np.arange(m) for m in dims_], indexing="ij"), axis=-1).reshape(-1, d)
    ptype = np.full(grid.shape[0], int(NodeType.FLUID), dtype=np.int32)
    keep = np.ones(grid.shape[0], dtype=bool)
    if spec.get("mask") == "dam":
        wall = (grid[:, 1] < 3) | ((grid[:, 0] < 3) & (grid[:, 1] < 9))
        fluid = (grid[:, 0] >= 3) & (grid[:, 0] < 101) & (grid[:, 1] >= 3) & (grid[:, 1] < 53)
        keep = wall | fluid
        ptype[wall] = int(NodeType.SOLID_WALL)
    if spec.get("walls") == "ldc":
        dims_a = np.array(dims_)
        shell = ((grid < 3) | (grid >= dims_a - 3)).any(axis=1)
        lid = grid[:, 1] >= dims_a[1] - 3
        ptype[shell] = int(NodeType.SOLID_WALL)
        ptype[lid] = int(NodeType.MOVING_WALL)
    grid, ptype = grid[keep], ptype[keep]
    n = grid.shape[0]
    fluid_mask = ptype == int(NodeType.FLUID)
    pos0 = (grid + 0.5) * dx
    pos0[fluid_mask] += 0.25 * dx * rng.standard_normal((int(fluid_mask.sum()), d))
    vel_std, acc_std = (2.0e-4 * dx, 1.0e-7 * dx) if quiet else (0.05 * dx, 5.0e-4 * dx)
    t_total = input_seq_length + n_future
    vel = vel_std * rng.standard_normal((n, d))
    vel[ptype == int(NodeType.SOLID_WALL)] = 0.0
    lid_v = np.zeros(d)
    lid_v[0] = 1.0e-3 * dx
    vel[ptype == int(NodeType.MOVING_WALL)] = lid_v
    frames = [pos0]
    for _ in range(t_total - 1):
        vel = vel + np.where(fluid_mask[:, None], acc_std * rng.standard_normal((n, d)), 0.0)
        nxt = frames[-1] + vel
        if spec["periodic"]:
            nxt = np.mod(nxt, box)
        else:
            nxt = np.clip(nxt, 1e-6 * dx, box - 1e-6 * dx)
        frames.append(nxt)
    if spec["periodic"]:
        frames[0] = np.mod(frames[0], box)
    else:
        frames[0] = np.clip(frames[0], 1e-6 * dx, box - 1e-6 * dx)
    positions = np.stack(frames, axis=1).astype(dtype)  # (N, T, d)
    if spec["periodic"]:  # the cast may round up onto the box edge
        edge = positions >= box.astype(dtype)
        positions[edge] = 0
    force = None
    if spec["force"] == "rpf":
        lo = [1.0] + [0.0] * (d - 1)
        hi = [-1.0] + [0.0] * (d - 1)
        force = PiecewiseForce(axis=1, threshold=float(box[1] / 2), lo=lo, hi=hi)
    elif spec["force"] == "gravity":
        g = [0.0] * d
        g[1] = -1.0
        force = PiecewiseForce.constant(g)
    metadata = {
        "solver": "synthetic", "dim": d, "dx": dx, "dt": 1.0, "write_every": 1,
        "num_particles_max": int(n),
        "periodic_boundary_conditions": [bool(spec["periodic"])] * d,
        "bounds": [[0.0, float(b)] for b in box],
        "default_connectivity_radius": _radius(dx),
        "vel_mean": [0.0] * d, "vel_std": [float(vel_std)] * d,
        "acc_mean": [0.0] * d, "acc_std": [float(acc_std)] * d,
    }
    return dict(name=name, metadata=metadata, box=box, positions=positions, particle_type=ptype, force=force,
                multiplier=spec["multiplier"], input_seq_length=input_seq_length)


class SyntheticDataset:
    """Dataset-shaped wrapper (``__len__``, ``__getitem__`` -> ``(pos (N, T, d), particle_type)``,
    ``metadata``, ``input_seq_length``, ``num_samples``) like ``H5Dataset`` for a test split."""

    def __init__(self, name, input_seq_length=6, n_rollout_steps=20, n_trajs=1, seed=0, dtype=np.float32, dims=None):
        self.cases = [make_case(name, input_seq_length, n_rollout_steps, seed + i, dtype, dims) for i in range(n_trajs)]
        self.metadata = self.cases[0]["metadata"]
        self.input_seq_length = input_seq_length
        self.num_samples = n_trajs
        self.external_force_fn = self.cases[0]["force"]

    def __len__(self):
        return self.num_samples

    def __getitem__(self, i):
        c = self.cases[i]
        return c["positions"], c["particle_type"]
