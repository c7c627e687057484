"""Datasets: host-side mirror of ``lagrangebench/data/data.py`` for the caller side of the rollout.

``H5Dataset`` keeps the reference constructor and sample layout (``data.py:43-51,199-266``):
``dataset[i] -> (positions (N, T, d) float32, particle_type (N,))`` with the reference's
windowing of train / valid / test splits.  The HDF5 files are read with the pure-Python
reader of :mod:`lagrangebench_b200.h5lite` (h5py is not part of this image); there is no
download path (no network) -- ``dataset_path`` must exist.

The reference loads a JAX ``force_fn`` from the dataset's ``force.py`` (``data.py:87-101``).
Here the file is executed too (``force_from_file``; with ``jax.numpy`` standing in as NumPy when
JAX is not installed), probed over the domain and turned into the :class:`PiecewiseForce` the
feature kernel evaluates on the device -- constant fields and ``jnp.where(r[axis] > t, a, b)``
fields, which is what the shipped datasets contain; anything else stays a Python callable
(evaluated on the host side of the per-step loop).  Only when a dataset has no ``force.py`` does
the name-keyed ``dataset_force`` table apply (reverse Poiseuille flow: ``+x`` below ``y = 1``,
``-x`` above, ``notebooks/data_gen.ipynb`` cells 11 / 13 of the reference).
"""

import bisect
import json
import os
import os.path as osp
import re
import warnings

import numpy as np

from .case_setup import PiecewiseForce
from .h5lite import H5File
from .utils import NodeType

DATASET_DIRS = {
    "tgv2d": "2D_TGV_2500_10kevery100", "rpf2d": "2D_RPF_3200_20kevery100", "ldc2d": "2D_LDC_2708_10kevery100",
    "dam2d": "2D_DAM_5740_20kevery100", "tgv3d": "3D_TGV_8000_10kevery100", "rpf3d": "3D_RPF_8000_10kevery100",
    "ldc3d": "3D_LDC_8160_10kevery100",
}


def get_dataset_name_from_path(path):
    """``{2D|3D}_{ABC}_...`` -> ``abc2d`` / ``abc3d``; otherwise the directory name (``data.py:272-298``)."""
    directory = osp.basename(osp.normpath(path))
    m = re.search(r"(?:2D|3D)_[A-Z]{3}", directory)
    if m is None:
        warnings.warn(f"Dataset directory {directory} does not follow the lagrangebench convention "
                      "{2D|3D}_{TGV|RPF|LDC|DAM}; pass the dataset name explicitly.")
        return directory
    dim, abc = m.group(0).split("_")
    return f"{abc}{dim}".lower()


def dataset_force(name, metadata):
    """External force of a shipped dataset as a device-evaluable :class:`PiecewiseForce`, or None."""
    d = int(metadata["dim"])
    if name in ("rpf2d", "rpf3d"):  # jnp.where(r[:, 1] > 1.0, -1.0, 1.0) * e_x  (data_gen.ipynb cells 11, 13)
        return PiecewiseForce(axis=1, threshold=1.0, lo=[1.0] + [0.0] * (d - 1), hi=[-1.0] + [0.0] * (d - 1))
    return None


def _load_force_module(path):
    """Execute a dataset's ``force.py`` (``data.py:87-95``).  Without JAX, ``jax`` / ``jax.numpy``
    resolve to NumPy for the duration of the import (the files use a handful of array functions)."""
    import importlib.util
    import sys
    import types

    shims = {}
    try:
        import jax  # noqa: F401
    except ImportError:
        fake = types.ModuleType("jax")
        fake.numpy = np
        fake.vmap = lambda f, *a, **k: (lambda x: np.stack([np.asarray(f(row)) for row in x]))
        fake.jit = lambda f, *a, **k: f
        shims = {"jax": fake, "jax.numpy": np}
    saved = {k: sys.modules.get(k) for k in shims}
    sys.modules.update(shims)
    try:
        spec = importlib.util.spec_from_file_location("force_module", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod.force_fn


def force_from_file(path, metadata, n_probe=4096, seed=0):
    """``force.py`` -> :class:`PiecewiseForce` when the field is constant or switches once along one
    axis (checked on probe points, threshold located by bisection, the side the threshold itself falls on
    verified); otherwise the function itself, vectorised over ``(N, d)`` positions."""
    fn = _load_force_module(path)
    d = int(metadata["dim"])
    bounds = np.asarray(metadata["bounds"], dtype=np.float64)
    lo_b, hi_b = bounds[:, 0], bounds[:, 1]
    rng = np.random.default_rng(seed)

    def f(r):
        return np.asarray(fn(np.asarray(r, dtype=np.float64)), dtype=np.float64).reshape(d)

    def vectorised(pos):
        import torch

        p = torch.as_tensor(pos).detach().cpu().numpy()
        out = np.stack([f(row) for row in p]) if len(p) else np.zeros((0, d))
        return torch.as_tensor(out, dtype=torch.float32)

    probes = lo_b + (hi_b - lo_b) * rng.random((n_probe, d))
    vals = np.stack([f(p) for p in probes])
    uniq = np.unique(vals, axis=0)
    if len(uniq) == 1:
        return PiecewiseForce.constant(uniq[0].tolist())
    if len(uniq) == 2:
        for axis in range(d):
            is_a = np.all(vals == uniq[0], axis=1)
            xa, xb = probes[is_a, axis], probes[~is_a, axis]
            if xa.max() < xb.min():
                lo_v, hi_v, left, right = uniq[0], uniq[1], xa.max(), xb.min()
            elif xb.max() < xa.min():
                lo_v, hi_v, left, right = uniq[1], uniq[0], xb.max(), xa.min()
            else:
                continue
            base = 0.5 * (lo_b + hi_b)
            for _ in range(200):  # bisection on the switching coordinate
                mid = 0.5 * (left + right)
                if mid == left or mid == right:
                    break
                q = base.copy()
                q[axis] = mid
                if np.array_equal(f(q), lo_v):
                    left = mid
                else:
                    right = mid
            # PiecewiseForce is hi for r[axis] > threshold: the threshold is the last coordinate with the low value
            cand = PiecewiseForce(axis, left, lo_v.tolist(), hi_v.tolist())
            check = np.concatenate([probes, [np.where(np.arange(d) == axis, left, base),
                                             np.where(np.arange(d) == axis, right, base)]])
            ok = all(np.array_equal(f(p), np.asarray(cand.hi if p[axis] > cand.threshold else cand.lo)) for p in check)
            if ok:
                return cand
    warnings.warn(f"{path}: force field is not piecewise-constant along one axis; it stays a host callable")
    return vectorised


def numpy_collate(batch):
    """Collate helper for torch dataloaders (``data/utils.py:48-58``)."""
    if isinstance(batch[0], np.ndarray):
        return np.stack(batch)
    if isinstance(batch[0], (tuple, list)):
        return type(batch[0])(numpy_collate(samples) for samples in zip(*batch))
    return np.asarray(batch)


class H5Dataset:
    """HDF5 simulation trajectories (``data.py:33-269``).

    * ``split == "train"``: samples are windows of ``input_seq_length + 1 + extra_seq_length``
      consecutive frames, every start position of every trajectory;
    * ``"valid"`` / ``"test"``: every trajectory is cut into ``sequence_length // (input_seq_length +
      extra_seq_length)`` consecutive sub-trajectories of that length (``extra_seq_length`` = rollout
      length of interest, must be > 0).
    """

    def __init__(self, split, dataset_path, name=None, input_seq_length=6, extra_seq_length=0,
                 nl_backend="jaxmd_vmap"):
        dataset_path = osp.normpath(dataset_path)
        self.name = get_dataset_name_from_path(dataset_path) if name is None else name
        if not osp.exists(dataset_path):
            raise FileNotFoundError(f"{dataset_path} does not exist (datasets cannot be downloaded here)")
        if split not in ("train", "valid", "test"):
            raise ValueError(f"unknown split {split!r}")
        if input_seq_length < 2:
            raise ValueError("input_seq_length must be at least 2 (one past velocity needs two positions)")
        self.dataset_path = dataset_path
        self.file_path = osp.join(dataset_path, split + ".h5")
        self.input_seq_length = input_seq_length
        self.nl_backend = nl_backend
        with open(osp.join(dataset_path, "metadata.json")) as f:
            self.metadata = json.load(f)
        force_fn_path = osp.join(dataset_path, "force.py")
        if osp.exists(force_fn_path):  # data.py:87-95
            self.external_force_fn = force_from_file(force_fn_path, self.metadata)
        else:  # the reference raises for dam2d / rpf2d / rpf3d (data.py:96-101); RPF's field is known from its generator
            if self.name == "dam2d":
                raise FileNotFoundError(f"External force function not found in {dataset_path} (force.py).")
            self.external_force_fn = dataset_force(self.name, self.metadata)
        self.db_hdf5 = None
        with H5File(self.file_path) as f:
            self.traj_keys = list(f.keys())
            self.sequence_length = f[f"{self.traj_keys[0]}/position"].shape[0]
        if split == "train":
            self.subseq_length = input_seq_length + 1 + extra_seq_length
            samples_per_traj = self.sequence_length - self.subseq_length + 1
            self._keylen_cumulative = list(np.cumsum([samples_per_traj] * len(self.traj_keys)))
            self.num_samples = int(samples_per_traj * len(self.traj_keys))
            self.getter = self.get_window
        else:
            if extra_seq_length <= 0:
                raise ValueError("valid / test splits need extra_seq_length > 0 (the rollout length of interest)")
            self.subseq_length = input_seq_length + extra_seq_length
            self._split_valid_traj_into_n = self.sequence_length // self.subseq_length
            self.num_samples = self._split_valid_traj_into_n * len(self.traj_keys)
            self.getter = self.get_trajectory
        if self.sequence_length < self.subseq_length:
            raise ValueError(f"trajectories have {self.sequence_length} frames, a sample needs {self.subseq_length}: "
                             "reduce input_seq_length or extra_seq_length")

    def _open_hdf5(self):
        if self.db_hdf5 is None:
            self.db_hdf5 = H5File(self.file_path)
        return self.db_hdf5

    def _matscipy_pad(self, pos_input, particle_type):
        pad = self.metadata["num_particles_max"] - pos_input.shape[0]
        pos_input = np.pad(pos_input, ((0, pad), (0, 0), (0, 0)), mode="constant", constant_values=0.0)
        particle_type = np.pad(particle_type, (0, pad), mode="constant", constant_values=int(NodeType.PAD_VALUE))
        return pos_input, particle_type

    def _sample(self, traj_idx, frame_from, frame_to):
        db = self._open_hdf5()
        key = self.traj_keys[traj_idx]
        pos = db[f"{key}/position"][frame_from:frame_to].transpose((1, 0, 2))  # (T, N, d) -> (N, T, d)
        particle_type = db[f"{key}/particle_type"][:]
        if self.nl_backend == "matscipy":
            pos, particle_type = self._matscipy_pad(pos, particle_type)
        return np.ascontiguousarray(pos), particle_type

    def get_trajectory(self, idx):
        """A (sub-)trajectory of a validation / test file (``data.py:199-226``)."""
        if self._split_valid_traj_into_n > 1:
            traj_idx = idx // self._split_valid_traj_into_n
            frame_from = (idx % self._split_valid_traj_into_n) * self.subseq_length
            frame_to = frame_from + self.subseq_length
        else:
            traj_idx, frame_from, frame_to = idx, 0, self.sequence_length
        return self._sample(traj_idx, frame_from, frame_to)

    def get_window(self, idx):
        """A training window (``data.py:228-255``)."""
        traj_idx = bisect.bisect(self._keylen_cumulative, idx)
        el_idx = idx - (self._keylen_cumulative[traj_idx - 1] if traj_idx != 0 else 0)
        assert el_idx >= 0
        return self._sample(traj_idx, el_idx, el_idx + self.subseq_length)

    def __getitem__(self, idx):
        return self.getter(int(idx))

    def __len__(self):
        return self.num_samples

    def close(self):
        if self.db_hdf5 is not None:
            self.db_hdf5.close()
            self.db_hdf5 = None


def _named(name):
    class _Named(H5Dataset):
        __doc__ = f"{name} dataset (``data.py:301-445``)."

        def __init__(self, split, dataset_path=osp.join("datasets", DATASET_DIRS[name]), input_seq_length=6,
                     extra_seq_length=0, nl_backend="jaxmd_vmap"):
            super().__init__(split, dataset_path, name=name, input_seq_length=input_seq_length,
                             extra_seq_length=extra_seq_length, nl_backend=nl_backend)

    _Named.__name__ = _Named.__qualname__ = name[:3].upper() + name[3:].upper()
    return _Named


TGV2D, TGV3D, RPF2D, RPF3D = _named("tgv2d"), _named("tgv3d"), _named("rpf2d"), _named("rpf3d")
LDC2D, LDC3D, DAM2D = _named("ldc2d"), _named("ldc3d"), _named("dam2d")
