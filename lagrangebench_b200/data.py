"""Datasets: host-side mirror of ``lagrangebench/data/data.py`` for the caller side of the rollout.

``H5Dataset`` keeps the reference constructor and sample layout (``data.py:43-51,199-266``):
``dataset[i] -> (positions (N, T, d) float32, particle_type (N,))`` with the reference's
windowing of train / valid / test splits.  The HDF5 files are read with the pure-Python
reader of :mod:`lagrangebench_b200.h5lite` (h5py is not part of this image); there is no
download path (no network) -- ``dataset_path`` must exist.

The reference loads a JAX ``force_fn`` from the dataset's ``force.py`` (``data.py:87-101``).
JAX is not available here and the feature kernel evaluates forces on the device, so the
force fields of the shipped datasets are provided as :class:`PiecewiseForce` objects keyed
by dataset name (``dataset_force``): reverse Poiseuille flow pushes ``+x`` in the lower
half of the box and ``-x`` in the upper half, the dam break has constant gravity along
``-y`` (``notebooks/data_gen.ipynb`` cells 11-19 of the reference).
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
    if name in ("rpf2d", "rpf3d"):
        y_mid = 0.5 * (metadata["bounds"][1][0] + metadata["bounds"][1][1])
        return PiecewiseForce(axis=1, threshold=y_mid, lo=[1.0] + [0.0] * (d - 1), hi=[-1.0] + [0.0] * (d - 1))
    if name == "dam2d":
        g = [0.0] * d
        g[1] = -1.0
        return PiecewiseForce.constant(g)
    return None


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
