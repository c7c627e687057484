"""The caller side of the step loop: HDF5 reader + ``H5Dataset`` windowing (``lagrangebench/data/data.py``).

Pinned against the reference's own Lennard-Jones fixture files when the reference tree is present
(this container), and against files written by ``tests/h5write.py`` everywhere."""

import json
import os

import numpy as np
import pytest

from h5write import write_h5
from lagrangebench_b200.data import H5Dataset, dataset_force, get_dataset_name_from_path, numpy_collate
from lagrangebench_b200.h5lite import H5File

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/tests/3D_LJ_3_1214every1"
needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "valid.h5")), reason="reference fixtures absent")


def _make_dataset(tmp_path, n_traj=2, t=40, n=5, d=2, chunk_rows=None, name="2D_TGV_5_test"):
    rng = np.random.default_rng(0)
    root = tmp_path / name
    root.mkdir()
    trajs = {}
    for split in ("train", "valid", "test"):
        groups = {}
        for k in range(n_traj):
            groups[f"{k:05d}"] = {"position": rng.random((t, n, d)).astype(np.float32),
                                  "particle_type": (np.arange(n) % 3).astype(np.int32)}
        write_h5(str(root / f"{split}.h5"), groups, chunk_rows)
        trajs[split] = groups
    meta = {"dim": d, "dx": 0.1, "dt": 1.0, "write_every": 1, "num_particles_max": n + 2,
            "periodic_boundary_conditions": [True] * d, "bounds": [[0.0, 1.0]] * d,
            "default_connectivity_radius": 0.15, "vel_mean": [0.0] * d, "vel_std": [1.0] * d,
            "acc_mean": [0.0] * d, "acc_std": [1.0] * d}
    (root / "metadata.json").write_text(json.dumps(meta))
    return str(root), trajs


@pytest.mark.parametrize("chunk_rows", [None, 7])
def test_reader_roundtrip_contiguous_and_gzip_chunked(tmp_path, chunk_rows):
    path, trajs = _make_dataset(tmp_path, chunk_rows=chunk_rows)
    with H5File(os.path.join(path, "valid.h5")) as f:
        assert f.keys() == ["00000", "00001"]
        assert f.keys("00001") == ["particle_type", "position"]
        assert "00001/position" in f and "00002/position" not in f
        for key, dsets in trajs["valid"].items():
            pos = f[f"{key}/position"]
            assert pos.shape == dsets["position"].shape and pos.dtype == np.float32 and len(pos) == 40
            assert np.array_equal(pos[:], dsets["position"])
            assert np.array_equal(pos[3:17], dsets["position"][3:17])  # straddles chunk boundaries
            assert np.array_equal(pos[-1], dsets["position"][-1]) and np.array_equal(pos[::5], dsets["position"][::5])
            pt = f[f"{key}/particle_type"]
            assert pt.dtype == np.int32 and np.array_equal(pt[:], dsets["particle_type"])
        with pytest.raises(KeyError):
            f["nope/position"]


def test_h5dataset_windowing_matches_the_reference_rules(tmp_path):
    path, trajs = _make_dataset(tmp_path, chunk_rows=7)
    # valid / test: every trajectory cut into sequence_length // (isl + extra) consecutive pieces
    ds = H5Dataset("valid", path, input_seq_length=6, extra_seq_length=10)
    assert ds.name == "tgv2d" and ds.external_force_fn is None and ds.traj_keys == ["00000", "00001"]
    assert (ds.sequence_length, ds.subseq_length, len(ds)) == (40, 16, 2 * (40 // 16))
    pos, ptype = ds[3]  # trajectory 1, second piece
    assert pos.shape == (5, 16, 2) and pos.dtype == np.float32 and ptype.shape == (5,)
    assert np.array_equal(pos, trajs["valid"]["00001"]["position"][16:32].transpose(1, 0, 2))
    # train: sliding windows of isl + 1 + extra frames
    tr = H5Dataset("train", path, input_seq_length=6, extra_seq_length=2)
    assert tr.subseq_length == 9 and len(tr) == 2 * (40 - 9 + 1)
    pos, _ = tr[32 + 4]  # window 4 of trajectory 1
    assert np.array_equal(pos, trajs["train"]["00001"]["position"][4:13].transpose(1, 0, 2))
    # matscipy backend pads to num_particles_max with PAD particles (data.py:183-197)
    pad = H5Dataset("test", path, input_seq_length=6, extra_seq_length=10, nl_backend="matscipy")
    pos, ptype = pad[0]
    assert pos.shape == (7, 16, 2) and ptype.tolist() == [0, 1, 2, 0, 1, -1, -1] and (pos[5:] == 0).all()
    with pytest.raises(ValueError):
        H5Dataset("valid", path, input_seq_length=6, extra_seq_length=0)
    batch = numpy_collate([ds[0], ds[1]])
    assert batch[0].shape == (2, 5, 16, 2) and batch[1].shape == (2, 5)


def test_dataset_names_and_forces():
    assert get_dataset_name_from_path("datasets/2D_RPF_3200_20kevery100/") == "rpf2d"
    assert get_dataset_name_from_path("/x/3D_LDC_8160_10kevery100") == "ldc3d"
    with pytest.warns(UserWarning):
        assert get_dataset_name_from_path("/x/mydata") == "mydata"
    meta = {"dim": 2, "bounds": [[0.0, 1.0], [0.0, 2.0]]}
    f = dataset_force("rpf2d", meta)
    assert (f.axis, f.threshold, f.lo, f.hi) == (1, 1.0, [1.0, 0.0], [-1.0, 0.0])
    assert dataset_force("dam2d", meta) is None and dataset_force("tgv2d", meta) is None  # dam2d: force.py or nothing


FORCE_FILES = {
    # the forms the reference's datasets ship (single-particle ``force_fn(r)``, vmapped in features.py:105-107)
    "rpf": ("import jax.numpy as jnp\n\n\ndef force_fn(r):\n"
            "    return jnp.where(r[1] > 1.0, jnp.array([-1.0, 0.0]), jnp.array([1.0, 0.0]))\n"),
    "dam": "import jax.numpy as jnp\n\n\ndef force_fn(r):\n    return jnp.array([0.0, -1.0])\n",
    "dam_up": "import jax.numpy as jnp\n\n\ndef force_fn(r):\n    return jnp.ones_like(r) * jnp.array([0.0, 1.0])\n",
    "ge": ("import jax.numpy as jnp\n\n\ndef force_fn(r):\n"
           "    return jnp.where(r[0] >= 0.25, jnp.array([0.0, 2.0]), jnp.array([0.0, -2.0]))\n"),
    "smooth": "import jax.numpy as jnp\n\n\ndef force_fn(r):\n    return jnp.array([jnp.sin(r[1]), 0.0])\n",
}


def test_force_py_is_read_and_probed(tmp_path):
    """``data.py:87-101``: the dataset's own ``force.py`` decides the force -- sign, threshold and the side
    the threshold itself falls on -- not a name-keyed table."""
    from lagrangebench_b200.case_setup import PiecewiseForce
    from lagrangebench_b200.data import force_from_file

    meta = {"dim": 2, "bounds": [[0.0, 1.0], [0.0, 2.0]]}
    paths = {}
    for k, src in FORCE_FILES.items():
        paths[k] = tmp_path / f"force_{k}.py"
        paths[k].write_text(src)
    f = force_from_file(str(paths["rpf"]), meta)
    assert isinstance(f, PiecewiseForce) and (f.axis, f.threshold, f.lo, f.hi) == (1, 1.0, [1.0, 0.0], [-1.0, 0.0])
    f = force_from_file(str(paths["dam"]), meta)
    assert isinstance(f, PiecewiseForce) and f.lo == f.hi == [0.0, -1.0]
    assert force_from_file(str(paths["dam_up"]), meta).lo == [0.0, 1.0]  # the generator notebook's sign
    f = force_from_file(str(paths["ge"]), meta)  # ">=": the threshold is the last double below 0.25
    assert isinstance(f, PiecewiseForce) and f.axis == 0 and f.threshold == np.nextafter(0.25, 0.0)
    assert f(np.array([[0.25, 1.0]])).tolist() == [[0.0, 2.0]] and f(np.array([[0.2, 1.0]])).tolist() == [[0.0, -2.0]]
    with pytest.warns(UserWarning):
        g = force_from_file(str(paths["smooth"]), meta)
    assert not isinstance(g, PiecewiseForce)
    assert np.allclose(np.asarray(g(np.array([[0.3, 0.5]]))), [[np.sin(0.5), 0.0]])


@needs_ref
def test_reader_on_the_reference_fixture_files():
    """``tests/3D_LJ_3_1214every1/*.h5`` as written by h5py: float32 (T, 3, 3), gzip chunks of 203 frames."""
    gold = np.load(os.path.join(HERE, "golden", "lj3d_valid_position.npy"))
    with H5File(os.path.join(REF, "valid.h5")) as f:
        assert f.keys() == ["00000"]
        pos = f["00000/position"]
        assert pos.shape == (405, 3, 3) and pos.dtype == np.float32
        assert np.array_equal(pos[:], gold) and np.array_equal(pos[190:215], gold[190:215])
        assert f["00000/particle_type"][:].tolist() == [0, 0, 0]
    with H5File(os.path.join(REF, "train.h5")) as f:
        assert f["00000/position"].shape == (1214, 3, 3)


@needs_ref
def test_h5dataset_on_the_reference_fixture():
    """The split sizes the reference's rollout test relies on (``tests/rollout_test.py:28-43``)."""
    gold = np.load(os.path.join(HERE, "golden", "lj3d_valid_position.npy"))
    ds = H5Dataset("valid", REF, name="lj3d", input_seq_length=3, extra_seq_length=100)
    assert ds.sequence_length == 405 and len(ds) == 405 // 103
    pos, ptype = ds[1]
    assert np.array_equal(pos, gold[103:206].transpose(1, 0, 2)) and ptype.tolist() == [0, 0, 0]
    assert ds.metadata["num_particles_max"] == 3 and ds.metadata["default_connectivity_radius"] == 3.0
