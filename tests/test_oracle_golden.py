"""Pin the oracle against every golden vector the reference's own tests hold for the path.

Replays ``tests/case_test.py:14-206`` (3-particle periodic box) and
``tests/rollout_test.py:74-195`` (Lennard-Jones rollout identity with a cheating model)
of the reference against ``oracle/``.  Values below are the reference's, copied as data.
"""

import json
import os

import numpy as np
import pytest

import oracle
from oracle import case as ocase
from oracle import rollout as orollout
from oracle import space as ospace

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")

METADATA = {  # tests/case_test.py:14-23
    "num_particles_max": 3,
    "periodic_boundary_conditions": [True, True, True],
    "default_connectivity_radius": 0.3,
    "bounds": [[0.0, 1.0], [0.0, 1.0], [0.0, 1.0]],
    "acc_mean": [0.0, 0.0, 0.0],
    "acc_std": [1.0, 1.0, 1.0],
    "vel_mean": [0.0, 0.0, 0.0],
    "vel_std": [1.0, 1.0, 1.0],
}
POSITION = np.array(  # tests/case_test.py:40-64, (N, T, d) = (3, 5, 3)
    [
        [[0.5, 0.5, 0.5]] * 5,
        [[0.7, 0.5, 0.5], [0.9, 0.5, 0.5], [0.1, 0.5, 0.5], [0.3, 0.5, 0.5], [0.5, 0.5, 0.5]],
        [[0.8, 0.6, 0.5], [0.8, 0.6, 0.5], [0.9, 0.6, 0.5], [0.2, 0.6, 0.5], [0.6, 0.6, 0.5]],
    ]
)
PTYPE = np.array([0, 0, 0])


def make_case(dtype):
    return ocase.case_builder(
        np.array([1.0, 1.0, 1.0]), METADATA, input_seq_length=3,
        cfg_neighbors={"backend": "jaxmd_vmap", "multiplier": 1.25},
        cfg_model={"isotropic_norm": False, "magnitude_features": False},
        noise_std=0.0, external_force_fn=None, dtype=dtype)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_allocate_golden(dtype):
    case = make_case(dtype)
    _, features, target, nbrs = case.allocate(None, (POSITION, PTYPE))
    # tests/case_test.py:77-82
    assert (nbrs.idx == np.array([[0, 1, 2, 2, 1, 3], [0, 1, 1, 2, 2, 3]])).all()
    assert not nbrs.did_buffer_overflow
    # tests/case_test.py:86-100
    assert np.isclose(target["vel"], [[0, 0, 0], [0.2, 0, 0], [0.3, 0, 0]]).all()
    assert np.isclose(target["acc"], [[0, 0, 0], [0, 0, 0], [0.2, 0, 0]], atol=1e-7).all()
    # tests/case_test.py:102-114
    assert np.isclose(features["vel_hist"],
                      [[0, 0, 0, 0, 0, 0], [0.2, 0, 0, 0.2, 0, 0], [0, 0, 0, 0.1, 0, 0]], atol=1e-7).all()
    # tests/case_test.py:116-137
    disp = np.array([[0, 0, 0], [0, 0, 0], [-0.2, 0.1, 0], [0, 0, 0], [0.2, -0.1, 0], [0, 0, 0]]) / 0.3
    dist = ((disp**2).sum(-1, keepdims=True)) ** 0.5
    assert np.isclose(features["rel_disp"], disp, atol=1e-6).all()
    assert np.isclose(features["rel_dist"], dist, atol=1e-6).all()


def test_preprocess_golden():
    case = make_case(np.float32)
    _, _, _, nbrs = case.allocate(None, (POSITION, PTYPE))
    _, _, _, nbrs_new = case.preprocess(None, (POSITION, PTYPE), 0.0, nbrs, 0)  # case_test.py:139-148
    assert (nbrs.idx == nbrs_new.idx).all()
    _, _, target, _ = case.preprocess(None, (POSITION, PTYPE), 0.0, nbrs, 1)  # case_test.py:150-163
    assert np.isclose(target["acc"], [[0, 0, 0], [0, 0, 0], [0.1, 0, 0]], atol=1e-7).all()


def test_integrate_golden():
    case = make_case(np.float32)
    acc = {"acc": np.array([[0.0, 0.0, 0.0], [0.0, 0.0, 0.0], [0.2, 0.0, 0.0]])}  # case_test.py:195-206
    new_pos = case.integrate(acc, POSITION[:, :3])
    assert np.isclose(new_pos, POSITION[:, 3]).all()


@pytest.mark.parametrize("n_extrap_steps", [0, 5, 10])
def test_lj_rollout_identity(n_extrap_steps):
    """tests/rollout_test.py:74-195 with x64 (rollout_test.py:14)."""
    with open(os.path.join(GOLDEN, "lj3d_metadata.json")) as f:
        metadata = json.load(f)
    pos_tnd = np.load(os.path.join(GOLDEN, "lj3d_valid_position.npy"))
    isl, n_rollout = 3, 100
    # H5Dataset valid split: first chunk of subseq_length = isl + extra (data.py:136-141,199-211)
    positions = pos_tnd[: isl + n_rollout].transpose(1, 0, 2)  # (N, T, d) float32
    ptype = np.zeros(positions.shape[0], dtype=np.int32)
    bounds = np.array(metadata["bounds"])
    box = bounds[:, 1] - bounds[:, 0]
    case = ocase.case_builder(box, metadata, isl, noise_std=0.0, dtype=np.float64)
    disp, shift = ospace.periodic(box)
    pos64 = positions.astype(np.float64)
    vels = disp(pos64[:, 1:], pos64[:, :-1])
    accs = vels[:, 1:] - vels[:, :-1]
    stats = case.normalization_stats["acceleration"]
    accs = (accs - stats["mean"]) / stats["std"]

    # "Wrong setup" check (rollout_test.py:117-127)
    pred_pos = shift(pos64[:, isl - 1], vels[:, isl - 2] + (stats["mean"] + accs[:, 0] * stats["std"]))
    assert np.isclose(pred_pos.astype(np.float32), positions[:, isl], atol=1e-7).all()

    def model_apply(params, state, sample):  # CheatingModel, rollout_test.py:92-106
        i = state["counter"]
        i_c = min(i, accs.shape[1] - 1)
        return {"acc": accs[:, i_c]}, {"counter": i + 1}

    _, nbrs = case.allocate_eval((positions[:, :isl], ptype))
    # the LJ box is smaller than 3 cutoffs: all-pairs candidate path
    assert nbrs.cell_list_capacity is None
    pred, _ = orollout.eval_batched_rollout(
        model_apply, case, None, {"counter": isl - 2}, (positions[None], ptype[None]), nbrs,
        n_rollout_steps=n_rollout, t_window=isl, n_extrap_steps=n_extrap_steps)
    assert pred.shape[1] == n_rollout + n_extrap_steps
    target = positions[:, isl:isl + n_rollout].transpose(1, 0, 2)
    mse = orollout.mse(case.displacement, pred[0, :n_rollout].astype(np.float64), target.astype(np.float64))
    assert np.isclose(mse.mean(), 0.0, atol=1e-6)
    full = np.concatenate([positions.transpose(1, 0, 2)[:isl], pred[0]], axis=0)
    assert np.isclose(full[100, 0], positions.transpose(1, 0, 2)[100, 0], atol=1e-6).all()


def test_gns_param_count():
    """docs/pages/baselines.rst:55,62 publish 161K / 1.2M parameters."""
    from oracle import gns

    # RPF-2D: K*d + d = 12 node features (vel_hist 10 + force 2), edges d+1 = 3
    assert gns.num_params(gns.init_params(12, 3, 2, latent=128, num_mp_steps=10)) == 1211794
    assert gns.num_params(gns.init_params(12, 3, 2, latent=64, num_mp_steps=5)) == 161042


def test_segment_sum_drops_pad():
    from oracle import gns

    data = np.arange(12, dtype=np.float64).reshape(6, 2)
    out = gns.segment_sum(data, np.array([0, 2, 2, 3, 0, 3]), 3)
    assert np.array_equal(out, [[8, 10], [0, 0], [6, 8]])


def test_num_particles_keeps_padding_out_of_the_search():
    """``case.py:182-190`` + ``data.py:183-197``: trailing PAD_VALUE rows (all at the origin) are
    neither senders nor receivers; the list of the real particles is unchanged, the pad value is
    the padded row count."""
    from lagrangebench_b200 import synthetic

    c = synthetic.make_case("tgv2d", 6, 0, 0, np.float32)
    orac = ocase.case_builder(c["box"], c["metadata"], 6, cfg_neighbors={"multiplier": 1.25}, dtype=np.float32)
    n_pad = 77
    pos = np.concatenate([c["positions"][:, :6], np.zeros((n_pad, 6, 2), np.float32)])
    ptype = np.concatenate([c["particle_type"], np.full(n_pad, -1, np.int32)])
    _, padded = orac.allocate_eval((pos, ptype))
    _, plain = orac.allocate_eval((c["positions"][:, :6], c["particle_type"]))
    e = plain.n_edges
    assert padded.n_edges == e and padded.max_occupancy == plain.max_occupancy
    assert np.array_equal(padded.idx[:, :e], plain.idx[:, :e])
    assert (padded.idx[:, e:] == pos.shape[0]).all()
    _, again = orac.preprocess_eval((pos, ptype), padded)
    assert np.array_equal(again.idx, padded.idx) and not again.did_buffer_overflow


def test_torch_twin_of_the_forward_agrees_with_the_numpy_oracle():
    """oracle/gns_torch.py (the CPU baseline bench.py times) is the same restatement as oracle/gns.py."""
    from lagrangebench_b200 import synthetic
    from oracle import gns as ogns
    from oracle import gns_torch

    c = synthetic.make_case("ldc3d", 6, 0, 0, np.float32, dims=(10, 9, 8))
    orac = ocase.case_builder(c["box"], c["metadata"], 6, cfg_neighbors={"multiplier": 2.0}, dtype=np.float32)
    ptype = c["particle_type"].copy()
    ptype[-3:] = -1  # padding rows: hk.Embed takes the last row of the table
    feats, _ = orac.allocate_eval((c["positions"], ptype))
    params = ogns.init_params(21, 4, 3, num_mp_steps=4, seed=2)
    ref = ogns.forward(params, feats, ptype, 4, np.float64)["acc"]
    f32 = ogns.forward(params, feats, ptype, 4, np.float32)["acc"]
    twin = gns_torch.forward(gns_torch.pack(params), feats, ptype, 4)["acc"]
    scale = np.abs(ref).max()
    assert np.abs(twin - ref).max() <= 1e-5 * scale
    assert np.abs(twin - f32).max() <= 1e-5 * scale
