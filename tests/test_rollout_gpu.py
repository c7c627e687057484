"""Rollout loop parity: the reference's Lennard-Jones identity test on the GPU path, the
device-resident engine against the oracle loop, and the overflow -> re-allocate -> retry
contract (evaluate/rollout.py:135-151)."""

import json
import os

import numpy as np
import pytest
import torch

from helpers import build_pair, rel_err
from lagrangebench_b200 import GNS, MetricsComputer, RolloutEngine, case_builder, eval_rollout, infer, synthetic
from lagrangebench_b200.evaluate import SimpleLoader, _eval_batched_rollout
from oracle import gns as ogns
from oracle import rollout as orollout

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("n_extrap_steps", [0, 5, 10])
def test_lj_rollout_identity(n_extrap_steps):
    """reference tests/rollout_test.py:74-195 through case.preprocess_eval / integrate kernels."""
    with open(os.path.join(GOLDEN, "lj3d_metadata.json")) as f:
        metadata = json.load(f)
    pos_tnd = np.load(os.path.join(GOLDEN, "lj3d_valid_position.npy"))
    isl, n_rollout = 3, 100
    positions = pos_tnd[: isl + n_rollout].transpose(1, 0, 2)  # (N, T, d)
    ptype = np.zeros(3, np.int32)
    box = np.array([5.0, 5.0, 5.0])
    case = case_builder(box, metadata, isl, noise_std=0.0, dtype="float64")
    pos64 = torch.as_tensor(positions, dtype=torch.float64)
    vels = case.displacement(pos64[:, 1:], pos64[:, :-1])
    accs = vels[:, 1:] - vels[:, :-1]
    st = case.normalization_stats["acceleration"]
    accs = ((accs - torch.as_tensor(st["mean"])) / torch.as_tensor(st["std"])).to(torch.float32)

    def model_apply(params, state, sample):  # CheatingModel (rollout_test.py:92-106)
        i = state["counter"]
        return {"acc": accs[:, min(i, accs.shape[1] - 1)]}, {"counter": i + 1}

    _, nbrs = case.allocate_eval((positions[:, :isl], ptype))
    mc = MetricsComputer(["mse"], case.displacement, metadata, isl)
    pred, metrics, _ = _eval_batched_rollout(model_apply, case, None, {"counter": isl - 2},
                                             (positions[None], ptype[None]), nbrs, mc, n_rollout, isl, n_extrap_steps)
    assert pred.shape[1] == n_rollout + n_extrap_steps
    assert np.isclose(float(metrics[0]["mse"].mean()), 0.0, atol=1e-6)
    full = np.concatenate([positions.transpose(1, 0, 2)[:isl], pred[0].cpu().numpy()])
    assert np.isclose(full[100, 0], positions.transpose(1, 0, 2)[100, 0], atol=1e-6).all()


def _setup(name, dtype, n_steps, mp=2, seed=0):
    c, ours, orac = build_pair(name, dtype, n_future=n_steps, seed=seed)
    d = c["metadata"]["dim"]
    f_cpu, nb_cpu = orac.allocate_eval((c["positions"][:, :6], c["particle_type"]))
    node_in = sum(f_cpu[k].reshape(f_cpu[k].shape[0], -1).shape[1] for k in ("vel_hist", "bound", "force") if k in f_cpu)
    params = ogns.init_params(node_in, d + 1, d, num_mp_steps=mp, seed=seed + 2)
    return c, ours, orac, params, GNS(d, 128, 2, mp, 16), nb_cpu


@pytest.mark.parametrize("name,dtype", [("tgv2d", "float32"), ("ldc3d", "float64")])
def test_engine_rollout_matches_oracle_loop(name, dtype):
    n_steps = 4
    c, ours, orac, params, model, nb_cpu = _setup(name, dtype, n_steps)
    npd = np.float32 if dtype == "float32" else np.float64

    def oracle_apply(p, state, sample):
        feats, ptype = sample
        f32 = {k: (np.asarray(v).astype(np.float32) if np.asarray(v).dtype.kind == "f" else v) for k, v in feats.items()}
        return ogns.forward(p, f32, ptype, 2, np.float32), state

    ref, _ = orollout.eval_batched_rollout(oracle_apply, orac, params, {}, (c["positions"][None].astype(npd),
                                           c["particle_type"][None]), nb_cpu, n_steps, 6)
    engine = RolloutEngine(ours, model, params, steps_per_sync=3)
    window = torch.as_tensor(c["positions"][:, :6]).cuda().contiguous()
    targets = torch.as_tensor(c["positions"][:, 6:6 + n_steps]).permute(1, 0, 2)
    preds, _ = engine.run(window, c["particle_type"], targets, n_steps)
    got = preds.cpu().numpy()
    # 3 + 1 steps: one host sync per chunk (plus one per re-allocation), not one per step
    assert engine.n_launch_calls == 2 + engine.n_reallocations
    # accelerations agree to ~1e-6 relative; positions additionally carry their own rounding
    # (a few ulp of the box size in the position dtype)
    tol = 1e-6 * c["metadata"]["dx"] + 8 * np.finfo(npd).eps * float(np.max(c["box"]))
    assert np.abs(got - ref[0]).max() <= tol, "positions drift from the oracle loop"
    kin = np.isin(c["particle_type"], [1, 2, -1])
    if kin.any():  # kinematic particles follow the ground truth exactly (rollout.py:64-69)
        assert np.array_equal(got[:, kin], c["positions"][kin, 6:6 + n_steps].transpose(1, 0, 2))
    assert np.array_equal(window[:, -1].cpu().numpy(), got[-1])


def test_engine_equals_per_step_loop_bitwise():
    """Device-resident loop and the generic per-step loop launch the same kernels."""
    n_steps = 3
    c, ours, _, params, model, _ = _setup("rpf2d", "float32", n_steps)
    batch = (c["positions"][None], c["particle_type"][None])
    _, nbrs = ours.allocate_eval((c["positions"][:, :6], c["particle_type"]))
    slow, _, _ = _eval_batched_rollout(model.apply, ours, params, {}, batch, nbrs, None, n_steps, 6)
    engine = RolloutEngine(ours, model, params)
    fast, _, _ = _eval_batched_rollout(model.apply, ours, params, {}, batch, None, None, n_steps, 6, engine=engine)
    assert torch.equal(slow, fast)


def test_per_step_calls_replay_the_cached_graph_bitwise():
    """steps_per_sync = 1: from the second sight of the same buffers on, every call replays the step
    graph the library kept; the trajectory must equal the one-call rollout bit for bit, and the
    library must report the same number of kernel launches per step either way."""
    from lagrangebench_b200 import _cabi

    lib = _cabi.load()
    n_steps = 8
    c, ours, _, params, model, _ = _setup("tgv2d", "float32", n_steps)
    ptype = torch.as_tensor(c["particle_type"]).cuda().to(torch.int32)
    targets = torch.as_tensor(c["positions"][:, 6:6 + n_steps]).permute(1, 0, 2).cuda().contiguous()
    w_ref = torch.as_tensor(c["positions"][:, :6]).cuda().contiguous()
    l0 = lib.lb200_launch_count()
    ref_engine = RolloutEngine(ours, model, params)
    ref, _ = ref_engine.run(w_ref, ptype, targets, n_steps)
    ref_launches = lib.lb200_launch_count() - l0 - ref_engine.n_launch_calls  # one init kernel per call
    window = torch.as_tensor(c["positions"][:, :6]).cuda().contiguous()
    engine = RolloutEngine(ours, model, params, steps_per_sync=1)
    tgt1 = torch.empty_like(targets[:1])
    out = torch.empty_like(ref)
    nbrs = None
    l0 = lib.lb200_launch_count()
    for t in range(n_steps):
        tgt1[0].copy_(targets[t])
        p, nbrs = engine.run(window, ptype, tgt1, 1, nbrs)
        out[t].copy_(p[0])
        del p  # the caching allocator hands the same output block to the next call
    got_launches = lib.lb200_launch_count() - l0 - engine.n_launch_calls
    assert torch.equal(out, ref)
    assert torch.equal(window, w_ref)
    if ref_engine.n_reallocations == 0 and engine.n_reallocations == 0:
        # replayed launches are counted like eager ones (bench.py's gpu_launches)
        assert got_launches == ref_launches, (got_launches, ref_launches)


def test_overflow_reallocates_and_retries_same_step():
    n_steps = 3
    c, ours, _, params, model, _ = _setup("tgv2d", "float32", n_steps)
    window0 = torch.as_tensor(c["positions"][:, :6]).cuda().contiguous()
    good_engine = RolloutEngine(ours, model, params)
    ref, _ = good_engine.run(window0.clone(), c["particle_type"], None, n_steps)
    assert good_engine.n_reallocations == 0
    # a neighbor list whose capacity is far too small for the cloud: first step overflows
    tiny = ours._lb200["neighbor_fn"].allocate(window0[:, -1].contiguous())
    tiny.idx = tiny.idx[:, :1000].contiguous()
    tiny.max_occupancy = 1000
    assert tiny.max_occupancy < good_engine._cfg.e_cap
    engine = RolloutEngine(ours, model, params)
    got, nbrs = engine.run(window0.clone(), c["particle_type"], None, n_steps, neighbors=tiny)
    assert engine.n_reallocations >= 1
    assert torch.equal(got, ref), "the retried step must reproduce the un-overflowed rollout"
    assert nbrs.max_occupancy >= good_engine._cfg.e_cap


def test_infer_and_eval_rollout_api(tmp_path):
    ds = synthetic.SyntheticDataset("rpf2d", 6, n_rollout_steps=5, n_trajs=3, seed=2)
    case = case_builder(ds.cases[0]["box"], ds.metadata, 6, cfg_neighbors={"multiplier": 1.25},
                        external_force_fn=ds.external_force_fn, dtype="float32")
    model = GNS(2, 128, 2, 2, 16)
    feats, _ = case.allocate_eval(ds[0])
    params, state = model.init(0, (feats, ds[0][1]))
    metrics = infer(model, case, ds, params=params, state=state, cfg_eval_infer={"batch_size": 2, "metrics": ["mse"],
                    "out_type": "pkl", "n_trajs": -1}, rollout_dir=str(tmp_path), n_rollout_steps=5, seed=0)
    assert sorted(metrics) == ["rollout_0", "rollout_1", "rollout_2"]
    assert metrics["rollout_0"]["mse"].shape == (5,) and "mse1" in metrics["rollout_0"]
    import pickle
    with open(tmp_path / "rollout_1.pkl", "rb") as f:
        payload = pickle.load(f)
    assert payload["predicted_rollout"].shape == (11, 3200, 2)
    assert payload["ground_truth_rollout"].shape == (11, 3200, 2)
    assert np.array_equal(payload["predicted_rollout"][:6], payload["ground_truth_rollout"][:6])
    # the trajectories of a batch run on concurrent engines (the reference's vmap axis): same numbers as one by one
    one_by_one = infer(model, case, ds, params=params, state=state, cfg_eval_infer={"batch_size": 1, "metrics": ["mse"],
                       "out_type": "none", "n_trajs": -1}, n_rollout_steps=5, seed=0)
    for k in metrics:
        assert torch.equal(torch.as_tensor(metrics[k]["mse"]).cpu(), torch.as_tensor(one_by_one[k]["mse"]).cpu())


def test_infer_from_an_hdf5_dataset_with_saved_checkpoint(tmp_path):
    """The reference's inference entry point end to end on files: a dataset directory
    (``metadata.json`` + gzip-chunked ``test.h5``) read by ``H5Dataset``, parameters restored from
    a ``save_haiku`` checkpoint (``load_ckp``), ``infer`` rolling out on the GPU; the result must
    equal the same rollout from the in-memory arrays."""
    import json

    from h5write import write_h5
    from lagrangebench_b200.data import H5Dataset
    from lagrangebench_b200.utils import save_haiku

    n_steps, n_traj = 5, 2
    cases = [synthetic.make_case("rpf2d", 6, n_steps, seed=5 + i, dtype=np.float32) for i in range(n_traj)]
    root = tmp_path / "2D_RPF_3200_20kevery100"
    root.mkdir()
    groups = {f"{i:05d}": {"position": np.ascontiguousarray(c["positions"].transpose(1, 0, 2)),
                           "particle_type": c["particle_type"].astype(np.int32)} for i, c in enumerate(cases)}
    write_h5(str(root / "test.h5"), groups, chunk_rows=4)
    (root / "metadata.json").write_text(json.dumps(cases[0]["metadata"]))
    ds = H5Dataset("test", str(root), input_seq_length=6, extra_seq_length=n_steps)
    assert ds.name == "rpf2d" and len(ds) == n_traj and ds.external_force_fn is not None
    assert np.array_equal(ds[1][0], cases[1]["positions"])
    case = case_builder(cases[0]["box"], ds.metadata, 6, cfg_neighbors={"multiplier": 1.25},
                        external_force_fn=ds.external_force_fn, dtype="float32")
    model = GNS(2, 128, 2, 2, 16)
    feats, _ = case.allocate_eval(ds[0])
    params, state = model.init(3, (feats, ds[0][1]))
    save_haiku(str(tmp_path / "ckp"), params, state)
    cfg = {"batch_size": 1, "metrics": ["mse"], "out_type": "none", "n_trajs": -1}
    from_files = infer(model, case, ds, load_ckp=str(tmp_path / "ckp"), cfg_eval_infer=cfg, n_rollout_steps=n_steps)
    mem = synthetic.SyntheticDataset("rpf2d", 6, n_rollout_steps=n_steps, n_trajs=n_traj, seed=5)
    from_memory = infer(model, case, mem, params=params, state=state, cfg_eval_infer=cfg, n_rollout_steps=n_steps)
    for k in ("rollout_0", "rollout_1"):
        a, b = torch.as_tensor(from_files[k]["mse"]).cpu(), torch.as_tensor(from_memory[k]["mse"]).cpu()
        assert a.shape == (n_steps,) and torch.equal(a, b)


def test_push_forward_unroll_reuses_the_rollout_path():
    """``train/strats.py:137-159``: the forward-only unroll of the pushforward trick is one rollout step --
    model forward, integrate, window shift, neighbor / feature update -- checked against the oracle's."""
    from lagrangebench_b200 import push_forward_build
    from oracle import gns as ogns

    c, ours, orac, params, model, _ = _setup("rpf2d", "float64", 0)
    ptype = c["particle_type"]
    cur_g = torch.as_tensor(c["positions"][:, :6]).cuda()
    cur_c = c["positions"][:, :6].copy()
    f_g, n_g = ours.allocate_eval((cur_g, ptype))
    f_c, n_c = orac.allocate_eval((cur_c, ptype))
    pf = push_forward_build(model.apply, ours)
    for _ in range(2):
        cur_g, n_g, f_g = pf(f_g, cur_g, ptype, n_g, params, {})
        f32 = {k: (np.asarray(v).astype(np.float32) if np.asarray(v).dtype.kind == "f" else v) for k, v in f_c.items()}
        pred = ogns.forward(params, f32, ptype, model._mp_steps, np.float32)
        nxt = orac.integrate(pred, cur_c)
        cur_c = np.concatenate([cur_c[:, 1:], nxt[:, None]], axis=1)
        f_c, n_c = orac.preprocess_eval((cur_c, ptype), n_c)
    assert cur_g.shape == cur_c.shape and np.abs(cur_g.cpu().numpy() - cur_c).max() <= 1e-8  # 1e-5 of an acceleration
    assert np.array_equal(n_g.idx.cpu().numpy(), n_c.idx)
    assert np.allclose(f_g["vel_hist"].cpu().numpy(), f_c["vel_hist"], atol=1e-4)


def test_host_buffer_mode_equals_device_buffers():
    """``run(..., targets=<pinned host tensor>, host_out=<pinned host tensor>)``: upload, steps and read-back ride
    on the engine's stream behind one synchronisation per chunk; same numbers as with device tensors."""
    n_steps = 5
    c, ours, _, params, model, _ = _setup("ldc3d", "float64", n_steps)
    targets = torch.as_tensor(c["positions"][:, 6:6 + n_steps]).permute(1, 0, 2).contiguous()
    w_dev = torch.as_tensor(c["positions"][:, :6]).cuda().contiguous()
    ref, _ = RolloutEngine(ours, model, params).run(w_dev, c["particle_type"], targets.cuda(), n_steps)
    h_targets, h_out = targets.pin_memory(), torch.empty_like(targets).pin_memory()
    w_host = torch.as_tensor(c["positions"][:, :6]).cuda().contiguous()
    eng = RolloutEngine(ours, model, params, steps_per_sync=2)
    got, _ = eng.run(w_host, c["particle_type"], h_targets, n_steps, host_out=h_out)
    assert torch.equal(got, ref) and torch.equal(h_out.cuda(), ref) and torch.equal(w_host, w_dev)
    # per-step calls with one-frame host buffers (bench.py's e2e leg)
    w_step = torch.as_tensor(c["positions"][:, :6]).cuda().contiguous()
    eng1 = RolloutEngine(ours, model, params, steps_per_sync=1)
    d_out, h1, nb = torch.empty_like(ref[:1]), torch.empty_like(targets[:1]).pin_memory(), None
    for t in range(n_steps):
        _, nb = eng1.run(w_step, c["particle_type"], h_targets[t:t + 1], 1, nb, out=d_out, host_out=h1)
        assert torch.equal(h1.cuda(), ref[t:t + 1])
