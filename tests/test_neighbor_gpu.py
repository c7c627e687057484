"""Parity of kernel (i) and the feature/integrate kernels with the oracle: bit-exact edge
lists (order included), exact float32 features."""

import json
import os

import numpy as np
import pytest
import torch

from helpers import build_pair
from lagrangebench_b200 import case_builder
from oracle import case as ocase

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")

METADATA = {  # reference tests/case_test.py:14-23
    "num_particles_max": 3, "periodic_boundary_conditions": [True, True, True],
    "default_connectivity_radius": 0.3, "bounds": [[0.0, 1.0]] * 3,
    "acc_mean": [0.0] * 3, "acc_std": [1.0] * 3, "vel_mean": [0.0] * 3, "vel_std": [1.0] * 3,
}
POSITION = np.array([  # tests/case_test.py:40-64
    [[0.5, 0.5, 0.5]] * 5,
    [[0.7, 0.5, 0.5], [0.9, 0.5, 0.5], [0.1, 0.5, 0.5], [0.3, 0.5, 0.5], [0.5, 0.5, 0.5]],
    [[0.8, 0.6, 0.5], [0.8, 0.6, 0.5], [0.9, 0.6, 0.5], [0.2, 0.6, 0.5], [0.6, 0.6, 0.5]]])
PTYPE = np.array([0, 0, 0])


@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_reference_case_test_vectors(dtype):
    """The reference's own test_allocate / test_preprocess_* / test_integrate, on the GPU path."""
    case = case_builder(np.ones(3), METADATA, 3, cfg_neighbors={"multiplier": 1.25},
                        cfg_model={"isotropic_norm": False, "magnitude_features": False}, noise_std=0.0,
                        dtype=dtype)
    key, features, target, nbrs = case.allocate(0, (POSITION, PTYPE))
    assert nbrs.idx.cpu().tolist() == [[0, 1, 2, 2, 1, 3], [0, 1, 1, 2, 2, 3]]
    assert not bool(nbrs.did_buffer_overflow)
    assert np.isclose(target["vel"].cpu(), [[0, 0, 0], [0.2, 0, 0], [0.3, 0, 0]]).all()
    assert np.isclose(target["acc"].cpu(), [[0, 0, 0], [0, 0, 0], [0.2, 0, 0]], atol=1e-7).all()
    assert np.isclose(features["vel_hist"].cpu(),
                      [[0, 0, 0, 0, 0, 0], [0.2, 0, 0, 0.2, 0, 0], [0, 0, 0, 0.1, 0, 0]], atol=1e-7).all()
    disp = np.array([[0, 0, 0], [0, 0, 0], [-0.2, 0.1, 0], [0, 0, 0], [0.2, -0.1, 0], [0, 0, 0]]) / 0.3
    assert np.isclose(features["rel_disp"].cpu(), disp, atol=1e-6).all()
    assert np.isclose(features["rel_dist"].cpu(), ((disp**2).sum(-1, keepdims=True)) ** 0.5, atol=1e-6).all()
    _, _, _, nbrs_new = case.preprocess(0, (POSITION, PTYPE), 0.0, nbrs, 0)
    assert torch.equal(nbrs.idx, nbrs_new.idx)
    _, _, target1, _ = case.preprocess(0, (POSITION, PTYPE), 0.0, nbrs, 1)
    assert np.isclose(target1["acc"].cpu(), [[0, 0, 0], [0, 0, 0], [0.1, 0, 0]], atol=1e-7).all()
    new_pos = case.integrate({"acc": np.array([[0.0, 0, 0], [0, 0, 0], [0.2, 0, 0]], np.float32)}, POSITION[:, :3])
    assert np.isclose(new_pos.cpu(), POSITION[:, 3]).all()


CASES = [("tgv2d", "float32"), ("tgv2d", "float64"), ("rpf2d", "float32"), ("dam2d", "float64"),
         ("ldc3d", "float32"), ("ldc3d", "float64"), ("rpf3d_8k", "float32"), ("ldc3d_28k", "float32")]


@pytest.mark.parametrize("name,dtype", CASES)
def test_edge_list_bit_exact(name, dtype):
    c, ours, orac = build_pair(name, dtype)
    sample = (c["positions"], c["particle_type"])
    f_gpu, n_gpu = ours.allocate_eval(sample)
    f_cpu, n_cpu = orac.allocate_eval(sample)
    assert n_gpu.max_occupancy == n_cpu.max_occupancy
    assert n_gpu.cell_list_capacity == n_cpu.cell_list_capacity
    assert n_gpu.n_edges == n_cpu.n_edges
    assert np.array_equal(n_gpu.idx.cpu().numpy(), n_cpu.idx), "edge list differs (order is part of the contract)"
    assert bool(n_gpu.did_buffer_overflow) == n_cpu.did_buffer_overflow is False
    # float32 features are exact casts of the same arithmetic
    for k in ("vel_hist", "bound", "force", "rel_disp", "rel_dist"):
        assert (k in f_gpu) == (k in f_cpu)
        if k in f_cpu:
            ref = np.asarray(f_cpu[k]).astype(np.float32)
            got = f_gpu[k].cpu().numpy()
            assert got.shape == ref.shape
            assert np.array_equal(got, ref), f"{k}: max diff {np.abs(got - ref).max()}"


def test_update_and_overflow_flag():
    """Fixed, too-small edge capacity + moved particles: same truncated list, same flag.

    List contents are only specified while no CELL overflows (jax-md then drops particles
    from its buffer through colliding scatter writes), so the edge capacity alone is shrunk
    (to 90 % of the true count) on both sides before the update."""
    c, ours, orac = build_pair("tgv2d", "float32", n_future=1)
    pt = c["particle_type"]
    _, n_gpu = ours.allocate_eval((c["positions"][:, :6], pt))
    _, n_cpu = orac.allocate_eval((c["positions"][:, :6], pt))
    small = int(0.9 * n_cpu.n_edges)
    n_cpu.max_occupancy, n_cpu.idx = small, n_cpu.idx[:, :small].copy()
    n_gpu.max_occupancy, n_gpu.idx = small, n_gpu.idx[:, :small].contiguous()
    sample = (c["positions"][:, 1:7], pt)
    _, u_gpu = ours.preprocess_eval(sample, n_gpu)
    _, u_cpu = orac.preprocess_eval(sample, n_cpu)
    assert u_cpu.did_buffer_overflow and not u_cpu.cell_overflow, "test setup"
    assert bool(u_gpu.did_buffer_overflow)
    assert u_gpu.n_edges == u_cpu.n_edges > small
    assert np.array_equal(u_gpu.idx.cpu().numpy(), u_cpu.idx)
    # sticky like jax-md's error code: a later in-capacity update keeps the flag
    big = ours._lb200["neighbor_fn"].allocate(torch.as_tensor(c["positions"][:, 5]).cuda())
    u_gpu.idx, u_gpu.max_occupancy = big.idx, big.max_occupancy
    _, again = ours.preprocess_eval(sample, u_gpu)
    assert bool(again.did_buffer_overflow) and again.n_edges <= again.max_occupancy


def test_all_pairs_path_lj_fixture():
    """Box smaller than three cutoffs: jax-md's all-pairs candidate path (LJ fixture)."""
    with open(os.path.join(GOLDEN, "lj3d_metadata.json")) as f:
        metadata = json.load(f)
    pos = np.load(os.path.join(GOLDEN, "lj3d_valid_position.npy"))[:3].transpose(1, 0, 2)
    box = np.array([5.0, 5.0, 5.0])
    ours = case_builder(box, metadata, 3, noise_std=0.0, dtype="float64")
    orac = ocase.case_builder(box, metadata, 3, noise_std=0.0, dtype=np.float64)
    ptype = np.zeros(3, np.int32)
    _, n_gpu = ours.allocate_eval((pos, ptype))
    _, n_cpu = orac.allocate_eval((pos, ptype))
    assert n_gpu.cell_list_capacity is None and n_cpu.cell_list_capacity is None
    assert np.array_equal(n_gpu.idx.cpu().numpy(), n_cpu.idx)


def test_magnitude_features_and_isotropic_norm():
    c, _, _ = build_pair("rpf2d", "float64")
    kw = dict(cfg_model={"magnitude_features": True, "isotropic_norm": True}, noise_std=3e-4)
    from helpers import oracle_force
    ours = case_builder(c["box"], c["metadata"], 6, external_force_fn=c["force"], dtype="float64", **kw)
    orac = ocase.case_builder(c["box"], c["metadata"], 6, external_force_fn=oracle_force(c["force"]),
                              dtype=np.float64, **kw)
    sample = (c["positions"], c["particle_type"])
    f_gpu, _ = ours.allocate_eval(sample)
    f_cpu, _ = orac.allocate_eval(sample)
    assert np.array_equal(f_gpu["vel_mag"].cpu().numpy(), f_cpu["vel_mag"].astype(np.float32))
    assert np.array_equal(f_gpu["vel_hist"].cpu().numpy(), f_cpu["vel_hist"].astype(np.float32))


def test_integrate_matches_oracle():
    c, ours, orac = build_pair("rpf2d", "float64")
    rng = np.random.default_rng(1)
    acc = rng.standard_normal((c["positions"].shape[0], 2)).astype(np.float32)
    got = ours.integrate({"acc": acc}, c["positions"]).cpu().numpy()
    ref = orac.integrate({"acc": acc}, c["positions"])
    assert np.array_equal(got, ref)
    got_v = ours.integrate({"vel": acc}, c["positions"]).cpu().numpy()
    assert np.array_equal(got_v, orac.integrate({"vel": acc}, c["positions"]))


# ---- the receiver-major view built straight from the cells (lb200_nbr_csr_build) vs list -> csr_build
def _csr_both_ways(ours, window, nbrs, n_receivers=0):
    """-> (rowptr, snd, rcv, edge_feat[perm]) of the list path and of the direct build, as numpy."""
    import ctypes as C

    from lagrangebench_b200 import _cabi

    lib = _cabi.load()
    n, e_cap = window.shape[0], nbrs.max_occupancy
    dev = window.device
    i32 = dict(dtype=torch.int32, device=dev)
    st = _cabi.stream()
    rowptr, perm, snd, rcv = torch.empty(n + 1, **i32), torch.empty(e_cap, **i32), torch.empty(e_cap, **i32), \
        torch.empty(e_cap, **i32)
    scratch = torch.empty(lib.lb200_csr_scratch_bytes(n, e_cap), dtype=torch.uint8, device=dev)
    _cabi.check(lib.lb200_csr_build(_cabi.ptr(nbrs.idx), n, e_cap, _cabi.ptr(rowptr), _cabi.ptr(perm), _cabi.ptr(snd),
                                    _cabi.ptr(rcv), _cabi.ptr(scratch), scratch.numel(), st))
    fc = ours._lb200["feature_cfg"](n)
    ef_list = torch.empty((e_cap, 4), dtype=torch.float32, device=dev)
    _cabi.check(lib.lb200_features(C.byref(fc), _cabi.ptr(window), None, _cabi.ptr(nbrs.idx), e_cap, None,
                                   _cabi.ptr(ef_list), st))
    rowptr2, tmp, snd2, rcv2 = torch.empty(n + 1, **i32), torch.empty(e_cap, **i32), torch.full((e_cap,), -7, **i32), \
        torch.full((e_cap,), -7, **i32)
    ef2 = torch.zeros((e_cap, 4), dtype=torch.float32, device=dev)
    stats = torch.zeros(4, **i32)
    tw, dim = window.shape[1], window.shape[2]
    last = window.data_ptr() + (tw - 1) * dim * window.element_size()
    _cabi.check(lib.lb200_nbr_csr_build(C.byref(nbrs._grid), C.c_void_p(last), tw * dim, nbrs.cell_list_capacity,
                                        n_receivers, _cabi.ptr(rowptr2), _cabi.ptr(snd2), _cabi.ptr(rcv2),
                                        _cabi.ptr(ef2), _cabi.ptr(tmp), e_cap, _cabi.ptr(stats),
                                        _cabi.ptr(nbrs._scratch), nbrs._scratch.numel(), st))
    e = int(rowptr[n])
    a = (rowptr.cpu().numpy(), snd[:e].cpu().numpy(), rcv[:e].cpu().numpy(), ef_list[perm[:e].long()].cpu().numpy())
    e2 = int(rowptr2[n])
    b = (rowptr2.cpu().numpy(), snd2[:e2].cpu().numpy(), rcv2[:e2].cpu().numpy(), ef2[:e2].cpu().numpy())
    return a, b, stats.cpu().numpy()


@pytest.mark.parametrize("name,dtype", [("tgv2d", "float32"), ("rpf2d", "float64"), ("dam2d", "float64"),
                                        ("ldc3d", "float32"), ("rpf3d_8k", "float64"), ("ldc3d_28k", "float64")])
def test_direct_csr_is_the_list_path_bit_for_bit(name, dtype):
    c, ours, _ = build_pair(name, dtype)
    window = torch.as_tensor(c["positions"][:, :6]).cuda().contiguous()
    nbrs = ours._lb200["neighbor_fn"].allocate(window[:, -1].contiguous())
    a, b, stats = _csr_both_ways(ours, window, nbrs)
    assert stats[0] == nbrs.n_edges and stats[2] == 0
    for x, y, what in zip(a, b, ("rowptr", "snd", "rcv", "edge features")):
        assert np.array_equal(x, y), what
    # domain decomposition: only the first n_receivers particles receive
    n_recv = window.shape[0] // 3
    a, b, _ = _csr_both_ways(ours, window, nbrs, n_receivers=n_recv)
    e = a[0][n_recv]
    assert b[0][-1] == e and np.array_equal(b[0][:n_recv + 1], a[0][:n_recv + 1])
    for x, y in zip(a[1:], b[1:]):
        assert np.array_equal(x[:e], y)


@pytest.mark.parametrize("name,dims,dtype", [("tgv2d", (5, 5), "float32"), ("tgv2d", (7, 5), "float64"),
                                             ("rpf3d_8k", (5, 5, 5), "float64"), ("ldc3d", (6, 5, 7), "float32"),
                                             ("tgv2d", (4, 4), "float64")])
def test_small_random_clouds_with_coincident_and_border_particles(name, dims, dtype):
    """Smallest cell grids (3 cells per side; (4, 4) falls back to all pairs), uniformly random positions instead
    of a lattice, particles ON the lower box border, in the last ulp below the upper one, and coincident pairs
    (distance 0): the list equals the oracle's bit for bit and the direct view equals the list path."""
    c, ours, orac = build_pair(name, dtype, dims=dims, multiplier=2.0)
    rng = np.random.default_rng(sum(dims))
    n, t, d = c["positions"].shape
    box = c["box"].astype(c["positions"].dtype)
    pos = (rng.random((n, d)) * c["box"]).astype(c["positions"].dtype)
    pos[0] = 0
    pos[1] = np.nextafter(box, np.zeros_like(box))
    pos[2, 0] = 0
    pos[3] = pos[4]                                  # coincident pair
    pos[5] = pos[4] + np.asarray(1e-7, pos.dtype)    # and an almost coincident one
    window = np.repeat(pos[:, None], 6, axis=1)
    _, n_gpu = ours.allocate_eval((window, c["particle_type"]))
    _, n_cpu = orac.allocate_eval((window, c["particle_type"]))
    assert n_gpu.n_edges == n_cpu.n_edges and n_gpu.max_occupancy == n_cpu.max_occupancy
    assert np.array_equal(n_gpu.idx.cpu().numpy(), n_cpu.idx)
    assert not bool(n_gpu.did_buffer_overflow)
    if n_gpu._grid.use_cells:
        w = torch.as_tensor(window).cuda().contiguous()
        a, b, stats = _csr_both_ways(ours, w, n_gpu)
        assert stats[0] == n_cpu.n_edges and stats[2] == 0
        for x, y, what in zip(a, b, ("rowptr", "snd", "rcv", "edge features")):
            assert np.array_equal(x, y), what


def test_direct_csr_dense_bucket_and_capacity_clamp():
    """In-degrees beyond the shared-memory fast path (a pile of particles), and a capacity smaller
    than the edge count: the view is the truncated bucket structure, the flag is raised."""
    c, ours, _ = build_pair("tgv2d", "float64")
    pos = c["positions"][:, :6].copy()
    rng = np.random.default_rng(5)
    pile = rng.choice(pos.shape[0], 150, replace=False)
    pos[pile] = pos[pile[0]] + 1e-4 * rng.standard_normal((150, 6, 2))
    window = torch.as_tensor(pos).cuda().contiguous()
    nbrs = ours._lb200["neighbor_fn"].allocate(window[:, -1].contiguous())
    a, b, stats = _csr_both_ways(ours, window, nbrs)
    assert np.diff(a[0]).max() >= 150 and stats[2] == 0
    for x, y, what in zip(a, b, ("rowptr", "snd", "rcv", "edge features")):
        assert np.array_equal(x, y), what
    # shrink the capacity below the edge count
    small = int(0.8 * nbrs.n_edges)
    nbrs.max_occupancy, nbrs.idx = small, nbrs.idx[:, :small].contiguous()
    _, b2, stats2 = _csr_both_ways(ours, window, nbrs)
    assert stats2[2] & 1 and stats2[0] == stats[0]
    assert b2[0][-1] == small and np.array_equal(b2[0], np.minimum(b[0], small))
    whole = b[0][np.searchsorted(b[0], small, side="right") - 1]  # end of the last bucket that fits entirely
    assert whole > 0.7 * small
    assert np.array_equal(b2[1][:whole], b[1][:whole]) and np.array_equal(b2[2][:whole], b[2][:whole])


@pytest.mark.parametrize("name,dtype", [("tgv2d", "float32"), ("ldc3d", "float64")])
def test_padding_particles_stay_out_of_the_search(name, dtype):
    """``num_particles`` (case.py:182-190): rows of type PAD_VALUE appended by the loader
    (data.py:183-197, all at the origin) are neither senders nor receivers."""
    c, ours, orac = build_pair(name, dtype)
    n_pad = 321
    pos = np.concatenate([c["positions"][:, :6], np.zeros((n_pad,) + c["positions"][:, :6].shape[1:],
                                                            c["positions"].dtype)])
    ptype = np.concatenate([c["particle_type"], np.full(n_pad, -1, c["particle_type"].dtype)])
    n_real = c["positions"].shape[0]
    f_gpu, n_gpu = ours.allocate_eval((pos, ptype))
    f_cpu, n_cpu = orac.allocate_eval((pos, ptype))
    _, n_plain = orac.allocate_eval((c["positions"][:, :6], c["particle_type"]))
    assert n_cpu.n_edges == n_plain.n_edges
    idx = n_gpu.idx.cpu().numpy()
    assert np.array_equal(idx, n_cpu.idx)
    real = idx[0] < pos.shape[0]
    assert real.sum() == n_plain.n_edges and idx[:, real].max() < n_real
    assert np.array_equal(idx[:, real], n_plain.idx[:, :n_plain.n_edges])
    _, u_gpu = ours.preprocess_eval((pos, ptype), n_gpu)
    assert np.array_equal(u_gpu.idx.cpu().numpy(), n_cpu.idx)
    # the direct view leaves the padding rows without buckets
    window = torch.as_tensor(pos).cuda().to(f_gpu["abs_pos"].dtype).contiguous()
    a, b, _ = _csr_both_ways(ours, window, n_gpu)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    assert b[0][n_real] == b[0][-1]


def test_matscipy_padded_h5dataset_sample(tmp_path):
    """``H5Dataset(..., nl_backend="matscipy")`` pads every sample to ``num_particles_max`` with particles of
    type -1 at the origin (``data.py:183-197,222-223``); ``preprocess`` passes ``num_particles`` on
    (``case.py:182-190``).  The padded sample's edge list is the oracle's, and a rollout through ``infer``
    leaves the padding rows where the trajectory has them."""
    import json

    from h5write import write_h5
    from lagrangebench_b200 import GNS, infer, synthetic
    from lagrangebench_b200.data import H5Dataset

    c = synthetic.make_case("tgv2d", 6, 3, seed=3, dtype=np.float32)
    n = c["positions"].shape[0]
    root = tmp_path / "2D_TGV_2500_10kevery100"
    root.mkdir()
    write_h5(str(root / "test.h5"), {"00000": {"position": np.ascontiguousarray(c["positions"].transpose(1, 0, 2)),
                                               "particle_type": c["particle_type"].astype(np.int32)}}, chunk_rows=4)
    meta = dict(c["metadata"], num_particles_max=n + 137)
    (root / "metadata.json").write_text(json.dumps(meta))
    ds = H5Dataset("test", str(root), input_seq_length=6, extra_seq_length=3, nl_backend="matscipy")
    pos, ptype = ds[0]
    assert pos.shape[0] == n + 137 and (ptype[n:] == -1).all() and (pos[n:] == 0).all()
    kw = dict(cfg_neighbors={"multiplier": 1.25, "backend": "matscipy"}, noise_std=0.0)
    ours = case_builder(c["box"], meta, 6, dtype="float32", **kw)
    orac = ocase.case_builder(c["box"], meta, 6, dtype=np.float32, **kw)
    _, n_gpu = ours.allocate_eval((pos[:, :6], ptype))
    _, n_cpu = orac.allocate_eval((pos[:, :6], ptype))
    idx = n_gpu.idx.cpu().numpy()
    assert np.array_equal(idx, n_cpu.idx)
    assert idx[idx < n + 137].max() < n  # no padding row is a sender or a receiver
    model = GNS(2, 128, 2, 2, 16)
    feats, _ = ours.allocate_eval((pos[:, :6], ptype))
    params, state = model.init(0, (feats, ptype))
    out = infer(model, ours, ds, params=params, state=state, cfg_eval_infer={"batch_size": 1, "metrics": ["mse"],
                "out_type": "none", "n_trajs": 1}, n_rollout_steps=3)
    assert np.isfinite(np.asarray(torch.as_tensor(out["rollout_0"]["mse"]).cpu())).all()
