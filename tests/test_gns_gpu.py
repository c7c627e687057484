"""Parity of the GNS forward (kernels ii-iv) with the oracle.

Tolerance: BASELINE.json asks for accelerations within 1e-5 relative in float32.  It is
checked as max|a - ref| <= 1e-5 * max|ref| against the float64 oracle, and the float32
oracle's own distance to float64 is printed next to it for scale."""

import numpy as np
import pytest
import torch

from helpers import build_pair, elem_rel_err, rel_err
from lagrangebench_b200 import GNS
from oracle import gns as ogns

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _forward_both(name, dtype, num_mp_steps=10, seed=0):
    c, ours, orac = build_pair(name, dtype, seed=seed)
    sample = (c["positions"], c["particle_type"])
    f_gpu, _ = ours.allocate_eval(sample)
    f_cpu, _ = orac.allocate_eval(sample)
    d = c["metadata"]["dim"]
    node_in = sum(f_cpu[k].reshape(f_cpu[k].shape[0], -1).shape[1] for k in ("vel_hist", "bound", "force") if k in f_cpu)
    params = ogns.init_params(node_in, d + 1, d, num_mp_steps=num_mp_steps, seed=seed + 11, perturb=True)
    model = GNS(d, 128, 2, num_mp_steps, 16)
    out, _ = model.apply(params, {}, (f_gpu, torch.as_tensor(c["particle_type"]).cuda()))
    # the oracle consumes the float32 features the network sees (jmp casts inputs to f32)
    f32 = {k: (np.asarray(v).astype(np.float32) if np.asarray(v).dtype.kind == "f" else v) for k, v in f_cpu.items()}
    ref64 = ogns.forward(params, f32, c["particle_type"], num_mp_steps, np.float64)["acc"]
    ref32 = ogns.forward(params, f32, c["particle_type"], num_mp_steps, np.float32)["acc"]
    return out["acc"].cpu().numpy(), ref64, ref32, (c, ours, f_gpu, params, model)


@pytest.mark.parametrize("name,dtype", [("tgv2d", "float32"), ("rpf2d", "float64"), ("dam2d", "float32"),
                                        ("ldc3d", "float64"), ("rpf3d_8k", "float32")])
def test_forward_parity(name, dtype):
    got, ref64, ref32, _ = _forward_both(name, dtype)
    assert got.shape == ref64.shape and np.isfinite(got).all()
    e_gpu, e_cpu32 = rel_err(got, ref64), rel_err(ref32, ref64)
    el_max, el_p999 = elem_rel_err(got, ref64)
    print(f"{name}/{dtype}: rel err GPU vs f64 oracle {e_gpu:.2e}; f32 oracle vs f64 oracle {e_cpu32:.2e}; "
          f"elementwise (|ref| > 1% of max): max {el_max:.2e}, p99.9 {el_p999:.2e}")
    assert e_gpu <= TOL
    assert rel_err(got, ref32) <= TOL
    assert el_max <= 1e-3 and el_p999 <= 1e-4  # elementwise, where the reference carries signal


def test_forward_parity_at_the_benchmarked_configuration():
    """LDC-3D 28 000 particles / 405 k edges with float64 positions -- the cloud bench.py times -- against
    the float64 ground truth (torch twin of the oracle: the NumPy one takes minutes at this size)."""
    from oracle import gns_torch

    c, ours, orac = build_pair("ldc3d_28k", "float64")
    sample = (c["positions"], c["particle_type"])
    f_gpu, _ = ours.allocate_eval(sample)
    f_cpu, _ = orac.allocate_eval(sample)
    params = ogns.init_params(21, 4, 3, num_mp_steps=10, seed=5, perturb=True)
    model = GNS(3, 128, 2, 10, 16)
    out, _ = model.apply(params, {}, (f_gpu, torch.as_tensor(c["particle_type"]).cuda()))
    got = out["acc"].cpu().numpy()
    f32 = {k: (np.asarray(v).astype(np.float32) if np.asarray(v).dtype.kind == "f" else v) for k, v in f_cpu.items()}
    ref64 = gns_torch.forward(gns_torch.pack(params, np.float64), f32, c["particle_type"], 10)["acc"]
    e_gpu = rel_err(got, ref64)
    el_max, el_p999 = elem_rel_err(got, ref64)
    print(f"ldc3d_28k/float64: rel err {e_gpu:.2e}; elementwise max {el_max:.2e}, p99.9 {el_p999:.2e}")
    assert got.shape == (28000, 3) and e_gpu <= TOL and el_max <= 1e-3 and el_p999 <= 1e-4


def test_fp16_split_range_is_guarded():
    """Activations beyond the fp16 range (|x| > 65504) cannot be split into hi/lo halves.
    * What the WEIGHTS bound (latents, the edge MLPs' hidden layers) is checked when they are packed: such
      a model runs on the float32 CUDA-core kernels, with a warning, and still matches the oracle.
    * What depends on the DATA (the aggregate of a crowded receiver) is guarded on the device: the
      tensor-core path raises instead of returning numbers, the float32 kernels deliver the result."""
    from lagrangebench_b200 import case_builder, models, synthetic

    got, ref64, _, (c, ours, f_gpu, params, model) = _forward_both("tgv2d", "float32", num_mp_steps=3)
    key = "gns/~_processor/layer_norm"  # LayerNorm of the first edge update
    ptype = torch.as_tensor(c["particle_type"]).cuda()
    f_cpu = {k: (v.cpu().numpy() if isinstance(v, torch.Tensor) else v) for k, v in f_gpu.items()}
    # (1) edge latents of order 1e6: caught on the host
    big = {k: dict(v) for k, v in params.items()}
    big[key] = {"scale": params[key]["scale"] * np.float32(2.0e5), "offset": params[key]["offset"]}
    assert not models.pack_params(big, 3, 2).fp16_safe
    with pytest.warns(UserWarning, match="fp16 split"):
        out, _ = GNS(2, 128, 2, 3, 16).apply(big, {}, (f_gpu, ptype))
    ref = ogns.forward(big, f_cpu, c["particle_type"], 3, np.float64)["acc"]
    assert rel_err(out["acc"].cpu().numpy(), ref) <= 1e-4
    # (2) 150 particles in one spot, every message about +1000: the aggregate passes 65504
    pos = c["positions"][:, :6].copy()
    rng = np.random.default_rng(5)
    pile = rng.choice(pos.shape[0], 150, replace=False)
    pos[pile] = pos[pile[0]] + 1e-4 * rng.standard_normal((150, 6, 2)).astype(pos.dtype)
    f_pile, _ = ours.allocate_eval((pos, c["particle_type"]))
    shifted = {k: dict(v) for k, v in params.items()}
    shifted[key] = {"scale": params[key]["scale"], "offset": params[key]["offset"] + np.float32(1000.0)}
    assert models.pack_params(shifted, 3, 2).fp16_safe
    with pytest.raises(FloatingPointError):
        GNS(2, 128, 2, 3, 16).apply(shifted, {}, (f_pile, ptype))
    simt = GNS(2, 128, 2, 3, 16)
    simt.edge_impl = "simt"
    out, _ = simt.apply(shifted, {}, (f_pile, ptype))
    fp_cpu = {k: (v.cpu().numpy() if isinstance(v, torch.Tensor) else v) for k, v in f_pile.items()}
    ref = ogns.forward(shifted, fp_cpu, c["particle_type"], 3, np.float64)["acc"]
    assert np.isfinite(out["acc"].cpu().numpy()).all() and rel_err(out["acc"].cpu().numpy(), ref) <= 1e-4
    # (3) small activations: fp16 subnormals keep the split accurate orders of magnitude below 1
    small = {k: dict(v) for k, v in params.items()}
    small[key] = {"scale": params[key]["scale"] * np.float32(1.0e-3), "offset": params[key]["offset"] * np.float32(1.0e-3)}
    out_s, _ = model.apply(small, {}, (f_gpu, ptype))
    ref_s = ogns.forward(small, f_cpu, c["particle_type"], 3, np.float64)["acc"]
    assert rel_err(out_s["acc"].cpu().numpy(), ref_s) <= TOL


@pytest.mark.parametrize("latent,mp", [(64, 5), (96, 2)])
def test_narrower_latent_sizes_run_zero_padded(latent, mp):
    """``GNS(latent_size=64, num_mp_steps=5)`` is the reference's published GNS-5-64 (``baselines.rst:55``): it runs
    on the 128-wide kernels, every latent dimension zero-padded and LayerNorm dividing by the true width, on the
    tensor-core path and on the float32 CUDA-core path."""
    c, ours, orac = build_pair("rpf2d", "float32")
    sample = (c["positions"], c["particle_type"])
    f_gpu, _ = ours.allocate_eval(sample)
    f_cpu, _ = orac.allocate_eval(sample)
    params = ogns.init_params(12, 3, 2, latent=latent, num_mp_steps=mp, seed=4, perturb=True)
    ref = ogns.forward(params, f_cpu, c["particle_type"], mp, np.float64)["acc"]
    for impl in ("tc", "simt"):
        model = GNS(2, latent, 2, mp, 16)
        model.edge_impl = impl
        out, _ = model.apply(params, {}, (f_gpu, c["particle_type"]))
        err = rel_err(out["acc"].cpu().numpy(), ref)
        print(f"latent {latent}, {mp} MP steps, {impl}: rel err {err:.2e}")
        assert err <= TOL
    with pytest.raises(NotImplementedError):
        GNS(2, 256, 2, mp, 16)


def test_forward_single_mp_step_and_type_embedding():
    """One message-passing step; particle types other than FLUID exercise the embedding."""
    got, ref64, _, _ = _forward_both("ldc3d", "float32", num_mp_steps=1)
    assert rel_err(got, ref64) <= TOL


def test_apply_accepts_plain_feature_dict_in_any_edge_order():
    """model.apply on a plain dict (no packed buffers) with a shuffled, padded edge list:
    the receiver-major view must make the result independent of list order."""
    got, ref64, _, (c, ours, f_gpu, params, model) = _forward_both("tgv2d", "float32")
    rng = np.random.default_rng(3)
    e_cap = f_gpu["senders"].shape[0]
    perm = torch.as_tensor(rng.permutation(e_cap)).cuda()
    plain = {k: f_gpu[k].clone() for k in ("vel_hist", "rel_disp", "rel_dist", "senders", "receivers")}
    for k in ("rel_disp", "rel_dist", "senders", "receivers"):
        plain[k] = plain[k][perm].contiguous()
    out, _ = model.apply(params, {}, (plain, c["particle_type"]))
    shuffled = out["acc"].cpu().numpy()
    assert rel_err(shuffled, ref64) <= TOL
    # same receiver buckets in a different within-bucket order: float32 sums may differ in the last bits only
    assert rel_err(shuffled, got) <= 5e-6


def test_high_degree_receiver_straddles_many_tiles():
    """A receiver with hundreds of in-edges spans several message tiles (carry path)."""
    rng = np.random.default_rng(0)
    n, d, deg = 500, 2, 700
    snd = np.concatenate([rng.integers(0, n, deg), np.arange(n), rng.integers(0, n, 3000)]).astype(np.int32)
    rcv = np.concatenate([np.full(deg, 7), np.arange(n), rng.integers(0, n, 3000)]).astype(np.int32)
    pad = np.full(100, n, np.int32)
    snd, rcv = np.concatenate([snd, pad]), np.concatenate([rcv, pad])
    e = snd.shape[0]
    feats = {"vel_hist": rng.standard_normal((n, 10)).astype(np.float32),
             "rel_disp": rng.standard_normal((e, d)).astype(np.float32),
             "rel_dist": rng.random((e, 1)).astype(np.float32), "senders": snd, "receivers": rcv}
    ptype = rng.integers(0, 4, n).astype(np.int32)
    params = ogns.init_params(10, d + 1, d, num_mp_steps=2, seed=4)
    model = GNS(d, 128, 2, 2, 16)
    out, _ = model.apply(params, {}, (feats, ptype))
    ref = ogns.forward(params, feats, ptype, 2, np.float64)["acc"]
    assert rel_err(out["acc"].cpu().numpy(), ref) <= TOL


@pytest.mark.parametrize("n,e_real,n_pad", [(1, 1, 0), (2, 3, 5), (31, 31, 1), (33, 64, 0), (40, 65, 7), (200, 0, 4),
                                           (593, 2368, 0), (600, 18945, 3)])
def test_tiny_and_ragged_graphs(n, e_real, n_pad):
    """Edge cases of the tiling: one particle, edge counts at and around the 32-edge tile, fewer tiles than workers,
    one tile more than a whole round of the 148 x 4 workers, nodes without in-edges, an EMPTY edge list
    (jraph's segment_sum of nothing = 0), padding edges -- both kernel paths against the oracle."""
    rng = np.random.default_rng(n * 7 + e_real)
    d = 3
    snd = rng.integers(0, n, e_real).astype(np.int32)
    rcv = rng.integers(0, max(1, n - n // 4), e_real).astype(np.int32)  # the last quarter of the nodes receives nothing
    pad = np.full(n_pad, n, np.int32)
    snd, rcv = np.concatenate([snd, pad]), np.concatenate([rcv, pad])
    e = snd.shape[0]
    feats = {"vel_hist": rng.standard_normal((n, 15)).astype(np.float32),
             "rel_disp": rng.standard_normal((e, d)).astype(np.float32),
             "rel_dist": rng.random((e, 1)).astype(np.float32), "senders": snd, "receivers": rcv}
    ptype = rng.integers(0, 3, n).astype(np.int32)
    params = ogns.init_params(15, d + 1, d, num_mp_steps=3, seed=2, perturb=True)
    ref = ogns.forward(params, feats, ptype, 3, np.float64)["acc"]
    for impl in ("tc", "simt"):
        model = GNS(d, 128, 2, 3, 16)
        model.edge_impl = impl
        out, _ = model.apply(params, {}, (feats, ptype))
        got = out["acc"].cpu().numpy()
        assert got.shape == (n, d) and np.isfinite(got).all()
        assert rel_err(got, ref) <= TOL, (impl, rel_err(got, ref))


def test_forward_is_deterministic():
    got1, _, _, (c, ours, f_gpu, params, model) = _forward_both("rpf2d", "float32")
    out2, _ = model.apply(params, {}, (f_gpu, c["particle_type"]))
    assert np.array_equal(got1, out2["acc"].cpu().numpy()), "aggregation must be bitwise reproducible"


def test_tensor_core_kernel_matches_cuda_core_kernel():
    """The tcgen05 message kernel (fp16 hi/lo split, fp32 accumulate) against the fp32
    CUDA-core kernel on the same inputs: both sit at float32 rounding level."""
    got_tc, ref64, _, (c, ours, f_gpu, params, model) = _forward_both("rpf3d_8k", "float32")
    assert model.edge_impl == "tc"
    model.edge_impl = "simt"
    out, _ = model.apply(params, {}, (f_gpu, c["particle_type"]))
    got_simt = out["acc"].cpu().numpy()
    assert rel_err(got_tc, got_simt) <= 5e-6
    assert rel_err(got_tc, ref64) <= TOL and rel_err(got_simt, ref64) <= TOL


def test_tcgen05_selftest_tmem_operand_and_rescaled_accumulate():
    """The hardware features (also: MN-major B operand, fp16 subnormal inputs) the pipelined message kernel relies on: the A operand read
    from tensor memory must equal the same operand read from shared memory, and
    scale-input-d = 11 must compute A B + D * 2^-11 exactly."""
    from lagrangebench_b200 import _cabi

    lib = _cabi.load()
    out = torch.full((5,), -1.0, device="cuda")
    _cabi.check(lib.lb200_tc_selftest(_cabi.ptr(out), _cabi.stream()))
    torch.cuda.synchronize()
    ts_vs_ss, scaled, mag, mn_vs_k, subnormal = out.cpu().tolist()
    print(f"tc selftest: |D_ts - D_ss| {ts_vs_ss:.3e}  |D_scaled - expected| {scaled:.3e}  max|D| {mag:.3f}  "
          f"|D_mn - D_ss| {mn_vs_k:.3e}  |D_sub 2^18 - D_ss| {subnormal:.3e}")
    assert mag > 1.0
    assert ts_vs_ss == 0.0
    assert scaled <= 1e-6 * mag
    assert mn_vs_k == 0.0      # B operand in the MN-major (edge-contiguous) layout
    assert subnormal == 0.0    # fp16 subnormal inputs are honoured


def test_pipelined_message_kernel_matches_first_tensor_core_kernel():
    """v2 (weights in TMEM, one rescaled accumulator, two-tile pipeline) against v1 (weights in
    shared memory, two accumulators): same split-precision scheme, same summation order of
    the segmented sum -- they may differ in the last float32 bits of the GEMMs only."""
    from lagrangebench_b200 import _cabi

    if _cabi.load().lb200_version() % 2 == 0:
        pytest.skip("product build: the first-generation kernels are compiled with LB200_BUILD_CROSSCHECK=1 only")
    got_v2, ref64, _, (c, ours, f_gpu, params, model) = _forward_both("ldc3d", "float64")
    model.edge_impl = "tc1"
    out, _ = model.apply(params, {}, (f_gpu, c["particle_type"]))
    got_v1 = out["acc"].cpu().numpy()
    print(f"v2 vs v1 {rel_err(got_v2, got_v1):.2e}; v2 vs f64 {rel_err(got_v2, ref64):.2e}; v1 vs f64 {rel_err(got_v1, ref64):.2e}")
    assert rel_err(got_v2, got_v1) <= 5e-6
    assert rel_err(got_v2, ref64) <= TOL
