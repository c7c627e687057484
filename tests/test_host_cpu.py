"""CPU-only checks: the C-ABI library loads and exports every declared symbol, and the host
logic (grid geometry, weight packing, checkpoints, synthetic clouds) behaves."""

import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from lagrangebench_b200 import _cabi, models, synthetic, utils
from lagrangebench_b200.defaults import merged
from oracle import gns as ogns
from oracle import partition as opartition

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "lb200.h")).read()
    declared = set(re.findall(r"\b(lb200_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    lib = _cabi.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/lb200.h but not exported"
    assert declared == set(_cabi.EXPORTED)
    assert lib.lb200_version() >= 100
    assert b"invalid" in lib.lb200_error_string(-1)


@pytest.mark.parametrize("box,r", [([1.0, 1.0], 0.029), ([1.0, 2.0], 0.036), ([1.2, 1.0, 0.85], 0.072),
                                   ([5.48, 2.12], 0.029), ([1.0, 1.0, 1.0], 0.3), ([5.0, 5.0, 5.0], 3.0)])
def test_grid_init_matches_oracle(box, r):
    lib = _cabi.load()
    g = _cabi.Grid()
    b = (C.c_double * 3)(*(box + [1.0] * (3 - len(box))))
    _cabi.check(lib.lb200_grid_init(C.byref(g), 100, len(box), 0, 1, b, r))
    assert bool(g.use_cells) == opartition.use_cell_list(np.array(box), r)
    if g.use_cells:
        cell_size, cps = opartition.cell_dimensions(np.array(box), r)
        assert list(g.cells_per_side)[: len(box)] == cps.tolist()
        assert np.array_equal(np.array(list(g.cell_size)[: len(box)], dtype=np.float32), cell_size)
        assert g.n_cells == int(np.prod(cps))


def test_grid_init_rejects_bad_arguments():
    lib = _cabi.load()
    g = _cabi.Grid()
    b = (C.c_double * 3)(1.0, 1.0, 1.0)
    assert lib.lb200_grid_init(C.byref(g), 10, 4, 0, 1, b, 0.1) == -1
    assert lib.lb200_grid_init(C.byref(g), 10, 2, 0, 1, b, -0.1) == -1
    assert lib.lb200_grid_init(C.byref(g), 0, 2, 0, 1, b, 0.1) == -1


def test_pack_params_layout():
    params = ogns.init_params(12, 3, 2, num_mp_steps=3, seed=1)
    pk = models.pack_params(params, 3, 2, device="cpu")
    blob = pk.blob.numpy()
    w0 = params["gns/~_processor/MLP_2/~/linear_0"]["w"]  # edge MLP of step 1
    o = pk.proc_edge[1]
    assert np.array_equal(blob[o.w0:o.w0 + w0.size].reshape(w0.shape), w0)
    enc = params["gns/~_encoder/MLP/~/linear_0"]["w"]
    padded = blob[pk.enc_node.w0:pk.enc_node.w0 + 64 * 128].reshape(64, 128)
    assert np.array_equal(padded[: enc.shape[0]], enc) and not padded[enc.shape[0]:].any()
    assert pk.dec.ln_scale == -1 and pk.node_in_total == 28 and pk.embed_size == 16
    for off in (o.w0, o.b0, o.w1, o.b1, o.ln_scale, o.ln_offset, pk.embedding):
        assert off % 4 == 0  # 16-byte alignment for cp.async / float4 loads


def _tc_operands(blob, off, count):
    """``count`` 128x128 fp16 operands stored [k/8][m][k%8] (UMMA K-major core matrices) -> (count, m, k) float32."""
    halves = blob[off:off + count * 128 * 128 // 2].view(np.float16).reshape(count, 16, 128, 8)
    return halves.transpose(0, 2, 1, 3).reshape(count, 128, 128).astype(np.float32)


def test_pack_params_tensor_core_operands_reconstruct_the_weights():
    """hi + lo * 2^-11 must give back every weight to ~2^-22 relative (the split the tcgen05 kernels
    rely on), in the transposed (out feature, in feature) orientation, LayerNorm's mean folded into
    the second layer; the node encoder's first layer is zero-padded to 128 input columns."""
    params = ogns.init_params(12, 3, 2, num_mp_steps=2, seed=5, perturb=True)
    pk = models.pack_params(params, 2, 2, device="cpu")
    blob = pk.blob.numpy()

    def check(hi, lo, want):
        got = hi + lo / 2048.0
        assert np.abs(got - want).max() <= 2.0 ** -21 * max(1.0, np.abs(want).max())
        assert np.abs(lo).max() <= 1024.5  # |x - hi| <= half an fp16 ulp of x, times 2^11

    # processor edge MLP of step 1: W1e^T hi|lo, W2c^T hi|lo + b2c | scale | offset
    l0 = params["gns/~_processor/MLP_2/~/linear_0"]
    l1 = params["gns/~_processor/MLP_2/~/linear_1"]
    ln = params["gns/~_processor/layer_norm_2"]
    ops = _tc_operands(blob, pk.proc_edge[1].tc_w, 4)
    check(ops[0], ops[1], l0["w"][256:384].T)
    w2 = l1["w"].astype(np.float64)
    check(ops[2], ops[3], (w2 - w2.mean(axis=1, keepdims=True)).astype(np.float32).T)
    vec = blob[pk.proc_edge[1].tc_vec:pk.proc_edge[1].tc_vec + 384]
    assert np.allclose(vec[:128], l1["b"] - l1["b"].mean(), atol=1e-7)
    assert np.array_equal(vec[128:256], ln["scale"]) and np.array_equal(vec[256:], ln["offset"])
    # node encoder: W0pad^T, W1c^T, then the two halves of the first edge MLP's first layer
    e0 = params["gns/~_encoder/MLP/~/linear_0"]
    e1 = params["gns/~_encoder/MLP/~/linear_1"]
    first = params["gns/~_processor/MLP/~/linear_0"]
    ops = _tc_operands(blob, pk.enc_node.tc_w, 8)
    k_in = e0["w"].shape[0]
    check(ops[0][:, :k_in], ops[1][:, :k_in], e0["w"].T)
    assert not ops[0][:, k_in:].any() and not ops[1][:, k_in:].any()
    w1 = e1["w"].astype(np.float64)
    check(ops[2], ops[3], (w1 - w1.mean(axis=1, keepdims=True)).astype(np.float32).T)
    check(ops[4], ops[5], first["w"][:128].T)
    check(ops[6], ops[7], first["w"][128:256].T)
    vec = blob[pk.enc_node.tc_vec:pk.enc_node.tc_vec + 640]
    assert np.array_equal(vec[:128], e0["b"]) and np.array_equal(vec[512:640], first["b"])
    # edge encoder: only the second layer goes through the tensor cores
    ee1 = params["gns/~_encoder/MLP_1/~/linear_1"]
    ops = _tc_operands(blob, pk.enc_edge.tc_w, 2)
    we = ee1["w"].astype(np.float64)
    check(ops[0], ops[1], (we - we.mean(axis=1, keepdims=True)).astype(np.float32).T)
    assert pk.enc_edge.b0 == pk.enc_edge.w0 + 4 * 128  # W0[4][128] | b0[128] contiguous (encoder kernel)


def test_pack_params_accepts_alternate_embed_name_and_rejects_other_widths():
    params = ogns.init_params(10, 3, 2, num_mp_steps=1, seed=0)
    params["gns/embed"] = params.pop("gns/~/embed")
    assert models.pack_params(params, 1, 2, device="cpu").embed_size == 16
    assert models.GNS(2, 64, 2, 5, 16)._latent_size == 64  # narrower models run zero-padded
    with pytest.raises(NotImplementedError):
        models.GNS(2, 256, 2, 5, 16)
    with pytest.raises(NotImplementedError):
        models.GNS(2, 128, 3, 5, 16)  # num_mlp_layers != 2


def test_init_params_count_matches_published():
    assert utils.get_num_params(models.init_params(12, 2)) == 1211794  # docs/pages/baselines.rst:62


def test_checkpoint_roundtrip(tmp_path):
    params = models.init_params(10, 2, num_mp_steps=2, seed=3)
    utils.save_haiku(str(tmp_path), params, {}, metadata_ckp={"step": 7, "loss": 0.5})
    loaded, state, _, step = utils.load_haiku(str(tmp_path))
    assert step == 7 and state == {} and sorted(loaded) == sorted(params)
    for k in params:
        for leaf in params[k]:
            assert np.array_equal(loaded[k][leaf], params[k][leaf])


def test_defaults_merge_and_node_type():
    assert merged("neighbors", {"multiplier": 2.0}) == {"backend": "b200", "multiplier": 2.0}
    assert merged("eval.infer", None)["batch_size"] == 2
    m = utils.get_kinematic_mask(torch.tensor([0, 1, 2, 3, -1]))
    assert m.tolist() == [False, True, True, False, True]


def test_broadcast_helpers():
    t = {"a": torch.arange(3), "b": (torch.zeros(2, 2),)}
    b = utils.broadcast_to_batch(t, 4)
    assert b["a"].shape == (4, 3) and b["b"][0].shape == (4, 2, 2)
    assert torch.equal(utils.broadcast_from_batch(b, 2)["a"], t["a"])


@pytest.mark.parametrize("name,n", [("tgv2d", 2500), ("rpf2d", 3200), ("dam2d", 5740), ("ldc3d", 8160)])
def test_synthetic_shapes(name, n):
    c = synthetic.make_case(name, n_future=2)
    assert c["positions"].shape == (n, 8, c["metadata"]["dim"])
    box = c["box"]
    assert (c["positions"] >= 0).all() and (c["positions"] < box.astype(np.float32)).all()
    assert c["particle_type"].dtype == np.int32


def test_no_cpu_path():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU path"):
        _cabi.require_cuda()


def test_ctypes_structs_mirror_the_header():
    """sizeof / offsetof of every struct that crosses the C ABI, from a C program compiled against
    include/lb200.h, against the ctypes mirrors of _cabi.py."""
    import subprocess
    import tempfile

    pairs = [("lb200_grid", _cabi.Grid), ("lb200_feature_cfg", _cabi.FeatureCfg), ("lb200_mlp_off", _cabi.MlpOff),
             ("lb200_gns_cfg", _cabi.GnsCfg), ("lb200_integrate_cfg", _cabi.IntegrateCfg),
             ("lb200_rollout_cfg", _cabi.RolloutCfg), ("lb200_shard", _cabi.Shard)]
    lines = ["#include <stdio.h>", "#include <stddef.h>", '#include "lb200.h"', "int main(void) {"]
    for cname, ct in pairs:
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in ct._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["  return 0;", "}"]
    with tempfile.TemporaryDirectory() as tmp:
        src, exe = os.path.join(tmp, "layout.c"), os.path.join(tmp, "layout")
        with open(src, "w") as f:
            f.write("\n".join(lines))
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", exe, src])
        got = dict(ln.split() for ln in subprocess.check_output([exe], text=True).splitlines())
    for cname, ct in pairs:
        assert int(got[cname]) == C.sizeof(ct), cname
        for fname, _ in ct._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(ct, fname).offset, f"{cname}.{fname}"


def test_metrics_computer_values_against_the_oracle():
    """mse / mae with horizon slices, e_kin and the rollout averages (metrics.py:88-160,233-252) on a
    non-trivial periodic rollout pair."""
    from lagrangebench_b200 import MetricsComputer, averaged_metrics
    from lagrangebench_b200.case_setup import _make_space
    from oracle import metrics as ometrics
    from oracle import space as ospace

    rng = np.random.default_rng(3)
    box = np.array([1.0, 2.0, 0.5])
    meta = {"dt": 0.001, "write_every": 100, "dx": 0.05, "dim": 3}
    disp_t, _ = _make_space(box.tolist(), True, torch.float64)
    disp_np, _ = ospace.periodic(box)
    evals_t, evals_np = {}, {}
    for r in range(3):
        target = rng.random((27, 40, 3)) * box
        pred = np.mod(target + 0.3 * rng.standard_normal(target.shape) * np.linspace(0, 1, 27)[:, None, None], box)
        mc = MetricsComputer(["mse", "mae", "e_kin"], disp_t, meta, 6, stride=5)
        got = mc(torch.as_tensor(pred), torch.as_tensor(target))
        ref = ometrics.compute(["mse", "mae", "e_kin"], disp_np, meta, pred, target, stride=5)
        assert set(got) == set(ref) and {"mse5", "mse20", "mae10"} <= set(got) and "mse50" not in got
        for k in ref:
            if k == "e_kin":
                for kk in ("predicted", "target", "mse"):
                    np.testing.assert_allclose(np.asarray(got[k][kk]), ref[k][kk], rtol=1e-12)
            else:
                np.testing.assert_allclose(got[k].numpy(), ref[k], rtol=1e-12)
        evals_t[f"rollout_{r}"], evals_np[f"rollout_{r}"] = got, ref
    avg, avg_ref = averaged_metrics(evals_t), ometrics.averaged_metrics(evals_np)
    assert set(avg) == set(avg_ref) and "val/loss" in avg and "val/stdloss" in avg and "val/e_kin" in avg
    for k in avg_ref:
        assert abs(avg[k] - avg_ref[k]) <= 1e-12 * max(1.0, abs(avg_ref[k]))


def test_ghost_selection_bookkeeping():
    """select_ghosts: who sends what where, for periodic and walled cut axes, from all ranks' counts."""
    from lagrangebench_b200.domain import SlabDomain, select_ghosts

    box, halo = [1.0, 4.0, 1.0], 0.1
    rng = np.random.default_rng(0)
    for periodic in (True, False):
        world = 4
        coords = [torch.as_tensor(r + rng.random(30 + 5 * r)) for r in range(world)]  # slab r = [r, r + 1)
        doms = [SlabDomain(box, 1, world, r, halo, periodic=periodic) for r in range(world)]
        counts = []
        for r in range(world):
            ml, mr = doms[r].halo_masks(coords[r])
            counts.append([coords[r].numel(), int(ml.sum()) if doms[r].has_left else 0,
                           int(mr.sum()) if doms[r].has_right else 0])
        sels = [select_ghosts(doms[r], coords[r], counts_all=counts) for r in range(world)]
        for r in range(world):
            s, d = sels[r], doms[r]
            assert d.has_left == (periodic or r > 0) and d.has_right == (periodic or r < world - 1)
            assert s["send_left"].numel() == counts[r][1] and s["send_right"].numel() == counts[r][2]
            if d.has_left:
                assert bool((coords[r][s["send_left"]] < d.lo + halo).all())
                left = sels[d.left]
                # my left-going rows are the left neighbour's from-right block: behind its rows and its from-left block
                assert s["dst_row_left"] == counts[d.left][0] + left["n_ghost_left"]
                assert left["n_ghost_right"] == counts[r][1]
            else:
                assert s["n_ghost_left"] == 0 and s["send_left"].numel() == 0
            if d.has_right:
                assert s["dst_row_right"] == counts[d.right][0]
                assert sels[d.right]["n_ghost_left"] == counts[r][2]
            assert s["n_loc_max"] == max(c[0] + sels[i]["n_ghost_left"] + sels[i]["n_ghost_right"]
                                         for i, c in enumerate(counts))


def test_neighbor_list_survives_the_batch_round_trip():
    """utils.broadcast_to_batch / broadcast_from_batch (rollout.py:122,178) keep the grid handle."""
    from lagrangebench_b200.case_setup import NeighborList

    g = _cabi.Grid()
    g.n = g.n_valid = 5
    nl = NeighborList(None, torch.arange(12, dtype=torch.int32).view(2, 6), torch.zeros(4, dtype=torch.int32),
                      torch.zeros(5, 3), 3, 6, None, g)
    back = utils.broadcast_from_batch(utils.broadcast_to_batch(nl, 2), 1)
    assert back._grid is g and torch.equal(back.idx, nl.idx) and back.max_occupancy == 6
    assert back.idx.shape == (2, 6) and not bool(back.did_buffer_overflow)


def test_fp16_range_bounds_of_the_weights():
    """Static guard of the split-precision kernels: haiku-initialised weights sit far inside the fp16 range,
    a LayerNorm scale of 1e4 does not -- that model is routed to the float32 kernels."""
    params = models.init_params(15, 3, 128, 10, 16, seed=0)
    b = models.fp16_range_bounds(params, 10)
    assert max(b.values()) < 5000 and abs(b["edge latents"] - 11 * np.sqrt(127.0)) < 1e-6
    assert models.pack_params(params, 10, 3, device="cpu").fp16_safe
    params["gns/~_processor/layer_norm_4"]["scale"] = params["gns/~_processor/layer_norm_4"]["scale"] * 1.0e4
    pk = models.pack_params(params, 10, 3, device="cpu")
    assert not pk.fp16_safe and "edge latents" in pk.fp16_report
    with pytest.warns(UserWarning, match="fp16 split"):
        cfg = models.gns_cfg(pk, 100, 1000, 15, 15)
    assert cfg.edge_impl == models.EDGE_IMPL["simt"]


def test_load_haiku_reads_a_checkpoint_written_the_reference_way(tmp_path):
    """``utils.py:50-58`` of the reference, restated here as an independent writer: the leaves of
    ``jax.tree_leaves(params)`` (nested dict, keys sorted at every level) ``np.save``d back to back into
    ``params_array.npy``, the structure with every leaf replaced by 0 pickled into ``params_tree.pkl`` --
    here as a Mapping class of a module that does not exist at load time (haiku's FlatMapping)."""
    import json
    import pickle
    import sys
    import types

    params = ogns.init_params(12, 3, 2, num_mp_steps=2, seed=9)

    def leaves(tree):  # jax.tree_util order for dicts: sorted keys, depth first
        if isinstance(tree, dict):
            return [x for k in sorted(tree) for x in leaves(tree[k])]
        return [tree]

    ckp = tmp_path / "best"
    ckp.mkdir()
    with open(ckp / "params_array.npy", "wb") as f:
        for x in leaves(params):
            np.save(f, x, allow_pickle=False)
    mod = types.ModuleType("haiku_like_structures")

    class FlatMapping(dict):
        pass

    FlatMapping.__module__, FlatMapping.__qualname__ = "haiku_like_structures", "FlatMapping"
    mod.FlatMapping = FlatMapping
    sys.modules["haiku_like_structures"] = mod
    try:
        struct = FlatMapping({k: FlatMapping({kk: 0 for kk in v}) for k, v in params.items()})
        with open(ckp / "params_tree.pkl", "wb") as f:
            pickle.dump(struct, f)
        with open(ckp / "state_tree.pkl", "wb") as f:
            pickle.dump({}, f)
        open(ckp / "state_array.npy", "wb").close()
    finally:
        del sys.modules["haiku_like_structures"]  # the loader must cope without the class
    (ckp / "metadata_ckp.json").write_text(json.dumps({"step": 1234, "loss": 0.5}))
    got, state, opt, step = utils.load_haiku(str(ckp))
    assert step == 1234 and state == {} and opt is None
    assert sorted(got) == sorted(params) and type(got) is dict
    for k in params:
        for kk in params[k]:
            assert np.array_equal(got[k][kk], params[k][kk]), (k, kk)
    assert utils.get_num_params(got) == ogns.num_params(params)


def test_zero_padding_of_a_narrower_model_is_the_same_function():
    """``models._pad_latent``: the 64-wide GNS-5-64 embedded in 128-wide arrays.  Evaluated with the oracle's
    forward (LayerNorm over the true width emulated by its own padded statistics) the outputs agree."""
    from oracle import features as ofeatures  # noqa: F401

    rng = np.random.default_rng(0)
    width, mp, n, e = 64, 3, 40, 200
    params = ogns.init_params(10, 3, 2, latent=width, num_mp_steps=mp, seed=3, perturb=True)
    padded = models._pad_latent(params, mp, width, 128)
    assert padded["gns/~_processor/MLP/~/linear_0"]["w"].shape == (384, 128)
    assert padded["gns/~_processor/MLP_1/~/linear_0"]["w"].shape == (256, 128)
    assert padded["gns/~_decoder/MLP/~/linear_1"]["w"].shape == (128, 2)
    w = padded["gns/~_processor/MLP/~/linear_0"]["w"]
    assert np.array_equal(w[:64, :64], params["gns/~_processor/MLP/~/linear_0"]["w"][:64])
    assert np.array_equal(w[128:192, :64], params["gns/~_processor/MLP/~/linear_0"]["w"][64:128])
    assert not w[64:128].any() and not w[:, 64:].any()
    pk = models.pack_params(params, mp, 2, device="cpu")
    assert pk.latent == 64 and models.gns_cfg(pk, n, e, 10, 10).latent == 64
