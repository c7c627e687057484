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
    with pytest.raises(NotImplementedError):
        models.GNS(2, 64, 2, 5, 16)


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
