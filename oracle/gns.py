"""Oracle: GNS encode-process-decode forward (test infrastructure only).

Restates ``lagrangebench/models/gns.py:18-171`` with ``build_mlp``
(``lagrangebench/models/utils.py:100-115``) in NumPy, in float64 (ground truth) or
float32 ("reference-precision twin").  The arithmetic of the third-party pieces is
restated from their published semantics (not vendored in the reference):

  * ``hk.Linear``: ``y = x @ w + b`` with ``w: (in, out)``; ``hk.nets.MLP``: ReLU between
    layers, no final activation (``models/utils.py:105-110`` passes no ``activation=``);
  * ``hk.LayerNorm(axis=-1, create_scale=True, create_offset=True)``:
    ``(x - mean) * rsqrt(var_biased + 1e-5) * scale + offset``;
  * ``hk.Embed(9, 16)``: row lookup (``gns.py:61-63``);
  * ``jraph.GraphNetwork`` (``gns.py:117-119``): ``sent = nodes[senders]``,
    ``recv = nodes[receivers]``, ``e' = update_edge_fn(edges, sent, recv)``,
    ``agg_r = segment_sum(e', receivers, N)`` (the sender aggregate is computed and
    ignored, ``gns.py:105``), ``n' = update_node_fn(nodes, agg_r)``; edges whose index is
    the pad value ``N`` fall outside every segment and are dropped.
  * residuals outside the block (``gns.py:120-122``): ``nodes += n'``, ``edges += e'``.

PARITY UNPINNED by the reference (it has no GNS test); see ``oracle/__init__.py``.
Parameter naming follows haiku's module-path convention for this class; the loader in
``lagrangebench_b200`` accepts both ``gns/embed`` and ``gns/~/embed`` spellings.
"""

import numpy as np


def mlp_names(prefix, idx):
    s = "" if idx == 0 else f"_{idx}"
    return (f"{prefix}/MLP{s}/~/linear_0", f"{prefix}/MLP{s}/~/linear_1", f"{prefix}/layer_norm{s}")


def init_params(node_in, edge_in, dim, latent=128, num_mp_steps=10, embed=16, num_types=9,
                seed=0, perturb=True):
    """Random parameters in haiku's layout (``{module_path: {name: array}}``), float32.

    haiku defaults: Linear ``w ~ TruncatedNormal(1/sqrt(fan_in))``, ``b = 0``; LayerNorm
    ``scale = 1``, ``offset = 0``; Embed ``TruncatedNormal(1)``.  With ``perturb`` the
    biases / scales / offsets are made non-trivial so that a bug in any of them shows."""
    rng = np.random.default_rng(seed)

    def trunc_normal(shape, std):
        x = rng.standard_normal(shape)
        bad = np.abs(x) > 2.0
        while bad.any():
            x[bad] = rng.standard_normal(int(bad.sum()))
            bad = np.abs(x) > 2.0
        return (x * std).astype(np.float32)

    def linear(fan_in, fan_out):
        b = (0.1 * rng.standard_normal(fan_out)).astype(np.float32) if perturb else np.zeros(fan_out, np.float32)
        return {"w": trunc_normal((fan_in, fan_out), 1.0 / np.sqrt(fan_in)), "b": b}

    def layer_norm(n):
        if perturb:
            return {"scale": (1.0 + 0.1 * rng.standard_normal(n)).astype(np.float32),
                    "offset": (0.1 * rng.standard_normal(n)).astype(np.float32)}
        return {"scale": np.ones(n, np.float32), "offset": np.zeros(n, np.float32)}

    params = {"gns/~/embed": {"embeddings": trunc_normal((num_types, embed), 1.0)}}

    def add_mlp(prefix, idx, fan_in, out, ln=True):
        l0, l1, lnn = mlp_names(prefix, idx)
        params[l0] = linear(fan_in, latent)
        params[l1] = linear(latent, out)
        if ln:
            params[lnn] = layer_norm(out)

    add_mlp("gns/~_encoder", 0, node_in + embed, latent)
    add_mlp("gns/~_encoder", 1, edge_in, latent)
    for m in range(num_mp_steps):
        add_mlp("gns/~_processor", 2 * m, 3 * latent, latent)
        add_mlp("gns/~_processor", 2 * m + 1, 2 * latent, latent)
    add_mlp("gns/~_decoder", 0, latent, dim, ln=False)
    return params


def num_params(params):
    return int(sum(np.prod(a.shape) for mod in params.values() for a in mod.values()))


def _mlp(params, prefix, idx, x, dtype, ln=True):
    l0, l1, lnn = mlp_names(prefix, idx)
    w0, b0 = params[l0]["w"].astype(dtype), params[l0]["b"].astype(dtype)
    w1, b1 = params[l1]["w"].astype(dtype), params[l1]["b"].astype(dtype)
    h = np.maximum(x @ w0 + b0, 0)
    y = h @ w1 + b1
    if ln:
        scale, offset = params[lnn]["scale"].astype(dtype), params[lnn]["offset"].astype(dtype)
        mean = y.mean(axis=-1, keepdims=True)
        var = ((y - mean) ** 2).mean(axis=-1, keepdims=True)
        y = (y - mean) / np.sqrt(var + dtype.type(1e-5)) * scale + offset
    return y


def segment_sum(data, segment_ids, num_segments):
    """``jax.ops.segment_sum``: rows whose id is outside ``[0, num_segments)`` are dropped;
    rows are accumulated in ascending edge order within each segment."""
    keep = (segment_ids >= 0) & (segment_ids < num_segments)
    data, segment_ids = data[keep], segment_ids[keep]
    out = np.zeros((num_segments,) + data.shape[1:], dtype=data.dtype)
    if data.shape[0] == 0:
        return out
    order = np.argsort(segment_ids, kind="stable")
    sid = segment_ids[order]
    starts = np.flatnonzero(np.concatenate(([True], sid[1:] != sid[:-1])))
    out[sid[starts]] = np.add.reduceat(data[order], starts, axis=0)
    return out


def node_edge_inputs(features, particle_type, params, dtype):
    """``GNS._transform`` + type embedding (``gns.py:135-169``)."""
    nodes = np.concatenate(
        [np.asarray(features[k], dtype=dtype).reshape(features["vel_hist"].shape[0], -1)
         for k in ["vel_hist", "vel_mag", "bound", "force"] if k in features], axis=-1)
    edges = np.concatenate(
        [np.asarray(features[k], dtype=dtype) for k in ["rel_disp", "rel_dist"] if k in features], axis=-1)
    key = "gns/~/embed" if "gns/~/embed" in params else "gns/embed"
    emb = params[key]["embeddings"].astype(dtype)
    n_types = emb.shape[0]
    pt = np.asarray(particle_type)
    pt = np.clip(np.where(pt < 0, pt + n_types, pt), 0, n_types - 1)  # hk.Embed indexes like NumPy: -1 is the last row
    nodes = np.concatenate([nodes, emb[pt]], axis=-1)
    return nodes, edges


def forward(params, features, particle_type, num_mp_steps=10, dtype=np.float32, return_latents=False):
    """``GNS.__call__`` (``gns.py:159-171``) -> ``{"acc": (N, dim)}``.

    Only real edges (index < N) are processed; pad edges cannot influence the output
    (they are dropped by ``segment_sum``)."""
    dtype = np.dtype(dtype)
    n = features["vel_hist"].shape[0]
    nodes_in, edges_in = node_edge_inputs(features, particle_type, params, dtype)
    senders = np.asarray(features["senders"])
    receivers = np.asarray(features["receivers"])
    real = (senders < n) & (receivers < n)
    senders, receivers, edges_in = senders[real], receivers[real], edges_in[real]
    h = _mlp(params, "gns/~_encoder", 0, nodes_in, dtype)
    e = _mlp(params, "gns/~_encoder", 1, edges_in, dtype)
    for m in range(num_mp_steps):
        x = np.concatenate([h[senders], h[receivers], e], axis=-1)  # gns.py:97-100
        e_new = _mlp(params, "gns/~_processor", 2 * m, x, dtype)
        agg = segment_sum(e_new, receivers, n)
        h_new = _mlp(params, "gns/~_processor", 2 * m + 1, np.concatenate([h, agg], axis=-1), dtype)
        h = h_new + h
        e = e_new + e
    acc = _mlp(params, "gns/~_decoder", 0, h, dtype, ln=False)
    if return_latents:
        return {"acc": acc}, h, e
    return {"acc": acc}
