"""Oracle twin: the GNS forward in float32 torch-CPU ops (test / baseline infrastructure only).

Same restatement as :mod:`oracle.gns` (``lagrangebench/models/gns.py:61-171`` with
``build_mlp`` of ``models/utils.py:100-115``, jraph's gather / ``segment_sum``), written against
``torch`` so that the dense layers run on every host core (``torch.set_num_threads``): the
"reference-precision" CPU baseline ``bench.py`` times when JAX is not installed.  It follows the
reference's data flow literally -- gather ``h[senders]``, ``h[receivers]``, concatenate with the
edge latents, 384-wide first layer, scatter-add over receivers -- i.e. none of this repo's
restructuring.  Checked against :func:`oracle.gns.forward` in ``tests/test_oracle_golden.py``.
"""

import numpy as np
import torch

from .gns import mlp_names


def _t(a, dtype=np.float32):
    return torch.as_tensor(np.ascontiguousarray(a, dtype=dtype))


def pack(params, dtype=np.float32):
    """NumPy parameter tree -> the same tree of torch tensors (done once, outside the timed loop).
    ``dtype=np.float64`` gives the double-precision ground truth of the large parity tests."""
    return {mod: {k: _t(v, dtype) for k, v in leaves.items()} for mod, leaves in params.items()}


def _mlp(tp, prefix, idx, x, ln=True):
    l0, l1, lnn = mlp_names(prefix, idx)
    h = torch.relu(torch.addmm(tp[l0]["b"], x, tp[l0]["w"]))
    y = torch.addmm(tp[l1]["b"], h, tp[l1]["w"])
    if ln:
        y = torch.nn.functional.layer_norm(y, (y.shape[-1],), tp[lnn]["scale"], tp[lnn]["offset"], 1e-5)
    return y


def forward(tp, features, particle_type, num_mp_steps=10):
    """``GNS.__call__`` (``gns.py:159-171``) -> ``{"acc": (N, dim) ndarray}`` in the dtype of ``tp = pack(params)``."""
    n = features["vel_hist"].shape[0]
    key = "gns/~/embed" if "gns/~/embed" in tp else "gns/embed"
    emb = tp[key]["embeddings"]
    dt = np.float64 if emb.dtype == torch.float64 else np.float32
    nodes = torch.cat([_t(features[k], dt).reshape(n, -1) for k in ("vel_hist", "vel_mag", "bound", "force")
                       if k in features], dim=1)
    edges = torch.cat([_t(features[k], dt) for k in ("rel_disp", "rel_dist")], dim=1)
    pt = torch.as_tensor(np.asarray(particle_type)).long()
    pt = torch.where(pt < 0, pt + emb.shape[0], pt).clamp_(0, emb.shape[0] - 1)
    nodes = torch.cat([nodes, emb[pt]], dim=1)
    senders = torch.as_tensor(np.asarray(features["senders"])).long()
    receivers = torch.as_tensor(np.asarray(features["receivers"])).long()
    real = (senders < n) & (receivers < n)
    senders, receivers, edges = senders[real], receivers[real], edges[real]
    h = _mlp(tp, "gns/~_encoder", 0, nodes)
    e = _mlp(tp, "gns/~_encoder", 1, edges)
    for m in range(num_mp_steps):
        x = torch.cat([h[senders], h[receivers], e], dim=1)  # gns.py:97-100
        e_new = _mlp(tp, "gns/~_processor", 2 * m, x)
        agg = torch.zeros((n, e_new.shape[1]), dtype=e_new.dtype).index_add_(0, receivers, e_new)
        h = _mlp(tp, "gns/~_processor", 2 * m + 1, torch.cat([h, agg], dim=1)) + h
        e = e_new + e
    return {"acc": _mlp(tp, "gns/~_decoder", 0, h, ln=False).numpy()}
