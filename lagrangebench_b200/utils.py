"""General utilities mirroring ``lagrangebench/utils.py`` for the rollout path."""

import enum
import json
import os
import pickle

import numpy as np
import torch


class NodeType(enum.IntEnum):
    """Particle types (``lagrangebench/utils.py:17-25``)."""

    PAD_VALUE = -1
    FLUID = 0
    SOLID_WALL = 1
    MOVING_WALL = 2
    RIGID_BODY = 3
    SIZE = 9


def get_kinematic_mask(particle_type):
    """True for obstacle / padding particles (``utils.py:28-35``)."""
    pt = torch.as_tensor(particle_type)
    return (pt == NodeType.SOLID_WALL) | (pt == NodeType.MOVING_WALL) | (pt == NodeType.PAD_VALUE)


def _tree_map(fn, tree):
    if isinstance(tree, dict):
        return type(tree)((k, _tree_map(fn, v)) for k, v in tree.items())
    if isinstance(tree, (tuple, list)):
        return type(tree)(_tree_map(fn, v) for v in tree)
    if hasattr(tree, "tree_map"):
        return tree.tree_map(fn)
    return fn(tree)


def broadcast_to_batch(sample, batch_size):
    """``utils.py:38-41``: add a leading batch axis of size ``batch_size`` to every leaf."""
    assert batch_size > 0

    def rep(x):
        x = torch.as_tensor(x)
        return x.unsqueeze(0).expand(batch_size, *x.shape).contiguous()

    return _tree_map(rep, sample)


def broadcast_from_batch(batch, index):
    """``utils.py:44-47``."""
    assert index >= 0
    return _tree_map(lambda x: x[index], batch)


def resolve_dtype(dtype):
    """'float32' / 'float64' / numpy / torch dtype -> torch dtype."""
    if isinstance(dtype, torch.dtype):
        out = dtype
    else:
        out = {"float32": torch.float32, "float64": torch.float64}[np.dtype(dtype).name]
    if out not in (torch.float32, torch.float64):
        raise ValueError("dtype must be float32 or float64 (lagrangebench/defaults.py:22)")
    return out


def tree_leaves_sorted(tree):
    """Leaves in ``jax.tree_leaves`` order (dict keys sorted)."""
    if isinstance(tree, dict):
        out = []
        for k in sorted(tree):
            out += tree_leaves_sorted(tree[k])
        return out
    if isinstance(tree, (tuple, list)):
        out = []
        for v in tree:
            out += tree_leaves_sorted(v)
        return out
    return [tree]


def _tree_unflatten_sorted(tree, leaves):
    if isinstance(tree, dict):
        return {k: _tree_unflatten_sorted(tree[k], leaves) for k in sorted(tree)}
    if isinstance(tree, (tuple, list)):
        return type(tree)(_tree_unflatten_sorted(v, leaves) for v in tree)
    return next(leaves)


def save_pytree(ckp_dir, pytree_obj, name):
    """``utils.py:50-58``: leaves ``np.save``d back to back into one file, structure pickled."""
    os.makedirs(ckp_dir, exist_ok=True)
    with open(os.path.join(ckp_dir, f"{name}_array.npy"), "wb") as f:
        for x in tree_leaves_sorted(pytree_obj):
            np.save(f, np.asarray(x), allow_pickle=False)
    with open(os.path.join(ckp_dir, f"{name}_tree.pkl"), "wb") as f:
        pickle.dump(_tree_map(lambda t: 0, pytree_obj), f)


class _TreeUnpickler(pickle.Unpickler):
    """``*_tree.pkl`` holds only the STRUCTURE of the pytree.  Checkpoints written with older haiku
    versions pickle it as ``haiku._src.data_structures.FlatMapping`` (a Mapping); without haiku installed
    that class -- and any other mapping class of a missing module -- is read as a plain dict."""

    class _Mapping(dict):
        def __setstate__(self, state):  # FlatMapping pickles {"_structure"/"_leaves"} or its dict
            if isinstance(state, dict):
                inner = state.get("_mapping", state.get("mapping", state))
                self.update(inner if isinstance(inner, dict) else state)

        def __reduce_ex__(self, protocol):
            return (dict, (dict(self),))

    def find_class(self, module, name):
        try:
            return super().find_class(module, name)
        except (ImportError, AttributeError):
            return _TreeUnpickler._Mapping


def _plain(tree):
    if isinstance(tree, dict):
        return {k: _plain(v) for k, v in tree.items()}
    if isinstance(tree, (tuple, list)):
        return type(tree)(_plain(v) for v in tree)
    return tree


def load_pytree(model_dir, name):
    """``utils.py:99-109``."""
    with open(os.path.join(model_dir, f"{name}_tree.pkl"), "rb") as f:
        tree_struct = _plain(_TreeUnpickler(f).load())
    n_leaves = len(tree_leaves_sorted(tree_struct))
    with open(os.path.join(model_dir, f"{name}_array.npy"), "rb") as f:
        flat = [np.load(f) for _ in range(n_leaves)]
    return _tree_unflatten_sorted(tree_struct, iter(flat))


def save_haiku(ckp_dir, params, state, opt_state=None, metadata_ckp=None):
    """Parameter / state part of ``utils.py:61-96`` (no optimizer state on this path)."""
    save_pytree(ckp_dir, params, "params")
    save_pytree(ckp_dir, state if state is not None else {}, "state")
    with open(os.path.join(ckp_dir, "metadata_ckp.json"), "w") as f:
        json.dump(metadata_ckp or {"step": 0, "loss": float("inf")}, f)


def load_haiku(model_dir):
    """``utils.py:112-128`` -> ``(params, state, opt_state, step)``; ``opt_state`` is not
    loaded (cloudpickled optax state, training only)."""
    params = load_pytree(model_dir, "params")
    state_tree = os.path.join(model_dir, "state_tree.pkl")
    state = load_pytree(model_dir, "state") if os.path.exists(state_tree) else {}
    step = 0
    meta = os.path.join(model_dir, "metadata_ckp.json")
    if os.path.exists(meta):
        with open(meta) as fp:
            step = json.load(fp).get("step", 0)
    return params, state, None, step


def get_num_params(params):
    return int(sum(np.prod(np.asarray(p).shape) for p in tree_leaves_sorted(params)))


def set_seed(seed):
    """``utils.py:144-161`` without the JAX key: returns ``(key, seed_worker, generator)``."""
    import random

    np.random.seed(seed)
    random.seed(seed)
    torch.manual_seed(seed)

    def seed_worker(_):
        worker_seed = torch.initial_seed() % 2**32
        np.random.seed(worker_seed)
        random.seed(worker_seed)

    generator = torch.Generator()
    generator.manual_seed(seed)
    return seed, seed_worker, generator
