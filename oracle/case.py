"""Oracle: ``case_builder`` (test infrastructure only).

Restates ``lagrangebench/case_setup/case.py:62-269`` in NumPy: ``allocate`` /
``preprocess`` (noise-free: ``add_gns_noise`` is a training-only strategy,
``train/strats.py``, out of scope), ``allocate_eval`` / ``preprocess_eval``
(``case.py:162-228``), ``_compute_target`` (``case.py:143-160``) and the semi-implicit
Euler ``integrate`` (``case.py:230-259``).
Pinned by ``tests/case_test.py:72-206``.
"""

from types import SimpleNamespace

import numpy as np

from . import features as ofeatures
from . import partition, space

DEFAULT_NEIGHBORS = {"backend": "jaxmd_vmap", "multiplier": 1.25}  # defaults.py:172-174
DEFAULT_MODEL = {"isotropic_norm": False, "magnitude_features": False}  # defaults.py:51-53


def case_builder(box, metadata, input_seq_length, cfg_neighbors=None, cfg_model=None,
                 noise_std=3e-4, external_force_fn=None, dtype=np.float64):
    cfg_neighbors = {**DEFAULT_NEIGHBORS, **(cfg_neighbors or {})}
    cfg_model = {**DEFAULT_MODEL, **(cfg_model or {})}
    dtype = np.dtype(dtype)
    stats = ofeatures.get_dataset_stats(metadata, cfg_model["isotropic_norm"], noise_std, dtype)
    pbc = list(metadata["periodic_boundary_conditions"])
    if any(pbc):  # case.py:104-108: periodic in all directions or not at all
        displacement_fn, shift_fn = space.periodic(np.asarray(box, dtype=dtype))
    else:
        displacement_fn, shift_fn = space.free()
    neighbor_fn = partition.neighbor_list(
        displacement_fn, np.asarray(box), metadata["default_connectivity_radius"],
        capacity_multiplier=cfg_neighbors["multiplier"], dtype=dtype)
    feature_transform = ofeatures.physical_feature_builder(
        metadata["bounds"], stats, metadata["default_connectivity_radius"], displacement_fn,
        pbc, cfg_model["magnitude_features"], external_force_fn)

    def _compute_target(pos_input):  # case.py:143-160
        current_velocity = displacement_fn(pos_input[:, 1], pos_input[:, 0])
        next_velocity = displacement_fn(pos_input[:, 2], pos_input[:, 1])
        acc = next_velocity - current_velocity
        a, v = stats["acceleration"], stats["velocity"]
        return {"acc": (acc - a["mean"]) / a["std"], "vel": (next_velocity - v["mean"]) / v["std"],
                "pos": pos_input[:, -1]}

    def _preprocess(sample, neighbors=None, is_allocate=False, mode="train", unroll_steps=0):
        pos_input = np.asarray(sample[0], dtype=dtype)
        most_recent_position = pos_input[:, input_seq_length - 1]
        num_particles = int((np.asarray(sample[1]) != -1).sum())  # case.py:182
        if is_allocate:
            neighbors = neighbor_fn.allocate(most_recent_position, num_particles=num_particles)
        else:
            neighbors = neighbors.update(most_recent_position, num_particles=num_particles)
        feats = feature_transform(pos_input[:, :input_seq_length], neighbors)
        if mode == "train":
            begin = input_seq_length - 2 + unroll_steps
            return feats, _compute_target(pos_input[:, begin:begin + 3]), neighbors
        return feats, neighbors

    def allocate(key, sample, noise_std=0.0, unroll_steps=0):
        assert noise_std == 0.0, "the oracle restates the noise-free path only"
        f, t, n = _preprocess(sample, is_allocate=True, unroll_steps=unroll_steps)
        return key, f, t, n

    def preprocess(key, sample, noise_std, neighbors, unroll_steps=0):
        assert noise_std == 0.0, "the oracle restates the noise-free path only"
        f, t, n = _preprocess(sample, neighbors, unroll_steps=unroll_steps)
        return key, f, t, n

    def allocate_eval(sample):
        return _preprocess(sample, is_allocate=True, mode="eval")

    def preprocess_eval(sample, neighbors):
        return _preprocess(sample, neighbors, mode="eval")

    def integrate(normalized_in, position_sequence):  # case.py:230-259
        position_sequence = np.asarray(position_sequence, dtype=dtype)
        if "pos" in normalized_in:
            return np.asarray(normalized_in["pos"])
        most_recent_position = position_sequence[:, -1]
        if "vel" in normalized_in:
            v = stats["velocity"]
            new_velocity = v["mean"] + np.asarray(normalized_in["vel"]) * v["std"]
        else:
            a = stats["acceleration"]
            acceleration = a["mean"] + np.asarray(normalized_in["acc"]) * a["std"]
            most_recent_velocity = displacement_fn(most_recent_position, position_sequence[:, -2])
            new_velocity = most_recent_velocity + acceleration
        return shift_fn(most_recent_position, new_velocity)

    return SimpleNamespace(allocate=allocate, preprocess=preprocess, allocate_eval=allocate_eval,
                           preprocess_eval=preprocess_eval, integrate=integrate,
                           displacement=displacement_fn, normalization_stats=stats)
