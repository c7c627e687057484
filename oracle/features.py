"""Oracle: physical feature transform (test infrastructure only).

Restates ``lagrangebench/case_setup/features.py:14-128``
(``physical_feature_builder`` / ``feature_transform``) and
``lagrangebench/data/utils.py:9-45`` (``get_dataset_stats``) in NumPy.
Pinned by ``tests/case_test.py:102-137`` (``vel_hist``, ``rel_disp``, ``rel_dist``).
"""

import numpy as np

from . import space


def get_dataset_stats(metadata, is_isotropic_norm, noise_std, dtype=np.float32):
    """``data/utils.py:9-45``."""
    dtype = np.dtype(dtype)
    acc_mean = np.array(metadata["acc_mean"], dtype=dtype)
    acc_std = np.array(metadata["acc_std"], dtype=dtype)
    vel_mean = np.array(metadata["vel_mean"], dtype=dtype)
    vel_std = np.array(metadata["vel_std"], dtype=dtype)
    if is_isotropic_norm:
        acc_mean = np.mean(acc_mean) * np.ones_like(acc_mean)
        acc_std = np.sqrt(np.mean(acc_std**2)) * np.ones_like(acc_std)
        vel_mean = np.mean(vel_mean) * np.ones_like(vel_mean)
        vel_std = np.sqrt(np.mean(vel_std**2)) * np.ones_like(vel_std)
    ns = dtype.type(noise_std)
    return {
        "acceleration": {"mean": acc_mean, "std": np.sqrt(acc_std**2 + ns**2)},
        "velocity": {"mean": vel_mean, "std": np.sqrt(vel_std**2 + ns**2)},
    }


def physical_feature_builder(bounds, normalization_stats, connectivity_radius, displacement_fn,
                             pbc, magnitude_features=False, external_force_fn=None):
    """``features.py:14-128``.  ``external_force_fn`` maps one position ``(d,)`` to ``(d,)``."""
    velocity_stats = normalization_stats["velocity"]

    def feature_transform(pos_input, nbrs):
        features = {}
        n = pos_input.shape[0]
        dtype = pos_input.dtype
        most_recent_position = pos_input[:, -1]
        # features.py:68-78
        velocity_sequence = displacement_fn(pos_input[:, 1:], pos_input[:, :-1])
        normalized = (velocity_sequence - velocity_stats["mean"].astype(dtype)) / velocity_stats["std"].astype(dtype)
        features["abs_pos"] = pos_input
        features["vel_hist"] = normalized.reshape(n, -1)
        if magnitude_features:  # features.py:80-85
            features["vel_mag"] = np.sqrt(space.sum_sq(normalized))
        if not any(pbc):  # features.py:87-103
            boundaries = np.array(bounds, dtype=dtype)
            lower = most_recent_position - boundaries[:, 0][None]
            upper = boundaries[:, 1][None] - most_recent_position
            dist = np.concatenate([lower, upper], axis=1)
            features["bound"] = np.clip(dist / dtype.type(connectivity_radius), -1.0, 1.0).astype(dtype)
        if external_force_fn is not None:  # features.py:105-107
            features["force"] = np.stack([np.asarray(external_force_fn(p)) for p in most_recent_position]).astype(dtype)
        receivers, senders = nbrs.idx  # features.py:110
        features["senders"] = senders
        features["receivers"] = receivers
        # JAX gathers clamp the pad index N to N-1 (features.py:115-117)
        r_c = np.minimum(receivers, n - 1)
        s_c = np.minimum(senders, n - 1)
        disp = displacement_fn(most_recent_position[r_c], most_recent_position[s_c])
        rel = disp / dtype.type(connectivity_radius)
        features["rel_disp"] = rel
        features["rel_dist"] = space.distance(rel)[:, None]
        return features

    return feature_transform
