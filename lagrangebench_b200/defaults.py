"""Default configs, mirroring the keys of ``lagrangebench/defaults.py:7-176`` that the
rollout hot path reads.  Plain nested dicts (omegaconf is not a dependency); every public
function merges a user dict over these, as the reference does (``case.py:91-98``,
``rollout.py:345-349``)."""

import copy

defaults = {
    "seed": 0,
    "dtype": "float64",  # defaults.py:22
    "model": {
        "name": None,
        "input_seq_length": 6,
        "num_mp_steps": 10,
        "num_mlp_layers": 2,
        "latent_dim": 128,
        "magnitude_features": False,
        "isotropic_norm": False,
    },
    "train": {"noise_std": 3.0e-4},
    "eval": {
        "n_rollout_steps": 20,
        "rollout_dir": None,
        "infer": {
            "n_trajs": -1,
            "metrics_stride": 1,
            "batch_size": 2,
            "metrics": ["mse"],
            "out_type": "pkl",
            "n_extrap_steps": 0,
        },
    },
    "neighbors": {"backend": "b200", "multiplier": 1.25},  # defaults.py:172-174 (backend slot)
}


def merged(section, override):
    """``OmegaConf.merge(defaults.<section>, override)`` for plain dicts."""
    base = defaults
    for key in section.split("."):
        base = base[key]
    out = copy.deepcopy(base)
    if override is not None:
        for k, v in dict(override).items():
            out[k] = v
    return out
