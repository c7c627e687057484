"""Shared builders for the parity tests: the same seeded inputs through the CUDA path
(``lagrangebench_b200``) and the oracle (``oracle/``)."""

import numpy as np

from lagrangebench_b200 import case_builder, synthetic
from oracle import case as ocase


def oracle_force(force):
    if force is None:
        return None
    lo, hi = np.array(force.lo), np.array(force.hi)
    return lambda r: hi if r[force.axis] > force.threshold else lo


def build_pair(name, dtype="float32", seed=0, n_future=0, multiplier=None, dims=None, isl=6):
    npdtype = np.float32 if dtype == "float32" else np.float64
    c = synthetic.make_case(name, isl, n_future, seed, npdtype, dims)
    mult = multiplier if multiplier is not None else c["multiplier"]
    kw = dict(cfg_neighbors={"multiplier": mult}, noise_std=3e-4)
    ours = case_builder(c["box"], c["metadata"], isl, external_force_fn=c["force"], dtype=dtype, **kw)
    orac = ocase.case_builder(c["box"], c["metadata"], isl, external_force_fn=oracle_force(c["force"]),
                              dtype=npdtype, **kw)
    return c, ours, orac


def rel_err(a, b):
    """max |a - b| / max |b| -- the scale-relative error the 1e-5 bar is stated in."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def elem_rel_err(a, b, floor=1e-2):
    """Elementwise relative error |a - b| / |b| over the elements that carry signal (|b| above ``floor`` of
    the largest): returns (max, 99.9th percentile).  Reported beside the scale-relative ``rel_err``."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    keep = np.abs(b) > floor * np.abs(b).max()
    r = np.abs(a - b)[keep] / np.abs(b)[keep]
    return float(r.max()), float(np.percentile(r, 99.9))
