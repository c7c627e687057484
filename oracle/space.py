"""Oracle: displacement / shift functions (test infrastructure only).

Restates ``jax_sph.jax_md.space`` (jax-sph 0.0.3, a fork of jax-md; not vendored in the
reference) as used at ``lagrangebench/case_setup/case.py:104-110``:

  * ``space.periodic(side)``: ``disp(a, b) = mod(a - b + side/2, side) - side/2`` and
    ``shift(r, dr) = mod(r + dr, side)`` with floor-mod semantics (``jnp.mod``);
  * ``space.free()``: ``a - b`` and ``r + dr``;
  * ``space.distance(dR) = sqrt(sum(dR**2))`` with 0 mapped to 0 (safe mask).

Pinned by the reference's ``tests/case_test.py:116-137`` (periodic wrap of rel_disp)
and ``tests/case_test.py:195-206`` (integrate through ``shift``).

NumPy evaluates elementwise ops without FMA contraction; the CUDA kernels use
``__fmul_rn``/``__fadd_rn`` (or the f64 twins) in the same order to stay bit-identical.
"""

import numpy as np


def floor_mod(x, side):
    """``jnp.mod`` for floats: C ``fmod`` (exact), then ``+ side`` where the remainder is
    non-zero and has the wrong sign.  ``side > 0`` here, so that is ``r < 0``."""
    r = np.fmod(x, side)
    return np.where(r < 0, r + side, r)


def periodic(side):
    side = np.asarray(side)
    half = side * side.dtype.type(0.5)

    def displacement(a, b):
        return floor_mod((a - b) + half, side) - half

    def shift(r, dr):
        return floor_mod(r + dr, side)

    return displacement, shift


def free():
    def displacement(a, b):
        return a - b

    def shift(r, dr):
        return r + dr

    return displacement, shift


def sum_sq(dr):
    """Sum of squares over the last axis, accumulated left to right (x*x + y*y) + z*z."""
    acc = dr[..., 0] * dr[..., 0]
    for k in range(1, dr.shape[-1]):
        acc = acc + dr[..., k] * dr[..., k]
    return acc


def distance(dr):
    d2 = sum_sq(dr)
    out = np.zeros_like(d2)
    pos = d2 > 0
    out[pos] = np.sqrt(d2[pos])
    return out
