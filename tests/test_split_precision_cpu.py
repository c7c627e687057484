"""NumPy emulation of the split-precision scheme the tcgen05 kernels use (csrc/gns_tc2.cu):

    x = hi + lo,      hi = fp16(x),  lo  = fp16(x - hi)              (activations, unscaled)
    w = hi + lo'/2^11, hi = fp16(w), lo' = fp16((w - hi) * 2^11)     (weights, host-packed)
    acc = (W_lo' X_hi) * 2^-11 + W_hi X_lo + W_hi X_hi               (fp32 accumulate, lo*lo dropped)

It pins on the CPU why three fp16 MMA passes reach float32-level accuracy (BASELINE.json asks for
1e-5 on the accelerations after 10 message-passing layers) and why a plain fp16 GEMM would not."""

import numpy as np


def _split_activation(x):
    hi = x.astype(np.float16)
    lo = (x - hi.astype(np.float32)).astype(np.float16)  # may be an fp16 subnormal: the tensor core honours it
    return hi.astype(np.float64), lo.astype(np.float64)


def _split_weight(w):
    hi = w.astype(np.float16)
    lo = ((w - hi.astype(np.float32)) * np.float32(2048.0)).astype(np.float16)
    return hi.astype(np.float64), lo.astype(np.float64)


def _gemm_split(w, x):
    """(out, in) x (in, edges): products of fp16 values are exact in the tensor core's accumulator."""
    w_hi, w_lo = _split_weight(w)
    x_hi, x_lo = _split_activation(x)
    acc = (w_lo @ x_hi).astype(np.float32) * np.float32(2.0 ** -11)      # first pass, then scale-input-d = 11
    acc = (acc.astype(np.float64) + w_hi @ x_lo).astype(np.float32)
    return (acc.astype(np.float64) + w_hi @ x_hi).astype(np.float32)


def _rel(a, ref):
    return np.abs(a - ref).max() / np.abs(ref).max()


def test_three_pass_split_reaches_float32_accuracy():
    rng = np.random.default_rng(0)
    w = (rng.standard_normal((128, 128)) / np.sqrt(128)).astype(np.float32)   # hk.Linear init scale
    x = rng.standard_normal((128, 256)).astype(np.float32)                    # LayerNorm'd latents
    x[:, :32] *= 1e-3                                                         # small rows: lo is subnormal there
    ref = w.astype(np.float64) @ x.astype(np.float64)
    err_split = _rel(_gemm_split(w, x), ref)
    err_f32 = _rel((w @ x).astype(np.float32), ref)
    err_f16 = _rel(w.astype(np.float16).astype(np.float64) @ x.astype(np.float16).astype(np.float64), ref)
    assert err_split <= 4e-7, err_split          # float32 rounding level (2^-24 = 6e-8 per element)
    assert err_split <= 4 * max(err_f32, 6e-8)   # as good as a float32 GEMM
    assert err_f16 >= 1e-4                       # a single fp16 pass is three orders of magnitude off


def test_unscaled_low_half_costs_nothing_measurable():
    """Dropping the 2^11 factor on the activations' low halves (kNoScale) moves them into fp16's
    subnormal range for |x| < 0.125; the absolute error stays at 2^-25 per element."""
    rng = np.random.default_rng(1)
    x = (rng.standard_normal(200000) * np.exp(rng.uniform(-12, 3, 200000))).astype(np.float32)
    x = x[np.abs(x) < 6.0e4]                     # fp16 range (latents are LayerNorm'd: O(1))
    hi, lo = _split_activation(x)
    err = np.abs(hi + lo - x.astype(np.float64))
    assert (err <= np.maximum(2.0 ** -22 * np.abs(x), 2.0 ** -25) * 1.0001).all()


def test_ten_layers_stay_within_the_parity_budget():
    """Residual MLP + LayerNorm stack as in the processor: the split GEMMs keep the result within
    1e-5 of float64 after ten layers, where fp16 GEMMs alone drift past it."""
    rng = np.random.default_rng(2)
    h32 = rng.standard_normal((128, 64)).astype(np.float32)
    h64, h16 = h32.astype(np.float64), h32.astype(np.float64)
    for _ in range(10):
        w1 = (rng.standard_normal((128, 128)) / np.sqrt(128)).astype(np.float32)
        w2 = (rng.standard_normal((128, 128)) / np.sqrt(128)).astype(np.float32)

        def block(h, mm):
            y = mm(w2, np.maximum(mm(w1, h), 0))
            y = (y - y.mean(axis=0)) / np.sqrt(y.var(axis=0) + 1e-5)
            return y + h

        h64 = block(h64, lambda w, x: w.astype(np.float64) @ x)
        h32 = block(h32, lambda w, x: _gemm_split(w, x.astype(np.float32))).astype(np.float32)
        h16 = block(h16, lambda w, x: w.astype(np.float16).astype(np.float64) @ x.astype(np.float16).astype(np.float64))
    assert _rel(h32, h64) <= 1e-5 / 3
    assert _rel(h16, h64) >= 1e-4
