"""CPU: the algebra behind the LayerNorm folded into the fused chain's GEMMs (dit_chain.cu), in numpy.
  (LN(x)(1 + s) + t) W^T + b  ==  (rstd / c) * ((c x (1 + s)) W^T)  -  rstd * mean * u  +  v,   u = W (1 + s),  v = W t + b
for ANY per-row scale c > 0, and the same with the operand rounded to fp16 / e4m3-like precision to the expected error."""
import numpy as np


def _case(rng, rows=64, D=256, N=96, outlier=False):
    x = rng.standard_normal((rows, D)) * rng.uniform(0.5, 30.0, size=(rows, 1)) + rng.uniform(-3, 3, size=(rows, 1))
    if outlier:
        x[:, 7] *= 80.0                                # a "massive activation" channel
    s, t = 0.3 * rng.standard_normal(D), 0.2 * rng.standard_normal(D)
    W, b = rng.standard_normal((N, D)) / np.sqrt(D), 0.1 * rng.standard_normal(N)
    return x, s, t, W, b


def _reference(x, s, t, W, b):
    mean = x.mean(-1, keepdims=True)
    var = (x * x).mean(-1, keepdims=True) - mean * mean
    rstd = 1.0 / np.sqrt(np.maximum(var, 0.0) + 1e-6)
    n = (x - mean) * rstd * (1.0 + s) + t
    return n @ W.T + b, mean, rstd


def test_folded_form_equals_layernorm_then_linear_for_any_row_scale():
    rng = np.random.default_rng(0)
    x, s, t, W, b = _case(rng)
    want, mean, rstd = _reference(x, s, t, W, b)
    u, v = W @ (1.0 + s), W @ t + b
    for c in (np.ones((x.shape[0], 1)), rstd, rstd * rng.uniform(0.3, 3.0, size=rstd.shape), 8.0 * rstd):
        a = c * x * (1.0 + s)                                   # what the producing epilogue writes
        got = (rstd / c) * (a @ W.T) - (rstd * mean) * u + v     # what the consuming epilogue computes
        np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-9)


def test_stale_row_scale_keeps_the_16_bit_operand_in_range_and_accurate():
    """c = 1 / std of the row at the PREVIOUS LayerNorm: x has since moved by a residual update, so c is off by tens of percent --
    still the operand stays O(|1 + s| * sqrt(D)) and fp16 rounding of it costs what rounding the normalised rows costs."""
    rng = np.random.default_rng(1)
    x, s, t, W, b = _case(rng, outlier=True)
    want, mean, rstd = _reference(x, s, t, W, b)
    u, v = W @ (1.0 + s), W @ t + b
    x_prev = x + 0.3 * rng.standard_normal(x.shape) * x.std(-1, keepdims=True)
    c = 1.0 / x_prev.std(-1, keepdims=True)
    a = c * x * (1.0 + s)
    assert np.abs(a).max() < 65504.0 / 100                      # far from the fp16 limit even with an 80x outlier channel
    a16 = a.astype(np.float16).astype(np.float64)
    got = (rstd / c) * (a16 @ W.T) - (rstd * mean) * u + v
    n16 = (((x - mean) * rstd * (1.0 + s) + t).astype(np.float16).astype(np.float64)) @ W.T + b      # rounding the normalised rows instead
    err_fold = np.abs(got - want).max()
    err_norm = np.abs(n16 - want).max()
    assert err_fold < 4.0 * err_norm + 1e-6                     # same error class
    assert err_fold < 2e-2 * np.abs(want).max()


def test_weight_scale_folds_into_gate_and_bias_of_ff2():
    """e4m3 level 2: gate * ((h * G) (W / sw)^T * sw / G + b)  ==  (gate * sw / G) * ((h G)(W / sw)^T + b * G / sw)."""
    rng = np.random.default_rng(2)
    h, W, b = np.abs(rng.standard_normal((32, 128))), rng.standard_normal((48, 128)), rng.standard_normal(48)
    gate = rng.standard_normal(48)
    sw, G = np.abs(W).max(-1) / 448.0, 16.0
    acc = (h * G) @ (W / sw[:, None]).T
    want = gate * (h @ W.T + b)
    got = (gate * sw / G) * (acc + b * G / sw)
    np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-10)
