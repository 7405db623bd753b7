"""CPU guard for the constants of the fused loss kernel's small-logit form (dh_fused_loss_kernel.cuh): for a label-0
logit x <= -0.7 the kernel evaluates sigmoid(x)^2 softplus(x) / ln 2 as e^3 P(e) and its derivative as e^3 Q(e), e = exp(x).
The coefficients are read from the header and evaluated here in float32 Horner form, the way the kernel does, against the
float64 functions: the 1e-5 parity bar leaves no room for a mistyped digit."""
import os
import re

import numpy as np

from conftest import PKG

HDR = os.path.join(PKG, "csrc", "dh_fused_loss_kernel.cuh")


def _coefficients(func):
    with open(HDR) as f:
        src = f.read()
    body = src[src.index("void %s(" % func):]
    body = body[:body.index("\n}\n")]
    return body


def _horner32(coeffs_high_to_low, e):
    p = np.full_like(e, np.float32(coeffs_high_to_low[0]))
    for c in coeffs_high_to_low[1:]:
        p = (p.astype(np.float64) * e.astype(np.float64) + np.float64(np.float32(c))).astype(np.float32)  # one rounding, like fma
    return p


def _pairs(body, var):
    """pack2(c, c) constants of the Horner chain of `var`, highest degree first."""
    first = re.search(r"f32x2 %s = fma2\(pack2\(([-0-9.e]+)f, [-0-9.e]+f\), e, pack2\(([-0-9.e]+)f" % var, body)
    rest = re.findall(r"\b%s = fma2\(%s, e, pack2\(([-0-9.e]+)f" % (var, var), body)
    return [float(first.group(1)), float(first.group(2))] + [float(v) for v in rest]


def test_switch_point_and_range():
    with open(HDR) as f:
        src = f.read()
    cut = float(re.search(r"constexpr float kSmallLogit = ([-0-9.]+)f;", src).group(1))
    assert cut == -0.7 and np.exp(cut) < 0.5          # the fits below cover e in [0, 0.5]


def test_value_polynomial():
    p = _pairs(_coefficients("stream_pair_g2_small"), "p")
    assert len(p) == 8
    e = np.linspace(0.0, 0.5, 200001).astype(np.float32)
    got = _horner32(p, e).astype(np.float64)
    e64 = e.astype(np.float64)
    want = np.ones_like(e64) / np.log(2.0)
    m = e64 > 1e-8
    want[m] = np.log1p(e64[m]) / (e64[m] * (1.0 + e64[m]) ** 2) / np.log(2.0)
    assert np.max(np.abs(got - want) / want) < 1.0e-6
    # the whole term at a few logits, against the textbook form
    x = np.array([-0.7, -1.0, -2.0, -4.595, -10.0, -30.0])
    ex = np.exp(x)
    s = ex / (1.0 + ex)
    term = s * s * np.log1p(ex) / np.log(2.0)
    approx = ex ** 3 * _horner32(p, ex.astype(np.float32)).astype(np.float64)
    assert np.allclose(approx, term, rtol=2e-6, atol=0)


def test_gradient_polynomial_and_shared_value_chain():
    body = _coefficients("stream_pair_g2_small_grad")
    p, q = _pairs(body, "p"), _pairs(body, "q")
    assert p == _pairs(_coefficients("stream_pair_g2_small"), "p"), "forward and gradient kernels must share the value polynomial"
    assert len(q) == 8
    e = np.linspace(0.0, 0.5, 200001).astype(np.float32)
    got = _horner32(q, e).astype(np.float64)
    e64 = e.astype(np.float64)
    want = np.full_like(e64, 3.0)
    m = e64 > 1e-8
    s = e64[m] / (1.0 + e64[m])
    want[m] = s * s * (2.0 * (1.0 - s) * np.log1p(e64[m]) + s) / e64[m] ** 3   # d/dx [sigmoid^2 softplus] / e^3
    assert np.max(np.abs(got - want) / want) < 3.0e-6
    # derivative of the textbook form by central differences in float64
    x = np.array([-0.7, -1.5, -3.0, -4.595, -8.0])
    f = lambda t: (1.0 / (1.0 + np.exp(-t))) ** 2 * np.logaddexp(0.0, t)
    h = 1e-5
    num = (f(x + h) - f(x - h)) / (2 * h)
    ex = np.exp(x)
    approx = ex ** 3 * _horner32(q, ex.astype(np.float32)).astype(np.float64)
    assert np.allclose(approx, num, rtol=1e-5, atol=0)
