#!/usr/bin/env python
"""Fits the MUFU-free polynomial GELU / GELU' used by the bf16 fused-MLP epilogues (csrc/gelu_poly.cuh).

    Phi(x)   ~= sat( 0.5 + xc * P(xc^2) ),   xc = clamp(x, -A, A)        gelu(x)  = x * Phi(x)
    gelu'(x) ~=      0.5 + xc * Q(xc^2),     xc = clamp(x, -B, B)

Weighted least squares on Chebyshev nodes with a few re-weighting rounds (near-minimax).  Prints the monomial
coefficients in t = x^2 (lowest degree first) and the fp32 evaluation error against the erf forms."""
import numpy as np
from numpy.polynomial import chebyshev as Ch, polynomial as P
from scipy.special import erf


def fit_odd(f, A, deg_t, n=40001):
    x = np.linspace(1e-6, A, n)
    t = x * x
    y = f(x) / x
    u = 2 * t / (A * A) - 1
    w = np.ones_like(x)
    for _ in range(40):
        c = Ch.chebfit(u, y, deg_t, w=w * x)
        err = (Ch.chebval(u, c) - y) * x
        w = w * (1 + 5 * np.abs(err) / np.abs(err).max())
    p = Ch.cheb2poly(c)
    a, b = 2 / (A * A), -1.0
    mono = np.zeros(1)
    for k, ck in enumerate(p):
        mono = P.polyadd(mono, ck * P.polypow([b, a], k))
    return mono, np.abs(err).max()


def horner32(m, t):
    acc = np.float32(m[-1]) * np.ones_like(t)
    for ck in m[-2::-1]:
        acc = acc * t + np.float32(ck)
    return acc


Phi0 = lambda x: 0.5 * erf(x / np.sqrt(2))                                          # Phi - 0.5
dG0 = lambda x: 0.5 * erf(x / np.sqrt(2)) + x * np.exp(-x * x / 2) / np.sqrt(2 * np.pi)  # gelu' - 0.5

A, B = 4.0, 4.5
mP, eP = fit_odd(Phi0, A, 7)
mP = mP * (1 + 2.5 * eP)        # overshoot so that the saturation makes Phi(+-A) exactly 1 / 0
mQ, eQ = fit_odd(dG0, B, 9)
x = np.linspace(-8, 8, 400001).astype(np.float32)
xc = np.clip(x, -A, A)
phi = np.clip(np.float32(0.5) + xc * horner32(mP, xc * xc), 0, 1)
g = x * phi
ref = x.astype(np.float64) * (0.5 + Phi0(x.astype(np.float64)))
print("gelu : max abs err %.3e   max |Phi err| %.3e" % (np.abs(g - ref).max(), np.abs(phi - (0.5 + Phi0(x.astype(np.float64)))).max()))
xc = np.clip(x, -B, B)
d = np.float32(0.5) + xc * horner32(mQ, xc * xc)
print("gelu': max abs err %.3e" % np.abs(d - (0.5 + dG0(x.astype(np.float64)))).max())
print("P:", ", ".join("%.9ef" % c for c in mP))
print("Q:", ", ".join("%.9ef" % c for c in mQ))

# ---- tanh form of the forward: Phi(x) = 1/2 + 1/2 tanh(x (c0 + c1 x^2 + c2 x^4)), the polynomial fitted to atanh(erf) ----
from scipy.optimize import least_squares

xs = np.linspace(-6, 6, 24001)
PhiT = 0.5 * (1 + erf(xs / np.sqrt(2)))


def tanh_model(c, x):
    t = np.minimum(x * x, 80.0)
    return 0.5 * (1 + np.tanh(x * (c[0] + t * (c[1] + t * c[2]))))


c = np.array([0.79788456, 0.79788456 * 0.044715, 0.0])
w = np.ones_like(xs)
for _ in range(30):
    c = least_squares(lambda c: (tanh_model(c, xs) - PhiT) * np.maximum(np.abs(xs), 0.3) * w, c, xtol=1e-15, ftol=1e-15).x
    e = np.abs((tanh_model(c, xs) - PhiT) * np.maximum(np.abs(xs), 0.3))
    w = w * (1 + 3 * e / e.max())
xw = np.linspace(-40, 40, 80001)
print("tanh form: c =", ", ".join("%.9ef" % v for v in c), " max |Phi err| %.2e (|x| <= 40, exact tanh)" %
      np.abs(tanh_model(c, xw) - 0.5 * (1 + erf(xw / np.sqrt(2)))).max())
