#!/usr/bin/env python3
"""numpy model of Activation1d as two banded-Toeplitz GEMMs + a pointwise SnakeBeta,
checked against the closed form of SURVEY Appendix A1.  The 12-tap filters are the same for every channel, so
  u = F_up (2T x T) @ x,   s = u + sin^2(a u) / (b + 1e-9),   y = F_dn (T x 2T) @ s
with replicate padding folded into the edge rows of F_up / F_dn (only the first / last few rows differ from the band)."""
import numpy as np

rng = np.random.default_rng(1)
T, C = 300, 8
x = rng.standard_normal((T, C))
alpha, beta = rng.standard_normal(C) * 0.3, rng.standard_normal(C) * 0.3
a, b = np.exp(alpha), np.exp(beta)
# kaiser_sinc_filter1d(cutoff 0.25, half_width 0.3, 12) taps, symmetric
half = np.array([0.0020290, 0.0093895, -0.0255435, -0.0576574, 0.1285726, 0.4432098])
f = np.concatenate([half, half[::-1]])
cl = lambda i, n: np.clip(i, 0, n - 1)

# closed form (Appendix A1)
u = np.zeros((2 * T, C))
n = np.arange(T)
for j in range(6):
    u[0::2] += 2 * x[cl(n - 3 + j, T)] * f[11 - 2 * j]
    u[1::2] += 2 * x[cl(n - 2 + j, T)] * f[10 - 2 * j]
s = u + np.sin(a * u) ** 2 / (b + 1e-9)
y_ref = np.zeros((T, C))
for k in range(12):
    y_ref += s[cl(2 * n + k - 5, 2 * T)] * f[k]

# Toeplitz form
F_up = np.zeros((2 * T, T))
for t in range(T):
    for j in range(6):
        F_up[2 * t, cl(t - 3 + j, T)] += 2 * f[11 - 2 * j]
        F_up[2 * t + 1, cl(t - 2 + j, T)] += 2 * f[10 - 2 * j]
F_dn = np.zeros((T, 2 * T))
for t in range(T):
    for k in range(12):
        F_dn[t, cl(2 * t + k - 5, 2 * T)] += f[k]
u2 = F_up @ x
s2 = u2 + np.sin(a * u2) ** 2 / (b + 1e-9)
y2 = F_dn @ s2
print("Toeplitz form vs closed form: max |err| %.2e" % np.abs(y2 - y_ref).max())
# band structure: interior rows of F_up touch 6 inputs, of F_dn 12; only rows within 5 / 6 samples of an edge differ
bw_up = max(np.ptp(np.flatnonzero(F_up[r])) + 1 for r in range(12, 2 * T - 12))
bw_dn = max(np.ptp(np.flatnonzero(F_dn[r])) + 1 for r in range(12, T - 12))
print("interior band widths: up %d inputs per up-sampled sample, down %d" % (bw_up, bw_dn))
# a 128-output tile needs up-sampled samples 2*t0-5 .. 2*t0+261 (267) and inputs t0-5 .. t0+133 (139)
assert np.abs(y2 - y_ref).max() < 1e-12
print("OK")
