#!/usr/bin/env python3
"""numpy model of the polyphase tensor-core depthwise conv planned in DESIGN.md section 10 (item 1): checks the index
algebra (phase windows, row shifts, block-diagonal weight blocks) against the direct convolution before any kernel
is written.  One MMA of the model = D[u, n] += sum_k A_q[u, k] * B_q[n, k] with M = 128 rows u (frames P*u + s),
N = 16*P columns n = s*16 + c, K = 16 input channels of one group."""
import numpy as np

KW, PAD, P, G = 31, 15, 4, 16
M = 128
rng = np.random.default_rng(0)
L, C = 777, 32                       # ragged length, two 16-channel groups
x = rng.standard_normal((L, C)).astype(np.float32)
w = rng.standard_normal((KW, C)).astype(np.float32)   # tap-major, as the library packs it
ref = np.zeros((L, C), np.float32)
for j in range(KW):
    lo, hi = max(0, PAD - j), min(L, L + PAD - j)
    ref[lo:hi] += x[lo + j - PAD:hi + j - PAD] * w[j]

T_TILE = M * P                       # 512 output frames per tile
out = np.zeros_like(ref)
n_mma = 0
for T0 in range(0, L, T_TILE):
    rows = M + (KW - 1 + P - 1) // P                      # 136 rows per phase window
    for g in range(C // G):
        cs = slice(g * G, (g + 1) * G)
        # phase windows: X_pi[u', c'] = x[T0 - PAD + P*u' + pi, c'] (zero outside [0, L): the TMA OOB fill)
        Xp = np.zeros((P, rows, G), np.float32)
        for pi in range(P):
            f = T0 - PAD + P * np.arange(rows) + pi
            ok = (f >= 0) & (f < L)
            Xp[pi, ok] = x[f[ok], cs]
        D = np.zeros((M, P * G), np.float32)              # accumulator: lane u, column s*16 + c
        for q in range(KW - 1 + P):                       # input shift q = s + j
            pi, h = q % P, q // P
            A = Xp[pi, h:h + M]                           # same buffer, descriptor start advanced by h rows
            B = np.zeros((P * G, G), np.float32)          # B_q[(s, c), c'] = delta(c, c') * w[q - s][c]
            for s in range(P):
                j = q - s
                if 0 <= j < KW:
                    B[s * G + np.arange(G), np.arange(G)] = w[j, cs]
            D += A @ B.T
            n_mma += 1
        for s in range(P):                                # epilogue: frame T0 + P*u + s
            t = T0 + P * np.arange(M) + s
            ok = t < L
            out[t[ok], cs] = D[ok, s * G:(s + 1) * G]
err = np.abs(out - ref).max()
print("polyphase P=%d: %d MMAs of %dx%dx%d for L=%d C=%d, max |err| vs direct conv %.2e" % (P, n_mma, M, P * G, G, L, C, err))
assert err < 1e-4
print("OK")
