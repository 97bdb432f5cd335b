"""CPU checks (fp64, no GPU) of the algebra behind three fused kernels - what the CUDA code computes, restated in numpy /
torch and compared with the direct form of the reference op.

  * TapGemm::lnu_*  : conv_3's epilogue recomputes the ConvNeXt inner residual u from the residual-stream slab
                      (prob_generator.py:107-111, 136, 162)
  * TapGemm::raff_* : the MLP-branch LayerNorm + modulate applied algebraically around mlp.0
                      (prob_generator.py:146-149, 163-164)
  * dwconv_tc.cu    : depthwise conv k=31 as Hankel (time series) x Toeplitz (taps) matrix products on a virtual position
                      axis, with the exact byte offsets the UMMA descriptors encode (prob_generator.py:81-88)
"""
import numpy as np
import torch


def test_inner_residual_recomputed_in_the_epilogue():
    g = torch.Generator().manual_seed(0)
    R, C = 37, 64
    h = torch.randn(R, C, generator=g, dtype=torch.float64) * 3 + 0.7
    acc, bias = torch.randn(R, C, generator=g, dtype=torch.float64), torch.randn(C, generator=g, dtype=torch.float64)
    w, b = torch.randn(C, generator=g, dtype=torch.float64), torch.randn(C, generator=g, dtype=torch.float64)
    shift, scale, gate = (torch.randn(C, generator=g, dtype=torch.float64) for _ in range(3))
    u = torch.nn.functional.layer_norm(h, (C,), w, b, 1e-6) * (1 + scale) + shift       # modulate(ln_conv(h))
    direct = h + gate * (u + (acc + bias))                                             # x + gate * (u + f(u))
    mean, var = h.mean(1, keepdim=True), h.var(1, unbiased=False, keepdim=True)
    rstd = (var + 1e-6).rsqrt()
    nm = -mean * rstd
    A, B = w * (1 + scale), b * (1 + scale) + shift                                    # the depthwise kernel's affine
    G, GA, GB = gate, gate * A, gate * (bias + B)                                      # its by-product table
    fused = h + G * acc + GA * (h * rstd + nm) + GB
    assert torch.allclose(fused, direct, rtol=1e-12, atol=1e-12)


def test_mlp_layernorm_applied_algebraically():
    g = torch.Generator().manual_seed(1)
    R, C, N = 29, 48, 40
    h = torch.randn(R, C, generator=g, dtype=torch.float64) * 2 - 0.4
    w, b = torch.randn(C, generator=g, dtype=torch.float64), torch.randn(C, generator=g, dtype=torch.float64)
    shift, scale = torch.randn(C, generator=g, dtype=torch.float64), torch.randn(C, generator=g, dtype=torch.float64)
    W0, b0 = torch.randn(N, C, generator=g, dtype=torch.float64), torch.randn(N, generator=g, dtype=torch.float64)
    u = torch.nn.functional.layer_norm(h, (C,), w, b, 1e-6) * (1 + scale) + shift
    direct = torch.nn.functional.silu(u @ W0.T + b0)
    A2, B2 = w * (1 + scale), b * (1 + scale) + shift
    c1, c2 = W0 @ A2, W0 @ B2 + b0                                                     # hoisted per (step, sample)
    s, q = h.sum(1, keepdim=True), (h * h).sum(1, keepdim=True)                        # what `rowstat` carries
    mean = s / C
    rstd = (q / C - mean * mean + 1e-6).rsqrt()
    acc = (A2 * h) @ W0.T                                                              # mlp.0 on conv_3's second output
    fused = torch.nn.functional.silu(rstd * acc + (-mean * rstd) * c1 + c2)
    assert torch.allclose(fused, direct, rtol=1e-10, atol=1e-10)


# ---- dwconv_tc.cu: constants and address arithmetic mirrored from the kernel
KW, PADW, NPOS, OWN, TS, XALLOC = 31, 15, 1024, 124, 992, 1064


def _geo(B, L):
    V = (L + 2 * PADW + 31) & ~31
    total = B * V
    return V, total, (total + TS - 1) // TS


def _chunk_of(pos):
    pt = pos // TS
    return 4 * pt + min((pos - pt * TS) >> 8, 3)


def _run_tile_model(B, L, C, seed):
    rng = np.random.default_rng(seed)
    u = rng.standard_normal((B, L, C))
    w = rng.standard_normal((KW, C)) * 0.2
    V, total, npt = _geo(B, L)
    up = np.zeros((B, L + 2 * PADW, C))
    up[:, PADW:PADW + L] = u
    ref = sum(up[:, k:k + L] * w[k] for k in range(KW))
    got = np.full((B, L, C), np.nan)
    pieces = {}
    m_idx, n_idx, k_idx = np.arange(128)[:, None, None], np.arange(16)[None, :, None], np.arange(16)[None, None, :]
    for pt in range(npt):
        P0 = pt * TS
        for ch in range(C):
            xs = np.zeros(XALLOC)                       # channel series (element units = 2 bytes); tail [1024, 1064) stays 0
            for pos in range(NPOS):
                p = P0 + pos
                b, f = p // V, p % V - PADW
                xs[pos] = u[b, f, ch] if b < B and 0 <= f < L else 0.0
            D = np.zeros((128, 16))
            for r in range(3):
                blk = np.zeros(256)                     # one Toeplitz block (128 elements) followed by the zero block
                for khalf in range(2):
                    for n in range(8):
                        for kk in range(8):
                            tap = 16 * r + khalf * 8 + kk - n
                            blk[khalf * 64 + n * 8 + kk] = w[tap, ch] if 0 <= tap < KW else 0.0
                # A descriptor: start + 32 r bytes, rows 16 B apart inside an 8-row group, groups SBO = 128 B, k halves LBO = 16 B
                a_el = 16 * r + 64 * (m_idx // 8) + 8 * (m_idx % 8) + 8 * (k_idx // 8) + (k_idx % 8)
                # B descriptor: n rows 16 B apart, k halves LBO = 128 B, the second n group SBO away = the zero block
                b_el = 128 * (n_idx // 8) + 8 * (n_idx % 8) + 64 * (k_idx // 8) + (k_idx % 8)
                D += (xs[a_el] * blk[b_el]).sum(-1)
            for m in range(OWN):
                pp = P0 + 8 * m
                if pp >= total:
                    break
                b, j0 = pp // V, pp % V
                nval = min(max(L - j0, 0), 8)
                if nval:
                    assert np.isnan(got[b, j0:j0 + nval, ch]).all()
                    got[b, j0:j0 + nval, ch] = D[m, :nval]
                if ch == 0:
                    key = (b, 4 * pt + m // 32 - _chunk_of(b * V))
                    pieces[key] = pieces.get(key, 0) + nval
    assert not np.isnan(got).any()
    assert np.abs(got - ref).max() < 1e-12
    for b in range(B):                                  # the statistics pieces of a sample cover its frames exactly once
        assert sum(n for (bb, k), n in pieces.items() if bb == b) == L
        assert all(0 <= k < V // 224 + 3 for (bb, k) in pieces if bb == b)


def test_hankel_toeplitz_depthwise_model():
    for B, L, C, seed in ((3, 100, 2, 0), (2, 7, 1, 1), (5, 190, 1, 2), (2, 1000, 1, 3)):
        _run_tile_model(B, L, C, seed)
