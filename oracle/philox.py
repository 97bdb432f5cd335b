"""CPU restatement (numpy) of the library's counter-based noise: Philox4x32-10 + Box-Muller, the documented
seed -> tensor map of include/flamed_b200.h (flm_durgen_sample / flm_denoiser_sample with NULL noise pointers).

TEST INFRASTRUCTURE ONLY.  Philox4x32-10 is the published generator of Salmon et al., "Parallel Random Numbers: As
Easy as 1, 2, 3" (SC'11, Random123); pinned here against the known-answer vectors of Random123's kat_vectors file
(tests/test_oracle_golden.py::test_philox_known_answers).  The reference itself draws with torch.randn on the CPU
(pva.py:101-102, prob_generator.py:440); this map is the B200 build's own device-side replacement.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr, key):
    """ctr: (..., 4) uint32, key: (2,) uint32 -> (..., 4) uint32"""
    c = [np.asarray(ctr[..., i], dtype=np.uint64) for i in range(4)]
    k0, k1 = int(key[0]), int(key[1])
    for _ in range(10):
        p0 = M0 * c[0]
        p1 = M1 * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c = [hi1 ^ c[1] ^ np.uint64(k0), lo1, hi0 ^ c[3] ^ np.uint64(k1), lo0]
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return np.stack(c, axis=-1).astype(np.uint32)


def normal(seed, tensor_id, n):
    """the n standard-normal values the library draws for (seed, tensor_id): float64 math, returned as float32"""
    groups = (n + 3) // 4
    g = np.arange(groups, dtype=np.uint64)
    ctr = np.stack([g & MASK, g >> np.uint64(32), np.full(groups, tensor_id, np.uint64), np.zeros(groups, np.uint64)],
                   axis=-1).astype(np.uint32)
    r = philox4x32_10(ctr, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))
    u = ((r >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -24)  # exact in fp32
    u = u.astype(np.float64)
    rad0, rad1 = np.sqrt(-2.0 * np.log(u[:, 0])), np.sqrt(-2.0 * np.log(u[:, 2]))
    z = np.stack([rad0 * np.cos(2 * np.pi * u[:, 1]), rad0 * np.sin(2 * np.pi * u[:, 1]),
                  rad1 * np.cos(2 * np.pi * u[:, 3]), rad1 * np.sin(2 * np.pi * u[:, 3])], axis=-1)
    return z.reshape(-1)[:n].astype(np.float32)
