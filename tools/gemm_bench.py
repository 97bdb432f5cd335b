#!/usr/bin/env python3
"""Micro-benchmark of the tap-GEMM kernels (CUDA events, back-to-back launches)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from flamed_tts_b200 import _lib  # noqa: E402
from flamed_tts_b200.engines import Context  # noqa: E402

ctx = Context.get("cuda:0")
lib = _lib.load_library()
reps = int(os.environ.get("REPS", 20))
cases = [  # mode, B, T, K, N, ntaps, dil, epi, out_bf16
    (1, 1, 8192, 1024, 1024, 1, 1, 1, 1),
    (1, 1, 76800, 1024, 1024, 1, 1, 1, 1),
    (1, 1, 76800, 1024, 1024, 1, 1, 5, 0),
    (1, 1, 76800, 1024, 1024, 1, 1, 0, 1),
    (1, 64, 1200, 1024, 256, 3, 1, 0, 0),
    (1, 1, 76800, 256, 1024, 1, 1, 0, 0),
    (1, 64, 6000, 512, 512, 7, 3, 0, 1),
    (1, 64, 120000, 64, 64, 7, 1, 0, 1),
    (1, 1, 350, 1024, 1024, 1, 1, 1, 1),
    (0, 1, 8192, 1024, 1024, 1, 1, 1, 0),
    (1, 1, 76800, 1024, 1024, 1, 1, 0, 0),   # 10: plain, fp32 out (direct path, 4 B/elt store)
    (1, 1, 76800, 1024, 1024, 1, 1, 4, 1),   # 11: codec-style skip, bf16 in place (transposed path, 2+2 B/elt)
    (1, 1, 76800, 1024, 1024, 1, 1, 4, 0),   # 12: skip, fp32 in place (transposed path, 4+4 B/elt)
    (1, 1, 76800, 1024, 1024, 1, 1, 2, 1),   # 13: SiLU bf16
    (1, 64, 1200, 1024, 1024, 1, 1, 5, 0),            # 14: gated residual, fp32 stream, 64 samples
    (1, 64, 1200, 1024, 1024, 1, 1, 5 | 0x100, 0),    # 15: + bf16 addend (conv_3)
    (1, 64, 1200, 1024, 1024, 1, 1, 5 | 0x200, 0),    # 16: gated residual, bf16 stream (mlp.2)
    (1, 64, 1200, 1024, 1024, 1, 1, 5 | 0x300, 0),    # 17: bf16 stream + bf16 addend (conv_3)
    (1, 64, 550, 1024, 1024, 1, 1, 5 | 0x300, 0),     # 18: same, short bucket
    (1, 64, 110000, 64, 64, 1, 1, 4, 1),              # 19: codec ResidualUnit k1 conv + skip, C=64
    (1, 64, 55000, 128, 128, 1, 1, 4, 1),             # 20: C=128
    (1, 1, 76800, 256, 1024, 1, 1, 0, 1),             # 21: proj_in shape (K=256), bf16 out: CTA-pair kernel
    (1, 1, 32768, 256, 1024, 1, 1, 0, 1),             # 22: the same at a bench-sized batch
    (1, 1, 32768, 1024, 1024, 1, 1, 0, 1),            # 23: plain K=1024 at a bench-sized batch
]
if len(sys.argv) > 1:
    cases = [cases[int(a)] for a in sys.argv[1:]]
for c in cases:
    ms = ctypes.c_float(0)
    _lib.check(lib.flm_tapgemm_bench(ctx.handle, *c, reps, ctypes.byref(ms), ctx.stream()))
    mode, B, T, K, N, ntaps, dil, epi, ob = c
    fl = 2.0 * B * T * K * N * ntaps
    print("mode=%s M=%d K=%d N=%d taps=%d epi=%d%s: %.4f ms  %.1f TFLOP/s" %
          ("bf16-tc" if mode else "fp32-fma", B * T, K, N, ntaps, epi & 0xff,
           ("+add" if epi & 0x100 else "") + ("+h16" if epi & 0x200 else "") + (" B%d" % B), ms.value, fl / ms.value / 1e9), flush=True)
