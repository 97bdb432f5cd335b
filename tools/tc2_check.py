#!/usr/bin/env python3
"""CTA-pair tcgen05 kernel (tapgemm_tc2.cu) against a torch fp32 restatement and the first-generation kernel."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from flamed_tts_b200 import _lib  # noqa: E402
from flamed_tts_b200.engines import Context  # noqa: E402

ctx = Context.get("cuda:0")
lib = _lib.load_library()
dev = torch.device("cuda:0")


def ptr(t):
    return None if t is None else t.data_ptr()


def reference(A, W, bias, T_out, off0, dil, epi, resid, addend, gate):
    B, T_in, K = A.shape
    ntaps, N, _ = W.shape
    Af, Wf = A.float(), W.float()
    acc = torch.zeros(B, T_out, N, device=dev, dtype=torch.float32)
    for tap in range(ntaps):
        sh = off0 + tap * dil
        lo, hi = max(0, -sh), min(T_out, T_in - sh)
        if hi > lo:
            acc[:, lo:hi] += Af[:, lo + sh:hi + sh] @ Wf[tap].T
    v = acc + bias
    if epi == 1:
        v = F.gelu(v)
    elif epi == 2:
        v = F.silu(v)
    elif epi == 3:
        v = F.relu(v)
    elif epi == 4:
        v = resid.float() + v
    elif epi == 5:
        if addend is not None:
            v = v + addend.float()
        v = resid.float() + gate[:, None, :] * v
    return v


def run(gen, A, W, bias, T_out, off0, dil, epi, resid, addend, gate):
    B, T_in, K = A.shape
    ntaps, N, _ = W.shape
    out = torch.full((B, T_out, N), float("nan"), device=dev, dtype=torch.bfloat16)
    r = None if resid is None else resid.clone()
    _lib.check(lib.flm_tapgemm_test_bf16(ctx.handle, gen, ptr(A), ptr(W), ptr(bias), B, T_in, T_out, K, N, ntaps, off0, dil,
                                         epi, ptr(out), ptr(r), ptr(addend), ptr(gate), ctx.stream()))
    torch.cuda.synchronize()
    return r if epi == 5 else out


cases = [  # B, T, K, N, ntaps, dil, epi, addend
    (1, 512, 128, 128, 1, 1, 0, False),
    (1, 4096, 1024, 1024, 1, 1, 1, False),
    (8, 1200, 1024, 1024, 1, 1, 2, False),
    (3, 300, 256, 512, 7, 3, 0, False),
    (2, 1000, 64, 64, 1, 1, 4, False),
    (4, 777, 128, 128, 7, 9, 3, False),
    (5, 333, 1024, 1024, 1, 1, 5, False),
    (5, 333, 1024, 1024, 1, 1, 5, True),
    (64, 1236, 1024, 1024, 1, 1, 5, True),
    (2, 550, 1024, 256, 3, 1, 0, False),
]
if len(sys.argv) > 1:
    cases = [cases[int(a)] for a in sys.argv[1:]]
ok = True
g = torch.Generator(device=dev).manual_seed(0)
for (B, T, K, N, ntaps, dil, epi, with_add) in cases:
    A = torch.randn(B, T, K, device=dev, generator=g).bfloat16()
    W = (torch.randn(ntaps, N, K, device=dev, generator=g) * (1.0 / (K * ntaps) ** 0.5)).bfloat16()
    bias = torch.randn(N, device=dev, generator=g) * 0.1
    resid = torch.randn(B, T, N, device=dev, generator=g).bfloat16() if epi in (4, 5) else None
    addend = torch.randn(B, T, N, device=dev, generator=g).bfloat16() if with_add else None
    gate = torch.randn(B, N, device=dev, generator=g) * 0.5 if epi == 5 else None
    off0 = -(ntaps // 2) * dil
    ref = reference(A, W, bias, T, off0, dil, epi, resid, addend, gate)
    res = {}
    for gen in (1, 2):
        o = run(gen, A, W, bias, T, off0, dil, epi, resid, addend, gate).float()
        bad = int((~torch.isfinite(o)).sum())
        err = float((o - ref).norm() / ref.norm())
        mx = float((o - ref).abs().max())
        res[gen] = o
        good = bad == 0 and err < 4e-3
        ok &= good
        print("B%d T%d K%d N%d taps%d dil%d epi%d add%d gen%d: rel-L2 %.3e max-abs %.3e nonfinite %d %s" %
              (B, T, K, N, ntaps, dil, epi, with_add, gen, err, mx, bad, "ok" if good else "FAIL"), flush=True)
    d = float((res[1] - res[2]).abs().max())
    print("   gen1 vs gen2 max-abs %.3e" % d, flush=True)
print("ALL OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
