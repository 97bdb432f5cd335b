#!/usr/bin/env python3
"""Attention of the prior decoders at bench shapes: the library's own kernel (flm_attention_bf16) against the
flash-attn library kernel it replaced (if installed) and torch SDPA with an additive mask."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from flamed_tts_b200.engines import Context, attention_bf16  # noqa: E402

ctx = Context.get("cuda:0")


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for B, S in ((26, 1465), (62, 760), (64, 1476), (13, 2640)):
    H, dh = 12, 32
    qkv = torch.randn(B, S, 3, H, dh, device="cuda", dtype=torch.bfloat16)
    lens = torch.randint(S * 3 // 4, S + 1, (B,), device="cuda", dtype=torch.int32)
    lens[0] = S
    own = timeit(lambda: attention_bf16(ctx, qkv, lens))
    fl = 4.0 * float((lens.double() * S).sum()) * H * dh  # useful FLOPs (valid keys only)
    line = "B=%d S=%d: own %.3f ms (%.0f TFLOP/s useful)" % (B, S, own, fl / own / 1e9)
    try:
        from flash_attn import flash_attn_with_kvcache
        fa = timeit(lambda: flash_attn_with_kvcache(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], cache_seqlens=lens, causal=False))
        line += "  flash-attn %.3f ms" % fa
    except Exception as e:  # noqa: BLE001
        line += "  flash-attn unavailable (%s)" % type(e).__name__
    print(line, flush=True)
