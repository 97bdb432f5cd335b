#!/usr/bin/env python3
"""Attention of the prior decoders (B=64, 12 heads x 32, S = L + 240) on the library back ends available here."""
import sys
import torch
import torch.nn.functional as F
from torch.nn.attention import SDPBackend, sdpa_kernel

dev = "cuda"
B, H, D = 64, 12, 32
for S in (1476, 890, 590):
    qkv = torch.randn(B, S, 3, H, D, device=dev, dtype=torch.bfloat16)
    q, k, v = (qkv[:, :, i].transpose(1, 2) for i in range(3))
    lens = torch.randint(S - 40, S + 1, (B,), device=dev)
    pad = torch.arange(S, device=dev)[None, :] >= lens[:, None]
    bias = torch.zeros(B, 1, 1, S, device=dev, dtype=torch.bfloat16).masked_fill_(pad[:, None, None, :], float("-inf"))

    def timeit(fn, n=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    fl = 4.0 * B * H * S * S * D
    res = {}
    for name, be in (("cudnn", SDPBackend.CUDNN_ATTENTION), ("flash", SDPBackend.FLASH_ATTENTION),
                     ("efficient", SDPBackend.EFFICIENT_ATTENTION)):
        for masked in (True, False):
            try:
                with sdpa_kernel(be):
                    ms = timeit(lambda: F.scaled_dot_product_attention(q, k, v, attn_mask=bias if masked else None))
                res["%s%s" % (name, "+mask" if masked else "")] = ms
            except Exception as e:  # noqa: BLE001
                res["%s%s" % (name, "+mask" if masked else "")] = "n/a (%s)" % str(e)[:40]
    try:
        from flash_attn import flash_attn_func, flash_attn_varlen_func
        qq, kk, vv = (qkv[:, :, i].contiguous() for i in range(3))
        res["flash_attn_func"] = timeit(lambda: flash_attn_func(qq, kk, vv))
        cu = torch.zeros(B + 1, device=dev, dtype=torch.int32)
        cu[1:] = torch.cumsum(lens, 0)
        tot = int(cu[-1])
        qv, kv, vv2 = (torch.randn(tot, H, D, device=dev, dtype=torch.bfloat16) for _ in range(3))
        res["flash_attn_varlen"] = timeit(lambda: flash_attn_varlen_func(qv, kv, vv2, cu, cu, S, S))
    except Exception as e:  # noqa: BLE001
        res["flash_attn"] = "n/a (%s)" % str(e)[:60]
    print("S=%d  (%.0f GFLOP)" % (S, fl / 1e9))
    for k2, v2 in res.items():
        print("   %-22s %s" % (k2, ("%.3f ms  %.0f TFLOP/s" % (v2, fl / v2 / 1e9)) if isinstance(v2, float) else v2))
