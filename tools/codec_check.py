#!/usr/bin/env python3
"""FaCodec decoder (bf16 mode): parity against the CPU oracle on a small batch and time per kernel class at a
bench-sized batch (CUDA events around every launch)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from flamed_tts_b200 import synthetic as W  # noqa: E402
from flamed_tts_b200.engines import CodecDecoderEngine, Context  # noqa: E402
from oracle import flamed_oracle as O  # noqa: E402

dsd = W.make_codec_decoder_state_dict(0)
ctx = Context.get("cuda:0")
dec = CodecDecoderEngine(ctx, dsd, "bf16")
g = torch.Generator().manual_seed(4)
lat, spk = torch.randn(2, 61, 256, generator=g), torch.randn(2, 256, generator=g)
with torch.inference_mode():
    ref = O.codec_decode(dsd, lat.transpose(1, 2), spk)
w = dec.decode(lat, spk).float().cpu()
e = float((w.double() - ref.double()).norm() / ref.double().norm())
print("decode B2 L61 bf16 vs oracle rel-L2 %.3e %s" % (e, "ok" if e < 3e-2 else "FAIL"))
ok = e < 3e-2
for B, L in ((26, 1225), (62, 520)):
    lat, spk = torch.randn(B, L, 256, generator=g).cuda(), torch.randn(B, 256, generator=g).cuda()
    for _ in range(2):
        dec.decode(lat, spk)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        dec.decode(lat, spk)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / 5 * 1000
    ctx.profile(True)
    for _ in range(3):
        dec.decode(lat, spk)
    prof = ctx.profile_read()
    ctx.profile(False)
    print("B%d L%d: %.2f ms per decode; per class (ms): %s" % (B, L, wall, {k: round(v["ms"] / 3, 2) for k, v in prof.items()}), flush=True)
print("ALL OK" if ok else "FAILED")
