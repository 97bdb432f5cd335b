#!/usr/bin/env python3
"""One denoiser velocity evaluation + one codec decode at a representative size (for ncu captures)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import yaml  # noqa: E402

from flamed_tts_b200 import synthetic as W  # noqa: E402
from flamed_tts_b200.engines import CodecDecoderEngine, Context, DenoiserEngine  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B, L = int(os.environ.get("PB", 32)), int(os.environ.get("PL", 1200))
prior = yaml.safe_load(open(os.path.join(ROOT, "configs", "prior.yaml")))
prob = yaml.safe_load(open(os.path.join(ROOT, "configs", "prob.yaml")))
ctx = Context.get("cuda:0")
sd = W.make_flamed_state_dict(prior, prob, 0)
den = DenoiserEngine(ctx, {k[len("prob_generator."):]: v for k, v in sd.items() if k.startswith("prob_generator.")}, prob, "bf16")
x = torch.randn(B, L, 256, device="cuda")
spk = torch.randn(B, 256, device="cuda")
for _ in range(2):
    v = den.forward(x, 0.5, spk)
torch.cuda.synchronize()
if os.environ.get("PROF", "1") == "1":
    ctx.profile(True)
    for _ in range(3):
        v = den.forward(x, 0.5, spk)
    for name, r in ctx.profile_read().items():
        ms = r["ms"] / 3
        print("%-20s launches/fwd %3d  %8.3f ms/fwd  %8.1f GB/s  %8.1f TFLOP/s" %
              (name, r["launches"] // 3, ms, r["bytes"] / 3 / ms / 1e6, r["flops"] / 3 / ms / 1e9))
    ctx.profile(False)
if os.environ.get("CODEC", "1") == "1":
    dec = CodecDecoderEngine(ctx, W.make_codec_decoder_state_dict(0), "bf16")
    Lc = int(os.environ.get("PLC", 300))
    w = dec.decode(torch.randn(B, Lc, 256, device="cuda"), spk)
    torch.cuda.synchronize()
print("ok", v.shape)
