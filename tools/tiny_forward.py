#!/usr/bin/env python3
"""a few tiny denoiser evaluations + an 8-step loop + a codec decode in bf16 (for compute-sanitizer runs)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import yaml  # noqa: E402

from flamed_tts_b200 import synthetic as W  # noqa: E402
from flamed_tts_b200.engines import CodecDecoderEngine, Context, DenoiserEngine  # noqa: E402

prior = yaml.safe_load(open(os.path.join(ROOT, "configs", "prior.yaml")))
prob = yaml.safe_load(open(os.path.join(ROOT, "configs", "prob.yaml")))
sd = W.make_flamed_state_dict(prior, prob, 0)
psd = {k[len("prob_generator."):]: v for k, v in sd.items() if k.startswith("prob_generator.")}
ctx = Context.get("cuda:0")
den = DenoiserEngine(ctx, psd, prob, "bf16")
g = torch.Generator().manual_seed(1)
for B, L in ((1, 7), (3, 97), (2, 300)):
    x, spk = torch.randn(B, L, 256, generator=g).cuda(), torch.randn(B, 256, generator=g).cuda()
    v = den.forward(x, 0.5, spk)
    torch.cuda.synchronize()
    print("forward", B, L, float(v.float().abs().mean()))
B, L, nfe = 2, 40, 4
cond, spk, noise = torch.relu(torch.randn(B, L, 256, generator=g)), torch.randn(B, 256, generator=g), torch.randn(B, L, 256, generator=g)
out = den.sample(cond, spk, noise, torch.linspace(0, 1, nfe + 1), 0.3, use_graph=False)
torch.cuda.synchronize()
print("sample", float(out.float().abs().mean()))
if os.environ.get("CODEC", "1") == "1":
    dec = CodecDecoderEngine(ctx, W.make_codec_decoder_state_dict(0), "bf16")
    wav = dec.decode(torch.randn(1, 12, 256, generator=g).cuda(), torch.randn(1, 256, generator=g).cuda())
    torch.cuda.synchronize()
    print("decode", tuple(wav.shape))
print("ok")
