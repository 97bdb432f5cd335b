#!/usr/bin/env python3
"""Tiny / ragged shapes through the bf16 hot path (denoiser velocity + codec decode) against the oracle."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import yaml  # noqa: E402

from flamed_tts_b200 import synthetic as W  # noqa: E402
from flamed_tts_b200.engines import CodecDecoderEngine, Context, DenoiserEngine  # noqa: E402
from oracle import flamed_oracle as O  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
prior = yaml.safe_load(open(os.path.join(ROOT, "configs", "prior.yaml")))
prob = yaml.safe_load(open(os.path.join(ROOT, "configs", "prob.yaml")))
sd = W.make_flamed_state_dict(prior, prob, 0)
psd = {k[len("prob_generator."):]: v for k, v in sd.items() if k.startswith("prob_generator.")}
dsd = W.make_codec_decoder_state_dict(0)
ctx = Context.get("cuda:0")
den = DenoiserEngine(ctx, psd, prob, "bf16")
dec = CodecDecoderEngine(ctx, dsd, "bf16")
g = torch.Generator().manual_seed(9)
ok = True
for B, L in ((1, 2), (2, 7), (1, 31), (3, 33), (2, 129), (5, 64)):
    x, spk = torch.randn(B, L, 256, generator=g), torch.randn(B, 256, generator=g)
    v = den.forward(x.cuda(), 0.41, spk.cuda()).float().cpu()
    with torch.inference_mode():
        ref = O.denoiser_forward(psd, "denoiser", x, torch.full((1, 1), 0.41), spk)
        wref = O.codec_decode(dsd, x.transpose(1, 2), spk)
    w = dec.decode(x.cuda(), spk.cuda()).float().cpu()  # channels-last latents
    ev = float((v - ref).norm() / ref.norm())
    ew = float((w - wref).norm() / wref.norm())
    good = ev < 2e-2 and ew < 5e-2 and bool(torch.isfinite(v).all()) and bool(torch.isfinite(w).all())
    ok &= good
    print("B%d L%d: velocity rel-L2 %.3e  wav rel-L2 %.3e  %s" % (B, L, ev, ew, "ok" if good else "FAIL"), flush=True)
print("ALL OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
