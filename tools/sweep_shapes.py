#!/usr/bin/env python3
"""Shape sweep of the engines (for compute-sanitizer memcheck): every L in a range, small B."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import yaml  # noqa: E402

from flamed_tts_b200 import synthetic as W  # noqa: E402
from flamed_tts_b200.engines import CodecDecoderEngine, Context, DenoiserEngine, DurationEngine  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
prior = yaml.safe_load(open(os.path.join(ROOT, "configs", "prior.yaml")))
prob = yaml.safe_load(open(os.path.join(ROOT, "configs", "prob.yaml")))
ctx = Context.get("cuda:0")
sd = W.make_flamed_state_dict(prior, prob, 0)
prec = os.environ.get("PREC", "bf16")
den = DenoiserEngine(ctx, {k[len("prob_generator."):]: v for k, v in sd.items() if k.startswith("prob_generator.")}, prob, prec)
dec = CodecDecoderEngine(ctx, W.make_codec_decoder_state_dict(0), prec)
dur = DurationEngine(ctx, {k[len("prior_generator.pva."):]: v for k, v in sd.items() if k.startswith("prior_generator.pva.")})
lo, hi, step = (int(x) for x in os.environ.get("LRANGE", "1,200,1").split(","))
B = int(os.environ.get("PB", 2))
g = torch.Generator(device="cuda").manual_seed(0)
for L in range(lo, hi, step):
    x = torch.randn(B, L, 256, device="cuda", generator=g)
    spk = torch.randn(B, 256, device="cuda", generator=g)
    prior_embs = torch.randn(B, 6, L, 384, device="cuda", generator=g)
    mask = torch.ones(B, L, dtype=torch.bool, device="cuda")
    c = den.cond_prepare(prior_embs, mask)
    ts = torch.linspace(0, 1, 3)
    lat = den.sample(c, spk, x, ts, 0.3, use_graph=False)
    w = dec.decode(lat, spk)
    P = max(1, L // 7)
    enc = torch.randn(B, P, 192, device="cuda", generator=g)
    m = torch.zeros(B, P, dtype=torch.bool, device="cuda")
    ph, si, _, _ = dur.sample(enc, m, torch.randn(B, P, device="cuda"), torch.randn(B, P, device="cuda"), torch.linspace(0, 1, 3), 0.3)
    out, tl = dur.length_regulate(enc, ph.clamp(max=20), si.clamp(max=3), torch.full((B,), P, device="cuda"))
    torch.cuda.synchronize()
    if not torch.isfinite(w).all():
        print("non-finite wav at L", L)
print("sweep done", lo, hi, step, flush=True)
