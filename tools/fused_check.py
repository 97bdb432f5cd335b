#!/usr/bin/env python3
"""One velocity evaluation of the bf16 denoiser (LayerNorm-fused depthwise conv, programmatic dependent launches)
against the CPU oracle at edge and bench sizes, reproducibility, and timing per kernel class at bench-sized batches.
usage: python tools/fused_check.py"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import yaml  # noqa: E402

from flamed_tts_b200 import synthetic as W  # noqa: E402
from flamed_tts_b200.engines import Context, DenoiserEngine  # noqa: E402
from oracle import flamed_oracle as O  # noqa: E402

prior = yaml.safe_load(open(os.path.join(ROOT, "configs", "prior.yaml")))
prob = yaml.safe_load(open(os.path.join(ROOT, "configs", "prob.yaml")))
sd = W.make_flamed_state_dict(prior, prob, 0)
psd = {k[len("prob_generator."):]: v for k, v in sd.items() if k.startswith("prob_generator.")}
ctx = Context.get("cuda:0")
den = DenoiserEngine(ctx, psd, prob, "bf16")


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


ok = True
for B, L in ((1, 7), (1, 20), (2, 33), (3, 333), (2, 1200), (1, 2400), (5, 64)):
    g = torch.Generator().manual_seed(B * 1000 + L)
    x, spk = torch.randn(B, L, 256, generator=g), torch.randn(B, 256, generator=g)
    with torch.inference_mode():
        ref = O.denoiser_forward(psd, "denoiser", x, torch.full((1, 1), 0.37), spk)
    v = den.forward(x.cuda(), 0.37, spk.cuda()).float().cpu()
    v2 = den.forward(x.cuda(), 0.37, spk.cuda()).float().cpu()
    e = rel(v, ref)
    good = e < 1e-2 and bool(torch.isfinite(v).all()) and torch.equal(v, v2)
    ok &= good
    print("B%d L%d: velocity vs oracle %.3e  reproducible %s  %s" % (B, L, e, torch.equal(v, v2), "ok" if good else "FAIL"), flush=True)

# 8 Euler steps through sample() (graph capture of dependent launches on the second sighting)
g = torch.Generator().manual_seed(3)
B, L, nfe = 2, 150, 8
cond, spk, noise = torch.relu(torch.randn(B, L, 256, generator=g)), torch.randn(B, 256, generator=g), torch.randn(B, L, 256, generator=g)
with torch.inference_mode():
    ref = O.denoiser_sample(sd, "prob_generator", cond, spk, noise, nfe, 0.3).transpose(1, 2)
ts = torch.linspace(0, 1, nfe + 1)
a = den.sample(cond, spk, noise, ts, 0.3, use_graph=False).cpu()
outs = [den.sample(cond, spk, noise, ts, 0.3, use_graph=True).cpu() for _ in range(3)]
good = rel(a, ref) < 1e-2 and all(torch.equal(a, o) for o in outs)
ok &= good
print("8-step loop: vs oracle %.3e, graph == direct %s  %s" % (rel(a, ref), all(torch.equal(a, o) for o in outs), "ok" if good else "FAIL"))

for B, L in ((26, 1225), (62, 520), (64, 1236)):
    g = torch.Generator().manual_seed(1)
    x, spk = torch.randn(B, L, 256, generator=g).cuda(), torch.randn(B, 256, generator=g).cuda()
    for _ in range(3):
        den.forward(x, 0.5, spk)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        den.forward(x, 0.5, spk)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / 10 * 1000
    ctx.profile(True)
    for _ in range(5):
        den.forward(x, 0.5, spk)
    prof = ctx.profile_read()
    ctx.profile(False)
    print("B%d L%d: %.3f ms per velocity; per class (ms per velocity): %s" % (
        B, L, wall, {k: round(v["ms"] / 5, 3) for k, v in prof.items()}), flush=True)
print("ALL OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
