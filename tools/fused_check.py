#!/usr/bin/env python3
"""Fused LN + depthwise conv + GroupNorm kernel (dwconv_fused.cu) against the unfused kernel chain and the CPU
oracle on one velocity evaluation, plus timing of both at a bench-sized batch.  usage: python tools/fused_check.py"""
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


def run(x, spk, t, fused):
    if fused:
        os.environ.pop("FLAMED_B200_NO_FUSED", None)
    else:
        os.environ["FLAMED_B200_NO_FUSED"] = "1"
    v = den.forward(x.cuda(), t, spk.cuda()).float().cpu()
    torch.cuda.synchronize()
    return v


ok = True
for B, L in ((1, 7), (1, 20), (2, 33), (3, 333), (2, 1200), (1, 2400), (5, 64)):
    g = torch.Generator().manual_seed(B * 1000 + L)
    x, spk = torch.randn(B, L, 256, generator=g), torch.randn(B, 256, generator=g)
    with torch.inference_mode():
        ref = O.denoiser_forward(psd, "denoiser", x, torch.full((1, 1), 0.37), spk)
    vf, vu = run(x, spk, 0.37, True), run(x, spk, 0.37, False)
    ef, eu, d = rel(vf, ref), rel(vu, ref), rel(vf, vu)
    good = ef < 1e-2 and bool(torch.isfinite(vf).all())
    ok &= good
    print("B%d L%d: fused vs oracle %.3e  unfused vs oracle %.3e  fused vs unfused %.3e  %s" % (B, L, ef, eu, d, "ok" if good else "FAIL"),
          flush=True)
    vf2 = run(x, spk, 0.37, True)
    if not torch.equal(vf, vf2):
        ok = False
        print("  NOT deterministic")

# timing at a bench-sized batch (profiler: CUDA events around every launch)
for B, L in ((26, 1225), (62, 520), (64, 1236)):
    g = torch.Generator().manual_seed(1)
    x, spk = torch.randn(B, L, 256, generator=g).cuda(), torch.randn(B, 256, generator=g).cuda()
    for fused in (True, False):
        if fused:
            os.environ.pop("FLAMED_B200_NO_FUSED", None)
        else:
            os.environ["FLAMED_B200_NO_FUSED"] = "1"
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
        print("B%d L%d fused=%s: %.3f ms per velocity; per class (ms per velocity): %s" % (
            B, L, fused, wall, {k: round(v["ms"] / 5, 3) for k, v in prof.items()}), flush=True)
print("ALL OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
