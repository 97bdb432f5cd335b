#!/usr/bin/env python3
"""one denoiser velocity evaluation at a bench-sized batch (for ncu captures).  env: B, L"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import yaml  # noqa: E402

from flamed_tts_b200 import synthetic as W  # noqa: E402
from flamed_tts_b200.engines import Context, DenoiserEngine  # noqa: E402

prior = yaml.safe_load(open(os.path.join(ROOT, "configs", "prior.yaml")))
prob = yaml.safe_load(open(os.path.join(ROOT, "configs", "prob.yaml")))
sd = W.make_flamed_state_dict(prior, prob, 0)
psd = {k[len("prob_generator."):]: v for k, v in sd.items() if k.startswith("prob_generator.")}
den = DenoiserEngine(Context.get("cuda:0"), psd, prob, "bf16")
B, L = int(os.environ.get("B", 26)), int(os.environ.get("L", 1225))
g = torch.Generator().manual_seed(1)
x, spk = torch.randn(B, L, 256, generator=g).cuda(), torch.randn(B, 256, generator=g).cuda()
for _ in range(int(os.environ.get("REPS", 2))):
    v = den.forward(x, 0.5, spk)
torch.cuda.synchronize()
print("ok", float(v.abs().mean()))
