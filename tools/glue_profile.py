#!/usr/bin/env python3
"""torch profiler over the PyTorch glue (prior FFT decoders) for one B=64 bucket."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402


class A:
    utterances, max_batch = 256, 64
    nsteps_durgen, nsteps_denoiser, temp_durgen, temp_denoiser = 16, 4, 0.3, 0.3


dev = torch.device("cuda:0")
cfg, model, enc, dec = bench.build_models(dev, "bf16")
wl, batches = bench.make_batches(A, 0, model, enc, dec, dev)
b = batches[1]
pg = model.prior_generator
L = 1100
x = torch.randn(64, L, 192, device=dev)
tgt = torch.full((64,), L, device=dev, dtype=torch.long)
tgt[::3] -= 37
pr = b["prompts"].to(dev)
for _ in range(2):
    pg.decode_priors(x, tgt, pr, pr.size(-1), bf16=True)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    pg.decode_priors(x, tgt, pr, pr.size(-1), bf16=True)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=60))
