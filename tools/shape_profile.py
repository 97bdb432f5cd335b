#!/usr/bin/env python3
"""Per-(kernel class, shape) device time of one bench back-bucket (CUDA events around every launch).
env: NFE (denoiser steps, default 8), UTTS (pool size, default 256), BUCKET (which front bucket, default 0 = longest)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from flamed_tts_b200.engines import Context  # noqa: E402

args = bench.parse_args(["--nsteps-denoiser", os.environ.get("NFE", "8"), "--utterances", os.environ.get("UTTS", "256")])
dev = torch.device("cuda:0")
cfg, model, enc, dec = bench.build_models(dev, "bf16")
model.set_noise_device("philox")
model.prob_generator.use_cuda_graph = False
wl, n = bench.global_workload(args, 1)
codes, timbres = bench.prompt_codes(wl, enc, dec, dev)
batches = bench.host_batches(wl, bench.rank_share(args, wl, 0, 1), codes, timbres)
sel = [batches[int(os.environ.get("BUCKET", 0))]]
for b in sel:
    b["dev"] = {k: b[k].to(dev) for k in ("phonemes", "src_lens", "prompts", "timbres")}
ctx = Context.get(dev)
for rep in range(2):
    ctx.profile(rep == 1)
    torch.manual_seed(0)
    bench.run_step(model, dec, sel, args, dev, False)
    torch.cuda.synchronize()
rows = ctx.profile_detail()
ctx.profile(False)
tot = sum(r["ms"] for r in rows)
print("%-18s %-42s %6s %10s %8s %9s %9s" % ("class", "shape", "n", "ms", "share", "TFLOP/s", "GB/s"))
for r in sorted(rows, key=lambda r: -r["ms"]):
    print("%-18s %-42s %6d %10.3f %7.1f%% %9.1f %9.1f" % (r["cls"], r["tag"], r["launches"], r["ms"], 100 * r["ms"] / tot,
                                                        r["flops"] / r["ms"] / 1e9, r["bytes"] / r["ms"] / 1e6))
print("total profiled ms: %.2f" % tot)
