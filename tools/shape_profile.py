#!/usr/bin/env python3
"""Per-(kernel class, shape) device time of one bench bucket (CUDA events around every launch).
env: BATCH (bucket index, default 3 = longest), NFE (default 8)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from flamed_tts_b200.engines import Context  # noqa: E402


class A:
    utterances, max_batch = 256, 64
    nsteps_durgen, nsteps_denoiser, temp_durgen, temp_denoiser = 16, int(os.environ.get("NFE", 8)), 0.3, 0.3


dev = torch.device("cuda:0")
cfg, model, enc, dec = bench.build_models(dev, "bf16")
model.set_noise_device("cuda")
model.prob_generator.use_cuda_graph = False
wl, batches = bench.make_batches(A, 0, model, enc, dec, dev)
sel = [batches[int(os.environ.get("BATCH", len(batches) - 1))]]
for b in sel:
    b["dev"] = {k: b[k].to(dev) for k in ("phonemes", "src_lens", "prompts", "timbres")}
ctx = Context.get(dev)
for rep in range(2):
    ctx.profile(rep == 1)
    torch.manual_seed(0)
    bench.run_step(model, dec, sel, A, dev, False)
    torch.cuda.synchronize()
rows = ctx.profile_detail()
ctx.profile(False)
tot = sum(r["ms"] for r in rows)
print("%-18s %-42s %6s %10s %8s %9s %9s" % ("class", "shape", "n", "ms", "share", "TFLOP/s", "GB/s"))
for r in sorted(rows, key=lambda r: -r["ms"]):
    print("%-18s %-42s %6d %10.3f %7.1f%% %9.1f %9.1f" % (r["cls"], r["tag"], r["launches"], r["ms"], 100 * r["ms"] / tot,
                                                        r["flops"] / r["ms"] / 1e9, r["bytes"] / r["ms"] / 1e6))
print("total profiled ms: %.2f" % tot)
