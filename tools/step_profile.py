#!/usr/bin/env python3
"""torch.profiler over ONE full bench step (all buckets): top kernels by device time, library vs own split."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402


args = bench.parse_args(["--nsteps-denoiser", os.environ.get("NFE", "128")])
dev = torch.device("cuda:0")
cfg, model, enc, dec = bench.build_models(dev, "bf16")
model.set_noise_device("philox")
wl, n = bench.global_workload(args, 1)
codes, timbres = bench.prompt_codes(wl, enc, dec, dev)
batches = bench.host_batches(wl, bench.rank_share(args, wl, 0, 1), codes, timbres)
A = args
for b in batches:
    b["dev"] = {k: b[k].to(dev) for k in ("phonemes", "src_lens", "prompts", "timbres")}
for _ in range(2):
    bench.run_step(model, dec, batches, A, dev, False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    e0.record()
    bench.run_step(model, dec, batches, A, dev, False)
    e1.record()
    torch.cuda.synchronize()
print("step device time %.1f ms" % e0.elapsed_time(e1))
rows = [(e.key, e.device_time_total / 1e3, e.count) for e in prof.key_averages() if e.device_time_total > 0]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
own = sum(r[1] for r in rows if "flm::" in r[0] or "flm" in r[0][:40])
print("sum of kernel time %.1f ms; own kernels %.1f ms; library/torch kernels %.1f ms" % (tot, own, tot - own))
for k, ms, n in rows[:45]:
    print("%9.2f ms %7d  %s" % (ms, n, k[:110]))
