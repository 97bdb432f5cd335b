#!/bin/bash
TAG=${1:-b}
mkdir -p gpurun_out
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench_exit=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print("value %.1f e2e %.1f ms %.1f launches %d clocks %s"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["gpu_launches"],d["clocks"]))
print("roofline",d["roofline"])
for k,v in d["kernels"].items(): print("  %-18s %6d launches %8.1f ms share %.3f %s %.1f frac %.3f"%(k,v["launches"],v["ms"],v["share_of_step"],v["unit"],v["achieved"],v["frac"]))
PY
tail -3 gpurun_out/${TAG}_bench.err
