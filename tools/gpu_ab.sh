#!/bin/bash
# microbench + bench with an env toggle A/B
TAG=${1:-ab}
mkdir -p gpurun_out
REPS=20 timeout 120 python tools/gemm_bench.py 1 3 13 16 17 18 > gpurun_out/${TAG}_gemm.txt 2>&1; cat gpurun_out/${TAG}_gemm.txt
timeout 300 python bench.py --no-cpu-baseline --no-profile > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench_exit=$?"
python -c "
import json;d=json.load(open('gpurun_out/${TAG}_bench.json'));print('value %.1f e2e %.1f ms %.1f launches %d'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['gpu_launches']))"
if [ -n "$ALT" ]; then
env $ALT timeout 300 python bench.py --no-cpu-baseline --no-profile > gpurun_out/${TAG}_bench_alt.json 2> gpurun_out/${TAG}_bench_alt.err; echo "bench_alt_exit=$?"
python -c "
import json;d=json.load(open('gpurun_out/${TAG}_bench_alt.json'));print('ALT value %.1f e2e %.1f ms %.1f launches %d'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['gpu_launches']))"
fi
