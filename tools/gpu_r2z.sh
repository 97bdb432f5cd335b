#!/bin/bash
# Final round-2 evidence on one GPU: parity tests, smoke, bench (both arms), configs 1/2, ncu launch list of a short bench
# run, ncu --set full of the loop kernels of one velocity evaluation.  Short timeouts everywhere.
TAG=${1:-r2z}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rA --timeout 300 > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest_exit=$?"
grep -E "passed|failed|error" gpurun_out/${TAG}_pytest_gpu.log | tail -2
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke_exit=$?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench_exit=$?"
python -c "
import json;d=json.load(open('gpurun_out/${TAG}_bench.json'));print('value %.1f e2e %.1f ms %.1f launches %d roofline %s eager %s cpu %s'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['gpu_launches'],d['roofline']['frac'],d['gpu_eager_baseline']['bf16_autocast']['value'],d['cpu_baseline']['value']))"
timeout 400 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref_exit=$?"; tail -c 300 gpurun_out/${TAG}_bench_ref.json
timeout 600 python tools/config12.py > gpurun_out/${TAG}_config12.jsonl 2> gpurun_out/${TAG}_config12.err; echo "config12_exit=$?"; cut -c1-260 gpurun_out/${TAG}_config12.jsonl
# launch list of a short bench run (per-launch gpu time; shares must agree with bench.py's live event timing)
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2500 -c 1500 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --utterances 64 --steps 1 --warmup 1 --no-cpu-baseline --no-eager-baseline --no-profile --no-e2e > gpurun_out/${TAG}_ncu_bench.log 2>&1; echo "ncu_list=$?"
python tools/ncu_summarize.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches_summary.txt 2>&1; rm -f gpurun_out/${TAG}_launches.csv
head -16 gpurun_out/${TAG}_launches_summary.txt
# full captures out of the second of two velocity evaluations at 32 x 1200 rows
for V in "tapgemm_tc2_kernel:21:3" "dwconv_ln_kernel:5:1" "gn_stream_kernel:5:1" "ln_mod_bf16_kernel:1:1"; do
  K=${V%%:*}; R=${V#*:}; SKIP=${R%%:*}; CNT=${R#*:}
  B=32 L=1200 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"$K" --launch-skip $SKIP -c $CNT \
    -o gpurun_out/${TAG}_ncu_$K python tools/one_velocity.py > gpurun_out/${TAG}_ncu_$K.log 2>&1; echo "ncu $K=$?"
  ncu -i gpurun_out/${TAG}_ncu_$K.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_$K.raw.csv 2>/dev/null
  rm -f gpurun_out/${TAG}_ncu_$K.ncu-rep
done
python tools/ncu_key_metrics.py gpurun_out/${TAG}_ncu_*.raw.csv > gpurun_out/${TAG}_ncu_full_key_metrics.txt 2>&1
grep -E "^==|^   (unnamed|void|flm)|gpu__time_duration|dram read" gpurun_out/${TAG}_ncu_full_key_metrics.txt | cut -c1-200
