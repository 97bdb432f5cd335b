#!/bin/bash
# Full GPU verification: parity tests, smoke, bench (both arms), 2 small ncu passes.  Short timeouts everywhere.
TAG=${1:-v}
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -x -q --timeout 250 > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest_exit=$?"; tail -4 gpurun_out/${TAG}_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke_exit=$?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 500 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench_exit=$?"
python -c "
import json;d=json.load(open('gpurun_out/${TAG}_bench.json'));print('value %.1f e2e %.1f ms %.1f launches %d cpu %s'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['gpu_launches'],d['cpu_baseline']))"
if [ -n "$REF" ]; then
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref_exit=$?"
fi
if [ -n "$NCU" ]; then
  # launch list of a short bench run (per-launch gpu time; shares must agree with bench.py's live event timing)
  timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2500 -c 1500 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --utterances 64 --steps 1 --warmup 1 --no-cpu-baseline --no-profile > gpurun_out/${TAG}_ncu_bench.log 2>&1; echo "ncu_list=$?"
  python tools/ncu_summarize.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches_summary.txt 2>&1; rm -f gpurun_out/${TAG}_launches.csv
  # full captures, two launches per kernel family of one velocity evaluation + codec decode
  for K in tapgemm_tc2_kernel dwconv_tma_kernel ln_mod_bf16_kernel gn_stream_kernel act1d_kernel; do
    SKIP=2; [ "$K" = tapgemm_tc2_kernel ] && SKIP=3
    PB=32 PL=1200 timeout 200 ncu --set full --clock-control none --import-source on -k regex:"$K" --launch-skip $SKIP -c 3 \
      -o gpurun_out/${TAG}_ncu_$K python tools/kernels_probe.py > gpurun_out/${TAG}_ncu_$K.log 2>&1; echo "ncu $K=$?"
    ncu -i gpurun_out/${TAG}_ncu_$K.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_$K.raw.csv 2>/dev/null
  done
fi
while [ $(du -sm gpurun_out | cut -f1) -gt 55 ]; do rm -f "$(ls -S gpurun_out/*.ncu-rep | head -1)"; done
