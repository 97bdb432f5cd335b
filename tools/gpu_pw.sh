#!/bin/bash
# is the GEMM power-limited?  long back-to-back runs with the SM clock / board power sampled next to them
TAG=${1:-pw}
mkdir -p gpurun_out
for D in 0; do  # (the debug variants 1 = no epilogue work, 2 = no MMAs existed only for the r2w measurement)
  for C in 3 1 17; do
    nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_throttle_reasons.active --format=csv,noheader -lms 100 > gpurun_out/${TAG}_smi_d${D}_c$C.txt &
    SMI=$!
    REPS=8000 timeout 200 python tools/gemm_bench.py $C 2>&1 | tee -a gpurun_out/${TAG}_gemm.txt
    kill $SMI
    echo "debug=$D case=$C  clocks/power (last samples under load):"; tail -8 gpurun_out/${TAG}_smi_d${D}_c$C.txt | head -6 | tr '\n' ';'; echo
  done
done
