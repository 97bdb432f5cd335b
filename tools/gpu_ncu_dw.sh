#!/bin/bash
# ncu --set full of the LN-fused depthwise kernel (persistent form)
TAG=${1:-r2e}
mkdir -p gpurun_out
for V in "mode2:FLAMED_B200_FUSED_MODE=2:dwconv_ln_gn"; do
  NAME=${V%%:*}; REST=${V#*:}; ENVV=${REST%%:*}; KER=${REST#*:}
  env $ENVV timeout 120 python tools/one_velocity.py > gpurun_out/${TAG}_plain_$NAME.log 2>&1 &&
  env $ENVV timeout 600 ncu --set full --clock-control none --import-source on -k regex:$KER -s 6 -c 1 -o gpurun_out/${TAG}_$NAME python tools/one_velocity.py > gpurun_out/${TAG}_ncu_$NAME.log 2>&1
  echo "$NAME exit=$?"
  ncu -i gpurun_out/${TAG}_$NAME.ncu-rep --page raw --csv > gpurun_out/${TAG}_$NAME.raw.csv 2>/dev/null
  ncu -i gpurun_out/${TAG}_$NAME.ncu-rep --page source --csv > gpurun_out/${TAG}_$NAME.source.csv 2>/dev/null
done
