#!/bin/bash
TAG=${1:-r2r}
mkdir -p gpurun_out
for V in 0 1 2 3 4; do
FLAMED_B200_ACT=$V timeout 200 python tools/codec_check.py > gpurun_out/${TAG}_codec_act$V.txt 2>&1; echo "variant $V exit=$?"; grep -E "decode|B26" gpurun_out/${TAG}_codec_act$V.txt | cut -c1-200
done
