#!/bin/bash
# K2 variant sweep + ncu --set full of the fastest one (no tests / bench)
mkdir -p gpurun_out
timeout 400 python tools/tune_fast.py ${TUNE_CONFIGS} > gpurun_out/tune.log 2>&1
BEST=$(grep '^cfg' gpurun_out/tune.log | grep 'same best True' | sort -t' ' -k4 -g | head -1 | sed 's/cfg \([0-9]*\):.*/\1/')
echo "best config: $BEST" >> gpurun_out/tune.log
MH_FAST_CONFIG=${NCU_CONFIG:-$BEST} timeout 400 ncu --set full --clock-control none --import-source on -k regex:cost_argmin -s 2 -c 1 \
  -o gpurun_out/k2_best -f python tools/ncu_k2.py 1048576 > gpurun_out/ncu_k2.log 2>&1
cat gpurun_out/tune.log; tail -2 gpurun_out/ncu_k2.log
