#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -m pytest tests -m gpu -x -q -k "dense" ) > gpurun_out/pytest_dense.log 2>&1; tail -3 gpurun_out/pytest_dense.log
timeout 200 python tools/tune_dense.py 1024 1023 > gpurun_out/tune_dense.log 2>&1; grep -v "variant 2" gpurun_out/tune_dense.log | tail -30
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cost_argmin -s 2 -c 1 -o gpurun_out/k2_tc -f python tools/ncu_k2.py 1048576 > gpurun_out/ncu_k2.log 2>&1; tail -2 gpurun_out/ncu_k2.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cost_dense_tiled -s 2 -c 2 -o gpurun_out/dense -f python tools/ncu_dense.py > gpurun_out/ncu_dense.log 2>&1; tail -2 gpurun_out/ncu_dense.log
