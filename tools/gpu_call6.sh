#!/bin/bash
mkdir -p gpurun_out
MH_DENSE_SKIP_TAIL=1 timeout 200 python tools/tune_dense.py 1024 1027 1031 1039 1055 2> gpurun_out/tune_dense2.err | grep "variant 1" > gpurun_out/tune_dense2.log; cat gpurun_out/tune_dense2.log; tail -3 gpurun_out/tune_dense2.err
