#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -m pytest tests -m gpu -x -q -k "dense" ) > gpurun_out/pytest_dense.log 2>&1; tail -5 gpurun_out/pytest_dense.log
timeout 200 python tools/tune_dense.py 1024 1023 700 > gpurun_out/tune_dense.log 2>&1; cat gpurun_out/tune_dense.log | tail -30
