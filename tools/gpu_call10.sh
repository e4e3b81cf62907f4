#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest.log 2>&1; grep -E "passed|failed|Error|assert " gpurun_out/pytest.log | tail -5
timeout 120 python tools/pair_time.py 8 2>&1 | tail -1 | tee gpurun_out/pair_time.log
MH_GC_THREADS=1 timeout 120 python tools/pair_time.py 8 2>&1 | tail -1 | tee -a gpurun_out/pair_time.log
