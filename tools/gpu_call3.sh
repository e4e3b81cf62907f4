#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/exactness_probe.py 61 50 53 > gpurun_out/exact.log 2>&1
grep -c "cfg" gpurun_out/exact.log; grep -v " 0 rows differ" gpurun_out/exact.log | head -20
timeout 300 python tools/tune_fast.py ${TUNE_CONFIGS} > gpurun_out/tune.log 2>&1; cat gpurun_out/tune.log
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest.log 2>&1; grep -E "passed|failed|Error|assert " gpurun_out/pytest.log | tail -8
