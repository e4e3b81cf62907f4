#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest.log 2>&1; grep -E "passed|failed|Error|assert " gpurun_out/pytest.log | tail -5
timeout 500 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; cat gpurun_out/bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['roofline_dense']['frac'], d['pair_e2e'], d['batched_pairs'])"
tail -3 gpurun_out/bench.err
