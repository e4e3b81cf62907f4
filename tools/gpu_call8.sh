#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest.log 2>&1; tail -4 gpurun_out/pytest.log
timeout 500 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; cat gpurun_out/bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['roofline_dense']['frac'], d['pair_e2e'])"
tail -3 gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1; tail -2 gpurun_out/bench_ncu.log | cut -c1-300; wc -l gpurun_out/launches.csv
