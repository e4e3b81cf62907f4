#!/bin/bash
# One GPU-box visit: parity tests, K2 variant sweep, ncu --set full of the fastest variant, bench line.
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest.log 2>&1
timeout 300 python tools/tune_fast.py ${TUNE_CONFIGS:-5 30 31 32 33 34 35 36 37 38 39 40 41 42 43 44 45 46} > gpurun_out/tune.log 2>&1
BEST=$(grep '^cfg' gpurun_out/tune.log | grep 'same best True' | sort -t' ' -k4 -g | head -1 | sed 's/cfg \([0-9]*\):.*/\1/')
echo "best config: $BEST" >> gpurun_out/tune.log
MH_FAST_CONFIG=$BEST timeout 400 ncu --set full --clock-control none --import-source on -k regex:cost_argmin -s 2 -c 1 \
  -o gpurun_out/k2_best -f python tools/ncu_k2.py 1048576 > gpurun_out/ncu_k2.log 2>&1
timeout 400 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err
tail -3 gpurun_out/pytest.log; cat gpurun_out/tune.log; tail -2 gpurun_out/ncu_k2.log; cat gpurun_out/bench.log
