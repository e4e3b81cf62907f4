#!/bin/bash
# One GPU-box visit (run from the repo root under gpurun): parity tests, the bench line, the ncu launch list of the bench and
# --set full captures of the two K2 kernels the rooflines are quoted on.  Outputs under gpurun_out/; summarise into profiles/ with
#   python tools/launch_summary.py gpurun_out/launches.csv profiles/rN_bench_launch_summary.md gpurun_out/bench.log
#   python tools/ncu_summary.py gpurun_out/k2_tc.ncu-rep profiles/rN_k2_cost_argmin_tc_ncu_full.md profiles/k2_fused_traffic.json
#   python tools/ncu_summary.py gpurun_out/dense.ncu-rep profiles/rN_k2_dense_ncu_full.md
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest.log 2>&1; grep -E "passed|failed|Error|assert " gpurun_out/pytest.log | tail -5
timeout 500 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; tail -c 600 gpurun_out/bench.log; tail -3 gpurun_out/bench.err
[ "$1" = "quick" ] && exit 0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1; wc -l gpurun_out/launches.csv
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cost_argmin_tc -s 2 -c 1 -o gpurun_out/k2_tc -f \
  python tools/ncu_k2.py 1048576 > gpurun_out/ncu_k2.log 2>&1; tail -2 gpurun_out/ncu_k2.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cost_dense_tiled -s 2 -c 2 -o gpurun_out/dense -f \
  python tools/ncu_dense.py > gpurun_out/ncu_dense.log 2>&1; tail -2 gpurun_out/ncu_dense.log
