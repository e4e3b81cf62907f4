#!/bin/bash
mkdir -p gpurun_out
nproc; lscpu | grep -E "Model name|^CPU\(s\)|Thread|MHz" | head -5
for t in 1 2 4 8 default; do
  if [ "$t" = default ]; then unset MH_GC_THREADS; else export MH_GC_THREADS=$t; fi
  timeout 120 python tools/pair_time.py 8 2>&1 | tail -1
done | tee gpurun_out/pair_time.log
