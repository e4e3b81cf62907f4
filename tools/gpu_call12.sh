#!/bin/bash
mkdir -p gpurun_out
timeout 80 python tools/batched_probe.py 4 1 > gpurun_out/probe1.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/probe1.log
MH_TRACE=1 timeout 100 python tools/batched_probe.py 8 2 > gpurun_out/probe2.log 2>&1; echo "rc=$?"; grep -v "iteration" gpurun_out/probe2.log | tail -12
MH_GC_THREADS=1 timeout 100 python tools/batched_probe.py 16 8 > gpurun_out/probe3.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/probe3.log
