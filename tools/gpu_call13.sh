#!/bin/bash
mkdir -p gpurun_out
PROBE_DUMP_S=70 timeout 100 python tools/batched_probe.py 32 8 > gpurun_out/probe4.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/probe4.log
