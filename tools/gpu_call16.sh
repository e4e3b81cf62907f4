#!/bin/bash
mkdir -p gpurun_out
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$2', {k:d[k] for k in ('value','ms_per_step')}, d['roofline']['kernel_ms'], d['e2e']['ms_per_step'])"; }
run 29521 base
MH_BENCH_DIAG_NO_BCAST=1 run 29522 no_bcast
MH_BENCH_DIAG_NO_ALLREDUCE=1 run 29523 no_allreduce
MH_BENCH_DIAG_NO_BCAST=1 MH_BENCH_DIAG_NO_ALLREDUCE=1 run 29524 neither
