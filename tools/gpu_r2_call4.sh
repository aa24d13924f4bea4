#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_launches_1025.csv python bench.py --workload bratu_1024_restart30 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-parity --no-ttt --extras none > gpurun_out/r2_ncu1025.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_spmmls.csv python tools/bench_kernels.py --ks 30,15,2 --only spmm_ls --reps 2 > gpurun_out/r2_ncu_spmmls.log 2>&1
tail -2 gpurun_out/r2_ncu1025.log
