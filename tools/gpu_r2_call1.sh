#!/bin/bash
# round 2, GPU call 1: fused stencil+Gram microbench (v2 shuffle, v3 TMA) and the latency-regime baseline (1024^2)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
MB=tools/microbench_stencil_gram
{
  echo "== validation"
  timeout 120 $MB 256 15 3
  timeout 120 $MB 512 31 3
  timeout 120 $MB 512 7 3
  timeout 120 $MB 512 20 3 16
  echo "== timing 4096"
  timeout 120 $MB 4096 30 2
  for ti in 32 64 128; do timeout 120 $MB 4096 30 3 $ti; done
  timeout 120 $MB 4096 30 4 64
  for k in 23 15 7 3; do timeout 120 $MB 4096 $k 3 64; timeout 120 $MB 4096 $k 4 64; done
} > gpurun_out/r2_mb1.log 2>&1
timeout 600 python bench.py --grid-nodes 1025 --restart 30 --iters 99 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_bench1025_base.json 2> gpurun_out/r2_bench1025_base.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_launches_1025_base.csv python bench.py --grid-nodes 1025 --restart 30 --iters 99 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2_ncu1025.log 2>&1
tail -5 gpurun_out/r2_mb1.log
