#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests_9.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_gputests_9.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ttt > gpurun_out/r2_bench_9.json 2> gpurun_out/r2_bench_9.err
tail -30 gpurun_out/r2_gputests_9.log | cut -c1-300
