#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests_6.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_gputests_6.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_6.json 2> gpurun_out/r2_bench_6.err
echo "bench rc=$?" >> gpurun_out/r2_bench_6.err
GNK_SPECULATE=0 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-ttt > gpurun_out/r2_bench_6_nospec.json 2> gpurun_out/r2_bench_6_nospec.err
tail -3 gpurun_out/r2_gputests_6.log
