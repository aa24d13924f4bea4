#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests_8.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_gputests_8.log
timeout 900 python bench.py --grid-nodes 2049 --restart 50 --iters 100 --ls cgls --steps 3 --warmup 3 --no-cpu-baseline --no-ttt --extras none > gpurun_out/r2_bench_8_cgls2048.json 2> gpurun_out/r2_bench_8_cgls2048.err
echo "rc=$?" >> gpurun_out/r2_bench_8_cgls2048.err
timeout 900 python bench.py --impl reference --grid-nodes 513 --iters 30 > gpurun_out/r2_bench_8_ref513.json 2> gpurun_out/r2_bench_8_ref513.err
echo "rc=$?" >> gpurun_out/r2_bench_8_ref513.err
tail -3 gpurun_out/r2_gputests_8.log; tail -c 400 gpurun_out/r2_bench_8_cgls2048.json; tail -c 600 gpurun_out/r2_bench_8_ref513.json
