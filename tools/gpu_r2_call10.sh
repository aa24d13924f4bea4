#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests_10.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_gputests_10.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ttt > gpurun_out/r2_bench_10.json 2> gpurun_out/r2_bench_10.err
GNK_PDL=0 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ttt --no-e2e --no-parity > gpurun_out/r2_bench_10_nopdl.json 2> gpurun_out/r2_bench_10_nopdl.err
GNK_SG_CW=10 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ttt --no-e2e --extras none > gpurun_out/r2_bench_10_cw10.json 2> gpurun_out/r2_bench_10_cw10.err
tail -5 gpurun_out/r2_gputests_10.log | cut -c1-300
