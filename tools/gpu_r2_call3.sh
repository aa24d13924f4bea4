#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests_3.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_gputests_3.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_3.json 2> gpurun_out/r2_bench_3.err
echo "bench rc=$?" >> gpurun_out/r2_bench_3.err
timeout 600 python tools/bench_kernels.py --ks 2,7,8,15,16,23,24,30 --only spmm,tsqr,spmm_ls --out gpurun_out/r2_kernels_3.json > gpurun_out/r2_kernels_3.log 2>&1
tail -3 gpurun_out/r2_gputests_3.log; tail -c 1500 gpurun_out/r2_bench_3.json
