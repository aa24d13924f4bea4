#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests_5.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_gputests_5.log
timeout 600 python tools/bench_kernels.py --ks 2,7,15,23,30 --only spmm_ls,cgs_dots --out gpurun_out/r2_kernels_5.json > gpurun_out/r2_kernels_5.log 2>&1
timeout 600 python tools/bench_kernels.py --m 1024 --ks 2,7,15,23,30 --only spmm_ls,cgs_dots,cgs_update,combine,residual,spmv_t,normalize > gpurun_out/r2_kernels_5_1024.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_5.json 2> gpurun_out/r2_bench_5.err
echo "bench rc=$?" >> gpurun_out/r2_bench_5.err
tail -3 gpurun_out/r2_gputests_5.log; cat gpurun_out/r2_kernels_5.log
