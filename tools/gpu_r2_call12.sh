#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for pair in 6 8; do
  GNK_SG_PAIR=$pair timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "stencil_gram" > gpurun_out/r2_pair${pair}_tests.log 2>&1
  echo "pytest rc=$?" >> gpurun_out/r2_pair${pair}_tests.log
  GNK_SG_PAIR=$pair timeout 600 python tools/bench_kernels.py --ks 8,15,16,23,24,30 --only spmm_ls > gpurun_out/r2_pair${pair}_kernels.log 2>&1
  GNK_SG_PAIR=$pair timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ttt --no-e2e --extras none > gpurun_out/r2_pair${pair}_bench.json 2> gpurun_out/r2_pair${pair}_bench.err
done
timeout 600 python tools/bench_kernels.py --ks 8,15,16,23,24,30 --only spmm_ls > gpurun_out/r2_pair0_kernels.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ttt --no-e2e --extras none > gpurun_out/r2_pair0_bench.json 2> gpurun_out/r2_pair0_bench.err
tail -3 gpurun_out/r2_pair6_tests.log gpurun_out/r2_pair8_tests.log | cut -c1-200; cat gpurun_out/r2_pair*_kernels.log
