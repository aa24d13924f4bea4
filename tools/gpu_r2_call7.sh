#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests_7.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_gputests_7.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ttt > gpurun_out/r2_bench_7.json 2> gpurun_out/r2_bench_7.err
GNK_LS_CHOLQR_MIN=9 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-ttt --extras none > gpurun_out/r2_bench_7_min9.json 2> gpurun_out/r2_bench_7_min9.err
GNK_LS_FUSED=0 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-ttt --extras none > gpurun_out/r2_bench_7_nofuse.json 2> gpurun_out/r2_bench_7_nofuse.err
tail -3 gpurun_out/r2_gputests_7.log
