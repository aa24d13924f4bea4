#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests_11.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_gputests_11.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_11.json 2> gpurun_out/r2_bench_11.err
echo "bench rc=$?" >> gpurun_out/r2_bench_11.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-parity --no-ttt --extras none > gpurun_out/r2_ncu_launches.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"stencil_gram_kernel|cholqr_refine_kernel|cholqr_factor" -c 5 -o gpurun_out/r2_ls_k30 python tools/bench_kernels.py --ks 30 --only spmm_ls --reps 1 > gpurun_out/r2_ncu_ls.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"dots_kernel|combine_kernel|update_kernel|residual_kernel|apply_kernel|normalize_kernel" -c 8 -o gpurun_out/r2_stream_k30 python tools/bench_kernels.py --ks 30 --only cgs_dots,combine,cgs_update,residual,spmv_t,normalize --reps 1 > gpurun_out/r2_ncu_stream.log 2>&1
timeout 1500 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_11_reference.json 2> gpurun_out/r2_bench_11_reference.err
echo "ref rc=$?" >> gpurun_out/r2_bench_11_reference.err
tail -4 gpurun_out/r2_gputests_11.log | cut -c1-300; tail -c 700 gpurun_out/r2_bench_11_reference.json
