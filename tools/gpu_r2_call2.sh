#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
MB=tools/microbench_stencil_gram
{
  echo "== validation of the partial-tile path"
  timeout 120 $MB 512 31 3 64 10
  timeout 120 $MB 520 20 3 64 12
  echo "== timing 4096, consumer warps"
  for cw in 8 10 12 16; do for k in 30 15; do timeout 120 $MB 4096 $k 3 64 $cw; timeout 120 $MB 4096 $k 4 64 $cw; done; done
} > gpurun_out/r2_mb2.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:stencil_gram_tma -c 1 -s 3 -o gpurun_out/r2_sg_v3_k30 $MB 4096 30 3 64 8 > gpurun_out/r2_ncu_sg.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:stencil_gram_tma -c 1 -s 3 -o gpurun_out/r2_sg_v4_k30 $MB 4096 30 4 64 8 >> gpurun_out/r2_ncu_sg.log 2>&1
tail -3 gpurun_out/r2_mb2.log
