#!/bin/bash
# per-launch times at N GPUs: tools/gpu_r2_plt.sh N tag G...
N=$1; tag=$2; shift 2
mkdir -p gpurun_out
for G in "$@"; do
  if [ "$N" = 1 ]; then
    timeout 300 python tools/per_launch_times.py $G > gpurun_out/plt_${tag}_$G.log 2>&1
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
      tools/per_launch_times.py $G > gpurun_out/plt_${tag}_$G.log 2>&1
  fi
  echo "G=$G rc=$?"
done
