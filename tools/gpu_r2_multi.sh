#!/bin/bash
# usage: tools/gpu_r2_multi.sh N tag
cd "$(dirname "$0")/.."
N=$1; TAG=$2
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q -k "$N" > gpurun_out/r2_multi_${TAG}_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_multi_${TAG}_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_multi_${TAG}_bench.json 2> gpurun_out/r2_multi_${TAG}_bench.err
echo "bench rc=$?" >> gpurun_out/r2_multi_${TAG}_bench.err
tail -5 gpurun_out/r2_multi_${TAG}_tests.log; tail -c 600 gpurun_out/r2_multi_${TAG}_bench.json
