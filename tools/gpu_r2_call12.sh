#!/bin/bash
# on an 8-GPU box: mailbox microbenchmark at 1/2/4/8 ranks, then the 4-GPU tests + bench
mkdir -p gpurun_out
timeout 200 python tools/microbench_mailbox.py > gpurun_out/r2_mailbox_n1.json 2> gpurun_out/r2_mailbox_n1.err
echo "mailbox n1 rc=$?"
for N in 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N \
    tools/microbench_mailbox.py > gpurun_out/r2_mailbox_n$N.json 2> gpurun_out/r2_mailbox_n$N.err
  echo "mailbox n$N rc=$?"
done
bash tools/gpu_r2_multi.sh 4 n4e
