"""The N>1 CUDA + NCCL path: torchrun with one rank per GPU (skipped when fewer than 2 GPUs are visible).
The sharded host logic itself is also covered on the CPU over gloo (tests/test_distributed_cpu.py)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_cuda_nccl(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "multi_worker.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    log = os.path.join(ROOT, "gpurun_out", f"multi_worker_w{world}.log")
    if os.path.isdir(os.path.dirname(log)):
        with open(log, "w") as f:
            f.write(p.stdout + "\n==== stderr ====\n" + p.stderr)
    print(p.stdout[-4000:])
    print(p.stderr[-6000:])
    assert p.returncode == 0 and "MULTI_OK" in p.stdout
