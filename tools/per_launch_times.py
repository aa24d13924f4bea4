"""per-launch device times of one profiled GNK solve, by kernel class and in launch order.  Development aid.

    python tools/per_launch_times.py [G] [iters]                      one GPU
    torchrun --nproc-per-node N ... tools/per_launch_times.py [G]     N GPUs (rank 0 prints)

Also prints the wall time between consecutive outer iterations of an UNprofiled solve (host clock in the callback,
which the solver calls after its one read-back per iteration) -- the profiled pass puts two events around every class,
the unprofiled one shows what the iteration really costs.
"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import gauss_newton_via_generalized_krylov_subspaces_b200 as g

G = int(sys.argv[1]) if len(sys.argv) > 1 else 4097
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 30
n = (G - 1) ** 2
rt = g.get_runtime()
pb = g.BratuPdeProblem(G, 5, 10)
y = pb.pde_operator(pb.u_true)
u0 = pb.u_true + 0.1 * np.random.RandomState(42).normal(size=n)
res = pb.make_res(y); res.y_col
jac = pb.make_jac()
x0 = pb.dev.resident(u0)
kw = dict(krylow_restart=iters, max_iter=iters + 1, x_on_device=True)

stamps = []
def cb(**k):
    stamps.append(time.perf_counter())

for _ in range(3):
    g.gauss_newton_krylow(res, x0, jac, callback=lambda **k: None, **kw)
for rep in range(2):
    stamps.clear()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    out = g.gauss_newton_krylow(res, x0, jac, callback=cb, **kw)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    if rank == 0:
        d = np.diff(np.array([t0] + stamps + [t1])) * 1e6
        print(f"unprofiled solve {1e3 * (t1 - t0):.2f} ms, nit {out.nit}; us between callbacks:\n  "
              + " ".join(f"{v:.0f}" for v in d), file=sys.stderr)

rt.begin_profile()
g.gauss_newton_krylow(res, x0, jac, callback=lambda **k: None, **kw)
rt.sync()
if rank == 0:
    for name, recs in rt.prof.items():
        us = [1e3 * a.elapsed_time(b) for a, b, _ in recs]
        print(f"{name:12s} n={len(us):3d} total {sum(us) / 1e3:7.3f} ms | " + " ".join(f"{v:.0f}" for v in us), file=sys.stderr)
rt.end_profile()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
