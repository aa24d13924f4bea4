"""How much of an outer iteration does the host spend WAITING for the device (time inside the scalar read-back) and how
much enqueuing?  If the wait is near zero the solve is host-bound at that size.  Development aid.

    python tools/host_wait.py [G] [restart] [iters]          (also under torchrun; rank 0 prints)
"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
import gauss_newton_via_generalized_krylov_subspaces_b200 as g
G = int(sys.argv[1]) if len(sys.argv) > 1 else 101
restart = int(sys.argv[2]) if len(sys.argv) > 2 else 30
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 99
rt = g.get_runtime()
pb = g.BratuPdeProblem(G, 5, 10)
y = pb.pde_operator(pb.u_true)
u0 = pb.u_true + 0.1 * np.random.RandomState(42).normal(size=pb.n)
res, jac = pb.make_res(y), pb.make_jac()
res.y_col
x0 = pb.dev.resident(u0)
kw = dict(krylow_restart=restart, max_iter=iters + 1, callback=lambda **k: None, x_on_device=True)
wait = [0.0, 0]
for name in ("read_end", "read"):
    orig = getattr(rt, name)
    def timed(*a, _o=orig, **k):
        t = time.perf_counter()
        r = _o(*a, **k)
        wait[0] += time.perf_counter() - t
        wait[1] += 1
        return r
    setattr(rt, name, timed)
import contextlib, io
for rep in range(4):
    wait[:] = [0.0, 0]
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        out = g.gauss_newton_krylow(res, x0, jac, **kw)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if rank == 0 and rep:
        print(f"G={G} world={world}: {1e6 * dt / out.nit:.1f} us per iteration, of which {1e6 * wait[0] / out.nit:.1f} us inside "
              f"the read-back ({wait[1] / out.nit:.2f} reads per iteration, {1e6 * wait[0] / max(wait[1], 1):.1f} us each)", file=sys.stderr)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
