"""Wall time of gauss_newton (full space, CGLS inner solve) on a Bratu grid.  Development aid.

    python tools/time_gauss_newton.py [G] [reps]         (GNK_CG_PIPELINE=0 for the synchronous CG loop)
"""
import os, sys, time, io, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import gauss_newton_via_generalized_krylov_subspaces_b200 as g

G = int(sys.argv[1]) if len(sys.argv) > 1 else 101
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
pb = g.BratuPdeProblem(G, 5, 10)
y = pb.pde_operator(pb.u_true)
u0 = pb.u_true + 0.1 * np.random.RandomState(42).normal(size=pb.n)
res, jac = pb.make_res(y), pb.make_jac()
for rep in range(reps + 1):
    its = []
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        out = g.gauss_newton(res, u0, jac, max_iter=6, callback=lambda x, nfev, cg_iter: its.append(cg_iter))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if rep:
        print(f"G={G} pipeline={os.environ.get('GNK_CG_PIPELINE', '1')}: {dt * 1e3:.1f} ms, nit={out.nit}, CG iterations {sum(its)} "
              f"({1e6 * dt / max(sum(its), 1):.1f} us per CG iteration)", flush=True)
