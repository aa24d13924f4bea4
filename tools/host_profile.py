"""cProfile of the host path in the launch-latency regime (Bratu 100^2, restart 30, 99 iterations).  Development aid."""
import cProfile, io, os, pstats, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gauss_newton_via_generalized_krylov_subspaces_b200 as g
G = int(sys.argv[1]) if len(sys.argv) > 1 else 101
pb = g.BratuPdeProblem(G, 5, 10)
y = pb.pde_operator(pb.u_true)
u0 = pb.u_true + 0.1 * np.random.RandomState(42).normal(size=pb.n)
res, jac = pb.make_res(y), pb.make_jac()
kw = dict(krylow_restart=30, max_iter=100, callback=lambda **k: None)
for _ in range(3):
    out = g.gauss_newton_krylow(res, u0, jac, **kw)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    out = g.gauss_newton_krylow(res, u0, jac, **kw)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
print(f"G={G}: nit={out.nit} {out.nit/dt:.0f} it/s  {1e6*dt/out.nit:.1f} us per iteration, launches per iteration "
      f"{g.get_runtime().launches()/ (8*out.nit):.1f}")
pr = cProfile.Profile(); pr.enable()
out = g.gauss_newton_krylow(res, u0, jac, **kw)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28); print(s.getvalue()[:6000])
