"""8192^2 sanity run on one GPU: 64-bit index paths, memory, property checks (no golden exists at this size:
the reference cannot run it on a 62 GB host).  Development aid."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gauss_newton_via_generalized_krylov_subspaces_b200 as g
G = int(sys.argv[1]) if len(sys.argv) > 1 else 8193
pb = g.BratuPdeProblem(G, 5, 10)
n = pb.n
y = pb.pde_operator(pb.u_true)
r0 = pb.make_res(y)(pb.u_true)
print("n =", n, " residual at the solution: max|F| =", float(np.max(np.abs(r0))), flush=True)
u0 = pb.u_true + 0.1 * np.random.RandomState(42).normal(size=n)
res, jac = pb.make_res(y), pb.make_jac()
losses = []
def cb(x, nfev, cg_iter): losses.append(res.loss(x))
for kw in (dict(max_iter=31), dict(max_iter=9, krylow_restart=50, ls_solver="cgls", cg_rtol=1e-10)):
    losses.clear(); torch.cuda.synchronize(); t0 = time.perf_counter()
    out = g.gauss_newton_krylow(res, u0, jac, callback=cb, **kw)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    mono = all(b <= a * (1 + 1e-12) for a, b in zip(losses, losses[1:]))
    print(kw, "nit", out.nit, "nfev", out.nrev, f"{out.nit/dt:.1f} it/s", "loss first/last %.6e %.6e" % (losses[0], losses[-1]),
          "monotone:", mono, "peak mem GB %.1f" % (torch.cuda.max_memory_allocated() / 1e9), flush=True)
    if "ls_solver" in kw:
        ref = g.gauss_newton_krylow(res, u0, jac, callback=lambda **k: None, max_iter=9, krylow_restart=50)
        print("   cgls(1e-10) vs qr: rel diff of x %.2e" % (np.max(np.abs(out.x - ref.x)) / np.max(np.abs(ref.x))), flush=True)
