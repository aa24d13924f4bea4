"""where does the end-to-end (host buffers in, host ndarray out) time go?  Development aid."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gauss_newton_via_generalized_krylov_subspaces_b200 as g
rt = g.get_runtime()
G = 4097; n = (G - 1) ** 2
pb = g.BratuPdeProblem(G, 5, 10)
y = pb.pde_operator(pb.u_true)
u0 = pb.u_true + 0.1 * np.random.RandomState(42).normal(size=n)
yp = torch.empty(n, dtype=torch.float64, pin_memory=True); yp.numpy()[:] = y
up = torch.empty(n, dtype=torch.float64, pin_memory=True); up.numpy()[:] = u0
def T(): torch.cuda.synchronize(); return time.perf_counter()
for it in range(4):
    t0 = T(); res = pb.make_res(yp.numpy()); res.y_col; t1 = T()
    jac = pb.make_jac(); x0d = pb.dev.resident(up.numpy()); t2 = T()
    out = g.gauss_newton_krylow(res, x0d, jac, max_iter=31, callback=lambda **k: None, x_on_device=True); t3 = T()
    xh = np.asarray(out.x); t4 = T()
    out2 = g.gauss_newton_krylow(res, up.numpy(), jac, max_iter=31, callback=lambda **k: None); t5 = T()
    print(f"iter {it}: make_res+y upload {1e3*(t1-t0):.1f} ms | u0 upload {1e3*(t2-t1):.1f} | solve(resident) {1e3*(t3-t2):.1f} | "
          f"x download {1e3*(t4-t3):.1f} | solve(host in/out) {1e3*(t5-t4):.1f}", file=sys.stderr)
