"""Device time of gnk_tsqr_ls on panels of 33..56 columns: the wide tensor-pipe path (gram_cgls.cu: gnk_cholqr_wide_try)
against the Householder TSQR it replaces there.  Development aid.

    python tools/bench_wide_ls.py [n_rows] [k,k,...]     (default 8388608 rows = one rank's slab of BASELINE config 5)
"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gauss_newton_via_generalized_krylov_subspaces_b200 as g
from gauss_newton_via_generalized_krylov_subspaces_b200.gauss_newton_krylow import tsqr_solve

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8388608
rt = g.get_runtime()
out = rt.zeros(2 * 256 + 8)
res = {"n_rows": n, "ms": {}}
ks = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [31, 33, 40, 47, 50, 55]
for k in ks:
    A = torch.randn(k * n, dtype=torch.float64, device=rt.device)
    y = torch.randn(n, dtype=torch.float64, device=rt.device)
    row = {}
    for name, method in (("tensor_pipe", 0), ("householder", 1)):
        for _ in range(2):
            tsqr_solve(rt, A, n, n, k, y, -1.0, out, method=method)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            tsqr_solve(rt, A, n, n, k, y, -1.0, out, method=method)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        v = rt.read(out, 2 * k + 4)
        row[name] = {"ms": round(ms, 4), "frac_of_hbm": round(8.0 * n * (k + 1) / (ms * 1e-3) / 6549.4e9, 3),
                     "refused": bool(v[k + 2] < 0)}
    res["ms"][k] = row
    del A, y
print(json.dumps(res))
