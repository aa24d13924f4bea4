"""Numerical experiment behind csrc/cholqr.cu (test infrastructure, run by hand; needs ~25 GB and ~10 min at 4097):

    python tests/ls_numerics_experiment.py 1025 30 chol     # CholeskyQR2 as the projected least-squares solver
    python tests/ls_numerics_experiment.py 4097 30 ref1     # Cholesky + one refinement step

Replays the reference's Bratu run (oracle port, numpy/scipy, bit-identical operator) with the projected least squares
`linear_least_squares` (gauss_newton_krylow.py:16-36, LAPACK Householder QR) replaced, for k >= 8, by
  chol : R1 = chol(P^T P), B = P R1^{-1}, R2 = chol(B^T B), R = R2 R1, back substitution     (P = [A | y])
  refN : d0 from the normal equations with R1 = chol(A^T A), then N steps d += (R1^T R1)^{-1} A^T (y - A d)
and prints, per outer iteration, the deviation of the iterates from the reference's golden trace next to the tolerance
the GPU parity tests use (tests/golden_util.py:sensitivity_bound).  Results: at 1024^2
(cond(JV_k) ~ 2e3) and 4096^2 (cond ~ 3e4) both variants stay at the 2e-13 ... 1e-11 level of the LAPACK run; the
refinement step has relative size 1e-10 ... 1e-9 (1024^2) and 3e-8 ... 5e-7 (4096^2) = cond^2 eps.
"""
import sys
import time

import numpy as np
import scipy.linalg

sys.path.insert(0, __file__.rsplit("/tests/", 1)[0])
sys.path.insert(0, __file__.rsplit("/", 1)[0])
from golden_util import Golden, sensitivity_bound  # noqa: E402
from oracle import gnk_oracle as orc  # noqa: E402


def main():
    G, its, mode = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    gold, sens = Golden(f"bratu_g{G}").run("gnk_k30"), Golden(f"bratu_g{G}_sens").run("gnk_k30")
    o = orc.BratuOracle(G, 5, 10)
    y, u0 = o.operator(o.u_true), o.start_vector(seed=42)
    notes = []

    def cholqr2(A, yv, log=None):
        P = np.column_stack([A, yv])
        R1 = np.linalg.cholesky(P.T @ P).T
        B = P @ scipy.linalg.solve_triangular(R1, np.eye(R1.shape[0]))
        R = np.linalg.cholesky(B.T @ B).T @ R1
        k = A.shape[1]
        notes.append(np.linalg.cond(R[:k, :k]))
        return scipy.linalg.solve_triangular(R[:k, :k], R[:k, k])

    def refine(A, yv, log=None, steps=1):
        R1 = np.linalg.cholesky(A.T @ A).T
        solve = lambda v: scipy.linalg.solve_triangular(R1, scipy.linalg.solve_triangular(R1, v, trans="T"))
        d = solve(A.T @ yv)
        for _ in range(steps):
            dl = solve(A.T @ (yv - A @ d))
            d = d + dl
            notes.append(np.linalg.norm(dl) / np.linalg.norm(d))
        return d

    def ls(A, yv, log=None):
        if A.shape[1] < 8 or mode == "qr":
            return orc.ls_qr(A, yv, log)
        return cholqr2(A, yv, log) if mode == "chol" else refine(A, yv, log, int(mode[3:]))

    idx, xs = gold["sample_idx"], []
    t = time.time()
    out = orc.gnk(o.make_res(y), u0, o.make_jac(), restart=None, max_iter=its + 1, ls=ls,
                  callback=lambda x, nfev, cg_iter: xs.append(x[idx].copy()))
    print(f"{time.time() - t:.1f} s, nit={out['nit']} nfev={out['nfev']}")
    print("cond(R) per call" if mode == "chol" else "|delta|/|d| per call", ["%.2e" % v for v in notes])
    xs = np.array(xs)
    n = len(xs)
    dev = np.max(np.abs(xs - gold["xs"][:n]) / np.max(np.abs(gold["xs"][:n]), axis=1, keepdims=True), axis=1)
    tol = sensitivity_bound(gold, [sens])
    for i in range(n):
        print(f"iteration {i + 1:3d}  deviation {dev[i]:.2e}  tolerance {tol[i]:.2e}  {'ok' if dev[i] <= tol[i] else 'FAIL'}")


if __name__ == "__main__":
    main()
