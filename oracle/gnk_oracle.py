"""CPU oracle for the Gauss-Newton-Krylov hot path.  TEST INFRASTRUCTURE ONLY.

This module restates, in plain numpy/scipy, the algorithm of the reference
(mariusbaehr/gauss_newton_via_generalized_krylov_subspaces) for the path named
by BASELINE.json.  It is the *checker*: only ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import
it.  Nothing under ``gauss_newton_via_generalized_krylov_subspaces_b200/`` does.

Pinning: the reference has no tests, golden vectors or fixtures of its own
(SURVEY.md section 4), so this restatement is pinned against outputs of the
unmodified reference run in the build container: ``oracle/gen_golden.py``
imports ``/root/reference`` and writes ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` replays every fixture through this file.

Third-party arithmetic the reference leans on (not vendored, unpinned by the
reference; de-facto pin = this image: numpy 2.3.5, scipy 1.18.1):
  * scipy.linalg.qr / solve_triangular      (gauss_newton_krylow.py:30,35)
  * scipy.sparse.linalg.cg                  (gauss_newton.py:46,56) -- restated
    below from scipy 1.18.1 ``_isolve/iterative.py:cg`` (published algorithm:
    preconditioned conjugate gradients, x0 = 0, stop when ||r||_2 < max(atol,
    rtol*||b||_2) tested at the top of every iteration, maxiter = 10 n)
  * scipy.linalg.lstsq (gelsd)              (gauss_newton.py:116)

Layout notes: the Bratu unknown vector is the m x m interior grid flattened with
index ``i*m + j`` (i = x1 index, slow; j = x2 index, fast), m = grid_nodes - 1
(bratu_pde_problem.py:52-74).
"""
from __future__ import annotations

import numpy as np
import scipy.linalg
import scipy.sparse as sp


# ----------------------------------------------------------------------------
# Bratu problem, matrix free           (bratu_pde_problem.py:20-99)
# ----------------------------------------------------------------------------
class StencilJacobian:
    """J(u) = -(L + alpha D + lambda diag(e^u))  (bratu_pde_problem.py:88-96)
    kept as the three stencil constants and the e^u diagonal."""

    def __init__(self, prob, expu, transposed=False):
        self.prob = prob
        self.expu = expu
        self.transposed = transposed
        self.shape = (prob.n, prob.n)

    @property
    def T(self):
        return StencilJacobian(self.prob, self.expu, not self.transposed)

    def __matmul__(self, v):
        v = np.asarray(v)
        if v.ndim == 1:
            return self._apply(v[:, None])[:, 0]
        return self._apply(v)

    def __rmul__(self, s):  # "-1 * J" (gauss_newton.py:113)
        return ScaledOp(self, s)

    def _apply(self, V):
        p = self.prob
        m = p.m
        k = V.shape[1]
        W = np.ascontiguousarray(V.T).reshape(k, m, m)  # [col, i, j]
        out = np.empty_like(W)
        dg = 4.0 * p.c_lap - p.c_adv
        if self.expu is not None:
            dg = dg + p.lam * self.expu.reshape(1, m, m)
        np.multiply(W, dg, out=out)
        # i-1 / i+1 neighbours: the forward difference couples row i to i+1
        lo = -p.c_lap
        hi = p.c_adv - p.c_lap
        if self.transposed:
            lo, hi = hi, lo
        out[:, 1:, :] += lo * W[:, :-1, :]
        out[:, :-1, :] += hi * W[:, 1:, :]
        out[:, :, 1:] -= p.c_lap * W[:, :, :-1]
        out[:, :, :-1] -= p.c_lap * W[:, :, 1:]
        np.negative(out, out=out)
        return np.ascontiguousarray(out.reshape(k, m * m).T)

    def normal_diagonal(self):
        """diag(J^T J) = squared column norms (replaces the A.T @ A SpGEMM of
        gauss_newton.py:50-52)."""
        p = self.prob
        m = p.m
        dg = np.full((m, m), 4.0 * p.c_lap - p.c_adv)
        if self.expu is not None:
            dg = dg + p.lam * self.expu.reshape(m, m)
        lo = -p.c_lap
        hi = p.c_adv - p.c_lap
        if self.transposed:
            lo, hi = hi, lo
        out = dg * dg
        out[1:, :] += hi * hi      # row i-1 holds coefficient `hi` for column (i,j)
        out[:-1, :] += lo * lo     # row i+1 holds coefficient `lo`
        out[:, 1:] += p.c_lap ** 2
        out[:, :-1] += p.c_lap ** 2
        return out.reshape(-1)

    def tocsr(self):
        """Assembled CSR copy (used to check the matrix-free form)."""
        p = self.prob
        m = p.m
        e = np.ones(m)
        l1 = sp.diags_array((-e[:-1], 2 * e, -e[:-1]), offsets=(-1, 0, 1))
        lap = (sp.kron(l1, sp.eye(m)) + sp.kron(sp.eye(m), l1)) * p.c_lap
        adv = sp.kron(sp.diags_array((-e, e[:-1]), offsets=(0, 1)), sp.eye(m)) * p.c_adv
        J = lap + adv
        if self.expu is not None:
            J = J + p.lam * sp.diags(self.expu)
        J = (-1 * J).tocsr()
        return J.T.tocsr() if self.transposed else J


class ScaledOp:
    def __init__(self, op, s):
        self.op, self.s = op, s
        self.shape = op.shape

    @property
    def T(self):
        return ScaledOp(self.op.T, self.s)

    def __matmul__(self, v):
        return self.s * (self.op @ v)

    def normal_diagonal(self):
        return self.s ** 2 * self.op.normal_diagonal()


class BratuOracle:
    """-Lap(u) + alpha du/dx1 + lambda e^u = f on [lb,ub]^2, zero Dirichlet,
    5-point Laplacian, forward difference in the slow index
    (bratu_pde_problem.py:43-74)."""

    def __init__(self, grid_nodes, alpha, lam, lower=-3.0, upper=3.0, h=None, u_fn=None):
        self.G = int(grid_nodes)
        self.m = self.G - 1
        self.n = self.m * self.m
        self.alpha = alpha
        self.lam = lam
        self.h = (upper - lower) / grid_nodes if h is None else h
        # same host-side constants as the reference: h**-2, h**-1 via ** (:58,:67)
        self.c_lap = self.h ** -2
        self.c_adv = alpha * self.h ** -1
        t = np.linspace(lower, upper, grid_nodes + 1)[1:-1]
        self._t = t
        self._u_fn = u_fn
        self._u_true = None

    @property
    def u_true(self):
        # meshgrid(t,t) gives X[a,b]=t[b], Y[a,b]=t[a]; flatten("F") puts index
        # b*m + a  ->  u_true[i*m+j] = u(t[i], t[j])      (:69-74)
        if self._u_true is None:
            t = self._t
            x1 = t[:, None]
            x2 = t[None, :]
            if self._u_fn is None:
                U = np.exp(-10 * (x1 ** 2 + x2 ** 2))
            else:
                U = self._u_fn(np.broadcast_to(x1, (self.m, self.m)), np.broadcast_to(x2, (self.m, self.m)))
            self._u_true = np.ascontiguousarray(U).reshape(-1)
        return self._u_true

    def operator(self, u):
        """P(u) = L u + alpha D u + lambda e^u   (:76-83), evaluated in the SAME floating-point order as the
        reference so that y = P(u_true) -- an input of every parity run -- is bit-identical to the reference's:
          * laplace2d is a sorted CSR matrix with data {4 h^-2, -h^-2}; scipy's csr_matvec adds data*x one entry at a
            time in column order (i-1,j), (i,j-1), (i,j), (i,j+1), (i+1,j), no FMA;
          * ALPHA*partial_diff_x is COO with data -+ALPHA h^-1; coo_matvec adds -c u_ij, then +c u_i+1,j;
          * the three terms are then added left to right (:79-83).
        (h^-2 (4u - sum nb) of a smooth u cancels 4-5 digits, so a different order changes y by ~1e-13 relative, which
        the GNK trajectory amplifies to 1e-9 .. 1e-6 on the 1024^2 / 4096^2 grids.)"""
        m = self.m
        U = np.asarray(u, dtype=np.float64).reshape(m, m)
        c = self.c_lap
        z = np.zeros((m, m))
        up, left, right, down = z.copy(), z.copy(), z.copy(), z.copy()
        up[1:, :] = (-c) * U[:-1, :]
        left[:, 1:] = (-c) * U[:, :-1]
        right[:, :-1] = (-c) * U[:, 1:]
        down[:-1, :] = (-c) * U[1:, :]
        lap = up + left
        lap += (4.0 * c) * U
        lap += right
        lap += down
        adv = (-self.c_adv) * U
        adv[:-1, :] += self.c_adv * U[1:, :]
        out = lap + adv
        if self.lam != 0:
            out = out + self.lam * np.exp(U)
        return out.reshape(-1)

    def make_res(self, y):
        y = np.asarray(y, dtype=np.float64)
        return lambda u: y - self.operator(u)

    def make_jac(self):
        if self.lam == 0:
            return lambda u: StencilJacobian(self, None)
        return lambda u: StencilJacobian(self, np.exp(np.asarray(u, dtype=np.float64)))

    def make_error(self):
        return lambda u: float(np.linalg.norm(self.u_true - u))

    def start_vector(self, seed=42, noise=0.1):
        """u0 of the reference experiments (bratu_pde_test.py:34-37): legacy
        MT19937 global seed, drawn on the host."""
        rs = np.random.RandomState(seed)
        return self.u_true + noise * rs.normal(loc=0, scale=1, size=self.n)


# ----------------------------------------------------------------------------
# chained Rosenbrock                    (rosenbrock_problem.py:5-26)
# ----------------------------------------------------------------------------
def rosenbrock_res(x):
    a = 10.0 * (x[1:] - x[:-1] ** 2)
    b = 1.0 - x[:-1]
    return 2 ** 0.5 * np.concatenate([a, b])


def rosenbrock_jac(x, dense=False):
    p = x.shape[0]
    q = p - 1
    rows = np.concatenate([np.arange(q), np.arange(q), q + np.arange(q)])
    cols = np.concatenate([np.arange(q), 1 + np.arange(q), np.arange(q)])
    vals = 2 ** 0.5 * np.concatenate([-20.0 * x[:-1], np.full(q, 10.0), np.full(q, -1.0)])
    J = sp.coo_array((vals, (rows, cols)), shape=(2 * q, p)).tocsr()
    return J.toarray() if dense else J


# ----------------------------------------------------------------------------
# building blocks
# ----------------------------------------------------------------------------
class StepLengthFailure(RuntimeError):
    """armijo_goldstein.py:8-13,64-70"""

    def __init__(self, dnorm):
        super().__init__(f"armijo: 100 rejections, |d| = {dnorm}")
        self.dnorm = dnorm


def armijo(res, x, r, Jd_sq, args, d, max_trials=100, s0=1.0):
    """armijo_goldstein.py:47-72; ``Jd_sq`` = sum((J d)^2) is passed in.
    Returns (step, r_at_step, trials)."""
    s = s0
    prev = np.sum(r ** 2)
    for t in range(max_trials):
        rt = res(x + s * d, *args)
        cur = np.sum(rt ** 2)
        if prev - cur >= 0.5 * s * Jd_sq:
            return s, rt, t + 1
        s /= 2
    raise StepLengthFailure(float(np.linalg.norm(d)))


def ls_qr(A, y, log=None):
    """linear_least_squares, gauss_newton_krylow.py:30-35."""
    q, r = scipy.linalg.qr(A, mode="economic")
    for rkk in np.diagonal(r):
        if abs(rkk) <= 1e-8 and log is not None:
            log.append("A is rank deficient")
    return scipy.linalg.solve_triangular(r, q.T @ y)


def pcg(matvec, b, minv=None, rtol=1e-5, maxiter=None, x0=None):
    """scipy 1.18.1 sparse.linalg.cg restated (x0 = 0 unless given: then x = x0.copy(), r = b - A x0, and the
    tolerance stays relative to |b|).  Returns (x, n_callbacks)."""
    bn = np.linalg.norm(b)
    atol = max(0.0, rtol * bn)
    if bn == 0:
        return b, 0
    n = b.shape[0]
    if maxiter is None:
        maxiter = 10 * n
    if x0 is None:
        x = np.zeros_like(b)
        r = b.copy()
    else:
        x = np.array(x0, dtype=np.float64)
        r = b - matvec(x)
    rho_prev = None
    p = None
    its = 0
    for it in range(maxiter):
        if np.linalg.norm(r) < atol:
            break
        z = r if minv is None else minv * r
        rho = np.dot(r, z)
        if it > 0:
            p *= rho / rho_prev
            p += z
        else:
            p = z.copy()
        q = matvec(p)
        a = rho / np.dot(p, q)
        x += a * p
        r -= a * q
        rho_prev = rho
        its += 1
    return x, its


def cgls(A, y, rtol=1e-4, preconditioner=True, x0=None):
    """cg_least_squares, gauss_newton.py:11-60, including its quirk: with
    preconditioner=False an unpreconditioned CG runs first, is discarded, and
    the Jacobi-preconditioned CG always runs; cg_iter is the sum.  ``x0`` is
    the initial guess both runs are given (:46,:56)."""
    AT = A.T
    mv = lambda v: AT @ (A @ v)
    b = AT @ y
    total = 0
    if not preconditioner:
        _, its = pcg(mv, b, None, rtol, x0=x0)
        total += its
    if sp.issparse(A):
        diag = np.asarray(A.multiply(A).sum(axis=0)).reshape(-1)
    elif isinstance(A, np.ndarray):
        diag = np.sum(A * A, axis=0)
    else:
        diag = A.normal_diagonal()
    x, its = pcg(mv, b, 1.0 / diag, rtol, x0=x0)
    return x, total + its


def _is_sparse_like(J):
    return sp.issparse(J) or isinstance(J, (StencilJacobian, ScaledOp))


# ----------------------------------------------------------------------------
# solvers
# ----------------------------------------------------------------------------
def gnk(res, x0, jac, restart=None, args=(), tol=1e-8, max_iter=100, callback=None,
        version="res_old", reorth=1, ls=ls_qr, trace=None):
    """Gauss-Newton on generalized Krylov subspaces, gauss_newton_krylow.py:39-145
    with krylow.py:30-73 inlined.  Returns dict(x, success, nfev, njev, nit, log)."""
    log = []
    n = x0.shape[0]
    if restart is None:
        restart = max_iter
    cap = min(n, min(max_iter, restart) + 1)
    V = np.zeros((n, cap))

    def begin(x):
        if np.all(np.abs(x) <= 1e-8):  # krylow.py:31
            raise ValueError("x0 is not allowed to be 0 in the gauss_newton_krylow algorithm")
        nx = np.linalg.norm(x)
        V[:, 0] = x / nx
        return 1, np.array([nx])

    k, c = begin(x0)
    rk = lambda cc, *a: res(V[:, :cc.shape[0]] @ cc, *a)
    r_new = rk(c, *args)
    nfev, njev = 1, 1
    J = jac(x0, *args)
    success = False
    it = 0
    for it in range(1, max_iter):
        JV = J @ V[:, :k]
        r = r_new
        d = ls(-1 * JV, r, log)
        s, r_new, trials = armijo(rk, c, r, np.sum((JV @ d) ** 2), args, d)
        nfev += trials
        cprev = np.sum(c ** 2)
        c = c + s * d
        x = V[:, :k] @ c
        if trace is not None:
            trace.append(dict(it=it, k=k, step=s, trials=trials, loss=float(np.sum(r_new ** 2)),
                              dnorm2=float(np.sum(d ** 2)), cprev=float(cprev)))
        if callback is not None:
            callback(x=x, nfev=nfev, cg_iter=None)
        if s ** 2 * np.sum(d ** 2) <= tol ** 2 * cprev:
            success = True
            break
        J_old = J
        J = jac(x, *args)
        njev += 1
        if version == "res_old":
            Ju, ru = J, r
        elif version == "res_new":
            Ju, ru = J, r_new
        elif version == "jac_old_res_old":
            Ju, ru = J_old, r
        elif version == "jac_old_res_new":
            Ju, ru = J_old, r_new
        else:
            raise ValueError("Variable version must be in ['res_old','res_new','jac_old_res_old','jac_old_res_new']")
        if k == n:  # krylow.py:59
            log.append(f"spans entire space at iteration = {it}")
        else:
            w = -(Ju.T @ ru)
            for _ in range(reorth):
                w = w - V[:, :k] @ (V[:, :k].T @ w)
            if np.all(np.abs(w) <= 1e-8):  # krylow.py:66
                log.append(f"breakdown at iteration = {it}, basis.shape = ({n}, {k})")
            else:
                if k == V.shape[1]:
                    V = np.concatenate([V, np.zeros((n, V.shape[1]))], axis=1)
                V[:, k] = w / np.linalg.norm(w)
                k += 1
                c = np.append(c, 0.0)
        if it % restart == 0:
            k, c = begin(V[:, :k] @ c)
    x = V[:, :k] @ c
    return dict(x=x, success=success, nfev=nfev, njev=njev, nit=it, log=log)


def gn(res, x0, jac, args=(), tol=1e-8, max_iter=100, step_control=None, callback=None,
       cg_preconditioner=False, trace=None):
    """Full-space Gauss-Newton, gauss_newton.py:63-138."""
    x = np.array(x0, dtype=np.float64, copy=True)
    r = res(x, *args)
    nfev, njev = 1, 0
    success = False
    cg_iter = None
    it = 0
    for it in range(1, max_iter):
        J = jac(x, *args)
        njev += 1
        if _is_sparse_like(J):
            d, cg_iter = cgls(-1 * J, r, preconditioner=cg_preconditioner)
        else:
            J = np.asarray(J)
            d = scipy.linalg.lstsq(-1 * J, r)[0]
        if step_control is None:
            s, r, trials = armijo(res, x, r, np.sum((J @ d) ** 2), args, d)
        else:
            s, r, trials = step_control(res, x, r, J, args, d)
        nfev += trials
        xprev = np.sum(x ** 2)
        x += s * d
        if trace is not None:
            trace.append(dict(it=it, step=s, trials=trials, cg_iter=cg_iter))
        if callback is not None:
            callback(x=x, nfev=nfev, cg_iter=cg_iter)
        if s ** 2 * np.sum(d ** 2) <= tol ** 2 * xprev:
            success = True
            break
    return dict(x=x, success=success, nfev=nfev, njev=njev, nit=it)
