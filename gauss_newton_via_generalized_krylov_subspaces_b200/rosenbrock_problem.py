"""Chained Rosenbrock problem -- mirror of the reference's ``rosenbrock_problem.py`` (p = 1000 parameters,
1998 residuals, 2997 non-zeros).  ``res`` / ``jac`` are host callables with the reference's signatures and can be
called like the reference's.  When exactly these two functions are handed to ``gauss_newton_krylow`` the solver uses
their device-native twins (``RosenbrockDeviceProblem``: csrc/rosenbrock.cu), so the iterate never leaves HBM and
no Jacobian is built on or uploaded from the host; any other callable still runs on the host (it is user code)."""
import ctypes as C

import numpy as np
import scipy.sparse

parameter_count = 1000  # N = 2p - 2 residuals


def res(x):
    x = np.asarray(x)
    head = x[:-1]
    return 2**0.5 * np.concatenate([10 * (x[1:] - head**2), 1 - head])


def jac(x):
    x = np.asarray(x)
    q = parameter_count - 1
    i = np.arange(q)
    rows = np.concatenate([i, i, q + i])
    cols = np.concatenate([i, i + 1, i])
    vals = 2**0.5 * np.concatenate([-20.0 * x[:-1], np.full(q, 10.0), np.full(q, -1.0)])
    return scipy.sparse.coo_array((vals, (rows, cols)), shape=(2 * q, parameter_count))


x_exact = np.ones(parameter_count)


def error(x):
    return np.linalg.norm(np.asarray(x) - x_exact)


class _DeviceCsr:
    """J(x) of the chained Rosenbrock problem on the device: fixed index arrays (shared), per-evaluation values."""

    is_sparse = True
    transposed = False
    scale = 1.0

    def __init__(self, owner, val, val_t):
        self.o, self.val, self.val_t = owner, val, val_t

    def matmat(self, V, ldv, k, JV, ldjv):
        from . import _lib
        from .device import ptr
        o, rt = self.o, self.o.rt
        _lib.check(rt.lib.gnk_spmm_csr(rt.ctx, o.n_res, ptr(o.rowptr), ptr(o.col), ptr(self.val), ptr(V), ldv, 0, k, 1.0,
                                       ptr(JV), ldjv, 0, rt.stream), "gnk_spmm_csr")

    def neg_rmatvec(self, r, w):
        from . import _lib
        from .device import ptr
        o, rt = self.o, self.o.rt
        _lib.check(rt.lib.gnk_spmm_csr(rt.ctx, o.p_glob, ptr(o.rowptr_t), ptr(o.col_t), ptr(self.val_t), ptr(r), 0, 0, 1,
                                       -1.0, ptr(w), 0, 0, rt.stream), "gnk_spmm_csr(T)")


class RosenbrockDeviceProblem:
    """Problem adapter used by ``gauss_newton_krylow`` for (res, jac) of this module (rosenbrock_problem.py:8-19 on
    the device): F = res(x) by gnk_rosenbrock_residual, the values of J(x) and J(x)^T in CSR form by
    gnk_rosenbrock_jacobian, products by the CSR kernels.  Same arithmetic, bit for bit, as the host functions."""

    distributed = False
    device_native = True

    def __init__(self, x0):
        from .device import get_runtime, make_layout
        from .partition import flat_layout_fields
        self.rt = rt = get_runtime()
        self.p_glob = p = int(np.asarray(x0).shape[0]) if not hasattr(x0, "n_global") else int(x0.n_global)
        if p != parameter_count:  # the reference bakes parameter_count into jac (:15-19)
            raise ValueError(f"rosenbrock_problem.jac is defined for {parameter_count} parameters, got {p}")
        q = p - 1
        self.n_res = 2 * q
        self.sol_fields = flat_layout_fields(p)
        self.sol = make_layout(self.sol_fields)
        self.res_fields = flat_layout_fields(self.n_res)
        self.res_lay = make_layout(self.res_fields)
        i = np.arange(q, dtype=np.int32)
        rowptr = np.concatenate([2 * np.arange(q + 1), 2 * q + 1 + np.arange(q)]).astype(np.int32)
        col = np.concatenate([np.stack([i, i + 1], 1).reshape(-1), i]).astype(np.int32)
        rowptr_t = np.concatenate([[0], 3 * np.arange(1, q + 1) - 1, [3 * q]]).astype(np.int32)
        col_t = np.empty(3 * q, dtype=np.int32)
        col_t[0], col_t[1] = 0, q
        j = np.arange(1, q, dtype=np.int32)
        col_t[3 * j - 1], col_t[3 * j], col_t[3 * j + 1] = j - 1, j, q + j
        col_t[3 * q - 1] = q - 1
        t = rt.torch
        self.rowptr, self.col, self.rowptr_t, self.col_t = (t.from_numpy(a).to(rt.device)
                                                             for a in (rowptr, col, rowptr_t, col_t))
        self.sqrt2 = 2 ** 0.5

    @staticmethod
    def match(res_fn, jac_fn, args):
        return res_fn is res and jac_fn is jac and not args

    def new_sol(self):
        return self.rt.zeros(self.sol_fields["ld"])

    def new_res(self):
        return self.rt.zeros(self.res_fields["ld"])

    def upload_x(self, x_host, out):
        self.rt.upload(np.asarray(x_host, dtype=np.float64).reshape(-1), out[:self.p_glob])

    def download_global(self, t):
        return self.rt.download(t[:self.p_glob])

    def sumsq(self, vec, slot2, lay):
        from . import _lib
        from .device import ptr
        rt = self.rt
        _lib.check(rt.lib.gnk_norm_stats(rt.ctx, C.byref(lay), ptr(vec), ptr(slot2), rt.stream), "gnk_norm_stats")

    def residual(self, x, F, loss_slot, aux=None):
        from . import _lib
        from .device import ptr
        rt = self.rt
        _lib.check(rt.lib.gnk_rosenbrock_residual(rt.ctx, self.p_glob, self.sqrt2, ptr(x), ptr(F), rt.stream),
                   "gnk_rosenbrock_residual")
        self.sumsq(F, loss_slot, self.res_lay)

    def jacobian(self, x, aux=None):
        from . import _lib
        from .device import ptr
        rt = self.rt
        val, val_t = rt.empty(3 * (self.p_glob - 1)), rt.empty(3 * (self.p_glob - 1))
        _lib.check(rt.lib.gnk_rosenbrock_jacobian(rt.ctx, self.p_glob, self.sqrt2, ptr(x), ptr(val), ptr(val_t),
                                                  rt.stream), "gnk_rosenbrock_jacobian")
        return _DeviceCsr(self, val, val_t)
