"""numpy stand-in for libgnk_b200.so -- TEST INFRASTRUCTURE ONLY.

It implements the C ABI of include/gnk_b200.h on host memory so that the Python host logic (solver control
flow, buffer rotation, slab partition, halo exchange, reduction order) can be tested without a GPU, including
a world_size-2 run over gloo.  It is injected explicitly (``install()``); the product never imports it and never
falls back to it -- without the CUDA library and a device the package raises.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import gauss_newton_via_generalized_krylov_subspaces_b200.device as device  # noqa: E402
from gauss_newton_via_generalized_krylov_subspaces_b200 import _lib  # noqa: E402


def arr(p, n, dtype=np.float64):
    """view n items at the address held by a c_void_p"""
    addr = p.value if isinstance(p, C.c_void_p) else p
    if not addr:
        return None
    ct = {np.float64: C.c_double, np.int32: C.c_int32, np.int64: C.c_int64}[dtype]
    return np.ctypeslib.as_array((ct * int(n)).from_address(addr))


def isnull(p):
    return p is None or (isinstance(p, C.c_void_p) and not p.value) or p == 0


def obj(ref):
    return ref._obj if hasattr(ref, "_obj") else ref


class MockLib:
    def __init__(self, dist=None):
        self.dist = dist
        self.launches = 0
        self.rank, self.world = 0, 1

    # ---- lifecycle ----
    def gnk_abi_version(self):
        return 1

    def gnk_last_error(self):
        return b"mock"

    def gnk_launch_count(self, ctx):
        return self.launches

    def gnk_sm_count(self, ctx):
        return 148

    def gnk_scalars_fetch(self, ctx, d_src, count, h_dst, stream):
        arr(h_dst, count)[:] = arr(d_src, count)   # the mock has no streams: the copy is immediate
        return 0

    def gnk_scalars_wait(self, ctx):
        return 0

    # ---- bratu ----
    def _grid(self, lay, col, lo, hi):
        """rows [lo, hi) of a stored column as a 2-D view (row index relative to the first owned row)"""
        m = lay.m
        a = arr(col, lay.ld)
        return a[lay.off + lo * m: lay.off + hi * m].reshape(hi - lo, m)

    def gnk_bratu_residual(self, ctx, lay, prm, u, y, F, expu, depth, loss, stream):
        lay, prm = obj(lay), obj(prm)
        self.launches += 1
        rows, m = lay.rows, lay.m
        U = self._grid(lay, u, -depth - 1, rows + depth + 1)
        mid = U[1:-1]
        P = 4 * prm.c_lap * mid - prm.c_lap * (U[:-2] + U[2:]) + prm.c_adv * (U[2:] - mid)
        P[:, 1:] -= prm.c_lap * mid[:, :-1]
        P[:, :-1] -= prm.c_lap * mid[:, 1:]
        e = np.ones_like(mid)
        if prm.lam != 0:
            e = np.exp(mid)
            P += prm.lam * e
        Fv = self._grid(lay, y, -depth, rows + depth) - P
        if depth:
            if not lay.has_lo:
                Fv[0] = 0
            if not lay.has_hi:
                Fv[-1] = 0
        self._grid(lay, F, -depth, rows + depth)[:] = Fv
        if not isnull(expu):
            self._grid(lay, expu, -depth, rows + depth)[:] = e
        own = Fv[depth:depth + rows]
        arr(loss, 1)[0] = np.sum(own * own)
        return 0

    def gnk_stencil_apply(self, ctx, lay, prm, expu, inp, in_ld, k, sign, transpose, out, out_ld, out_off, stream):
        lay, prm = obj(lay), obj(prm)
        self.launches += 1
        rows, m = lay.rows, lay.m
        dg = 4 * prm.c_lap - prm.c_adv
        if not isnull(expu) and prm.lam != 0:
            dg = dg + prm.lam * self._grid(lay, expu, 0, rows)
        cu, cd = -prm.c_lap, prm.c_adv - prm.c_lap
        if transpose:
            cu, cd = cd, cu
        vin = arr(inp, in_ld * k)
        vout = arr(out, out_ld * (k - 1) + out_off + rows * m)
        for j in range(k):
            V = vin[j * in_ld + lay.off - m: j * in_ld + lay.off + (rows + 1) * m].reshape(rows + 2, m)
            mid = V[1:-1]
            o = dg * mid + cu * V[:-2] + cd * V[2:]
            o[:, 1:] -= prm.c_lap * mid[:, :-1]
            o[:, :-1] -= prm.c_lap * mid[:, 1:]
            vout[j * out_ld + out_off: j * out_ld + out_off + rows * m] = sign * o.reshape(-1)
        return 0

    def gnk_stencil_normal_diag(self, ctx, lay, prm, expu, out, stream):
        lay, prm = obj(lay), obj(prm)
        self.launches += 1
        rows, m = lay.rows, lay.m
        dg = np.full((rows, m), 4 * prm.c_lap - prm.c_adv)
        if not isnull(expu) and prm.lam != 0:
            dg = dg + prm.lam * self._grid(lay, expu, 0, rows)
        o = dg * dg
        cd2, cl2 = (prm.c_adv - prm.c_lap) ** 2, prm.c_lap ** 2
        o[1:] += cd2
        o[:-1] += cl2
        if lay.has_lo:
            o[0] += cd2
        if lay.has_hi:
            o[-1] += cl2
        o[:, 1:] += cl2
        o[:, :-1] += cl2
        self._grid(lay, out, 0, rows)[:] = o
        return 0

    # ---- basis ----
    def gnk_combine(self, ctx, lay, V, k, c, d, s, x, stream):
        return self.gnk_combine_step(ctx, lay, V, k, c, d, s, x, None, None, stream)

    def gnk_combine_step(self, ctx, lay, V, k, c, d, s, x, c_out, cprev2, stream):
        lay = obj(lay)
        self.launches += 1
        cv = arr(c, k).copy()
        coef = cv.copy()
        if not isnull(d):
            coef = coef + s * arr(d, k)
        Vm = arr(V, lay.ld * k).reshape(k, lay.ld)
        arr(x, lay.ld)[:] = coef @ Vm
        if not isnull(c_out):
            o = arr(c_out, k + 1)
            o[:k] = coef
            o[k] = 0.0
        if not isnull(cprev2):
            arr(cprev2, 1)[0] = np.dot(cv, cv)
        return 0

    def gnk_norm_stats(self, ctx, lay, x, stats, stream):
        lay = obj(lay)
        self.launches += 1
        v = arr(x, lay.ld)[lay.off:lay.off + lay.n_own]
        s = arr(stats, 2)
        s[0] = np.sum(v * v)
        s[1] = np.max(np.abs(v)) if v.size else 0.0
        return 0

    def gnk_normalize(self, ctx, lay, x, stats, atol, out, flag, stream):
        lay = obj(lay)
        self.launches += 1
        s = arr(stats, 2)
        if s[1] <= atol:
            arr(flag, 1, np.int32)[0] = 1
            return 0
        arr(flag, 1, np.int32)[0] = 0
        arr(out, lay.ld)[:] = arr(x, lay.ld) / np.sqrt(s[0])
        return 0

    def gnk_cgs_dots(self, ctx, lay, V, k, w, h, stream):
        lay = obj(lay)
        self.launches += 1
        Vm = arr(V, lay.ld * k).reshape(k, lay.ld)[:, lay.off:lay.off + lay.n_own]
        arr(h, k)[:] = Vm @ arr(w, lay.ld)[lay.off:lay.off + lay.n_own]
        return 0

    def gnk_cgs_update(self, ctx, lay, V, k, h, w, stats, stream):
        lay = obj(lay)
        self.launches += 1
        Vm = arr(V, lay.ld * k).reshape(k, lay.ld)[:, lay.off:lay.off + lay.n_own]
        wv = arr(w, lay.ld)[lay.off:lay.off + lay.n_own]
        wv -= arr(h, k) @ Vm
        if not isnull(stats):
            s = arr(stats, 2)
            s[0] = np.sum(wv * wv)
            s[1] = np.max(np.abs(wv)) if wv.size else 0.0
        return 0

    # ---- least squares: per-rank QR, gather of R factors, QR of the stack (the TSQR tree) ----
    def gnk_tsqr_ls(self, ctx, A, lda, n_rows, k, y, sign_a, out, stream):
        self.launches += 1
        Am = arr(A, lda * k).reshape(k, lda)[:, :n_rows].T * sign_a
        M = np.concatenate([Am, arr(y, n_rows)[:, None]], axis=1)
        R = np.linalg.qr(M, mode="r") if n_rows >= k + 1 else np.linalg.qr(
            np.concatenate([M, np.zeros((k + 1 - n_rows, k + 1))]), mode="r")
        if self.world > 1:
            stack = self._allgather(R.reshape(-1))
            R = np.linalg.qr(stack.reshape(-1, k + 1), mode="r")
        z = R[:k, k]
        import scipy.linalg
        try:
            d = scipy.linalg.solve_triangular(R[:k, :k], z)
        except Exception:  # exactly singular R: the CUDA kernel divides by zero and reports the zero diagonal
            d = np.full(k, np.inf)
        o = arr(out, 2 * k + 4)
        o[:k] = d
        o[k] = np.sum(z * z)
        o[k + 1] = R[k, k] ** 2
        o[k + 2] = np.sum(np.abs(np.diag(R)[:k]) <= 1e-8)
        o[k + 3] = np.sum(d * d)
        o[k + 4:2 * k + 4] = np.diag(R)[:k]
        return 0

    def gnk_tsqr_ls_method(self, ctx, method):
        prev = getattr(self, "ls_method", 0)
        self.ls_method = int(method)
        return prev

    def gnk_spmm_csr(self, ctx, n_rows, rowptr, col, val, inp, in_ld, in_off, k, sign, out, out_ld, out_off, stream):
        import scipy.sparse as sp
        self.launches += 1
        rp = arr(rowptr, n_rows + 1, np.int32)
        nnz = int(rp[-1])
        ci = arr(col, nnz, np.int32) if nnz else np.zeros(0, np.int32)
        va = arr(val, nnz) if nnz else np.zeros(0)
        ncol = int(ci.max()) + 1 if nnz else 1
        A = sp.csr_array((va, ci, rp), shape=(n_rows, ncol))
        vin = arr(inp, in_ld * (k - 1) + in_off + ncol)
        vout = arr(out, out_ld * (k - 1) + out_off + n_rows)
        for j in range(k):
            vout[j * out_ld + out_off: j * out_ld + out_off + n_rows] = sign * (
                A @ vin[j * in_ld + in_off: j * in_ld + in_off + ncol])
        return 0

    # ---- chained Rosenbrock (rosenbrock_problem.py:8-19) ----
    def gnk_rosenbrock_residual(self, ctx, p, sqrt2, x, F, stream):
        self.launches += 1
        xv = arr(x, p)
        head = xv[:-1]
        arr(F, 2 * p - 2)[:] = sqrt2 * np.concatenate([10 * (xv[1:] - head ** 2), 1 - head])
        return 0

    def gnk_rosenbrock_jacobian(self, ctx, p, sqrt2, x, val, val_t, stream):
        import scipy.sparse as sp
        self.launches += 1
        xv = arr(x, p)
        q = p - 1
        i = np.arange(q)
        rows = np.concatenate([i, i, q + i])
        cols = np.concatenate([i, i + 1, i])
        vals = sqrt2 * np.concatenate([-20.0 * xv[:-1], np.full(q, 10.0), np.full(q, -1.0)])
        A = sp.csr_array(sp.coo_array((vals, (rows, cols)), shape=(2 * q, p)))
        A.sort_indices()
        AT = sp.csr_array(A.T)
        AT.sort_indices()
        arr(val, 3 * q)[:] = A.data
        arr(val_t, 3 * q)[:] = AT.data
        return 0

    def gnk_csr_row_sumsq(self, ctx, n_rows, rowptr, val, out, stream):
        self.launches += 1
        rp = arr(rowptr, n_rows + 1, np.int32)
        va = arr(val, int(rp[-1])) if rp[-1] else np.zeros(0)
        o = arr(out, n_rows)
        for i in range(n_rows):
            o[i] = np.sum(va[rp[i]:rp[i + 1]] ** 2)
        return 0

    # ---- vector algebra ----
    def gnk_axpby(self, ctx, n, a, x, b, y, out, stream):
        self.launches += 1
        r = np.zeros(n)
        if a != 0:
            r = a * arr(x, n)
        if b != 0:
            r = r + b * arr(y, n)
        arr(out, n)[:] = r
        return 0

    def gnk_dot(self, ctx, n, x, y, out, stream):
        self.launches += 1
        arr(out, 1)[0] = np.dot(arr(x, n), arr(y, n))
        return 0

    def gnk_normalize_halo(self, ctx, lay, x, stats, atol, out, flag, stream):
        rc = self.gnk_normalize(ctx, lay, x, stats, atol, out, flag, stream)
        lo = obj(lay)
        if rc == 0 and self.world > 1 and lo.m > 0:
            return self.gnk_comm_halo_exchange(ctx, lay, out, lo.halo, stream)
        return rc

    def gnk_gram_cgls(self, ctx, A, lda, n_rows, k, y, sign_a, rtol, out, stream):
        """CG on the explicit normal equations, as the device kernel does (Gram matrices summed over the ranks)"""
        from oracle.gnk_oracle import pcg
        self.launches += 1
        Am = arr(A, lda * k).reshape(k, lda)[:, :n_rows].T * sign_a
        yv = arr(y, n_rows)
        G, b, yy = Am.T @ Am, Am.T @ yv, np.array([yv @ yv])
        if self.world > 1:
            packed = self._allgather(np.concatenate([G.reshape(-1), b, yy])).reshape(self.world, -1).sum(axis=0)
            G, b, yy = packed[:k * k].reshape(k, k), packed[k * k:k * k + k], packed[-1:]
        d, its = pcg(lambda v: G @ v, b, 1.0 / np.diag(G), rtol, maxiter=10 * k)
        o = arr(out, 2 * k + 5)
        o[:k] = d
        o[k] = d @ G @ d
        o[k + 1] = yy[0] - 2 * d @ b + d @ G @ d
        o[k + 2] = 0.0
        o[k + 3] = d @ d
        o[k + 4:2 * k + 4] = np.sqrt(np.diag(G))
        o[2 * k + 4] = its
        return 0

    def gnk_stencil_gram_ls(self, *a):
        return 1  # "not eligible": the host falls back to gnk_stencil_apply + gnk_tsqr_ls (which the mock implements)

    # ---- CGLS (scipy cg restated; single rank) ----
    def gnk_cgls(self, ctx, op, y, rtol, preconditioner, x, work, iters, stream):
        return self.gnk_cgls_x0(ctx, op, y, None, rtol, preconditioner, x, work, iters, stream)

    def gnk_cgls_x0(self, ctx, op, y, x0, rtol, preconditioner, x, work, iters, stream):
        op = obj(op)
        it_out = obj(iters)
        if op.kind == 0:
            lay, prm = op.lay, op.prm
            n, off, ld = lay.n_own, lay.off, lay.ld

            def apply(v, tr):
                vin = np.zeros(ld)
                vin[off:off + n] = v
                if self.world > 1:  # cgls.cu: Cg::apply exchanges one halo row of the input first
                    self.gnk_comm_halo_exchange(None, lay, C.c_void_p(vin.ctypes.data), 1, None)
                vout = np.zeros(ld)
                self.gnk_stencil_apply(None, lay, prm, C.c_void_p(op.d_expu or 0), C.c_void_p(vin.ctypes.data), ld, 1,
                                       op.sign, tr, C.c_void_p(vout.ctypes.data), ld, off, None)
                return vout[off:off + n].copy()
            yv = arr(y, ld)[off:off + n]
            dinv = np.zeros(ld)
            self.gnk_stencil_normal_diag(None, lay, prm, C.c_void_p(op.d_expu or 0), C.c_void_p(dinv.ctypes.data), None)
            minv = 1.0 / (op.sign ** 2 * dinv[off:off + n])
            xv = arr(x, ld)[off:off + n]
        else:
            import scipy.sparse as sp
            rp = arr(C.c_void_p(op.d_rowptr), op.n_res + 1, np.int32)
            nnz = int(rp[-1])
            A = op.sign * sp.csr_array((arr(C.c_void_p(op.d_val), nnz), arr(C.c_void_p(op.d_col), nnz, np.int32), rp),
                                       shape=(op.n_res, op.p))
            apply = lambda v, tr: (A.T @ v) if tr else (A @ v)  # noqa: E731
            yv = arr(y, op.n_res)
            minv = 1.0 / np.asarray(A.multiply(A).sum(axis=0)).reshape(-1)
            xv = arr(x, op.p)
        from oracle.gnk_oracle import pcg
        if self.world > 1:
            pcg = self._pcg_sharded
        b = apply(yv, 1)
        mv = lambda v: apply(apply(v, 0), 1)  # noqa: E731
        total = 0
        start = None if isnull(x0) else (arr(x0, ld)[off:off + n].copy() if op.kind == 0 else arr(x0, op.p).copy())
        if not preconditioner:
            _, its = pcg(mv, b, None, rtol, x0=start)
            total += its
        sol, its = pcg(mv, b, minv, rtol, x0=start)
        xv[:] = sol
        it_out.value = total + its
        return 0

    def _pcg_sharded(self, matvec, b, minv=None, rtol=1e-5, maxiter=None, x0=None):
        """oracle.gnk_oracle.pcg on row-sharded vectors: every inner product is summed over the ranks in rank order
        (cgls.cu: Cg::reduce), maxiter counts the GLOBAL unknowns"""
        dot = lambda u, v: float(self._allgather(np.array([np.dot(u, v)])).sum())  # noqa: E731
        bn = np.sqrt(dot(b, b))
        atol = max(0.0, rtol * bn)
        if bn == 0:
            return b, 0
        if maxiter is None:
            maxiter = 10 * int(self._allgather(np.array([float(b.shape[0])])).sum())
        if x0 is None:
            x, r = np.zeros_like(b), b.copy()
        else:
            x = np.array(x0, dtype=np.float64)
            r = b - matvec(x)
        rho_prev, p, its = None, None, 0
        for it in range(maxiter):
            if np.sqrt(dot(r, r)) < atol:
                break
            z = r if minv is None else minv * r
            rho = dot(r, z)
            p = z.copy() if it == 0 else p * (rho / rho_prev) + z
            q = matvec(p)
            a = rho / dot(p, q)
            x += a * p
            r -= a * q
            rho_prev = rho
            its += 1
        return x, its

    # ---- comm over torch.distributed (gloo) ----
    def _allgather(self, v):
        import torch
        t = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float64))
        outs = [torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(outs, t)
        return np.concatenate([o.numpy() for o in outs])

    def gnk_comm_size(self, ctx):
        return self.world

    def gnk_comm_allreduce(self, ctx, buf, count, op, stream):
        if self.world == 1:
            return 0
        v = arr(buf, count)
        g = self._allgather(v).reshape(self.world, count)
        acc = g[0].copy()
        for r in range(1, self.world):
            for i in range(count):
                if op == 1 or (op == 2 and i == 1):
                    acc[i] = max(acc[i], g[r, i])
                else:
                    acc[i] = acc[i] + g[r, i]
        v[:] = acc
        return 0

    def gnk_comm_halo_exchange(self, ctx, lay, col, depth, stream):
        import torch
        lay = obj(lay)
        if self.world == 1:
            return 0
        a = arr(col, lay.ld)
        cnt = depth * lay.m
        off = lay.off
        ops = []
        bufs = []
        if lay.has_lo:
            send = torch.from_numpy(a[off:off + cnt].copy())
            recv = torch.empty(cnt, dtype=torch.float64)
            ops += [self.dist.P2POp(self.dist.isend, send, self.rank - 1), self.dist.P2POp(self.dist.irecv, recv, self.rank - 1)]
            bufs.append((off - cnt, recv))
        if lay.has_hi:
            send = torch.from_numpy(a[off + (lay.rows - depth) * lay.m: off + lay.rows * lay.m].copy())
            recv = torch.empty(cnt, dtype=torch.float64)
            ops += [self.dist.P2POp(self.dist.isend, send, self.rank + 1), self.dist.P2POp(self.dist.irecv, recv, self.rank + 1)]
            bufs.append((off + lay.rows * lay.m, recv))
        for r in self.dist.batch_isend_irecv(ops):
            r.wait()
        for pos, recv in bufs:
            a[pos:pos + cnt] = recv.numpy()
        return 0

    def gnk_comm_allgather_owned(self, ctx, lay, col, full, counts, stream):
        import torch
        lay = obj(lay)
        own = arr(col, lay.ld)[lay.off:lay.off + lay.n_own]
        cs = [int(counts[r]) for r in range(self.world)]
        out = arr(full, sum(cs))
        if self.world == 1:
            out[:] = own
            return 0
        mx = max(cs)
        pad = np.zeros(mx)
        pad[:own.shape[0]] = own
        g = self._allgather(pad).reshape(self.world, mx)
        pos = 0
        for r in range(self.world):
            out[pos:pos + cs[r]] = g[r, :cs[r]]
            pos += cs[r]
        return 0


class MockRuntime(device.Runtime):
    def __init__(self):
        import torch
        self.torch = torch
        dist = torch.distributed
        self.lib = MockLib(dist)
        self.device = torch.device("cpu")
        self.device_index = 0
        self.ctx = C.c_void_p(0)
        self.rank, self.world = 0, 1
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            self.rank, self.world = dist.get_rank(), dist.get_world_size()
            self.lib.rank, self.lib.world = self.rank, self.world
        self._pinned = torch.empty(4096, dtype=torch.float64)
        self._pinned_np = self._pinned.numpy()
        self._pinned_ptr = self._pinned.data_ptr()
        self._pinned2 = torch.empty(1024, dtype=torch.float64)
        self._pinned2_np = self._pinned2.numpy()
        self._pinned2_ptr = self._pinned2.data_ptr()
        self._pinned_i = torch.empty(16, dtype=torch.int32)

    @property
    def stream(self):
        return 0

    def sync(self):
        pass

    def pinned(self, n, dtype=None):
        return self.torch.empty(int(n), dtype=dtype or self.torch.float64)

    def download(self, t):
        return t.detach().clone().numpy()

    def host_register(self, addr, nbytes):
        pass

    def launches(self):
        return self.lib.launches


def install():
    """route the package's host code to the numpy mock (tests only)"""
    device._runtime = MockRuntime()
    _lib.check = lambda rc, what="": None if rc == 0 else (_ for _ in ()).throw(_lib.GnkError(what))
    return device._runtime


def uninstall():
    device._runtime = None
