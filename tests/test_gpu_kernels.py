"""Kernel-level parity (-m gpu): every C-ABI entry point against the oracle / the reference's golden vectors.

Tolerances: all arithmetic is fp64; a kernel differs from the numpy/scipy evaluation only by summation order
and FMA contraction, so element-wise results are compared at 1e-13 relative to the vector's max-norm and
reductions over n terms at 1e-12 relative.
"""
import ctypes as C
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from golden_util import Golden, rel  # noqa: E402
from oracle import gnk_oracle as orc  # noqa: E402


@pytest.fixture(scope="module")
def g():
    import gauss_newton_via_generalized_krylov_subspaces_b200 as pkg
    import gauss_newton_via_generalized_krylov_subspaces_b200.device as device
    if os.environ.get("GNK_TEST_MOCK"):  # debugging aid for the test code itself; never set by the driver
        import mock_backend
        mock_backend.install()
        return pkg
    device._runtime = None
    pkg.get_runtime()  # raises without CUDA + the built library: no fallback
    return pkg


def _lib_mods():
    from gauss_newton_via_generalized_krylov_subspaces_b200 import _lib, device
    return _lib, device


def test_library_is_native(g):
    if os.environ.get("GNK_TEST_MOCK"):
        pytest.skip("mock")
    rt = g.get_runtime()
    assert rt.lib.gnk_abi_version() == 1
    assert rt.lib.gnk_sm_count(rt.ctx) >= 100
    import torch
    assert torch.cuda.get_device_capability(rt.device_index)[0] >= 10


# ------------------------------------------------------------------------------------------------
# Bratu stencil kernels vs the reference's scipy operators (golden) and the oracle
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["g11", "g10", "g26lin", "g33h1", "g101"])
def test_stencil_against_reference_golden(g, tag):
    z = Golden("kernels")
    G, a, l, h = z[f"{tag}/params"]
    pb = g.BratuPdeProblem(int(G), a, l, grid_resolution=None if h < 0 else h)
    u, V, r = z[f"{tag}/u"], z[f"{tag}/V"], z[f"{tag}/r"]
    assert rel(pb.u_true, z[f"{tag}/u_true"]) == 0.0
    assert rel(pb.pde_operator(u), z[f"{tag}/P"]) < 1e-13
    J = pb.make_jac()(u)
    assert rel(J @ V, z[f"{tag}/JV"]) < 1e-13
    assert rel(J.T @ r, z[f"{tag}/JTr"]) < 1e-13
    assert rel((-1 * J) @ r, -(J @ r)) == 0.0
    # diag(J^T J) through the C ABI
    _lib, device = _lib_mods()
    d = pb.dev
    out = d.new_col()
    _lib.check(d.rt.lib.gnk_stencil_normal_diag(d.rt.ctx, C.byref(d.lay), C.byref(d.prm), device.ptr(J.expu),
                                                device.ptr(out), d.rt.stream))
    assert rel(d.download_global(out), z[f"{tag}/JTJdiag"]) < 1e-13


@pytest.mark.parametrize("G,lam", [(12, 10.0), (9, 10.0), (130, 10.0), (258, 3.0), (131, 0.0)])
@pytest.mark.parametrize("depth", [0, 1])
def test_fused_residual(g, G, lam, depth):
    """F, e^u and sum(F^2) in one pass, on owned rows and on the halo rows (zero outside the domain)."""
    _lib, device = _lib_mods()
    o = orc.BratuOracle(G, 5, lam)
    pb = g.BratuPdeProblem(G, 5, lam)
    rs = np.random.RandomState(G)
    u = 0.5 * rs.normal(size=o.n)
    y = rs.normal(size=o.n)
    d = pb.dev
    x, ycol, F, E = d.new_col(), d.new_col(), d.new_col(), d.new_col()
    d.upload_x(u, x)
    d.upload_x(y, ycol)
    loss = d.rt.zeros(2)
    d.residual_into(x, ycol, F, E, loss, depth=depth)
    Fref = y - o.operator(u)
    assert rel(d.download_global(F), Fref) < 1e-13
    if lam != 0:
        assert rel(d.download_global(E), np.exp(u)) < 1e-14
    got = float(d.rt.read(loss, 1)[0])
    assert abs(got - np.sum(Fref ** 2)) <= 1e-12 * np.sum(Fref ** 2)
    # single rank: the halo rows lie outside the domain -> F must be exactly zero there
    f = d.fields
    full = d.rt.download(F)
    assert np.all(full[:f["off"]] == 0.0) and np.all(full[f["off"] + f["n_own"]:] == 0.0)


def test_residual_is_zero_at_solution(g):
    pb = g.BratuPdeProblem(200, 5, 10)
    y = pb.pde_operator(pb.u_true)
    r = pb.make_res(y)(pb.u_true)
    assert np.max(np.abs(r)) == 0.0  # the same kernel evaluated both sides


# ------------------------------------------------------------------------------------------------
# basis kernels
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,k", [(1, 1), (2, 1), (7, 3), (1000, 7), (1001, 16), (4097, 17), (100003, 33), (65536, 103), (20000, 255)])
def test_basis_kernels(g, n, k):
    _lib, device = _lib_mods()
    from gauss_newton_via_generalized_krylov_subspaces_b200.partition import flat_layout_fields
    rt = g.get_runtime()
    lib = rt.lib
    f = flat_layout_fields(n)
    lay = device.make_layout(f)
    ld = f["ld"]
    rs = np.random.RandomState(n + k)
    Vh = rs.normal(size=(k, n))
    c, dd, w = rs.normal(size=k), rs.normal(size=k), rs.normal(size=n)
    V = rt.zeros(k * ld)
    for j in range(k):
        rt.upload(Vh[j], V[j * ld:j * ld + n])
    dc, ddv, dw, x = rt.zeros(256), rt.zeros(256), rt.zeros(ld), rt.zeros(ld)
    rt.upload(c, dc[:k])
    rt.upload(dd, ddv[:k])
    rt.upload(w, dw[:n])
    # combine
    _lib.check(lib.gnk_combine(rt.ctx, C.byref(lay), device.ptr(V), k, device.ptr(dc), device.ptr(ddv), 0.25,
                               device.ptr(x), rt.stream))
    ref = (c + 0.25 * dd) @ Vh
    assert rel(rt.download(x[:n]), ref) < 1e-13
    _lib.check(lib.gnk_combine(rt.ctx, C.byref(lay), device.ptr(V), k, device.ptr(dc), None, 0.0, device.ptr(x),
                               rt.stream))
    assert rel(rt.download(x[:n]), c @ Vh) < 1e-13
    # stats + normalize
    st = rt.zeros(2)
    flag = rt.zeros(1, dtype=rt.torch.int32)
    _lib.check(lib.gnk_norm_stats(rt.ctx, C.byref(lay), device.ptr(dw), device.ptr(st), rt.stream))
    s = rt.read(st, 2)
    assert abs(s[0] - np.sum(w * w)) <= 1e-13 * np.sum(w * w) and s[1] == np.max(np.abs(w))
    out = rt.zeros(ld)
    _lib.check(lib.gnk_normalize(rt.ctx, C.byref(lay), device.ptr(dw), device.ptr(st), 1e-8, device.ptr(out),
                                 device.ptr(flag), rt.stream))
    assert rt.read_i32(flag)[0] == 0
    assert rel(rt.download(out[:n]), w / np.linalg.norm(w)) < 1e-15
    tiny = rt.zeros(ld)
    rt.upload(1e-9 * w / np.max(np.abs(w)), tiny[:n])
    _lib.check(lib.gnk_norm_stats(rt.ctx, C.byref(lay), device.ptr(tiny), device.ptr(st), rt.stream))
    _lib.check(lib.gnk_normalize(rt.ctx, C.byref(lay), device.ptr(tiny), device.ptr(st), 1e-8, device.ptr(out),
                                 device.ptr(flag), rt.stream))
    assert rt.read_i32(flag)[0] == 1  # breakdown: max|w| <= 1e-8 (krylow.py:66)
    # Gram-Schmidt halves
    h = rt.zeros(256)
    _lib.check(lib.gnk_cgs_dots(rt.ctx, C.byref(lay), device.ptr(V), k, device.ptr(dw), device.ptr(h), rt.stream))
    href = Vh @ w
    assert rel(rt.read(h, k), href) < 1e-12
    rt.upload(href, h[:k])
    _lib.check(lib.gnk_cgs_update(rt.ctx, C.byref(lay), device.ptr(V), k, device.ptr(h), device.ptr(dw),
                                  device.ptr(st), rt.stream))
    wref = w - href @ Vh
    assert rel(rt.download(dw[:n]), wref) < 1e-12
    s = rt.read(st, 2)
    assert abs(s[0] - np.sum(wref ** 2)) <= 1e-12 * np.sum(wref ** 2)
    assert abs(s[1] - np.max(np.abs(wref))) <= 1e-12 * np.max(np.abs(wref))
    # axpby / dot
    _lib.check(lib.gnk_axpby(rt.ctx, n, 2.0, device.ptr(dw), -0.5, device.ptr(x), device.ptr(out), rt.stream))
    assert rel(rt.download(out[:n]), 2.0 * wref - 0.5 * (c @ Vh)) < 1e-13
    _lib.check(lib.gnk_dot(rt.ctx, n, device.ptr(dw), device.ptr(x), device.ptr(st), rt.stream))
    assert abs(rt.read(st, 1)[0] - np.dot(wref, c @ Vh)) <= 1e-11 * np.linalg.norm(wref) * np.linalg.norm(c @ Vh)


def test_reductions_are_deterministic(g):
    _lib, device = _lib_mods()
    from gauss_newton_via_generalized_krylov_subspaces_b200.partition import flat_layout_fields
    rt = g.get_runtime()
    n, k = 1 << 20, 9
    f = flat_layout_fields(n)
    lay = device.make_layout(f)
    V = rt.torch.randn(k * f["ld"], dtype=rt.torch.float64, device=rt.device)
    w = rt.torch.randn(f["ld"], dtype=rt.torch.float64, device=rt.device)
    h = rt.zeros(128)
    outs = []
    for _ in range(3):
        _lib.check(rt.lib.gnk_cgs_dots(rt.ctx, C.byref(lay), device.ptr(V), k, device.ptr(w), device.ptr(h), rt.stream))
        outs.append(rt.read(h, k))
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])


# ------------------------------------------------------------------------------------------------
# Householder TSQR least squares
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,k", [(2, 1), (3, 2), (5, 4), (40, 7), (300, 8), (1000, 15), (5000, 16), (100000, 30),
                                 (20000, 31), (7777, 40), (30000, 63), (9000, 64), (4000, 100), (250, 103)])
def test_tsqr_least_squares(g, n, k):
    rs = np.random.RandomState(n * 7 + k)
    A = rs.normal(size=(n, k)) @ (np.eye(k) + 0.3 * rs.normal(size=(k, k)))
    y = rs.normal(size=n)
    x = g.linear_least_squares(A, y)
    xr = np.linalg.lstsq(A, y, rcond=None)[0]
    cond = np.linalg.cond(A)
    assert rel(x, xr) < 1e-13 * max(cond, 10.0)


@pytest.mark.parametrize("n,k", [(576, 104), (576, 177), (1000, 255), (300, 200), (16000, 130)])
def test_wide_panel_least_squares(g, n, k):
    """more than 103 columns: the single-CTA Householder QR (csrc/tsqr.cu: dense_qr_ls_kernel)"""
    rs = np.random.RandomState(n + k)
    A = rs.normal(size=(n, k)) @ (np.eye(k) + 0.1 * rs.normal(size=(k, k)))
    y = rs.normal(size=n)
    x = g.linear_least_squares(A, y)
    xr = np.linalg.lstsq(A, y, rcond=None)[0]
    assert rel(x, xr) < 1e-13 * max(np.linalg.cond(A), 10.0)
    v = _raw_ls(g, A, y, -1.0)
    Ad = -A @ v[:k]
    assert abs(v[k] - Ad @ Ad) <= 1e-10 * (Ad @ Ad) and abs(v[k + 1] - np.sum((y - Ad) ** 2)) <= 1e-10 * (y @ y)
    assert v[k + 2] == 0 and abs(v[k + 3] - v[:k] @ v[:k]) <= 1e-12 * (v[:k] @ v[:k])
    assert np.allclose(np.abs(v[k + 4:2 * k + 4]), np.abs(np.diag(np.linalg.qr(A, mode="r"))), rtol=1e-10)


@pytest.mark.parametrize("n,k", [(16384, 1), (16385, 2), (20001, 7), (33333, 8), (50001, 12), (65537, 15), (70001, 16),
                                 (40003, 23), (100001, 24), (262147, 31)])
def test_tsqr_warp_autonomous_leaf(g, n, k):
    """large panels (>= 16384 rows, <= 32 columns) take the warp-autonomous leaf: odd row counts (8-byte tails of the
    16-byte copies), every column-slot count, partial last tiles, warps without tiles."""
    rs = np.random.RandomState(n + k)
    A = rs.normal(size=(n, k)) @ (np.eye(k) + 0.3 * rs.normal(size=(k, k)))
    y = rs.normal(size=n)
    x = g.linear_least_squares(A, y)
    xr = np.linalg.lstsq(A, y, rcond=None)[0]
    assert rel(x, xr) < 1e-13 * max(np.linalg.cond(A), 10.0)
    # scaling by a power of two changes no rounding: identical bits, and graded columns stay accurate
    x2 = g.linear_least_squares(A * 2.0 ** 40, y * 2.0 ** 40)
    assert np.array_equal(x, x2)


# ------------------------------------------------------------------------------------------------
# CholeskyQR2 on the FP64 tensor pipe (csrc/cholqr.cu): same entry point, large even-row panels of 9..32 columns
# ------------------------------------------------------------------------------------------------
def _raw_ls(g, A, y, sign, householder=False, method=None):
    """gnk_tsqr_ls on a host panel; returns the 2k+4 result doubles.  method: 0 automatic (refinement form for
    well-conditioned panels), 1 Householder, 2 CholeskyQR2 (gnk_tsqr_ls_method)."""
    _lib, device = _lib_mods()
    from gauss_newton_via_generalized_krylov_subspaces_b200.gauss_newton_krylow import tsqr_solve
    rt = g.get_runtime()
    n, k = A.shape
    lda = (n + 15) // 16 * 16
    dA = rt.zeros(lda * k)
    for j in range(k):
        rt.upload(np.ascontiguousarray(A[:, j]), dA[j * lda:j * lda + n])
    dy = rt.zeros(lda)
    rt.upload(y, dy[:n])
    out = rt.zeros(2 * 256 + 8)
    tsqr_solve(rt, dA, lda, n, k, dy, sign, out, householder=householder, method=method)
    return rt.read(out, 2 * k + 4).copy()


@pytest.mark.parametrize("n,k", [(16384, 2), (40000, 3), (65536, 5), (30002, 7), (16384, 8), (16386, 9), (20000, 15), (50002, 16), (65536, 17), (30000, 23),
                                 (100000, 24), (262144, 31), (1 << 20, 30), (16384 + 190, 12)])
def test_cholqr2_matches_householder_and_lstsq(g, n, k):
    """every block count (1 .. 4 blocks of 8 columns), partial last steps, CTAs without rows; tolerance: both
    factorisations are backward stable, so d agrees to eps * cond, the scalars to 1e-11 relative."""
    rs = np.random.RandomState(n + 31 * k)
    A = rs.normal(size=(n, k)) @ (np.eye(k) + 0.3 * rs.normal(size=(k, k)))
    A *= np.exp(rs.uniform(-4, 4, size=k))[None, :]          # graded columns: CholeskyQR is scaling invariant
    y = rs.normal(size=n)
    cond = np.linalg.cond(A / np.linalg.norm(A, axis=0))
    for sign in (1.0, -1.0):
        vh = _raw_ls(g, A, y, sign, method=1)
        xr = np.linalg.lstsq(sign * A, y, rcond=None)[0]
        for method in (0, 2):                                  # refinement form / second CholeskyQR2 pass
            vc = _raw_ls(g, A, y, sign, method=method)
            assert vc[k + 2] == 0 and vh[k + 2] == 0
            assert rel(vc[:k], xr) < 1e-13 * max(cond, 10.0), (method, rel(vc[:k], xr), cond)
            assert rel(vc[:k], vh[:k]) < 1e-13 * max(cond, 10.0)
            # ||R d||^2, LS residual^2, ||d||^2: the refinement form reports ||A d0||^2 of the normal-equation
            # solution (include/gnk_b200.h), which agrees to cond^2 eps
            for i, tol in ((k, 1e-11 if method == 2 else 1e-9), (k + 1, 1e-11), (k + 3, 1e-11)):
                assert abs(vc[i] - vh[i]) <= tol * abs(vh[i]), (method, i, vc[i], vh[i])
            assert np.allclose(vc[k + 4:], np.abs(vh[k + 4:]), rtol=1e-11 if method == 2 else 1e-9)
    # run-to-run deterministic (fixed reduction order) and exact under power-of-two scaling
    assert np.array_equal(_raw_ls(g, A, y, 1.0, False), _raw_ls(g, A, y, 1.0, False))
    assert np.array_equal(_raw_ls(g, A * 2.0 ** 30, y * 2.0 ** 30, 1.0, False)[:k], _raw_ls(g, A, y, 1.0, False)[:k])


def test_cholqr2_moderately_ill_conditioned(g):
    """cond(A) = 1e5: a single Cholesky pass would lose cond^2 eps = 1e-6; the second pass restores eps * cond."""
    rs = np.random.RandomState(11)
    n, k = 40000, 20
    U, _ = np.linalg.qr(rs.normal(size=(n, k)))
    W, _ = np.linalg.qr(rs.normal(size=(k, k)))
    A = (U * np.logspace(0, -5, k)) @ W.T
    y = rs.normal(size=n)
    vh = _raw_ls(g, A, y, 1.0, method=1)
    xr = np.linalg.lstsq(A, y, rcond=None)[0]
    assert rel(vh[:k], xr) < 1e-16 * 1e5 * 200
    for method in (0, 2):
        vc = _raw_ls(g, A, y, 1.0, method=method)
        if method == 0 and vc[k + 2] == -1:
            # pivot ratios ~1e-10 sit on the refinement form's floor: the default chain (Gram pass + refinement pass, no
            # second CholeskyQR2 pass) may decline, and the caller falls back to the Householder path
            continue
        assert vc[k + 2] == 0
        assert rel(vc[:k], xr) < 1e-16 * 1e5 * 200, (method, rel(vc[:k], xr))
        assert abs(vc[k] - vh[k]) <= 1e-5 * vh[k] and abs(vc[k + 1] - vh[k + 1]) <= 1e-10 * vh[k + 1]


def test_cholqr2_refuses_and_falls_back(g, capsys):
    """consistent system (y in range(A)), nearly dependent columns, exactly repeated column: the Gram matrix is
    numerically singular -> sentinel (d = 0, out[k+2] = -1); the public entry points fall back to Householder."""
    rs = np.random.RandomState(12)
    n, k = 20000, 10
    A = rs.normal(size=(n, k))
    x0 = rs.normal(size=k)
    v = _raw_ls(g, A, A @ x0, 1.0, householder=False)
    assert v[k + 2] == -1 and np.all(v[:k] == 0) and v[k] == 0 and v[k + 3] == 0
    assert rel(g.linear_least_squares(A, A @ x0), x0) < 1e-12
    A2 = A.copy()
    A2[:, 7] = A2[:, 2] + 1e-9 * rs.normal(size=n)
    y = rs.normal(size=n)
    assert _raw_ls(g, A2, y, 1.0, householder=False)[k + 2] == -1
    xh = _raw_ls(g, A2, y, 1.0, householder=True)
    assert xh[k + 2] == 0 and rel(g.linear_least_squares(A2, y), xh[:k]) == 0.0
    A3 = A.copy()
    A3[:, 4] = A3[:, 1]
    capsys.readouterr()
    g.linear_least_squares(A3, y)
    assert capsys.readouterr().out.count("A is rank deficient") == 1
    # an all-zero panel must refuse as well (no positive pivot), not produce NaNs silently
    assert _raw_ls(g, np.zeros((16384, 9)), np.zeros(16384), 1.0, householder=False)[9 + 2] == -1


def test_cholqr2_norm_preservation_at_benchmark_size(g):
    """size-independent properties at 4096^2 rows, k = 30: R^T R = M^T M (diag R against the Cholesky factor of the
    torch-computed Gram matrix), ||Q^T y||^2 + resid^2 = ||y||^2, agreement with the Householder path."""
    from gauss_newton_via_generalized_krylov_subspaces_b200.gauss_newton_krylow import tsqr_solve
    rt = g.get_runtime()
    n, k = 4096 * 4096, 30
    A = rt.torch.randn(k * n, dtype=rt.torch.float64, device=rt.device)
    y = rt.torch.randn(n, dtype=rt.torch.float64, device=rt.device)
    out = rt.zeros(256)
    tsqr_solve(rt, A, n, n, k, y, 1.0, out, method=2)
    v = rt.read(out, 2 * k + 4).copy()
    tsqr_solve(rt, A, n, n, k, y, 1.0, out)                 # automatic: the refinement form on this panel
    v0 = rt.read(out, 2 * k + 4).copy()
    tsqr_solve(rt, A, n, n, k, y, 1.0, out, householder=True)
    vh = rt.read(out, 2 * k + 4).copy()
    assert v0[k + 2] == 0 and rel(v0[:k], vh[:k]) < 1e-12
    assert abs(v0[k] + v0[k + 1] - vh[k] - vh[k + 1]) < 1e-12 * (vh[k] + vh[k + 1])
    M = rt.torch.cat([A.view(k, n), y.view(1, n)], 0)
    Gm = (M @ M.T).cpu().numpy()          # torch here is the checker, not the product
    assert v[k + 2] == 0
    assert rel(v[:k], vh[:k]) < 1e-12
    assert abs(v[k] + v[k + 1] - Gm[k, k]) < 1e-12 * Gm[k, k]
    assert np.allclose(v[k + 4:] ** 2, np.diag(np.linalg.cholesky(Gm[:k, :k])) ** 2, rtol=1e-10)
    assert np.allclose(v[k + 4:], np.abs(vh[k + 4:]), rtol=1e-12)


@pytest.mark.parametrize("n,k", [(16384, 32), (20000, 33), (40002, 39), (65536, 40), (30000, 41), (50000, 47),
                                 (16386, 48), (100000, 50), (262144, 55)])
def test_wide_gram_qr_matches_householder_and_lstsq(g, n, k):
    """33..56 panel columns (krylow_restart up to 55) on the tensor-pipe path: wide Gram kernel + dense Cholesky +
    refinement pass (csrc/gram_cgls.cu: gnk_cholqr_wide_try), every block count (5..7 blocks of 8 columns), both
    register-tile widths of the refinement kernel, partial last steps; same bars as the narrow path."""
    rs = np.random.RandomState(n + 31 * k)
    A = rs.normal(size=(n, k)) @ (np.eye(k) + 0.3 * rs.normal(size=(k, k)))
    A *= np.exp(rs.uniform(-4, 4, size=k))[None, :]          # graded columns: CholeskyQR is scaling invariant
    y = rs.normal(size=n)
    cond = np.linalg.cond(A / np.linalg.norm(A, axis=0))
    for sign in (1.0, -1.0):
        vh = _raw_ls(g, A, y, sign, method=1)
        xr = np.linalg.lstsq(sign * A, y, rcond=None)[0]
        vc = _raw_ls(g, A, y, sign, method=0)
        assert vc[k + 2] == 0 and vh[k + 2] == 0
        assert rel(vc[:k], xr) < 1e-13 * max(cond, 10.0), (rel(vc[:k], xr), cond)
        assert rel(vc[:k], vh[:k]) < 1e-13 * max(cond, 10.0)
        for i, tol in ((k, 1e-9), (k + 1, 1e-11), (k + 3, 1e-11)):
            assert abs(vc[i] - vh[i]) <= tol * abs(vh[i]), (i, vc[i], vh[i])
        assert np.allclose(vc[k + 4:], np.abs(vh[k + 4:]), rtol=1e-9)
        assert not np.array_equal(vc[:k], vh[:k]) or k == 0   # the two calls really took different paths
    assert np.array_equal(_raw_ls(g, A, y, 1.0), _raw_ls(g, A, y, 1.0))
    assert np.array_equal(_raw_ls(g, A * 2.0 ** 30, y * 2.0 ** 30, 1.0)[:k], _raw_ls(g, A, y, 1.0)[:k])


def test_wide_gram_qr_refuses_and_falls_back(g):
    """consistent system / nearly dependent columns / all-zero panel at 40 columns: sentinel, then Householder"""
    rs = np.random.RandomState(13)
    n, k = 20000, 40
    A = rs.normal(size=(n, k))
    x0 = rs.normal(size=k)
    v = _raw_ls(g, A, A @ x0, 1.0)
    assert v[k + 2] == -1 and np.all(v[:k] == 0) and v[k] == 0 and v[k + 3] == 0
    assert rel(g.linear_least_squares(A, A @ x0), x0) < 1e-12
    A2 = A.copy()
    A2[:, 37] = A2[:, 2] + 1e-9 * rs.normal(size=n)
    y = rs.normal(size=n)
    assert _raw_ls(g, A2, y, 1.0)[k + 2] == -1
    xh = _raw_ls(g, A2, y, 1.0, householder=True)
    assert xh[k + 2] == 0 and rel(g.linear_least_squares(A2, y), xh[:k]) == 0.0
    assert _raw_ls(g, np.zeros((16384, 35)), np.zeros(16384), 1.0)[35 + 2] == -1


def test_tsqr_scalar_block_and_rank_deficiency(g, capsys):
    _lib, device = _lib_mods()
    from gauss_newton_via_generalized_krylov_subspaces_b200.gauss_newton_krylow import tsqr_solve
    rt = g.get_runtime()
    n, k = 5000, 6
    rs = np.random.RandomState(5)
    A = rs.normal(size=(n, k))
    y = rs.normal(size=n)
    lda = 5008
    dA = rt.zeros(lda * k)
    for j in range(k):
        rt.upload(A[:, j].copy(), dA[j * lda:j * lda + n])
    dy = rt.zeros(lda)
    rt.upload(y, dy[:n])
    out = rt.zeros(256)
    tsqr_solve(rt, dA, lda, n, k, dy, -1.0, out)  # min || -A d - y ||
    v = rt.read(out, 2 * k + 4)
    d = np.linalg.lstsq(-A, y, rcond=None)[0]
    assert rel(v[:k], d) < 1e-12
    assert abs(v[k] - np.sum((A @ d) ** 2)) < 1e-11 * np.sum((A @ d) ** 2)       # ||R d||^2 = ||A d||^2
    assert abs(v[k + 1] - np.sum((-A @ d - y) ** 2)) < 1e-11 * np.sum(y ** 2)     # LS residual
    assert v[k + 2] == 0 and abs(v[k + 3] - np.sum(d * d)) < 1e-12 * np.sum(d * d)
    assert np.allclose(np.abs(v[k + 4:2 * k + 4]), np.abs(np.diag(np.linalg.qr(A, mode="r"))), rtol=1e-11)
    # a repeated column -> exactly one |r_kk| <= 1e-8, one print (gauss_newton_krylow.py:32-34)
    A2 = A.copy()
    A2[:, 4] = A2[:, 1]
    g.linear_least_squares(A2, y)
    assert capsys.readouterr().out.count("A is rank deficient") == 1


def test_tsqr_norm_preservation_large(g):
    """size-independent property at the benchmark's row count: ||R||_F = ||[A|y]||_F and R^T R = M^T M."""
    _lib, device = _lib_mods()
    from gauss_newton_via_generalized_krylov_subspaces_b200.gauss_newton_krylow import tsqr_solve
    rt = g.get_runtime()
    n, k = 4096 * 4096, 12
    A = rt.torch.randn(k * n, dtype=rt.torch.float64, device=rt.device)
    y = rt.torch.randn(n, dtype=rt.torch.float64, device=rt.device)
    out = rt.zeros(256)
    tsqr_solve(rt, A, n, n, k, y, 1.0, out)
    v = rt.read(out, 2 * k + 4)
    M = rt.torch.cat([A.view(k, n), y.view(1, n)], 0)
    Gm = (M @ M.T).cpu().numpy()          # torch here is the checker, not the product
    d = np.linalg.solve(Gm[:k, :k], Gm[:k, k])
    assert rel(v[:k], d) < 1e-9
    assert abs(v[k] + v[k + 1] - Gm[k, k]) < 1e-12 * Gm[k, k]   # ||Q^T y||^2 + resid^2 = ||y||^2
    assert np.allclose(v[k + 4:2 * k + 4] ** 2, np.diag(np.linalg.cholesky(Gm[:k, :k])) ** 2, rtol=1e-10)


# ------------------------------------------------------------------------------------------------
# CSR kernels and CGLS
# ------------------------------------------------------------------------------------------------
def test_csr_kernels_rosenbrock(g):
    _lib, device = _lib_mods()
    import scipy.sparse as sp
    z = Golden("kernels")
    x = z["rosen/x"]
    from gauss_newton_via_generalized_krylov_subspaces_b200 import rosenbrock_problem as rp
    assert rel(rp.res(x), z["rosen/res"]) == 0.0
    J = sp.csr_array(rp.jac(x))
    Jg = sp.csr_array((z["rosen/data"], z["rosen/indices"], z["rosen/indptr"]), shape=J.shape)
    assert abs(J - Jg).max() == 0.0
    rt = g.get_runtime()
    op = device.CsrJacobian(rt, rp.jac(x), True)
    rs = np.random.RandomState(0)
    V = rs.normal(size=(3, 1000))
    dV = rt.zeros(3 * 1008)
    for j in range(3):
        rt.upload(V[j], dV[j * 1008:j * 1008 + 1000])
    JV = rt.zeros(3 * 2000)
    op.matmat(dV, 1008, 3, JV, 2000)
    got = rt.download(JV).reshape(3, 2000)[:, :1998]
    assert rel(got, (J @ V.T).T) < 1e-14
    r = rs.normal(size=1998)
    dr, w = rt.zeros(2000), rt.zeros(1008)
    rt.upload(r, dr[:1998])
    op.neg_rmatvec(dr, w)
    assert rel(rt.download(w[:1000]), -(J.T @ r)) < 1e-14
    ss = rt.zeros(1008)
    _lib.check(rt.lib.gnk_csr_row_sumsq(rt.ctx, 1000, device.ptr(op.rowptr_t), device.ptr(op.val_t), device.ptr(ss),
                                        rt.stream))
    assert rel(rt.download(ss[:1000]), (J.T @ J).diagonal()) < 1e-14


@pytest.mark.parametrize("precond", [True, False])
def test_cgls_stencil_and_csr(g, precond):
    o = orc.BratuOracle(41, 5, 10)
    pb = g.BratuPdeProblem(41, 5, 10)
    rs = np.random.RandomState(1)
    u = 0.2 * rs.normal(size=o.n)
    y = rs.normal(size=o.n)
    xr, itr = orc.cgls(-1 * o.make_jac()(u), y, preconditioner=precond)
    xg, itg = g.cg_least_squares(-1 * pb.make_jac()(u), y, preconditioner=precond)
    assert abs(itg - itr) <= max(2, itr // 50)
    assert rel(xg, xr) < 1e-3  # both stop at rtol=1e-4 on the normal-equation residual
    from gauss_newton_via_generalized_krylov_subspaces_b200 import rosenbrock_problem as rp
    x = 1 + 0.1 * rs.normal(size=1000)
    r = rp.res(x)
    xr, itr = orc.cgls(-1 * orc.rosenbrock_jac(x), r, preconditioner=precond)
    xg, itg = g.cg_least_squares(-1 * rp.jac(x), r, preconditioner=precond)
    assert abs(itg - itr) <= 2
    assert rel(xg, xr) < 1e-3


@pytest.mark.parametrize("precond", [True, False])
def test_cgls_with_initial_guess_matches_reference_golden(g, precond):
    """cg_least_squares(A, y, x0=...): the reference forwards x0 to scipy's cg (gauss_newton.py:14,46,56).  Golden from
    the unmodified reference (oracle/gen_golden.py cgx0): iteration counts (rounding-sensitive: +-2) and the solution."""
    import scipy.sparse as sp
    from gauss_newton_via_generalized_krylov_subspaces_b200 import rosenbrock_problem as rp
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cg_x0.npz"))
    pb = g.BratuPdeProblem(34, 5, 10)
    x, it = g.cg_least_squares(-1 * pb.make_jac()(z["bratu/u"]), z["bratu/y"], x0=z["bratu/x0"], preconditioner=precond)
    assert abs(it - int(z[f"bratu/it_pre{int(precond)}"])) <= 2, (it, int(z[f"bratu/it_pre{int(precond)}"]))
    assert rel(x, z[f"bratu/x_pre{int(precond)}"]) < 1e-6
    x, it = g.cg_least_squares(sp.csr_array(-1 * rp.jac(z["rosen/x"])), rp.res(z["rosen/x"]), x0=z["rosen/x0"],
                               preconditioner=precond)
    assert abs(it - int(z[f"rosen/it_pre{int(precond)}"])) <= 1
    assert rel(x, z[f"rosen/x_pre{int(precond)}"]) < 1e-6
    # x0 = exact solution of a consistent system: zero iterations of the preconditioned run
    xs = np.random.RandomState(5).normal(size=1000)
    A = sp.csr_array(rp.jac(z["rosen/x"]))
    x, it = g.cg_least_squares(A, A @ xs, x0=xs, preconditioner=True)
    assert it == 0 and np.array_equal(x, xs)


@pytest.mark.parametrize("n,k", [(4096, 3), (20000, 15), (30002, 31), (16384, 32), (40000, 40), (65536, 50), (20480, 55)])
def test_gram_cgls_matches_the_reference_cg_least_squares(g, n, k):
    """gnk_gram_cgls (Gram matrix on the tensor pipe + the CG recurrence in one kernel on the k x k system) against the
    oracle's restatement of cg_least_squares(A, y, cg_rtol, preconditioner=True) (gauss_newton.py:11-60 / scipy cg)
    applied to the dense panel: same iteration count (+-1: the stopping test is evaluated on differently rounded
    residuals), same solution to the accuracy of the stopping rule, ||A d||^2 and ||y - A d||^2 consistent with d."""
    _lib, device = _lib_mods()
    rt = g.get_runtime()
    rs = np.random.RandomState(n + k)
    A = rs.normal(size=(n, k)) @ (np.eye(k) + 0.2 * rs.normal(size=(k, k)))
    A *= np.exp(rs.uniform(-2, 2, size=k))[None, :]
    y = rs.normal(size=n)
    lda = (n + 15) // 16 * 16
    dA = rt.zeros(lda * k)
    for j in range(k):
        rt.upload(np.ascontiguousarray(A[:, j]), dA[j * lda:j * lda + n])
    dy = rt.zeros(lda)
    rt.upload(y, dy[:n])
    for sign in (1.0, -1.0):
        for rtol in (1e-4, 1e-10):
            out = rt.zeros(256)
            _lib.check(rt.lib.gnk_gram_cgls(rt.ctx, device.ptr(dA), lda, n, k, device.ptr(dy), sign, rtol, device.ptr(out),
                                            rt.stream), "gnk_gram_cgls")
            v = rt.read(out, 2 * k + 5).copy()
            xr, itr = orc.cgls(sign * A, y, rtol=rtol, preconditioner=True)
            d = v[:k]
            assert abs(int(v[2 * k + 4]) - itr) <= 1, (int(v[2 * k + 4]), itr)
            assert rel(d, xr) < max(50 * rtol, 1e-9), (rel(d, xr), rtol)
            Ad = sign * A @ d
            assert abs(v[k] - Ad @ Ad) <= 1e-10 * (Ad @ Ad) and abs(v[k + 3] - d @ d) <= 1e-12 * (d @ d)
            assert abs(v[k + 1] - np.sum((y - Ad) ** 2)) <= 1e-9 * (y @ y)
            assert np.allclose(v[k + 4:2 * k + 4], np.linalg.norm(A, axis=0), rtol=1e-12)


# ------------------------------------------------------------------------------------------------
# public building blocks with the reference's signatures
# ------------------------------------------------------------------------------------------------
def test_generalized_krylow_subspace_public_api(g):
    rs = np.random.RandomState(3)
    x0 = rs.normal(size=500)
    ks = g.GeneralizedKrylowSubspace()
    c = ks.start(x0)
    assert c.shape == (1,) and abs(c[0] - np.linalg.norm(x0)) < 1e-13 * c[0]
    assert rel(ks.x(c), x0) < 1e-15
    V = [x0 / np.linalg.norm(x0)]
    for it in range(4):
        J = rs.normal(size=(700, 500))
        r = rs.normal(size=700)
        ks.update(J, r)
        w = -(J.T @ r)
        Vm = np.array(V).T
        w = w - Vm @ (Vm.T @ w)
        V.append(w / np.linalg.norm(w))
    B = ks.basis
    assert B.shape == (500, 5) and rel(B, np.array(V).T) < 1e-12
    assert ks.evaluate(lambda x, a: a * np.sum(x), np.ones(5), 2.0) == pytest.approx(2.0 * np.sum(B @ np.ones(5)))
    with pytest.raises(ValueError):
        g.GeneralizedKrylowSubspace().start(np.zeros(10))
    with pytest.raises(g.GeneralizedKrylowSubspaceBreakdown):
        ks.update(np.zeros((700, 500)), r)
    small = g.GeneralizedKrylowSubspace()
    small.start(np.array([1.0]))
    with pytest.raises(g.GeneralizedKrylowSubspaceSpansEntireSpace):
        small.update(np.array([[1.0], [2.0]]), np.array([1.0, 1.0]))


def test_armijo_goldstein_public_api(g):
    def res(x, t):
        return np.array([x[0] + 1, t * x[0] ** 2 + x[0] - 1])

    x = np.array([1.0])
    J = np.array([[1.0], [2 * 5 * x[0] + 1]])
    r = res(x, 5)
    d = np.linalg.lstsq(-J, r, rcond=None)[0]
    s, rn, it = g.armijo_goldstein(res, x, r, J, (5,), d)
    so, rno, ito = orc.armijo(res, x, r, np.sum((J @ d) ** 2), (5,), d)
    assert (s, it) == (so, ito) and np.array_equal(rn, rno)
    with pytest.raises(g.StepLengthConvergenceError):
        g.armijo_goldstein(res, x, r, J, (5,), -d)


@pytest.mark.parametrize("G,k,lam", [(129, 7, 10.0), (137, 8, 10.0), (145, 11, 10.0), (201, 15, 10.0), (257, 16, 3.0),
                                      (265, 23, 10.0), (513, 24, 10.0), (521, 31, 10.0), (1025, 30, 10.0),
                                      (193, 12, 0.0)])
def test_stencil_gram_ls_matches_apply_plus_tsqr(g, G, k, lam):
    """gnk_stencil_gram_ls (TMA-staged kernel: stencil + J V store + Gram matrix in one sweep, then the refinement
    pass) against gnk_stencil_apply + gnk_tsqr_ls: J V bit-identical, the result block equal to rounding (the two Gram
    matrices are summed in different orders), d against numpy's lstsq.  Covers 1..4 column blocks, rows that are not a
    multiple of the strip height, row lengths that are multiples of 8 but not of the 64-point tile (partial tiles,
    idle consumer warps), lam = 0 (no e^u plane)."""
    _lib, device = _lib_mods()
    from gauss_newton_via_generalized_krylov_subspaces_b200.gauss_newton_krylow import tsqr_solve
    pb = g.BratuPdeProblem(G, 5, lam)
    d = pb.dev
    rt, lib = d.rt, d.rt.lib
    n, ld, off = d.fields["n_own"], d.ld, d.fields["off"]
    assert pb.m % 8 == 0 and n >= 16384
    rs = np.random.RandomState(G * 100 + k)
    cap = k + 3
    V = rt.zeros(cap * ld)
    for j in range(cap):  # columns beyond k hold data too: they must not leak into the panel
        d.upload_x(rs.normal(size=pb.n), V[j * ld:(j + 1) * ld])
    u, r = d.new_col(), d.new_col()
    d.upload_x(0.3 * rs.normal(size=pb.n), u)
    d.upload_x(rs.normal(size=pb.n), r)
    E = None
    if lam != 0:
        E = d.new_col()
        d.residual_into(u, d.zero_col(), d.new_col(), E, d.scal_tmp, depth=0)
    ldjv = (n + 15) // 16 * 16
    JVa, JVb = rt.zeros(k * ldjv), rt.zeros(k * ldjv)
    d.apply(E, V, ld, k, -1.0, 0, JVa, ldjv, 0)
    a, b = rt.zeros(256), rt.zeros(256)
    tsqr_solve(rt, JVa, ldjv, n, k, r[off:], -1.0, a)
    rc = lib.gnk_stencil_gram_ls(rt.ctx, C.byref(d.lay), C.byref(d.prm), device.ptr(E), device.ptr(V), ld, cap, k,
                                 device.ptr(r), -1.0, device.ptr(JVb), ldjv, -1.0, device.ptr(b), rt.stream)
    assert rc == 0, rc
    va, vb = rt.read(a, 2 * k + 4), rt.read(b, 2 * k + 4)
    assert np.array_equal(rt.download(JVa), rt.download(JVb)), "J V differs from gnk_stencil_apply"
    JVh = rt.download(JVa).reshape(k, ldjv)[:, :n].T
    dref = np.linalg.lstsq(-JVh, rt.download(r[off:off + n]), rcond=None)[0]
    cond = np.linalg.cond(JVh)
    assert vb[k + 2] == 0 and va[k + 2] == 0
    assert rel(vb[:k], dref) < 1e-13 * max(cond, 10.0) and rel(vb[:k], va[:k]) < 1e-13 * max(cond, 10.0)
    assert np.allclose(vb[k:k + 4], va[k:k + 4], rtol=1e-9, atol=0)            # |A d0|^2, resid^2, ndef, |d|^2
    assert np.allclose(np.abs(vb[k + 4:]), np.abs(va[k + 4:]), rtol=1e-11)       # |diag R|
    # run-to-run deterministic
    b2 = rt.zeros(256)
    lib.gnk_stencil_gram_ls(rt.ctx, C.byref(d.lay), C.byref(d.prm), device.ptr(E), device.ptr(V), ld, cap, k,
                            device.ptr(r), -1.0, device.ptr(JVb), ldjv, -1.0, device.ptr(b2), rt.stream)
    assert np.array_equal(rt.read(b2, 2 * k + 4), vb)


def test_stencil_gram_ls_declines_ineligible_panels(g):
    """panels the fused tensor-pipe path does not take are answered with 1 and nothing is launched: small slabs, row
    lengths that are not a multiple of 8, more than 32 panel columns, the Householder path pinned"""
    _lib, device = _lib_mods()
    for G, k, pin in ((101, 9, 0), (134, 9, 0), (257, 40, 0), (257, 9, 1), (257, 3, 0)):
        pb = g.BratuPdeProblem(G, 5, 10)
        d = pb.dev
        rt, lib = d.rt, d.rt.lib
        n, ld = d.fields["n_own"], d.ld
        V, r, E = rt.zeros(k * ld), d.new_col(), d.new_col()
        ldjv = (n + 15) // 16 * 16
        JV, out = rt.zeros(k * ldjv), rt.zeros(256)
        prev = lib.gnk_tsqr_ls_method(rt.ctx, pin)
        l0 = rt.launches()
        rc = lib.gnk_stencil_gram_ls(rt.ctx, C.byref(d.lay), C.byref(d.prm), device.ptr(E), device.ptr(V), ld, k, k,
                                     device.ptr(r), -1.0, device.ptr(JV), ldjv, -1.0, device.ptr(out), rt.stream)
        lib.gnk_tsqr_ls_method(rt.ctx, prev)
        assert rc == 1 and rt.launches() == l0, (G, k, pin, rc)


def test_tsqr_degenerate_inputs(g, capsys):
    """zero column -> scipy's solve_triangular error; all-zero right-hand side -> d = 0; single row; k at the limit."""
    rs = np.random.RandomState(11)
    A = rs.normal(size=(3000, 5))
    y = rs.normal(size=3000)
    A0 = A.copy()
    A0[:, 2] = 0.0
    with pytest.raises(np.linalg.LinAlgError):
        g.linear_least_squares(A0, y)
    assert "A is rank deficient" in capsys.readouterr().out
    assert np.all(g.linear_least_squares(A, np.zeros(3000)) == 0.0)
    x = g.linear_least_squares(np.array([[2.0]]), np.array([3.0]))
    assert x.shape == (1,) and abs(x[0] - 1.5) < 1e-15
    B = rs.normal(size=(104, 103))                       # square-ish, k = 103 = the limit of the tiled TSQR
    xb = g.linear_least_squares(B, rs.normal(size=104))
    assert np.all(np.isfinite(xb))
    yb = rs.normal(size=200)
    B = rs.normal(size=(200, 104))                       # one more column: the single-CTA wide-panel QR
    assert rel(g.linear_least_squares(B, yb), np.linalg.lstsq(B, yb, rcond=None)[0]) < 1e-11
    from gauss_newton_via_generalized_krylov_subspaces_b200._lib import GnkError
    with pytest.raises(GnkError):                        # beyond GNK_MAX_BASIS - 1 = 255 columns
        g.linear_least_squares(rs.normal(size=(400, 256)), rs.normal(size=400))
    with pytest.raises(GnkError):                        # wide AND tall: the single-CTA path is for small panels
        g.linear_least_squares(rs.normal(size=(40000, 110)), rs.normal(size=40000))
