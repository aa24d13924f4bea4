"""The CPU oracle (oracle/gnk_oracle.py) against golden outputs of the UNMODIFIED reference (tests/golden/*.npz,
written by oracle/gen_golden.py in the build container).  This is what pins the oracle; the reference has no tests
or fixtures of its own.  Runs without a GPU.
"""
import numpy as np
import pytest

from golden_util import Golden, Recorder, check_trace, rel
from oracle import gnk_oracle as orc


@pytest.mark.parametrize("tag", ["g11", "g10", "g26lin", "g33h1", "g101"])
def test_stencil_operators_match_reference(tag):
    z = Golden("kernels")
    G, a, l, h = z[f"{tag}/params"]
    o = orc.BratuOracle(int(G), a, l, h=None if h < 0 else h)
    u, V, r = z[f"{tag}/u"], z[f"{tag}/V"], z[f"{tag}/r"]
    assert rel(o.u_true, z[f"{tag}/u_true"]) == 0.0
    assert rel(o.operator(u), z[f"{tag}/P"]) < 1e-14
    J = o.make_jac()(u)
    assert rel(J @ V, z[f"{tag}/JV"]) < 1e-14
    assert rel(J.T @ r, z[f"{tag}/JTr"]) < 1e-14
    assert rel(J.normal_diagonal(), z[f"{tag}/JTJdiag"]) < 1e-14
    assert abs(J.tocsr() @ r - J @ r).max() < 1e-9 * abs(J @ r).max()


def test_rosenbrock_matches_reference():
    z = Golden("kernels")
    x = z["rosen/x"]
    assert rel(orc.rosenbrock_res(x), z["rosen/res"]) == 0.0
    J = orc.rosenbrock_jac(x)
    assert np.array_equal(J.indptr, z["rosen/indptr"]) and np.array_equal(J.indices, z["rosen/indices"])
    assert rel(J.data, z["rosen/data"]) == 0.0


def _bratu(name, G, lam=10, h=None):
    gd = Golden(name)
    o = orc.BratuOracle(G, 5, lam, h=h)
    return gd, o, o.make_res(gd["y"]), o.make_jac(), o.make_error()


@pytest.mark.parametrize("rname,kw,tol", [
    ("gnk_res_old", dict(max_iter=45), 1e-12),
    ("gnk_res_new", dict(max_iter=45, version="res_new"), 1e-12),
    ("gnk_jac_old_res_old", dict(max_iter=25, version="jac_old_res_old"), 1e-12),
    ("gnk_jac_old_res_new", dict(max_iter=25, version="jac_old_res_new"), 1e-12),
    ("gnk_restart30", dict(max_iter=100, restart=30), 2e-9),   # restarts amplify rounding ~300x each
    ("gnk_restart7_res_new", dict(max_iter=40, restart=7, version="res_new"), 1e-9),
])
def test_gnk_bratu_g101(rname, kw, tol):
    gd, o, res, jac, err = _bratu("bratu_g101", 101)
    assert rel(o.operator(o.u_true), gd["y"]) < 1e-14 and rel(o.start_vector(), gd["u0"]) == 0.0
    gr = gd.run(rname)
    rec = Recorder(gr["sample_idx"], err)
    out = orc.gnk(res, gd["u0"], jac, callback=rec, **kw)
    n = len(rec.xnorm)
    check_trace(rec, gr, tol, upto=n)
    if n == len(gr["xnorm"]):
        assert (out["nit"], out["nfev"], out["njev"], out["success"]) == (
            int(gr["nit"]), int(gr["nfev"]), int(gr["njev"]), bool(gr["success"]))
        assert rel(out["x"], gr["x_final"]) < tol


@pytest.mark.parametrize("variant", ["cholqr2", "refine"])
def test_gram_least_squares_variants_follow_the_reference_trajectory(variant):
    """The CUDA library solves the projected least squares of the large panels with the Gram matrix + Cholesky and a
    second pass (csrc/cholqr.cu) instead of LAPACK's Householder QR.  This is the numpy emulation of both second
    passes inside the oracle's solver loop (the 1024^2 / 4096^2 versions are tests/ls_numerics_experiment.py): the
    iterates must stay on the reference's golden trajectory exactly like the LAPACK-based oracle does."""
    import scipy.linalg
    gd, o, res, jac, err = _bratu("bratu_g101", 101)
    steps = []

    def ls(A, y, log=None):
        k = A.shape[1]
        if variant == "cholqr2":
            P = np.column_stack([A, y])
            R1 = np.linalg.cholesky(P.T @ P).T
            B = P @ scipy.linalg.solve_triangular(R1, np.eye(k + 1))
            R = np.linalg.cholesky(B.T @ B).T @ R1
            return scipy.linalg.solve_triangular(R[:k, :k], R[:k, k])
        R1 = np.linalg.cholesky(A.T @ A).T
        solve = lambda v: scipy.linalg.solve_triangular(R1, scipy.linalg.solve_triangular(R1, v, trans="T"))
        d0 = solve(A.T @ y)
        delta = solve(A.T @ (y - A @ d0))
        steps.append(np.linalg.norm(delta) / np.linalg.norm(d0))
        return d0 + delta

    gr = gd.run("gnk_res_old")
    rec = Recorder(gr["sample_idx"], err)
    orc.gnk(res, gd["u0"], jac, callback=rec, max_iter=45, ls=ls)
    check_trace(rec, gr, 1e-12, upto=len(rec.xnorm))
    if variant == "refine":   # cond(JV_k) <= 100 on this grid: the refinement step is at rounding level
        assert max(steps) < 1e-10


@pytest.mark.parametrize("rname,kw", [("gnk_res_old", {}), ("gnk_res_new", dict(version="res_new"))])
def test_gnk_bratu_manufactured_solution(rname, kw):
    """bratu_pde_test.compare_manufactured_solution (:141-190): right-hand side from the continuous operator (sympy in
    oracle/gen_golden.py: manufactured; y travels in the fixture), all 99 iterations."""
    gd, o, res, jac, err = _bratu("bratu_g101_manufactured", 101)
    assert rel(o.start_vector(), gd["u0"]) == 0.0
    gr = gd.run(rname)
    rec = Recorder(gr["sample_idx"], err)
    out = orc.gnk(res, gd["u0"], jac, callback=rec, max_iter=100, **kw)
    check_trace(rec, gr, 1e-12)
    assert (out["nit"], out["nfev"], out["njev"], out["success"]) == (
        int(gr["nit"]), int(gr["nfev"]), int(gr["njev"]), bool(gr["success"])) == (99, 100, 100, False)
    # the error curve levels off at the discretisation error of the grid, far above the solver tolerance
    assert np.max(np.abs(np.array(rec.err) / gr["err"] - 1)) < 1e-10 and rec.err[-1] > 0.1


def test_gn_bratu_manufactured_solution():
    gd, o, res, jac, err = _bratu("bratu_g101_manufactured", 101)
    gr = gd.run("gn")
    rec = Recorder(gr["sample_idx"], err)
    out = orc.gn(res, gd["u0"], jac, callback=rec)
    assert (out["nit"], out["nfev"], out["njev"], out["success"]) == (4, 5, 4, True)
    assert rel(out["x"], gr["x_final"]) < 1e-10 and abs(rec.err[-1] - gr["err"][-1]) < 1e-9 * gr["err"][-1]
    assert all(abs(a - b) <= max(2, 0.02 * b) for a, b in zip(rec.cg, gr["cg_iter"]))


def test_gn_bratu_g101():
    gd, o, res, jac, err = _bratu("bratu_g101", 101)
    gr = gd.run("gn")
    rec = Recorder(gr["sample_idx"], err)
    out = orc.gn(res, gd["u0"], jac, callback=rec)
    assert (out["nit"], out["nfev"], out["njev"], out["success"]) == (4, 5, 4, True)
    assert rel(out["x"], gr["x_final"]) < 1e-10
    assert all(abs(a - b) <= max(2, 0.02 * b) for a, b in zip(rec.cg, gr["cg_iter"]))


def test_gnk_bratu_without_scaling_converges():
    gd, o, res, jac, err = _bratu("bratu_g101_h1", 101, h=1.0)
    gr = gd.run("gnk_res_new")
    rec = Recorder(gr["sample_idx"], err)
    out = orc.gnk(res, gd["u0"], jac, callback=rec, max_iter=100, version="res_new")
    check_trace(rec, gr, 1e-11)
    assert (out["nit"], out["nfev"], out["success"]) == (56, 57, True)


def test_gnk_bratu_linear_breakdown_logged():
    gd, o, res, jac, err = _bratu("bratu_g101_linear", 101, lam=0)
    u0 = -1 * jac(np.zeros(o.n)).T @ gd["y"]
    assert rel(u0, gd["u0"]) < 1e-14
    gr = gd.run("gnk_res_old")
    rec = Recorder(gr["sample_idx"], err)
    try:  # iteration 3 is decided by a 1-ulp loss difference (see tests/test_gpu_solvers.py)
        out = orc.gnk(res, gd["u0"], jac, callback=rec, max_iter=100)
        assert any("breakdown at iteration = 2, basis.shape = (10000, 2)" in m for m in out["log"])
    except orc.StepLengthFailure:
        pass
    check_trace(rec, gr, 1e-10, upto=2)


@pytest.mark.parametrize("tag,rname,kw", [("i", "gnk_res_old", {}), ("i", "gnk_res_new", dict(version="res_new")),
                                          ("ii", "gnk_res_new", dict(version="res_new")), ("iii", "gnk_res_old", {})])
def test_gnk_rosenbrock(tag, rname, kw):
    gd = Golden("rosenbrock")
    gr = gd.run(f"{tag}_{rname}")
    err = lambda x: np.linalg.norm(x - 1.0)  # noqa: E731
    rec = Recorder(gr["sample_idx"], err)
    out = orc.gnk(orc.rosenbrock_res, gd["x0_" + tag], orc.rosenbrock_jac, callback=rec, **kw)
    check_trace(rec, gr, 1e-10)
    assert (out["nit"], out["nfev"], out["njev"], out["success"]) == (
        int(gr["nit"]), int(gr["nfev"]), int(gr["njev"]), bool(gr["success"]))
    if tag == "ii":
        assert any("breakdown at iteration = 5, basis.shape = (1000, 5)" in m for m in out["log"])


def test_gn_rosenbrock_and_dense():
    gd = Golden("rosenbrock")
    gr = gd.run("i_gn")
    rec = Recorder(gr["sample_idx"], None)
    out = orc.gn(orc.rosenbrock_res, gd["x0_i"], orc.rosenbrock_jac, callback=rec)
    assert (out["nit"], out["nfev"]) == (int(gr["nit"]), int(gr["nfev"])) and list(rec.cg) == list(gr["cg_iter"])
    g3 = Golden("rosenbrock_3d").run("gn")
    xs = []
    out = orc.gn(orc.rosenbrock_res, np.array([-1.0, 1.0]), lambda x: orc.rosenbrock_jac(x, dense=True),
                 callback=lambda x, nfev, cg_iter: xs.append(x.copy()))
    assert (out["nit"], out["nfev"], out["njev"], out["success"]) == (19, 71, 19, True)
    assert np.max(np.abs(np.array(xs) - g3["xs"])) < 1e-13


def test_powell_armijo():
    def pres(x, tau):
        return np.array([x[0] + 1, tau * x[0] ** 2 + x[0] - 1])

    def pjac(x, tau):
        return np.array([[1], [2 * tau * x[0] + 1]])

    gd = Golden("powell")
    gr = gd.run("tau5_armijo_goldstein")
    xs = []
    out = orc.gn(pres, np.array([1.0]), pjac, args=(5,), max_iter=19, callback=lambda x, nfev, cg_iter: xs.append(x[0]))
    assert (out["nit"], out["nfev"], out["success"]) == (15, 16, True)
    assert np.max(np.abs(np.array(xs) - gr["xs"][:, 0])) < 1e-14
    with pytest.raises(orc.StepLengthFailure):
        orc.gn(pres, np.array([1.0]), pjac, args=(-5,), max_iter=19)
