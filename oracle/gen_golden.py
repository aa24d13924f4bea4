"""Generate tests/golden/*.npz by running the UNMODIFIED reference.

TEST INFRASTRUCTURE ONLY.  Runs in the build container (needs /root/reference,
which does not exist on the GPU box); the fixtures it writes are committed.

    python oracle/gen_golden.py small          # all n <= 1e4 configs (~1 min)
    python oracle/gen_golden.py g1025          # Bratu 1024^2 (~5 min)
    python oracle/gen_golden.py g4097          # Bratu 4096^2, 30 iterations (~10 min, 25 GB)
    python oracle/gen_golden.py ttt            # converging grid_resolution=1 runs at 256^2..1024^2 (~4 min)
    python oracle/gen_golden.py manufactured   # bratu_pde_test.compare_manufactured_solution (~1 min)

Per solver run a fixture stores: the RegressionResult fields, and per callback
|x|_2, error(x), loss 0.5|res(x)|^2, nfev, cg_iter and x sampled at fixed indices
(``sample_idx``); the final x in full when n <= 1e4.
"""
import contextlib
import io
import os
import sys
import time

import numpy as np

REF = "/root/reference"
sys.path.insert(0, REF)
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

from armijo_goldstein import StepLengthConvergenceError, armijo_goldstein  # noqa: E402
from bratu_pde_problem import BratuPdeProblem  # noqa: E402
from gauss_newton import gauss_newton  # noqa: E402
from gauss_newton_krylow import gauss_newton_krylow  # noqa: E402
import rosenbrock_problem  # noqa: E402


def sample_idx(n, count=64):
    return np.unique(np.linspace(0, n - 1, min(n, count)).astype(np.int64))


def run(method, res, x0, jac, error, args=(), loss_every=1, **kw):
    idx = sample_idx(x0.shape[0])
    rec = dict(xnorm=[], err=[], loss=[], nfev=[], cg_iter=[], xs=[])

    def cb(x, nfev, cg_iter):
        rec["xnorm"].append(np.linalg.norm(x))
        rec["err"].append(error(x) if error is not None else np.nan)
        if loss_every and (len(rec["xnorm"]) % loss_every == 0):
            rec["loss"].append(0.5 * np.sum(res(x, *args) ** 2))
        else:
            rec["loss"].append(np.nan)
        rec["nfev"].append(-1 if nfev is None else nfev)
        rec["cg_iter"].append(-1 if cg_iter is None else cg_iter)
        rec["xs"].append(np.array(x[idx], copy=True))

    buf = io.StringIO()
    t0 = time.perf_counter()
    raised = ""
    out = None
    with contextlib.redirect_stdout(buf):
        try:
            out = method(res, x0.copy(), jac, args=args, callback=cb, **kw)
        except StepLengthConvergenceError as e:
            raised = "StepLengthConvergenceError"
    wall = time.perf_counter() - t0
    d = dict(
        sample_idx=idx,
        xnorm=np.array(rec["xnorm"]), err=np.array(rec["err"]), loss=np.array(rec["loss"]),
        nfev_cb=np.array(rec["nfev"]), cg_iter=np.array(rec["cg_iter"]),
        xs=np.array(rec["xs"]).reshape(len(rec["xs"]), idx.shape[0]),
        raised=np.array(raised), stdout=np.array(buf.getvalue()), wall_s=np.array(wall),
    )
    if out is not None:
        d.update(success=np.array(out.success), nit=np.array(out.nit), nfev=np.array(out.nrev),
                 njev=np.array(out.njev), x_sample=np.array(out.x[idx]), x_norm_final=np.array(np.linalg.norm(out.x)))
        if x0.shape[0] <= 10000:
            d["x_final"] = np.array(out.x)
    return d


def save(name, runs, **extra):
    flat = dict(extra)
    for rname, d in runs.items():
        for k, v in d.items():
            flat[f"{rname}/{k}"] = v
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **flat)
    for rname, d in runs.items():
        print(f"{name}:{rname}: nit={d.get('nit')} nfev={d.get('nfev')} success={d.get('success')} "
              f"raised={d['raised']} callbacks={len(d['xnorm'])} err_last={d['err'][-1] if len(d['err']) else None} "
              f"wall={float(d['wall_s']):.1f}s", flush=True)


def bratu_setup(G, alpha, lam, h=None, linear_start=False):
    pb = BratuPdeProblem(G, alpha, lam, grid_resolution=h)
    y = pb.pde_operator(pb.u_true)
    res, jac, err = pb.make_res(y), pb.make_jac(), pb.make_error()
    if linear_start:
        u0 = -1 * jac(np.zeros((G - 1) ** 2)).T @ y
    else:
        np.random.seed(42)
        u0 = pb.u_true + 0.1 * np.random.normal(loc=0, scale=1, size=len(pb.u_true))
    return pb, y, res, jac, err, u0


def kernels_fixture():
    """operator-level known answers straight from the reference's scipy path."""
    flat = {}
    for tag, (G, a, l, h) in dict(g11=(11, 5, 10, None), g10=(10, 5, 10, None), g26lin=(26, 5, 0, None),
                                   g33h1=(33, 5, 10, 1.0), g101=(101, 5, 10, None)).items():
        pb = BratuPdeProblem(G, a, l, grid_resolution=h)
        rs = np.random.RandomState(7)
        n = (G - 1) ** 2
        u = 0.3 * rs.normal(size=n)
        V = rs.normal(size=(n, 3))
        r = rs.normal(size=n)
        J = pb.make_jac()(u)
        flat.update({f"{tag}/params": np.array([G, a, l, -1.0 if h is None else h]), f"{tag}/u": u, f"{tag}/V": V, f"{tag}/r": r,
                     f"{tag}/P": pb.pde_operator(u), f"{tag}/JV": J @ V, f"{tag}/JTr": J.T @ r,
                     f"{tag}/u_true": pb.u_true, f"{tag}/JTJdiag": (J.T @ J).diagonal()})
    x = np.random.RandomState(3).normal(size=1000)
    Jr = rosenbrock_problem.jac(x).tocsr()
    flat.update({"rosen/x": x, "rosen/res": rosenbrock_problem.res(x), "rosen/indptr": Jr.indptr,
                 "rosen/indices": Jr.indices, "rosen/data": Jr.data})
    np.savez_compressed(os.path.join(OUT, "kernels.npz"), **flat)
    print("kernels.npz written")


def small():
    kernels_fixture()
    # --- bratu_pde_test.compare (bratu_pde_test.py:22-50)
    pb, y, res, jac, err, u0 = bratu_setup(101, 5, 10)
    save("bratu_g101", dict(
        gnk_res_old=run(gauss_newton_krylow, res, u0, jac, err, max_iter=100),
        gnk_res_new=run(gauss_newton_krylow, res, u0, jac, err, max_iter=100, version="res_new"),
        gnk_jac_old_res_old=run(gauss_newton_krylow, res, u0, jac, err, max_iter=25, version="jac_old_res_old"),
        gnk_jac_old_res_new=run(gauss_newton_krylow, res, u0, jac, err, max_iter=25, version="jac_old_res_new"),
        gnk_restart30=run(gauss_newton_krylow, res, u0, jac, err, max_iter=100, krylow_restart=30),
        gnk_restart7_res_new=run(gauss_newton_krylow, res, u0, jac, err, max_iter=40, krylow_restart=7, version="res_new"),
        gn=run(gauss_newton, res, u0, jac, err),
        gn_precond=run(gauss_newton, res, u0, jac, err, cg_preconditioner=True),
    ), y=y, u0=u0)
    # --- compare_without_scaling (:76-103)
    pb, y, res, jac, err, u0 = bratu_setup(101, 5, 10, h=1)
    save("bratu_g101_h1", dict(
        gnk_res_old=run(gauss_newton_krylow, res, u0, jac, err, max_iter=100),
        gnk_res_new=run(gauss_newton_krylow, res, u0, jac, err, max_iter=100, version="res_new"),
        gn=run(gauss_newton, res, u0, jac, err),
    ), y=y, u0=u0)
    # --- compare_linear (:193-241)
    pb, y, res, jac, err, u0 = bratu_setup(101, 5, 0, linear_start=True)
    save("bratu_g101_linear", dict(
        gnk_res_old=run(gauss_newton_krylow, res, u0, jac, err, max_iter=100),
        gnk_res_new=run(gauss_newton_krylow, res, u0, jac, err, max_iter=100, version="res_new"),
        gn=run(gauss_newton, res, u0, jac, err),
    ), y=y, u0=u0)
    # --- compare_linear_small (:277-330)
    pb, y, res, jac, err, u0 = bratu_setup(25, 5, 0, linear_start=True)
    save("bratu_g25_linear", dict(
        gnk_res_old=run(gauss_newton_krylow, res, u0, jac, err, max_iter=100),
        gnk_res_new=run(gauss_newton_krylow, res, u0, jac, err, max_iter=200, version="res_new"),
        gn=run(gauss_newton, res, u0, jac, err),
    ), y=y, u0=u0)
    # --- odd m (m = 33), uneven splits
    pb, y, res, jac, err, u0 = bratu_setup(34, 5, 10)
    save("bratu_g34", dict(
        gnk_res_old=run(gauss_newton_krylow, res, u0, jac, err, max_iter=40, krylow_restart=12),
    ), y=y, u0=u0)
    # --- rosenbrock_test i/ii/iii (rosenbrock_test.py:20-31,79-88,101-111)
    rp = rosenbrock_problem
    np.random.seed(42)
    x0_i = rp.x_exact + 0.1 * np.random.normal(loc=0, scale=1, size=rp.parameter_count)
    x0_ii = 2 * rp.x_exact
    x0_iii = 2 * rp.x_exact
    x0_iii[2] = 1.99
    runs = {}
    for tag, x0 in (("i", x0_i), ("ii", x0_ii), ("iii", x0_iii)):
        runs[f"{tag}_gnk_res_old"] = run(gauss_newton_krylow, rp.res, x0, rp.jac, rp.error)
        runs[f"{tag}_gnk_res_new"] = run(gauss_newton_krylow, rp.res, x0, rp.jac, rp.error, version="res_new")
        runs[f"{tag}_gn"] = run(gauss_newton, rp.res, x0, rp.jac, rp.error)
    save("rosenbrock", runs, x0_i=x0_i, x0_ii=x0_ii, x0_iii=x0_iii)
    # --- rosenbrock_3d_test.py:20-36,74 (dense Jacobian -> lstsq path)
    import scipy.sparse

    def res2(x):
        return 2 ** 0.5 * np.concatenate([10 * (x[1:] - x[:-1] ** 2), 1 - x[:-1]])

    def jac2(x):
        b1 = 10 * scipy.sparse.eye(1, 2, k=1) - 20 * scipy.sparse.diags(x[:-1], shape=(1, 2))
        b2 = -scipy.sparse.eye(1, 2, k=0)
        return 2 ** 0.5 * scipy.sparse.block_array([[b1], [b2]]).todense()

    x0 = np.array([-1.0, 1.0])
    d = run(gauss_newton, res2, x0, jac2, None)
    save("rosenbrock_3d", dict(gn=d))
    # --- powell_divergence_test.py:17-61,87-102,192-204
    def pres(x, tau):
        return np.array([x[0] + 1, tau * x[0] ** 2 + x[0] - 1])

    def pjac(x, tau):
        return np.array([[1], [2 * tau * x[0] + 1]])

    def no_step_length_control(res, x, res_ev, jac_ev, args, descent_direction, *_):
        return 1, res(x + descent_direction, *args), 1

    state = dict(it=2)

    def too_small_steps(res, x, res_ev, jac_ev, args, descent_direction, *_):
        step_length = -1 / descent_direction[0] * 2 ** -state["it"]
        state["it"] += 1
        return step_length, res(x + step_length * descent_direction, *args), 1

    x0 = np.array([1.0])
    runs = {}
    for tau in (-5, 5):
        for ctl in (armijo_goldstein, too_small_steps, no_step_length_control):
            if tau == 5 and ctl is too_small_steps:
                continue
            state["it"] = 2
            runs[f"tau{tau}_{ctl.__name__}"] = run(gauss_newton, pres, x0, pjac, None, args=(tau,), max_iter=19,
                                                   step_length_control=ctl)
    save("powell", runs)


def g1025():
    pb, y, res, jac, err, u0 = bratu_setup(1025, 5, 10)
    save("bratu_g1025", dict(
        gnk_restart30=run(gauss_newton_krylow, res, u0, jac, err, max_iter=100, krylow_restart=30),
        gnk_k30=run(gauss_newton_krylow, res, u0, jac, err, max_iter=31),
    ), y_sample=y[sample_idx(y.shape[0])], u0_sample=u0[sample_idx(u0.shape[0])])


def g4097():
    pb, y, res, jac, err, u0 = bratu_setup(4097, 5, 10)
    save("bratu_g4097", dict(
        gnk_k30=run(gauss_newton_krylow, res, u0, jac, err, max_iter=31),
    ), y_sample=y[sample_idx(y.shape[0])], u0_sample=u0[sample_idx(u0.shape[0])])


def ttt():
    """time-to-tolerance configurations beyond compare_without_scaling (bratu_pde_test.py:76-103, grid_resolution=1):
    the same set-up on larger grids with version="res_new", which genuinely converges (tol=1e-8) within 100 iterations
    (res_old needs 106 at grid_nodes=257: more columns than max_iter=100 allows)."""
    for G in (257, 513, 1025):
        pb, y, res, jac, err, u0 = bratu_setup(G, 5, 10, h=1)
        save(f"bratu_g{G}_h1", dict(
            gnk_res_new=run(gauss_newton_krylow, res, u0, jac, err, loss_every=0, max_iter=100, version="res_new"),
        ), y_sample=y[sample_idx(y.shape[0])], u0_sample=u0[sample_idx(u0.shape[0])])


def manufactured():
    """bratu_pde_test.compare_manufactured_solution (:141-190; disabled in the reference's __main__, :337): the right-hand
    side is the CONTINUOUS operator applied to u = exp(-10 (x1^2 + x2^2)) (sympy, lambdified) instead of the discrete
    operator applied to u_true, so the solvers converge to the discrete solution u_h and the error curve levels off at
    the discretisation error.  y is stored in full: it is an input (no sympy needed to replay the run)."""
    import sympy as sp
    G, ALPHA, LAMBDA = 101, 5, 10
    pb = BratuPdeProblem(G, ALPHA, LAMBDA)
    sx, sy = sp.symbols("sp_x sp_y")
    su = sp.exp(-10 * (sx**2 + sy**2))
    sf = -sp.diff(su, sx, sx) - sp.diff(su, sy, sy) + ALPHA * sp.diff(su, sx) + LAMBDA * sp.exp(su)
    y = sp.lambdify((sx, sy), sf)(*pb.grid).flatten("F")
    res, jac, err = pb.make_res(y), pb.make_jac(), pb.make_error()
    np.random.seed(42)
    u0 = pb.u_true + 0.1 * np.random.normal(loc=0, scale=1, size=len(pb.u_true))
    save("bratu_g101_manufactured", dict(
        gn=run(gauss_newton, res, u0, jac, err),
        gnk_res_old=run(gauss_newton_krylow, res, u0, jac, err, max_iter=100),
        gnk_res_new=run(gauss_newton_krylow, res, u0, jac, err, max_iter=100, version="res_new"),
    ), y=y, u0=u0)


def cgx0():
    """cg_least_squares with an initial guess (gauss_newton.py:14,46,56 forward x0 to scipy's cg): the Bratu Jacobian at
    grid_nodes=34 and the chained Rosenbrock Jacobian, with and without the discarded unpreconditioned first run."""
    from gauss_newton import cg_least_squares
    flat = {}
    pb, y, res, jac, err, u0 = bratu_setup(34, 5, 10)
    rs = np.random.RandomState(11)
    A = -1 * jac(u0)
    r = res(u0)
    guess = 1e-3 * rs.normal(size=u0.shape[0])
    flat.update({"bratu/u": u0, "bratu/y": r, "bratu/x0": guess})
    for pre in (True, False):
        x, it = cg_least_squares(A, r, x0=guess.copy(), preconditioner=pre)
        flat[f"bratu/x_pre{int(pre)}"] = x
        flat[f"bratu/it_pre{int(pre)}"] = np.array(it)
    xr = 1.0 + 0.1 * rs.normal(size=1000)
    Ar = -1 * rosenbrock_problem.jac(xr)
    rr = rosenbrock_problem.res(xr)
    gr = 0.05 * rs.normal(size=1000)
    flat.update({"rosen/x": xr, "rosen/x0": gr})
    for pre in (True, False):
        x, it = cg_least_squares(Ar, rr, x0=gr.copy(), preconditioner=pre)
        flat[f"rosen/x_pre{int(pre)}"] = x
        flat[f"rosen/it_pre{int(pre)}"] = np.array(it)
    np.savez_compressed(os.path.join(OUT, "cg_x0.npz"), **flat)
    print("cg_x0.npz written:", {k: int(v) for k, v in flat.items() if "/it_" in k})


def sensitivity(G, name, runs_kw, seed=7):
    """The reference against ITSELF when the start vector u0 is perturbed by one unit in the last place (relative
    2^-52, random signs, seed 7).  The deviation of the perturbed trace from the unperturbed one is the conditioning of
    the reference's trajectory: no independent implementation (different summation order, FMA contraction, another
    exp) can be expected to reproduce the iterates more closely than a small multiple of this envelope.  (Perturbing y
    instead under-estimates it: |y| << |res(u0)| on these noise-dominated starts.)"""
    pb, y, res, jac, err, u0 = bratu_setup(G, 5, 10)
    sgn = np.random.RandomState(seed).choice([-1.0, 1.0], size=u0.shape[0])
    u0p = u0 * (1.0 + sgn * 2.0 ** -52)
    out = {}
    for rname, kw in runs_kw.items():
        out[rname] = run(gauss_newton_krylow, res, u0p, jac, err, loss_every=0, **kw)
    save(name, out)


def sensitivity_ls(G, name, runs_kw, seed=11):
    """The reference against ITSELF when every projected least-squares solution it computes
    (gauss_newton_krylow.py:89) is moved by ONE unit in the last place per component (random signs).  No independent
    implementation can reproduce the rounding of those k numbers, and the trajectory amplifies it: on the fine grids
    the first iterate x_1 = (c + d) v_0 is formed by a ~1e7-fold cancellation, so one ulp of d is ~1e-9 of x_1.
    (Perturbing u0 -- ``sensitivity`` above -- does not show this: the mathematical effect of a random 1-ulp change
    of 1.7e7 inputs averages out and the rounded d stays the same double.)"""
    import gauss_newton_krylow as ref_gnk
    pb, y, res, jac, err, u0 = bratu_setup(G, 5, 10)
    orig = ref_gnk.linear_least_squares
    out = {}
    try:
        for rname, kw in runs_kw.items():
            rs = np.random.RandomState(seed)

            def perturbed(A, yy):
                d = orig(A, yy)
                return d + rs.choice([-1.0, 1.0], size=d.shape[0]) * np.spacing(np.abs(d))

            ref_gnk.linear_least_squares = perturbed
            out[rname] = run(gauss_newton_krylow, res, u0, jac, err, loss_every=0, **kw)
    finally:
        ref_gnk.linear_least_squares = orig
    save(name, out)


def sensd4097():
    sensitivity_ls(4097, "bratu_g4097_sensd", dict(gnk_k30=dict(max_iter=31)))


def sensd1025():
    sensitivity_ls(1025, "bratu_g1025_sensd", dict(gnk_k30=dict(max_iter=31),
                                                   gnk_restart30=dict(max_iter=100, krylow_restart=30)))


def sens101():
    sensitivity(101, "bratu_g101_sens", dict(gnk_res_old=dict(max_iter=100),
                                             gnk_restart30=dict(max_iter=100, krylow_restart=30)))


def sens1025():
    sensitivity(1025, "bratu_g1025_sens", dict(gnk_k30=dict(max_iter=31),
                                               gnk_restart30=dict(max_iter=100, krylow_restart=30)))


def sens4097():
    sensitivity(4097, "bratu_g4097_sens", dict(gnk_k30=dict(max_iter=31)))


def sens4097b():
    """two more draws of the 1-ulp perturbation (other random sign patterns): one draw is ONE sample of the reference's
    conditioning; the parity bound uses the largest of the three envelopes per iteration"""
    sensitivity(4097, "bratu_g4097_sens2", dict(gnk_k30=dict(max_iter=31)), seed=8)
    sensitivity(4097, "bratu_g4097_sens3", dict(gnk_k30=dict(max_iter=31)), seed=9)


def sens1025b():
    sensitivity(1025, "bratu_g1025_sens2", dict(gnk_k30=dict(max_iter=31),
                                                gnk_restart30=dict(max_iter=100, krylow_restart=30)), seed=8)
    sensitivity(1025, "bratu_g1025_sens3", dict(gnk_k30=dict(max_iter=31),
                                                gnk_restart30=dict(max_iter=100, krylow_restart=30)), seed=9)


if __name__ == "__main__":
    for what in sys.argv[1:] or ["small"]:
        dict(small=small, g1025=g1025, g4097=g4097, kernels=kernels_fixture, sens101=sens101, sens1025=sens1025,
             sens4097=sens4097, sens4097b=sens4097b, sens1025b=sens1025b, sensd4097=sensd4097, sensd1025=sensd1025, ttt=ttt,
             cgx0=cgx0, manufactured=manufactured)[what]()
