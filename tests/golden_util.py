"""Helpers shared by the oracle and GPU parity tests (test infrastructure)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class Golden:
    def __init__(self, name):
        self.z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)

    def run(self, rname):
        pre = rname + "/"
        return {k[len(pre):]: self.z[k] for k in self.z.files if k.startswith(pre)}

    def __getitem__(self, k):
        return self.z[k]


class Recorder:
    """callback(x, nfev, cg_iter) recorder mirroring oracle/gen_golden.py:run."""

    def __init__(self, idx, error=None):
        self.idx = idx
        self.error = error
        self.xnorm, self.err, self.nfev, self.cg, self.xs = [], [], [], [], []

    def __call__(self, x, nfev, cg_iter):
        x = np.asarray(x)
        self.xnorm.append(np.linalg.norm(x))
        self.err.append(self.error(x) if self.error is not None else np.nan)
        self.nfev.append(-1 if nfev is None else nfev)
        self.cg.append(-1 if cg_iter is None else cg_iter)
        self.xs.append(x[self.idx].copy())


def _envelope(golden_run, perturbed_run):
    a, b = golden_run["xs"], perturbed_run["xs"]
    n = min(len(a), len(b))
    scale = np.max(np.abs(a[:n]), axis=1, keepdims=True)
    env = np.max(np.abs(a[:n] - b[:n]) / scale, axis=1)
    return np.concatenate([env, np.full(len(a) - n, env[-1] if n else 0.0)])


def sensitivity_bound(golden_run, perturbed_runs, ls_perturbed_run=None, floor=1e-10, factor=10.0, window=1):
    """Per-iteration tolerance for the large grids, from the conditioning of the REFERENCE's own trajectory.

    Fixtures (oracle/gen_golden.py), all produced by the unmodified reference:
    * ``perturbed_runs``: the reference re-run with its start vector u0 perturbed by one unit in the last place (random
      signs; several draws).  On the fine grids the outcome is bimodal: usually the rounded least-squares solution d of
      the first iterations stays the same double and the trace moves by ~1e-11; when it flips by one ulp (one draw of
      three at 4096^2) the first iterates move by 2e-9 .. 4e-9, because x_1 = (c + d) v_0 is formed by a ~1e7-fold
      cancellation;
    * ``ls_perturbed_run``: the reference re-run with every projected least-squares solution moved by one ulp per
      component (``sensitivity_ls``) -- the same mechanism, provoked deliberately: 1.1e-9 at iteration 1, 1.6e-9 at
      iterations 3-4, decaying to 1e-10 by iteration 9 and 2e-11 by iteration 13.
    No independent implementation (other summation order, FMA contraction, another exp) can reproduce the rounding of
    d: two backward-stable solvers of the same one-column problem differ by a handful of ulps (measured against
    LAPACK's result: 0 ulp at 4096^2 on one GPU, 1 ulp on 2 and 8 GPUs, 6 ulp at 1024^2).  So the deviation is held to
    ``factor`` (10, i.e. ten ulps of d) x the LARGEST deviation the reference shows against itself under a ONE-ulp
    change, at that iteration (+- ``window`` iterations), over all fixtures -- and never below ``floor`` = the 1e-10 bar
    of the north star, which is what applies wherever the reference's trajectory is that well determined (from
    iteration ~16 on at 4096^2, ~3 on at 1024^2, and at the end).  (Round 1 used 100 x the RUNNING maximum of a
    single u0 draw, which never came back down to 1e-10.)"""
    if isinstance(perturbed_runs, dict):
        perturbed_runs = [perturbed_runs]
    envs = [_envelope(golden_run, p) for p in perturbed_runs]
    if ls_perturbed_run is not None:
        envs.append(_envelope(golden_run, ls_perturbed_run))
    env = np.maximum.reduce(envs)
    n = len(env)
    win = np.array([env[max(0, i - window):min(n, i + window + 1)].max() for i in range(n)])
    return np.maximum(floor, factor * win)


def bound_for(name, rname, floor=1e-10):
    """the tolerance of golden run ``name``:``rname`` from every sensitivity fixture that exists for it"""
    gr = Golden(name).run(rname)
    perts = []
    for suffix in ("_sens", "_sens2", "_sens3"):
        if os.path.exists(os.path.join(GOLDEN, name + suffix + ".npz")):
            r = Golden(name + suffix).run(rname)
            if "xs" in r:
                perts.append(r)
    ls = None
    if os.path.exists(os.path.join(GOLDEN, name + "_sensd.npz")):
        r = Golden(name + "_sensd").run(rname)
        ls = r if "xs" in r else None
    return sensitivity_bound(gr, perts, ls, floor=floor)


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (den if den > 0 else 1.0))


def check_trace(rec, g, tol, upto=None, check_cg=False):
    """compare a Recorder with a golden run: same number of callbacks, same nfev
    sequence, iterates (|x|, sampled entries) within tol relative."""
    n = len(g["xnorm"]) if upto is None else upto
    assert len(rec.xnorm) >= n if upto is not None else len(rec.xnorm) == n, (len(rec.xnorm), n)
    assert list(rec.nfev[:n]) == list(g["nfev_cb"][:n])
    xs = np.array(rec.xs[:n])
    scale = np.max(np.abs(g["xs"][:n]), axis=1, keepdims=True)
    dv = np.max(np.abs(xs - g["xs"][:n]) / scale, axis=1)
    dnv = np.abs(np.array(rec.xnorm[:n]) - g["xnorm"][:n]) / g["xnorm"][:n]
    tolv = np.broadcast_to(np.asarray(tol, dtype=np.float64), (n,)) if np.ndim(tol) == 0 else np.asarray(tol)[:n]
    assert np.all(dv <= tolv) and np.all(dnv <= tolv), (list(zip(dv, tolv))[:8], float(dv.max()), float(dnv.max()))
    d = float(dv.max())
    print(f"[parity] {n} iterations: max deviation {d:.3e}, max deviation / tolerance {float(np.max(dv / tolv)):.3f}")
    if check_cg:
        assert list(rec.cg[:n]) == list(g["cg_iter"][:n])
    return d
