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


def sensitivity_bound(golden_run, perturbed_run, floor=1e-10, factor=30.0):
    """Per-iteration tolerance for the large grids: the unmodified reference, re-run with its start vector perturbed by
    one unit in the last place (oracle/gen_golden.py:sensitivity), moves by env_i at iteration i.  An independent
    implementation is held to max(floor, factor * running max of env): i.e. to the 1e-10 bar wherever the reference's
    own trajectory is that well determined, and to a fixed multiple of its 1-ulp conditioning where it is not (a few
    early iterations, where x = c * v0 is formed by cancellation).  Measured GPU/envelope ratios are <= 10 (printed by
    ``check_trace`` / reported by bench.py's ``parity.max_dev_over_bound``: 0.31 of the bound at 4096^2)."""
    a, b = golden_run["xs"], perturbed_run["xs"]
    n = min(len(a), len(b))
    scale = np.max(np.abs(a[:n]), axis=1, keepdims=True)
    env = np.max(np.abs(a[:n] - b[:n]) / scale, axis=1)
    return np.maximum(floor, factor * np.maximum.accumulate(env))


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
