"""Armijo-Goldstein backtracking -- mirror of the reference's ``armijo_goldstein.py``.

Rule (armijo_goldstein.py:47-72): with prev = sum(r^2) and g = sum((J d)^2), try s = s0, s0/2, ... (at most
``max_iter`` trials) and accept the first s with  prev - sum(res(x + s d)^2) >= 0.5 * s * g.

``armijo_goldstein`` keeps the reference's signature for host callers (custom drivers, the
``step_length_control`` plug-in point of ``gauss_newton``); its reductions run on the device.  The solvers
themselves use ``armijo_device``, where the trial residual norm comes straight out of the fused residual
kernel and only one scalar per trial reaches the host.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Callable, Tuple

import numpy as np

from . import _lib
from .device import CsrJacobian, DeviceVector, get_runtime, make_layout, ptr
from .partition import flat_layout_fields


class StepLengthConvergenceError(RuntimeError):
    message: str

    def __init__(self, message: str):
        super().__init__(message)
        self.message = message


def _failure(dnorm):
    return StepLengthConvergenceError(
        "The armijio_goldstein subroutine reached maximum iteration bound before principle was satisfied! Possible reasons:"
        + "\n- The max iteration count is not big enough to allow for a sufficiently small step size"
        + "\n- Or the descent direction is invalid."
        + f"Norm of descent_direction ={dnorm}."
    )


def armijo_device(trial_loss: Callable[[float], float], prev_loss: float, jac_dot_descent: Callable[[], float],
                  dnorm: Callable[[], float], max_iter: int = 100, initial_step_length: float = 1.0):
    """Core loop.  ``trial_loss(s)`` evaluates sum(res(x + s d)^2) on the device and returns it as a float;
    ``jac_dot_descent()`` returns sum((J d)^2) and is asked for after the first trial, so that one
    device->host read serves both.  Returns (step_length, trials)."""
    s = initial_step_length
    for it in range(max_iter):
        cur = trial_loss(s)
        if prev_loss - cur >= 0.5 * s * jac_dot_descent():
            return s, it + 1
        s /= 2
    raise _failure(dnorm())


def _device_sumsq(rt, host_vec):
    v = np.asarray(host_vec, dtype=np.float64).reshape(-1)
    f = flat_layout_fields(v.shape[0])
    t = rt.zeros(f["ld"])
    rt.upload(v, t[:v.shape[0]])
    out = rt.zeros(2)
    lay = make_layout(f)
    _lib.check(rt.lib.gnk_norm_stats(rt.ctx, C.byref(lay), ptr(t), ptr(out), rt.stream), "gnk_norm_stats")
    return float(rt.read(out, 1)[0])


def armijo_goldstein(res: Callable, x, res_ev, jac_ev, args: Tuple, descent_direction, max_iter: int = 100,
                     initial_step_length: float = 1.0) -> Tuple[float, Any, int]:
    """
    Parameters
    ----------
    res: Residual function, called as res(x, *args).
    x: Current argument value.
    res_ev: Evaluated residual at x.
    jac_ev: Evaluated jacobian of the residual at x (ndarray, scipy sparse, or a device Jacobian).
    descent_direction: Descent direction.
    max_iter: Maximum number of trials.
    initial_step_length: Initial step size.

    Returns
    -------
    step_length, the residual at x + step_length * descent_direction, number of trials used.
    """
    rt = get_runtime()
    x = np.asarray(x.materialize() if isinstance(x, DeviceVector) else x, dtype=np.float64)
    d = np.asarray(descent_direction, dtype=np.float64)
    prev_loss = _device_sumsq(rt, res_ev.materialize() if isinstance(res_ev, DeviceVector) else res_ev)
    if hasattr(jac_ev, "neg_rmatvec") and not isinstance(jac_ev, CsrJacobian):
        Jd = jac_ev @ d  # device stencil Jacobian
    else:
        import scipy.sparse as sp
        op = jac_ev if isinstance(jac_ev, CsrJacobian) else CsrJacobian(
            rt, jac_ev, isinstance(jac_ev, (sp.sparray, sp.spmatrix)))
        fi, fo = flat_layout_fields(op.p), flat_layout_fields(op.n_res)
        din, dout = rt.zeros(fi["ld"]), rt.zeros(fo["ld"])
        rt.upload(d.reshape(-1), din[:op.p])
        op.matmat(din, fi["ld"], 1, dout, fo["ld"])
        Jd = rt.download(dout[:op.n_res])
    g = _device_sumsq(rt, Jd)
    box = {}

    def trial(s):
        box["r"] = res(x + s * d, *args)
        r = box["r"]
        return _device_sumsq(rt, r.materialize() if isinstance(r, DeviceVector) else r)

    s, trials = armijo_device(trial, prev_loss, lambda: g, lambda: np.linalg.norm(d), max_iter, initial_step_length)
    return s, box["r"], trials
