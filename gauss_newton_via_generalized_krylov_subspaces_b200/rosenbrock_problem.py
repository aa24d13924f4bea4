"""Chained Rosenbrock problem -- mirror of the reference's ``rosenbrock_problem.py`` (p = 1000 parameters,
1998 residuals, 2997 non-zeros).  ``res`` / ``jac`` are host callables with the reference's signatures; the
solvers upload their outputs (CSR of J and of J^T) and run every solver operation on the device."""
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
