"""Result record returned by both solvers -- same fields and spelling as the reference's
``regression_result.py`` (note ``nrev``; ``full_njev`` is declared there but never set)."""
from __future__ import annotations

from typing import Optional

import numpy as np


class RegressionResult:
    """
    Attributes
    ----------
    method_name: Name of the method used ("gauss newton krylow" / "gauss newton").
    x: Solution for the regression parameters (host ndarray).
    success: If the algorithm terminated by its tolerance test.
    nrev: Count of residual evaluations.
    njev: Count of jacobian evaluations.
    nit: Count of iterations.
    """

    method_name: str
    x: np.ndarray
    success: bool
    nrev: Optional[int]
    full_njev: Optional[int]
    njev: Optional[int]
    nit: int

    def __init__(self, method_name, x, success, nrev, njev, nit):
        self.method_name = method_name
        self.x = x
        self.success = success
        self.nrev = nrev
        self.njev = njev
        self.nit = nit

    def __str__(self):
        how = "converged successfuly to" if self.success else "failed to terminate and stopped at"
        return (f"{self.method_name} {how} {self.x}. After {self.nit} iterations using {self.nrev} "
                f"evaluations of the residual, ")
