"""Newton iteration for small vector systems -- mirror of ``sem.rootfind``.

Off the hot path (only the inverse element map of point
location uses it, sem/mapping.py:171-172); kept so that ``Mapping.inv`` and
``DOFManager.find_elem_containing_point`` keep working.
"""
import numpy as np

__all__ = ["SolverFailure", "newton"]


class SolverFailure(Exception):
    """A non-linear solve did not reach its tolerance."""


def newton(f, x0, jac, it_max, tol):
    """Newton-Raphson on ``f(x) = 0`` from ``x0`` (updated in place like the
    reference, sem/rootfind.py:22-53); stops when ``|dx|_2 <= tol``."""
    x = x0[:]
    for _ in range(it_max):
        step = np.linalg.solve(np.atleast_2d(jac(x)), -np.atleast_1d(f(x)))
        x += step
        if np.isclose(np.linalg.norm(step), 0.0, atol=tol):
            return x
    raise SolverFailure("Maximum number of iterations exceeded before"
                        "tolerence could be met.")
