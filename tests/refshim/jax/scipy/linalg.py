"""`jax.scipy.linalg.solve_triangular` (TEST INFRASTRUCTURE): the reference calls it on host NumPy arrays
(deluca/igpc/lqr_solver.py:166), so this is scipy's."""
import numpy as np
import scipy.linalg as _sl


def solve_triangular(a, b, lower=False, check_finite=True, **kw):
    return _sl.solve_triangular(np.asarray(a), np.asarray(b), lower=lower, check_finite=check_finite)
