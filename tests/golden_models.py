"""Model builders shared by tests (numpy only; mirrors the Brocher relations of the reference,
_utils.py:45-75)."""
import numpy as np


def brocher_np(vs):
    vs = np.asarray(vs, dtype=np.float64)
    vp = 0.9409 + 2.0947 * vs - 0.8206 * vs**2 + 0.2683 * vs**3 - 0.0251 * vs**4
    rho = 1.6612 * vp - 0.4721 * vp**2 + 0.0671 * vp**3 - 0.0043 * vp**4 + 0.000106 * vp**5
    return vp, rho
