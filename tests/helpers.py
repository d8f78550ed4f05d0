"""Shared parity helpers: the tolerance rule of SURVEY.md §8(c) ("Parity noise floor")."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# north_star tolerance: 1e-9 relative in FP64.  Read as SURVEY.md §8(c) defines it:
# |new-ref| <= RTOL * max(|ref|, scale), scale = per-(model,f) column max|det| for determinants and
# per-(model, parameter kind) max-over-layers |grad| for gradients.
RTOL = 1e-9


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


def col_scaled_err(new, ref, axis):
    """max over entries of |new-ref| / max(|ref|, max_axis|ref|)."""
    new, ref = np.asarray(new, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    scale = np.max(np.abs(ref), axis=axis, keepdims=True)
    denom = np.maximum(np.abs(ref), scale)
    denom = np.where(denom == 0, 1.0, denom)
    return float(np.max(np.abs(new - ref) / denom))


def elementwise_err(new, ref):
    new, ref = np.asarray(new, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    denom = np.where(ref == 0, 1.0, np.abs(ref))
    return float(np.max(np.abs(new - ref) / denom))


def assert_det_close(new, ref, axis, rtol=RTOL, what="det"):
    err = col_scaled_err(new, ref, axis)
    assert err <= rtol, f"{what}: column-scaled error {err:.3e} > {rtol:g} (element-wise {elementwise_err(new, ref):.3e})"
    return err


def assert_grad_close(new, ref, rtol=RTOL, what="grad"):
    """gradient arrays [S,L] (scale = max over layers per model) or [L]."""
    err = col_scaled_err(new, ref, axis=-1)
    assert err <= rtol, f"{what}: layer-scaled error {err:.3e} > {rtol:g} (element-wise {elementwise_err(new, ref):.3e})"
    return err


# Trial velocities that coincide (to a few ulp) with a layer velocity are degenerate: there the
# reference evaluates sinh-like terms as (1 - exp(-2q))/2 with q ~ 1e-8, i.e. with a RELATIVE rounding
# error of ~1e-8 of its own, and the half-space vector is linear in sqrt(|k - k_beta|).  The reference's
# value at such a sample is only defined to ~1e-8; parity there is asserted at COINCIDENT_RTOL.
COINCIDENT_RTOL = 1e-7


def coincident_rows(c, *velocities, rel=1e-12):
    """boolean mask over trial velocities c[K] that coincide with any entry of the velocity arrays"""
    c = np.asarray(c, dtype=np.float64).reshape(-1, 1)
    v = np.concatenate([np.asarray(x, dtype=np.float64).reshape(-1) for x in velocities]).reshape(1, -1)
    return np.any(np.abs(c - v) <= rel * np.abs(v), axis=1)
