"""adsurf_b200 — B200-native Rayleigh secular-function path of ADsurf (determinant + adjoint).

Public surface (mirrors the reference's module functions for this path):
    adsurf_b200.ops            det_points / det_grid / det_grid_minmax / dmf_data_misfit / sensitivity
    adsurf_b200._cps.*         dltar_vector / dltar_matrix shims with the reference signatures
    adsurf_b200._model         iter_cal_grad / multi_cal_grad / iter_inversion / multi_inversion
    adsurf_b200._capi          ctypes binding of include/adsurf_b200.h
The arithmetic runs only in libadsurf_b200.so (hand-written sm_100a CUDA); there is no CPU path.
"""
from . import _capi  # noqa: F401
from .ops import (  # noqa: F401
    det_grid,
    det_grid_minmax,
    det_grid_signbits,
    det_points,
    dmf_data_misfit,
    sensitivity,
)

__version__ = "0.1.0"
