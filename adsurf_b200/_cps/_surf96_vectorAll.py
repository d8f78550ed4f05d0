"""Drop-in for the reference module ADsurf/_cps/_surf96_vectorAll.py (function `dltar_vector`).

Same call signature; the arithmetic runs in the CUDA kernels (float64).  `device` says where the
result tensor should live; computation is always on the GPU (there is no CPU fallback).
"""
from ._common import run_points

__all__ = ["dltar_vector"]


def dltar_vector(vc, t, d, alpha, beta, rho, ifunc, llw, device="cpu"):
    """Secular function at the picks (vc, t): [N] for 1-D model arrays, [S,N] for [S,L] arrays."""
    return run_points(vc, t, d, alpha, beta, rho, ifunc, llw, device)
