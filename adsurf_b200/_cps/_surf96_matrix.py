"""Drop-in for the reference module ADsurf/_cps/_surf96_matrix.py (function `dltar_matrix`).

Same call signature; the arithmetic runs in the CUDA kernels (float64).  `device` says where the
result tensor should live; computation is always on the GPU (there is no CPU fallback).
"""
from ._common import run_grid

__all__ = ["dltar_matrix"]


def dltar_matrix(vlist, tlist, d, a, b, rho, ifunc, llw, device="cpu"):
    """Secular function on the grid: [K,NF] for 1-D model arrays, [S,K,NF] for [S,L] arrays."""
    return run_grid(vlist, tlist, d, a, b, rho, ifunc, llw, device)
