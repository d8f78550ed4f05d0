"""Shared argument handling for the `dltar_*` shims.

Reference dispatchers: _cps/_surf96_vectorAll_gpu.py:194-218, _cps/_surf96_matrixAll_gpu.py:183-191.
Only ifunc == 2 (Rayleigh, Dunkin) with llw == -1 (no water layer) is executable in the reference's
torch path (the Love / fast-delta / water branches are broken there); the shims
raise NotImplementedError for the others instead of reproducing those crashes.
"""
import torch

from .. import ops


def check_mode(ifunc, llw):
    if ifunc != 2:
        raise NotImplementedError(
            "adsurf_b200 implements ifunc=2 (Rayleigh, Dunkin matrix) only; the reference's Love and "
            "fast-delta torch paths are not executable")
    if llw == 0:
        raise NotImplementedError("adsurf_b200: water-layer branch (llw=0) is not executable in the reference either")


def resolve_device(device, *tensors):
    """Where the caller wants the result.  The reference passes `device` straight to `.to()`
    (README shows device="Cuda", which torch rejects) — lower-case it here."""
    if isinstance(device, torch.device):
        return device
    return torch.device(str(device).lower())


def run_points(vc, t, d, alpha, beta, rho, ifunc, llw, device):
    check_mode(ifunc, llw)
    want = resolve_device(device)
    out = ops.det_points(vc, t, d, alpha, beta, rho, device=want if want.type == "cuda" else None)
    return out if want.type == "cuda" else out.to(want)


def run_grid(vlist, tlist, d, a, b, rho, ifunc, llw, device):
    check_mode(ifunc, llw)
    want = resolve_device(device)
    out = ops.det_grid(vlist, tlist, d, a, b, rho, device=want if want.type == "cuda" else None)
    return out if want.type == "cuda" else out.to(want)
