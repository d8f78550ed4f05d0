"""Modal dispersion curves on the GPU — two extractors.

`dispersion_curves` — the reference's own root search (`_surf96.getc/getsol/nevill`: ladder bracketing from the
previous period's root, Neville refinement to 1e-6 relative, mode bookkeeping with `clow`, `one = 1e-2`,
`onea = 1.5`, zeros for absent modes) run by the CUDA kernel `k_dispersion`, one warp per model (C ABI
`adsurf_dispersion_curves`).  Same brackets, same iterates: mode indices are the reference's and roots agree with
`_surf96.surf96` within 1e-6 km/s.  The search is sequential within a model, so the parallel axis is the ensemble:
the curves of 4096 models cost what one costs.  `cal_dispersion_curve(..., device="cuda")` calls it.

`dispersion_curves_dense` — a different algorithm for the same roots: one dense sweep `det[K, NF]` over a fixed
trial-velocity ladder (every sign change along c brackets a root; the n-th sign change at a period is mode n), then
all brackets of all periods and modes refined together by safeguarded regula falsi (Illinois), one pointwise-kernel
launch per iteration, to a relative width of `rtol` (default 1e-12).  Converged roots, i.e. zeros of the secular
function to rounding — they agree with the reference to ITS tolerance (1e-6 relative, ~3e-6 km/s), and brackets
come from a fixed ladder instead of one that follows the previous root, so two roots closer than `dc` may be
resolved at different periods than `getc` does.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _capi, ops
from ._surf96 import DispersionError, gtsolh


@torch.no_grad()
def dispersion_curves(t, d, a, b, rho, modes=(0,), dc=0.005, device=None, itype=0, dt=0.025, ifunc=2):
    """Phase (itype=0) or group (itype=1) velocities of the requested Rayleigh (ifunc=2, Dunkin's matrix) or Love
    (ifunc=1) modes by the reference's root search on the GPU (`surf96`'s ifunc, _surf96.py:693-760).

    t: periods [NF] (or [S,NF]); d, a, b, rho: thickness, Vp, Vs, density [L] (one model) or [S,L] (an ensemble).
    Returns c[len(modes), NF] (one model) or c[S, len(modes), NF]; 0 where a mode does not exist.
    Group velocity as the reference forms it (_surf96.py:735-758): two phase curves at t/(1+dt) and t/(1-dt),
    U = (f1 - f2) / (f1/c1 - f2/c2), 0 where the second curve has no root.
    Raises DispersionError if the fundamental mode of any model cannot be bracketed (as `_surf96.getc` does)."""
    if itype == 1:
        t_np = np.asarray(t, dtype=np.float64)
        t1, t2 = t_np / (1.0 + dt), t_np / (1.0 - dt)
        c1 = dispersion_curves(t1, d, a, b, rho, modes, dc, device, ifunc=ifunc)
        c2 = dispersion_curves(t2, d, a, b, rho, modes, dc, device, ifunc=ifunc)
        f1 = np.broadcast_to((1.0 / t1) if t_np.ndim == 1 else (1.0 / t1)[:, None, :], c1.shape)
        f2 = np.broadcast_to((1.0 / t2) if t_np.ndim == 1 else (1.0 / t2)[:, None, :], c1.shape)
        out = np.zeros_like(c1)
        ok = c2 > 0.0
        out[ok] = (f1[ok] - f2[ok]) / (f1[ok] / c1[ok] - f2[ok] / c2[ok])
        return out
    dev = ops._device_of(device=device)
    D, A_, B_, R = (torch.as_tensor(np.asarray(x, dtype=np.float64), device=dev) for x in (d, a, b, rho))
    single = D.dim() == 1
    if single:
        D, A_, B_, R = (x.reshape(1, -1) for x in (D, A_, B_, R))
    S = D.shape[0]
    T = torch.as_tensor(np.asarray(t, dtype=np.float64), device=dev)
    T = T.reshape(1, -1).expand(S, -1) if T.dim() == 1 else T
    if bool((B_[:, 0] <= 0).any()) and D.shape[1] < 3:  # layer 0 water (Vs <= 0): as in _surf96.getc
        raise NotImplementedError("adsurf_b200: a water layer needs a solid layer above the half-space")
    modes = [int(m) for m in modes]
    if ifunc not in (1, 2):
        raise NotImplementedError("adsurf_b200: Love (ifunc=1) and Rayleigh / Dunkin (ifunc=2) only")
    c, status = _capi.dispersion_curves(D.contiguous(), A_.contiguous(), B_.contiguous(), R.contiguous(),
                                        T.contiguous(), max(modes) + 1, float(dc), ifunc=ifunc)
    if bool(status.any()):
        raise DispersionError("failed to find root for fundamental mode")
    out = c[:, modes, :].cpu().numpy()
    return out[0] if single else out


@torch.no_grad()
def dispersion_curves_dense(t, d, a, b, rho, modes=(0,), dc=0.005, rtol=1e-12, max_iter=60, device=None):
    """Converged roots c[len(modes), len(t)] of the requested Rayleigh modes of ONE layered model (dense-grid
    bracketing + batched Illinois refinement; see the module docstring).

    t: periods (s); d, a, b, rho: thickness, Vp, Vs, density per layer (last = half-space)."""
    t_np = np.asarray(t, dtype=np.float64).reshape(-1)
    d_np, a_np, b_np, r_np = (np.asarray(x, dtype=np.float64).reshape(-1) for x in (d, a, b, rho))
    jmn = int(np.argmin(np.where(b_np > 0.01, b_np, np.inf)))
    c_low = 0.9 * float(gtsolh(a_np[jmn], b_np[jmn]))  # getc's starting value (_surf96.py:625-628)
    betmx = float(b_np.max())
    K = int(np.ceil((betmx - c_low) / dc)) + 2
    cgrid = c_low + dc * np.arange(K)
    dev = ops._device_of(device=device)
    T = torch.as_tensor(t_np, device=dev)
    C = torch.as_tensor(cgrid, device=dev)
    D, A_, B_, R = (torch.as_tensor(x, device=dev) for x in (d_np, a_np, b_np, r_np))
    det = ops.det_grid(C, T, D, A_, B_, R)                      # [K, NF]
    sg = torch.sign(det)
    # a bracket needs finite values of opposite sign at both ends (NaN or an exact zero at a ladder point is not a
    # sign change: it would shift the mode index of every higher root at that period)
    change = (sg[:-1] * sg[1:]) < 0                               # [K-1, NF] bracket k = (c_k, c_{k+1})
    order = torch.cumsum(change.to(torch.int32), dim=0)          # running count of sign changes along c
    NF = T.numel()
    modes = list(modes)
    M = len(modes)
    lo = torch.zeros((M, NF), dtype=torch.float64, device=dev)
    hi = torch.zeros_like(lo)
    ok = torch.zeros((M, NF), dtype=torch.bool, device=dev)
    for i, n in enumerate(modes):
        hit = change & (order == n + 1)                           # the (n+1)-th sign change per period
        has = hit.any(dim=0)
        idx = torch.argmax(hit.to(torch.int8), dim=0)
        lo[i], hi[i], ok[i] = C[idx], C[idx + 1], has
    # Illinois regula falsi on every bracket at once
    Tm = T.expand(M, NF).reshape(1, -1).contiguous()
    lo_f, hi_f = lo.reshape(1, -1).clone(), hi.reshape(1, -1).clone()
    lo_f[~ok.reshape(1, -1)] = c_low                               # dummy but valid brackets for absent modes
    hi_f[~ok.reshape(1, -1)] = c_low + dc

    def F(c):
        return ops.det_points(c.contiguous(), Tm, D, A_, B_, R)

    f_lo, f_hi = F(lo_f), F(hi_f)
    side = torch.zeros_like(lo_f, dtype=torch.int8)
    for _ in range(max_iter):
        denom = f_hi - f_lo
        x = torch.where(denom != 0, (lo_f * f_hi - hi_f * f_lo) / denom, 0.5 * (lo_f + hi_f))
        x = torch.where((x > lo_f) & (x < hi_f), x, 0.5 * (lo_f + hi_f))
        fx = F(x)
        left = torch.sign(fx) == torch.sign(f_lo)                 # root is in (x, hi)
        # Illinois: halve the retained end's value when the same end is retained twice in a row
        f_hi = torch.where(left & (side == 1), 0.5 * f_hi, f_hi)
        f_lo = torch.where(~left & (side == -1), 0.5 * f_lo, f_lo)
        side = torch.where(left, torch.ones_like(side), -torch.ones_like(side))
        lo_f = torch.where(left, x, lo_f)
        f_lo = torch.where(left, fx, f_lo)
        hi_f = torch.where(left, hi_f, x)
        f_hi = torch.where(left, f_hi, fx)
        if bool(torch.all((hi_f - lo_f) <= rtol * hi_f)):
            break
    # whichever end has the smaller |F| (an Illinois end can be stale when the iteration cap is hit)
    f_lo_true, f_hi_true = F(lo_f), F(hi_f)
    root = torch.where(f_lo_true.abs() <= f_hi_true.abs(), lo_f, hi_f)
    wide = ((hi_f - lo_f) > 1e3 * rtol * hi_f) & ok.reshape(1, -1)
    if bool(wide.any()):
        import warnings
        warnings.warn(f"dispersion_curves_dense: {int(wide.sum())} bracket(s) did not converge in {max_iter} iterations")
    root = root.reshape(M, NF)
    root = torch.where(ok & (root <= betmx), root, torch.zeros_like(root))
    return root.cpu().numpy()
