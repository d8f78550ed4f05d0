"""All-modes-at-once dispersion-curve extraction on the GPU.

The reference extracts a curve with ~13 *sequential* scalar determinant evaluations per period and
per mode (`_surf96.getc`: ladder bracketing + Neville refinement, period k starts from the root of
period k-1).  Here the same roots come out of two batched steps built on the dense kernels:

  1. one dense sweep `det[K, NF]` over the trial-velocity ladder c_k = c_low + k*dc (the reference's
     step `dc`, starting like `getc` 10 % below the slowest layer's half-space Rayleigh velocity and
     ending at max(Vs)): every sign change along c brackets a root, the n-th sign change at a period is
     mode n (the fundamental is the slowest root, overtones follow in increasing c);
  2. all brackets of all periods and modes are refined together by safeguarded regula falsi
     (Illinois), one pointwise-kernel launch per iteration, to a relative width of `rtol`.

Modes that do not exist at a period (no n-th sign change below max(Vs)) return 0, like the reference.
Differences from `surf96` by construction: brackets come from a fixed ladder instead of a ladder
that follows the previous period's root, so two roots closer than `dc` are missed by both but not
necessarily at the same periods; roots are converged to `rtol` (default 1e-12) whereas the reference
stops at 1e-6 relative, so values agree to the reference's own tolerance (~3e-6 km/s), not to 1e-6.
`adsurf_b200._surf96.surf96` remains the bit-for-bit-compatible path.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from ._surf96 import gtsolh


@torch.no_grad()
def dispersion_curves(t, d, a, b, rho, modes=(0,), dc=0.005, rtol=1e-12, max_iter=60, device=None):
    """Phase velocities c[len(modes), len(t)] of the requested Rayleigh modes of ONE layered model.

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
    change = sg[:-1] != sg[1:]                                    # [K-1, NF] bracket k = (c_k, c_{k+1})
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
    root = (0.5 * (lo_f + hi_f)).reshape(M, NF)
    root = torch.where(ok & (root <= betmx), root, torch.zeros_like(root))
    return root.cpu().numpy()
