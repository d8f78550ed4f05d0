"""CPU oracle for the ADsurf Rayleigh secular-function hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain-PyTorch (CPU, float64 by default) restatement of the reference's
Dunkin delta-matrix recursion and of the Determinant Misfit Function built on it.  It is the
checker for the CUDA product path; nothing under ``adsurf_b200/`` may import it.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs use it.

Parity pin: ``tests/golden/*.npz`` hold inputs and outputs (determinants and autograd gradients)
produced by the *unmodified reference code* (``/root/reference/ADsurf/_cps/*``) run in this
container in float64 (its ``numpy2tensor`` rebound to a float64 version + default dtype float64,
SURVEY.md §8c) and in its native float32; ``tests/golden/make_golden.py`` is the generating
script and ``tests/test_oracle_golden.py`` checks this oracle against every one of them.

Reference lines each function follows (paths relative to /root/reference/ADsurf):
  eigenfunction_terms  <- _cps/_surf96_vectorAll_gpu.py:102-192  (var_vector; twins var_matrix)
  dunkin_matrix_apply  <- _cps/_surf96_vectorAll_gpu.py:52-100   (dnka_vector) + :376-383 (E·CA)
  secular_rayleigh     <- _cps/_surf96_vectorAll_gpu.py:303-399  (dltar4_vector, llw=-1)
  det_points           <- dltar_vector  (_cps/_surf96_vectorAll_gpu.py:194, _vector_gpu.py:194)
  det_grid             <- dltar_matrix  (_cps/_surf96_matrixAll_gpu.py:183/242, _matrix_gpu.py:186/245)
  brocher_vp_rho       <- _model.py:431-433
  dmf_loss             <- _model.py:430-478 (multi_cal_grad.forward), :305-344 (iter_cal_grad.forward)

Differences from the reference that do not change values: boolean-mask scatter/gather is
expressed with ``torch.where`` (same selected values, far fewer ATen calls), the 5x5 matrix is
never materialised (the row-vector product is written out), and broadcasting replaces
``np.repeat``.  The operation order inside every formula is kept so float64 results agree with the
reference to rounding (the golden test asserts <= 1e-13 relative to the column scale).
"""
from __future__ import annotations

import math

import torch

TWOPI = 2.0 * math.pi


def eigenfunction_terms(p, q, ra, rb, wvno, xka, xkb, dpth):
    """Exponent-scaled layer eigenfunction products (var_vector, :102-192)."""
    zero = torch.zeros_like(wvno)
    one = torch.ones_like(wvno)

    # P part (:109-136)
    lt = wvno < xka
    gt = wvno > xka
    eq = wvno == xka
    sin_p = torch.sin(p)
    cos_p = torch.cos(p)
    fac = torch.where(gt & (p < 16.0), torch.exp(-2.0 * p), zero)
    sinp_ev = (1.0 - fac) * 0.5
    pex = torch.where(gt, p, zero)
    cosp = torch.where(lt, cos_p, torch.where(gt, (1.0 + fac) * 0.5, torch.where(eq, one, zero)))
    w = torch.where(lt, sin_p / ra, torch.where(gt, sinp_ev / ra, torch.where(eq, dpth, zero)))
    x = torch.where(lt, -ra * sin_p, torch.where(gt, ra * sinp_ev, zero))

    # S part (:138-167)
    lt = wvno < xkb
    gt = wvno > xkb
    eq = wvno == xkb
    sin_q = torch.sin(q)
    cos_q = torch.cos(q)
    fac = torch.where(gt & (q < 16.0), torch.exp(-2.0 * q), zero)
    sinq_ev = (1.0 - fac) * 0.5
    sex = torch.where(gt, q, zero)
    cosq = torch.where(lt, cos_q, torch.where(gt, (1.0 + fac) * 0.5, torch.where(eq, one, zero)))
    y = torch.where(lt, sin_q / rb, torch.where(gt, sinq_ev / rb, torch.where(eq, dpth, zero)))
    z = torch.where(lt, -rb * sin_q, torch.where(gt, rb * sinq_ev, zero))

    # products (:170-182)
    exa = pex + sex
    a0 = torch.where(exa < 60.0, torch.exp(-exa), zero)
    return (a0, cosp * cosq, cosp * y, cosp * z, cosq * w, cosq * x, x * y, x * z, w * y, w * z)


def dunkin_matrix_apply(e, wvno2, gam, gammk, rho, terms):
    """Row vector ``e`` (tuple of 5) times Dunkin's compound matrix (dnka_vector :52-100, :379)."""
    a0, cpcq, cpy, cpz, cqw, cqx, xy, xz, wy, wz = terms
    gamm1 = gam - 1.0
    twgm1 = gam + gamm1
    gmgmk = gam * gammk
    gmgm1 = gam * gamm1
    gm1sq = gamm1 * gamm1
    rho2 = rho * rho
    a0pq = a0 - cpcq
    t = -2.0 * wvno2

    c00 = cpcq - 2.0 * gmgm1 * a0pq - gmgmk * xz - wvno2 * gm1sq * wy
    c01 = (wvno2 * cpy - cqx) / rho
    c02 = -(twgm1 * a0pq + gammk * xz + wvno2 * gamm1 * wy) / rho
    c03 = (cpz - wvno2 * cqw) / rho
    c04 = -(2.0 * wvno2 * a0pq + xz + wvno2 * wvno2 * wy) / rho2
    c10 = (gmgmk * cpz - gm1sq * cqw) * rho
    c11 = cpcq
    c12 = gammk * cpz - gamm1 * cqw
    c13 = -wz
    c14 = c03
    c30 = (gm1sq * cpy - gmgmk * cqx) * rho
    c31 = -xy
    c32 = gamm1 * cpy - gammk * cqx
    c33 = cpcq
    c34 = c01
    c40 = -(2.0 * gmgmk * gm1sq * a0pq + gmgmk * gmgmk * xz + gm1sq * gm1sq * wy) * rho2
    c41 = c30
    c42 = -(gammk * gamm1 * twgm1 * a0pq + gam * gammk * gammk * xz + gamm1 * gm1sq * wy) * rho
    c43 = c10
    c44 = c00
    c20 = t * c42
    c21 = t * c32
    c22 = a0 + 2.0 * (cpcq - c00)
    c23 = t * c12
    c24 = t * c02

    e0, e1, e2, e3, e4 = e
    # same accumulation order as a 1x5 @ 5x5 dot product (j = 0..4)
    n0 = e0 * c00 + e1 * c10 + e2 * c20 + e3 * c30 + e4 * c40
    n1 = e0 * c01 + e1 * c11 + e2 * c21 + e3 * c31 + e4 * c41
    n2 = e0 * c02 + e1 * c12 + e2 * c22 + e3 * c32 + e4 * c42
    n3 = e0 * c03 + e1 * c13 + e2 * c23 + e3 * c33 + e4 * c43
    n4 = e0 * c04 + e1 * c14 + e2 * c24 + e3 * c34 + e4 * c44
    return (n0, n1, n2, n3, n4)


def secular_rayleigh(wvno, omega0, d, alpha, beta, rho):
    """dltar4_vector (:303-399) for llw=-1 on already-broadcastable tensors.

    ``wvno`` and ``omega0`` have the sample shape ``[S, ...]``; ``d, alpha, beta, rho`` are
    ``[S, L]``.  Returns the secular function on the sample shape.
    """
    nd = wvno.dim()
    L = d.shape[-1]

    def layer(par, m):  # [S] -> [S,1,...]
        return par[:, m].reshape((-1,) + (1,) * (nd - 1))

    omega = torch.clamp(omega0, min=1.0e-4)  # torch.max(omega, 1e-4) (:326)
    wvno2 = wvno * wvno
    # half-space (:328-347)
    xka = omega / layer(alpha, L - 1)
    xkb = omega / layer(beta, L - 1)
    ra = torch.sqrt((wvno + xka) * torch.abs(wvno - xka))
    rb = torch.sqrt((wvno + xkb) * torch.abs(wvno - xkb))
    t = layer(beta, L - 1) / omega
    gammk = 2.0 * t * t
    gam = gammk * wvno2
    gamm1 = gam - 1.0
    rho1 = layer(rho, L - 1)
    e = (
        rho1 * rho1 * (gamm1 * gamm1 - gam * gammk * ra * rb),
        -rho1 * ra,
        rho1 * (gamm1 - gammk * ra * rb),
        rho1 * rb,
        wvno2 - ra * rb,
    )
    # layers bottom-up (:349-383)
    for m in range(L - 2, -1, -1):
        xka = omega / layer(alpha, m)
        xkb = omega / layer(beta, m)
        t = layer(beta, m) / omega
        gammk = 2.0 * t * t
        gam = gammk * wvno2
        ra = torch.sqrt((wvno + xka) * torch.abs(wvno - xka))
        rb = torch.sqrt((wvno + xkb) * torch.abs(wvno - xkb))
        dpth = torch.ones_like(wvno) * layer(d, m)
        p = ra * dpth
        q = rb * dpth
        terms = eigenfunction_terms(p, q, ra, rb, wvno, xka, xkb, dpth)
        e = dunkin_matrix_apply(e, wvno2, gam, gammk, layer(rho, m), terms)
    return e[0]


def _as2d(x, dtype):
    x = torch.as_tensor(x)
    if not x.is_floating_point() or x.dtype != dtype:
        x = x.to(dtype)
    return x


def det_points(vc, t, d, alpha, beta, rho, dtype=torch.float64):
    """``dltar_vector(vc, t, d, alpha, beta, rho, ifunc=2, llw=-1)``.

    Batched: ``vc, t [S,N]``, params ``[S,L]`` -> ``[S,N]``; single model: ``[N]``, ``[L]`` -> ``[N]``.
    """
    vc, t, d, alpha, beta, rho = (_as2d(v, dtype) for v in (vc, t, d, alpha, beta, rho))
    single = d.dim() == 1
    if single:
        vc, t = vc.reshape(1, -1), t.reshape(1, -1)
        d, alpha, beta, rho = (v.reshape(1, -1) for v in (d, alpha, beta, rho))
    omega0 = TWOPI / t
    wvno = 1 / vc * omega0
    out = secular_rayleigh(wvno, omega0, d, alpha, beta, rho)
    return out[0] if single else out


def det_grid(vlist, tlist, d, alpha, beta, rho, dtype=torch.float64):
    """``dltar_matrix(vlist, tlist, d, a, b, rho, ifunc=2, llw=-1)``.

    Batched: ``vlist [S,K]``, ``tlist [S,NF]`` -> ``[S,K,NF]``; single: ``[K]``, ``[NF]`` -> ``[K,NF]``.
    """
    vlist, tlist, d, alpha, beta, rho = (_as2d(v, dtype) for v in (vlist, tlist, d, alpha, beta, rho))
    single = d.dim() == 1
    if single:
        vlist, tlist = vlist.reshape(1, -1), tlist.reshape(1, -1)
        d, alpha, beta, rho = (v.reshape(1, -1) for v in (d, alpha, beta, rho))
    S = d.shape[0]
    omega0 = (TWOPI / tlist).reshape(S, 1, -1)
    wvno = 1 / vlist.reshape(S, -1, 1) * omega0
    omega0 = omega0.expand_as(wvno)
    out = secular_rayleigh(wvno, omega0, d, alpha, beta, rho)
    return out[0] if single else out


def brocher_vp_rho(b):
    """Brocher (2005) Vp(Vs), rho(Vp) polynomials as written at _model.py:432-433."""
    a = 0.9409 + 2.0947 * b - 0.8206 * b**2 + 0.2683 * b**3 - 0.0251 * b**4
    rho = 1.6612 * a - 0.4721 * a**2 + 0.0671 * a**3 - 0.0043 * a**4 + 0.000106 * a**5
    return a, rho


def second_difference_matrix(n, dtype=torch.float64):
    """Tikhonov operator built at _model.py:407-409 / :415-417."""
    L0 = 2.0 * torch.eye(n, dtype=dtype)
    idx = torch.arange(n - 1)
    L0[idx, idx + 1] = -1.0
    L0[idx + 1, idx] = -1.0
    L0[0, 0] = 1.0
    L0[n - 1, n - 1] = 1.0
    return L0


def dmf_loss(vlist, tlist, Clist, d, b, a=None, rho=None, *, initial_method="Brocher",
             vp_vs_ratio=1.732, compress=True, normalized=True, compress_method="exp",
             damp_vertical=0.0, damp_horizontal=0.0, dtype=torch.float64):
    """Determinant Misfit Function per model, ``multi_cal_grad.forward`` (_model.py:430-478).

    ``vlist, tlist [S,N]`` observed picks, ``Clist [S,K]`` trial velocities, ``d, b [S,L]``.
    Differentiable w.r.t. ``b`` and ``d``.  Returns ``loss [S]``.
    """
    if initial_method == "Brocher":
        a, rho = brocher_vp_rho(b)
    elif initial_method == "Constant":
        a = b * vp_vs_ratio
    F = det_points(vlist, tlist, d, a, b, rho, dtype=dtype)
    if compress:
        if normalized:
            with torch.no_grad():
                det = det_grid(Clist, tlist, d, a, b, rho, dtype=dtype)
            F = F / (torch.max(det, dim=1).values - torch.min(det, dim=1).values)
            if compress_method == "log":
                F = torch.log(torch.abs(F) + 1)
            elif compress_method == "exp":
                F = (1e-1) ** torch.abs(F) - 1
        else:
            F = (1e-1) ** torch.abs(F) - 1
    F_all = torch.sum(torch.abs(F), dim=1) / F.shape[-1]
    out = F_all
    nL = b.shape[1]
    if damp_vertical > 0:
        Lv = second_difference_matrix(nL, dtype=b.dtype)
        mv = damp_vertical * torch.matmul(b, Lv.T) / nL
        out = out + torch.sum(torch.abs(mv), dim=1)
    if damp_horizontal > 0:
        Lh = second_difference_matrix(b.shape[0], dtype=b.dtype)
        mh = damp_horizontal * torch.matmul(Lh, b) / nL
        out = out + torch.sum(torch.abs(mh), dim=1)
    return out


def sign_change_indices(det):
    """Indices k (along the trial-velocity axis, dim -2) where sign(det[k]) != sign(det[k+1]).

    The reference's notebooks build the determinant image from ``torch.sign(F)``
    (examples/00_0 cell 7); mode/bracket indices are where that sign flips along c.
    Returns a boolean tensor of shape ``[..., K-1, NF]``.
    """
    s = torch.sign(det)
    return s[..., :-1, :] != s[..., 1:, :]
