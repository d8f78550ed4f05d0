"""Modal dispersion curves by root search — API mirror of the reference's ADsurf/_surf96.py.

    surf96(t, d, a, b, rho, mode=0, itype=0, ifunc=2, dc=0.005, dt=0.025) -> c[len(t)]

This is the classical CPS / disba solver (bracket a sign change of the secular function on a `dc`
ladder, refine with the hybrid Neville / bisection scheme to 1e-6 relative): strictly sequential
and data dependent (period k starts from the root of period k-1), so it stays on the host, as in
the reference (SURVEY.md §2 "Scalar forward + root search", §8f N1).  It is the root oracle of the
package: `cal_dispersion_curve` (the `_ADsurf` API) calls it.  The control flow follows the
reference so that the same brackets and the same iterates are produced:
    starting value   gtsolh  (_surf96.py:576-594)      mode bookkeeping  getc   (:596-691)
    bracketing       getsol  (:543-574)                refinement        nevill (:452-541)
The scalar secular function (`rayleigh_secular`) uses per-layer max-normalisation like the
reference's scalar path (:278, `normc`), written with the rank-structured layer product of
adsurf_math.cuh instead of a 5x5 matrix.  A batched, all-modes-at-once GPU extraction built on
the dense-grid kernel lives in `adsurf_b200.dispersion` (different algorithm, same roots).
"""
from __future__ import annotations

import numpy as np

TWOPI = 2.0 * np.pi


class DispersionError(Exception):
    pass


def _eig(k, kx, r, dm):
    """(cos-like, sin/r-like, -+r*sin-like, exponent) of one wave type (reference var(), :96-157)"""
    p = r * dm
    if k < kx:
        s = np.sin(p)
        return np.cos(p), s / r, -r * s, 0.0
    if k == kx:
        return 1.0, dm, 0.0, 0.0
    fac = np.exp(-2.0 * p) if p < 16.0 else 0.0
    s = (1.0 - fac) * 0.5
    return (1.0 + fac) * 0.5, s / r, r * s, p


def rayleigh_secular(wvno, omega, d, a, b, rho, llw=-1):
    """Normalised Rayleigh secular function for one (k, omega) (reference dltar4, :217-297).  llw = 0: layer 0 is
    water (Vs = 0): the solid stack ends above it and the water column closes the equation (:284-295)."""
    omega = max(omega, 1.0e-4)
    k2 = wvno * wvno
    t2 = -2.0 * k2
    L = len(d)
    xka, xkb = omega / a[L - 1], omega / b[L - 1]
    ra = np.sqrt((wvno + xka) * abs(wvno - xka))
    rb = np.sqrt((wvno + xkb) * abs(wvno - xkb))
    tt = b[L - 1] / omega
    gk = 2.0 * tt * tt
    gam = gk * k2
    g1 = gam - 1.0
    r1 = rho[L - 1]
    e0 = r1 * r1 * (g1 * g1 - gam * gk * ra * rb)
    e1 = -r1 * ra
    e2 = r1 * (g1 - gk * ra * rb)
    e3 = r1 * rb
    e4 = k2 - ra * rb
    for m in range(L - 2, llw, -1):
        xka, xkb = omega / a[m], omega / b[m]
        tt = b[m] / omega
        gk = 2.0 * tt * tt
        gam = gk * k2
        g1 = gam - 1.0
        ra = np.sqrt((wvno + xka) * abs(wvno - xka))
        rb = np.sqrt((wvno + xkb) * abs(wvno - xkb))
        dm, r1 = d[m], rho[m]
        cp, w, x, pex = _eig(wvno, xka, ra, dm)
        cq, y, z, sex = _eig(wvno, xkb, rb, dm)
        exa = pex + sex
        a0 = np.exp(-exa) if exa < 60.0 else 0.0
        ir = 1.0 / r1
        ar, br, kr = gam * gk * r1, g1 * g1 * r1, k2 * ir
        sa = e0 * ir + e2 * (t2 * gk) + e4 * ar
        sb = e0 * kr + e2 * (t2 * g1) + e4 * br
        h1, h2 = sb * y + e1 * cq, sa * cq + e3 * y
        h3, h4 = sa * z + e3 * cq, sb * cq + e1 * z
        ca = cp * h4 - x * h3 - sb * a0
        cb = cp * h2 - w * h1 - sa * a0
        n0 = a0 * e0 + ar * ca + br * cb
        n1 = cp * h1 - x * h2
        n2 = a0 * e2 + gk * ca + g1 * cb
        n3 = cp * h3 - w * h4
        n4 = a0 * e4 + ir * ca + kr * cb
        big = max(abs(n0), abs(n1), abs(n2), abs(n3), abs(n4))
        if big < 1.0e-40:
            big = 1.0
        e0, e1, e2, e3, e4 = n0 / big, n1 / big, n2 / big, n3 / big, n4 / big
    if llw == 0:
        xka = omega / a[0]
        ra = np.sqrt((wvno + xka) * abs(wvno - xka))
        cp, w, _, _ = _eig(wvno, xka, ra, d[0])
        return cp * e0 - rho[0] * w * e1
    return e0


def love_secular(wvno, omega, d, b, rho, llw=-1):
    """Normalised Love secular function for one (k, omega) (reference dltar1, :168-215, llw=-1): the SH
    stress-displacement pair carried up from the half-space, max-normalised after every layer."""
    L = len(d)
    xkb = omega / b[L - 1]
    rb = np.sqrt((wvno + xkb) * abs(wvno - xkb))
    e1, e2 = rho[L - 1] * rb, 1.0 / (b[L - 1] * b[L - 1])
    for m in range(L - 2, llw, -1):  # a water layer carries no SH motion
        xmu = rho[m] * b[m] * b[m]
        xkb = omega / b[m]
        rb = np.sqrt((wvno + xkb) * abs(wvno - xkb))
        cq, y, z, _ = _eig(wvno, xkb, rb, d[m])
        n1 = e1 * cq + e2 * xmu * z
        n2 = e1 * y / xmu + e2 * cq
        big = max(abs(n1), abs(n2))
        if big < 1.0e-40:
            big = 1.0
        e1, e2 = n1 / big, n2 / big
    return e1


def _delta(t, c, d, a, b, rho, ifunc=2):
    omega = TWOPI / t
    llw = 0 if b[0] <= 0.0 else -1
    if ifunc == 1:
        return love_secular(omega / c, omega, d, b, rho, llw)
    return rayleigh_secular(omega / c, omega, d, a, b, rho, llw)


def nevill(t, c1, c2, del1, del2, d, a, b, rho, ifunc=2):
    """Refine a bracketed root: Neville rational interpolation with bisection safeguards (:452-541)."""
    xs, ys = np.zeros(20), np.zeros(20)
    c3 = 0.5 * (c1 + c2)
    del3 = _delta(t, c3, d, a, b, rho, ifunc)
    nev, m = 1, 1
    for _ in range(98):
        if c3 < min(c1, c2) or c3 > max(c1, c2):  # estimate left the bracket: halve instead
            nev = 0
            c3 = 0.5 * (c1 + c2)
            del3 = _delta(t, c3, d, a, b, rho, ifunc)
        s13, s32 = del1 - del3, del3 - del2
        if np.sign(del3) * np.sign(del1) < 0.0:
            c2, del2 = c3, del3
        else:
            c1, del1 = c3, del3
        if abs(c1 - c2) <= 1.0e-6 * c1:
            break
        if np.sign(s13) != np.sign(s32):
            nev = 0
        ss1, ss2 = abs(del1), abs(del2)
        if 0.01 * ss1 > ss2 or 0.01 * ss2 > ss1 or nev == 0:  # values too unequal for a polynomial fit
            c3 = 0.5 * (c1 + c2)
            del3 = _delta(t, c3, d, a, b, rho, ifunc)
            nev, m = 1, 1
            continue
        if nev == 2:
            xs[m - 1], ys[m - 1] = c3, del3
        else:
            xs[0], ys[0], xs[1], ys[1] = c1, del1, c2, del2
            m = 1
        ok = True
        for kk in range(m):
            j = m - kk
            denom = ys[m] - ys[j]
            if abs(denom) < 1.0e-10 * abs(ys[m]):
                ok = False
                break
            xs[j - 1] = (-ys[j - 1] * xs[j] + ys[m] * xs[j - 1]) / denom
        if ok:
            c3 = xs[0]
            del3 = _delta(t, c3, d, a, b, rho, ifunc)
            nev = 2
            m = min(m + 1, 10)
        else:
            c3 = 0.5 * (c1 + c2)
            del3 = _delta(t, c3, d, a, b, rho, ifunc)
            nev, m = 1, 1
    return c3


def getsol(t1, c1, clow, dc, cm, betmx, ifirst, del1st, d, a, b, rho, ifunc=2):
    """Walk the dc ladder until the secular function changes sign, then refine (:543-574)."""
    del1 = _delta(t1, c1, d, a, b, rho, ifunc)
    if ifirst:
        del1st = del1
    idir = -1.0 if (not ifirst and np.sign(del1st) * np.sign(del1) < 0.0) else 1.0
    while True:
        c2 = c1 + idir * dc
        if c2 <= clow:
            idir = 1.0
            c1 = clow
            continue
        del2 = _delta(t1, c2, d, a, b, rho, ifunc)
        if np.sign(del1) != np.sign(del2):
            c1 = nevill(t1, c1, c2, del1, del2, d, a, b, rho, ifunc)
            return c1, del1st, c1 > betmx
        c1, del1 = c2, del2
        if c1 < cm or c1 >= betmx + dc:
            return c1, del1st, True


def gtsolh(a, b):
    """Half-space Rayleigh velocity by 5 Newton steps on the Rayleigh equation (:576-594)."""
    c = 0.95 * b
    gamma = b / a
    for _ in range(5):
        kappa = c / b
        k2 = kappa**2
        gk2 = (gamma * kappa) ** 2
        fac1, fac2 = np.sqrt(1.0 - gk2), np.sqrt(1.0 - k2)
        fr = (2.0 - k2) ** 2 - 4.0 * fac1 * fac2
        frp = -4.0 * (2.0 - k2) * kappa
        frp = frp + 4.0 * fac2 * gamma * gamma * kappa / fac1
        frp = frp + 4.0 * fac1 * kappa / fac2
        frp /= b
        c = c - fr / frp
    return c


def getc(t, d, a, b, rho, mode, ifunc, dc):
    """Phase velocities of mode `mode` at the periods t; 0 where the mode does not exist (:596-691)."""
    if ifunc not in (1, 2):  # 3 = fast delta matrix: the reference's own path fails on numpy >= 2 (np.complex_)
        raise NotImplementedError("adsurf_b200: Love (ifunc=1) and Rayleigh / Dunkin (ifunc=2) only")
    if b[0] <= 0.0 and len(d) < 3:
        raise NotImplementedError("adsurf_b200: a water layer needs a solid layer above the half-space")
    kmax = len(t)
    c, cg = np.zeros(kmax), np.zeros(kmax)
    betmx, betmn, jmn, jsol = -1.0e20, 1.0e20, 0, False
    for i in range(len(d)):
        if 0.01 < b[i] < betmn:
            betmn, jmn, jsol = b[i], i, False
        elif b[i] < 0.01 and a[i] < betmn:  # a fluid layer slower than every solid so far (:617-620)
            betmn, jmn, jsol = a[i], i, True
        betmx = max(betmx, b[i])
    # start a little below the slowest layer's Rayleigh velocity (the fluid's Vp if that is the slowest)
    cc = 0.9 * (betmn if jsol else gtsolh(a[jmn], b[jmn]))
    c1, cm = cc, cc
    one, onea = 1.0e-2, 1.5
    for iq in range(mode + 1):
        del1st = 0.0
        for k in range(kmax):
            ifirst = k == 0
            if ifirst and iq == 0:
                clow, c1 = cc, cc
            elif ifirst:
                clow, c1 = c1, c[0] + one * dc
            elif iq > 0:
                clow = c[k] + one * dc
                c1 = max(c[k - 1], clow)
            else:
                clow, c1 = cm, c[k - 1] - onea * dc
            c1, del1st, iret = getsol(t[k], c1, clow, dc, cm, betmx, ifirst, del1st, d, a, b, rho, ifunc)
            if iret:
                if iq == 0:
                    raise DispersionError("failed to find root for fundamental mode")
                cg[k:] = 0.0
                if iq == mode:
                    return cg
                c1 = 0.0
                break
            c[k] = c1
            cg[k] = c1
            c1 = 0.0
    return cg


def surf96(t, d, a, b, rho, mode=0, itype=0, ifunc=2, dc=0.005, dt=0.025):
    """Phase (itype=0) or group (itype=1) velocity dispersion curve of one mode (:693-760)."""
    t = np.asarray(t, dtype=np.float64)
    d, a, b, rho = (np.asarray(x, dtype=np.float64) for x in (d, a, b, rho))
    if itype != 1:
        return getc(t.copy(), d, a, b, rho, mode, ifunc, dc)
    t1, t2 = t / (1.0 + dt), t / (1.0 - dt)
    c1 = getc(t1, d, a, b, rho, mode, ifunc, dc)
    c2 = getc(t2, d, a, b, rho, mode, ifunc, dc)
    out = np.zeros_like(c1)
    ok = c2 > 0.0
    f1, f2 = 1.0 / t1[ok], 1.0 / t2[ok]
    out[ok] = (f1 - f2) / (f1 / c1[ok] - f2 / c2[ok])
    return out
