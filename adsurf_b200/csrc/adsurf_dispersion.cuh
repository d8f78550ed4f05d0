// adsurf_dispersion.cuh — modal dispersion curves by the classical CPS / disba root search.
// Shared by the CUDA kernel k_dispersion (one thread per model) and by the host-side check that tests/ builds from
// this same header (tests/host_math/host_math.cpp); the product library never runs it on the host.
//
// Same control flow as the reference's scalar solver (/root/reference/ADsurf/_surf96.py: gtsolh :576-594, getc
// :596-691, getsol :543-574, nevill :452-541, dltar4 :217-297 with its per-layer normalisation), so that the same
// brackets and the same iterates come out: the search is strictly sequential within a model (period k starts from the
// root of period k-1, mode n from mode n-1), so the parallel axis is the ensemble — 4096 curves cost what one costs.
#pragma once
#include <math.h>

#include "adsurf_math.cuh"

#if defined(__CUDACC__)
#define ADSURF_HDN __host__ __device__ __noinline__
#else
#define ADSURF_HDN inline
#endif

namespace adsurf {

ADSURF_HD void disp_sincos(double x, double* sn, double* cs) {
#if defined(__CUDA_ARCH__)
  sincos(x, sn, cs);
#else
  *sn = sin(x);
  *cs = cos(x);
#endif
}

// Same control flow as the reference's scalar solver (_surf96.py: gtsolh :576-594, getc :596-691, getsol
// :543-574, nevill :452-541, dltar4 :217-297 with its per-layer normalisation) so that the same brackets and the same
// iterates come out: the search is strictly sequential within a model (period k starts from the root of period
// k-1, mode n from mode n-1), so the parallel axis is the ensemble — 4096 curves cost what one costs.
struct DispModel {
  const double *d, *a, *b, *rho;
  int L;
};

// normalised Rayleigh secular function at (period t, phase velocity c): dltar4 with llw = -1
ADSURF_HDN double disp_delta(const DispModel& M, double t, double c) {
  double omega = kTwoPi / t;
  const double wvno = omega / c;
  omega = fmax(omega, 1.0e-4);
  const double k2 = wvno * wvno, t2 = -2.0 * k2;
  const int L = M.L;
  double xka = omega / M.a[L - 1], xkb = omega / M.b[L - 1];
  double ra = sqrt((wvno + xka) * fabs(wvno - xka));
  double rb = sqrt((wvno + xkb) * fabs(wvno - xkb));
  double tt = M.b[L - 1] / omega;
  double gk = 2.0 * tt * tt, gam = gk * k2, g1 = gam - 1.0, r1 = M.rho[L - 1];
  double e0 = r1 * r1 * (g1 * g1 - gam * gk * ra * rb);
  double e1 = -r1 * ra;
  double e2 = r1 * (g1 - gk * ra * rb);
  double e3 = r1 * rb;
  double e4 = k2 - ra * rb;
  for (int m = L - 2; m >= 0; --m) {
    xka = omega / M.a[m];
    xkb = omega / M.b[m];
    tt = M.b[m] / omega;
    gk = 2.0 * tt * tt;
    gam = gk * k2;
    g1 = gam - 1.0;
    ra = sqrt((wvno + xka) * fabs(wvno - xka));
    rb = sqrt((wvno + xkb) * fabs(wvno - xkb));
    const double dm = M.d[m];
    r1 = M.rho[m];
    // eigenfunction terms (var, :96-157)
    double cp, w, x, pex, cq, y, z, sex;
    {
      const double p = ra * dm;
      if (wvno < xka) { double sn, cs; disp_sincos(p, &sn, &cs); cp = cs; w = sn / ra; x = -ra * sn; pex = 0.0; }
      else if (wvno == xka) { cp = 1.0; w = dm; x = 0.0; pex = 0.0; }
      else { const double fac = p < 16.0 ? exp(-2.0 * p) : 0.0, sn = (1.0 - fac) * 0.5; cp = (1.0 + fac) * 0.5; w = sn / ra; x = ra * sn; pex = p; }
      const double q = rb * dm;
      if (wvno < xkb) { double sn, cs; disp_sincos(q, &sn, &cs); cq = cs; y = sn / rb; z = -rb * sn; sex = 0.0; }
      else if (wvno == xkb) { cq = 1.0; y = dm; z = 0.0; sex = 0.0; }
      else { const double fac = q < 16.0 ? exp(-2.0 * q) : 0.0, sn = (1.0 - fac) * 0.5; cq = (1.0 + fac) * 0.5; y = sn / rb; z = rb * sn; sex = q; }
    }
    const double exa = pex + sex;
    const double a0 = exa < 60.0 ? exp(-exa) : 0.0;
    // E <- E . CA in the rank-structured form (adsurf_math.cuh), then the reference's max-normalisation (:278)
    const double ir = 1.0 / r1, ar = gam * gk * r1, br = g1 * g1 * r1, kr = k2 * ir;
    const double sa = e0 * ir + e2 * (t2 * gk) + e4 * ar;
    const double sb = e0 * kr + e2 * (t2 * g1) + e4 * br;
    const double h1 = sb * y + e1 * cq, h2 = sa * cq + e3 * y;
    const double h3 = sa * z + e3 * cq, h4 = sb * cq + e1 * z;
    const double ca = cp * h4 - x * h3 - sb * a0;
    const double cb = cp * h2 - w * h1 - sa * a0;
    const double n0 = a0 * e0 + ar * ca + br * cb;
    const double n1 = cp * h1 - x * h2;
    const double n2 = a0 * e2 + gk * ca + g1 * cb;
    const double n3 = cp * h3 - w * h4;
    const double n4 = a0 * e4 + ir * ca + kr * cb;
    double big = fmax(fmax(fmax(fabs(n0), fabs(n1)), fmax(fabs(n2), fabs(n3))), fabs(n4));
    if (big < 1.0e-40) big = 1.0;
    e0 = n0 / big; e1 = n1 / big; e2 = n2 / big; e3 = n3 / big; e4 = n4 / big;
  }
  return e0;
}

ADSURF_HD double np_sign(double v) { return v > 0.0 ? 1.0 : (v < 0.0 ? -1.0 : (v == 0.0 ? 0.0 : v)); }

// refine a bracketed root: Neville rational interpolation with bisection safeguards (:452-541)
ADSURF_HDN double disp_nevill(const DispModel& M, double t, double c1, double c2, double del1, double del2) {
  double xs[20], ys[20];
  for (int i = 0; i < 20; ++i) xs[i] = ys[i] = 0.0;
  double c3 = 0.5 * (c1 + c2);
  double del3 = disp_delta(M, t, c3);
  int nev = 1, m = 1;
  for (int it = 0; it < 98; ++it) {
    if (c3 < fmin(c1, c2) || c3 > fmax(c1, c2)) {  // estimate left the bracket: halve instead
      nev = 0;
      c3 = 0.5 * (c1 + c2);
      del3 = disp_delta(M, t, c3);
    }
    const double s13 = del1 - del3, s32 = del3 - del2;
    if (np_sign(del3) * np_sign(del1) < 0.0) { c2 = c3; del2 = del3; }
    else { c1 = c3; del1 = del3; }
    if (fabs(c1 - c2) <= 1.0e-6 * c1) break;
    if (np_sign(s13) != np_sign(s32)) nev = 0;
    const double ss1 = fabs(del1), ss2 = fabs(del2);
    if (0.01 * ss1 > ss2 || 0.01 * ss2 > ss1 || nev == 0) {  // values too unequal for a polynomial fit
      c3 = 0.5 * (c1 + c2);
      del3 = disp_delta(M, t, c3);
      nev = 1;
      m = 1;
      continue;
    }
    if (nev == 2) { xs[m - 1] = c3; ys[m - 1] = del3; }
    else { xs[0] = c1; ys[0] = del1; xs[1] = c2; ys[1] = del2; m = 1; }
    bool ok = true;
    for (int kk = 0; kk < m; ++kk) {
      const int j = m - kk;
      const double denom = ys[m] - ys[j];
      if (fabs(denom) < 1.0e-10 * fabs(ys[m])) { ok = false; break; }
      xs[j - 1] = (-ys[j - 1] * xs[j] + ys[m] * xs[j - 1]) / denom;
    }
    if (ok) {
      c3 = xs[0];
      del3 = disp_delta(M, t, c3);
      nev = 2;
      m = m + 1 < 10 ? m + 1 : 10;
    } else {
      c3 = 0.5 * (c1 + c2);
      del3 = disp_delta(M, t, c3);
      nev = 1;
      m = 1;
    }
  }
  return c3;
}

// walk the dc ladder until the secular function changes sign, then refine (:543-574); returns iret
ADSURF_HD bool disp_getsol(const DispModel& M, double t1, double& c1, double clow, double dc, double cm, double betmx,
                            bool ifirst, double& del1st) {
  double del1 = disp_delta(M, t1, c1);
  if (ifirst) del1st = del1;
  double idir = (!ifirst && np_sign(del1st) * np_sign(del1) < 0.0) ? -1.0 : 1.0;
  for (;;) {
    const double c2 = c1 + idir * dc;
    if (c2 <= clow) {
      idir = 1.0;
      c1 = clow;
      continue;
    }
    const double del2 = disp_delta(M, t1, c2);
    if (np_sign(del1) != np_sign(del2)) {
      c1 = disp_nevill(M, t1, c1, c2, del1, del2);
      return c1 > betmx;
    }
    c1 = c2;
    del1 = del2;
    if (c1 < cm || c1 >= betmx + dc) return true;
  }
}

// half-space Rayleigh velocity by 5 Newton steps (:576-594)
ADSURF_HD double disp_gtsolh(double a, double b) {
  double c = 0.95 * b;
  const double gamma = b / a;
  for (int i = 0; i < 5; ++i) {
    const double kappa = c / b, k2 = kappa * kappa, gk2 = (gamma * kappa) * (gamma * kappa);
    const double fac1 = sqrt(1.0 - gk2), fac2 = sqrt(1.0 - k2);
    const double fr = (2.0 - k2) * (2.0 - k2) - 4.0 * fac1 * fac2;
    double frp = -4.0 * (2.0 - k2) * kappa;
    frp = frp + 4.0 * fac2 * gamma * gamma * kappa / fac1;
    frp = frp + 4.0 * fac1 * kappa / fac2;
    frp /= b;
    c = c - fr / frp;
  }
  return c;
}

// getc (:596-691) for modes 0 .. nmodes-1 of one model; cout [nmodes, NF]; c [NF] scratch (the previous mode's roots).
// Returns 0, or 1 when the fundamental mode cannot be bracketed (the reference raises DispersionError).
ADSURF_HD int disp_getc(const DispModel& M, const double* ts, int NF, int nmodes, double dc, double* cout, double* c) {
  const int L = M.L;
  double betmx = -1.0e20, betmn = 1.0e20;
  int jmn = 0;
  for (int i = 0; i < L; ++i) {
    const double v = M.b[i];
    if (v > 0.01 && v < betmn) { betmn = v; jmn = i; }
    betmx = fmax(betmx, v);
  }
  const double cc = 0.9 * disp_gtsolh(M.a[jmn], M.b[jmn]);  // start a little below the slowest layer's Rayleigh velocity
  double c1 = cc;
  const double cm = cc, one = 1.0e-2, onea = 1.5;
  for (int i = 0; i < nmodes * NF; ++i) cout[i] = 0.0;
  for (int k = 0; k < NF; ++k) c[k] = 0.0;
  for (int iq = 0; iq < nmodes; ++iq) {
    double* cg = cout + (size_t)iq * NF;
    double del1st = 0.0;
    for (int k = 0; k < NF; ++k) {
      const bool ifirst = k == 0;
      double clow;
      if (ifirst && iq == 0) { clow = cc; c1 = cc; }
      else if (ifirst) { clow = c1; c1 = c[0] + one * dc; }
      else if (iq > 0) { clow = c[k] + one * dc; c1 = fmax(c[k - 1], clow); }
      else { clow = cm; c1 = c[k - 1] - onea * dc; }
      const bool iret = disp_getsol(M, ts[k], c1, clow, dc, cm, betmx, ifirst, del1st);
      if (iret) {
        if (iq == 0) return 1;
        for (int kk = k; kk < NF; ++kk) cg[kk] = 0.0;
        c1 = 0.0;
        break;
      }
      c[k] = c1;
      cg[k] = c1;
      c1 = 0.0;
    }
  }
  return 0;
}

}  // namespace adsurf
