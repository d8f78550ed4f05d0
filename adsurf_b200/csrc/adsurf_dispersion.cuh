// adsurf_dispersion.cuh — modal dispersion curves by the classical CPS / disba root search.
// Shared by the CUDA kernel k_dispersion (one warp per model) and by the host-side check that tests/ builds from
// this same header (tests/host_math/host_math.cpp); the product library never runs it on the host.
//
// Same control flow as the reference's scalar solver (/root/reference/ADsurf/_surf96.py: gtsolh :576-594, getc
// :596-691, getsol :543-574, nevill :452-541, dltar4 :217-297 with its per-layer normalisation), so that the same
// brackets and the same iterates come out.  The search is strictly sequential within a model (period k starts from the
// root of period k-1, mode n from mode n-1) and so is the layer product of one evaluation; what is parallel inside an
// evaluation is the per-layer preparation (square roots, divisions, sincos, exp: most of its cost), which the lanes of
// the model's warp share out; the ensemble spreads over the warps.
#pragma once
#include <math.h>

#include "adsurf_math.cuh"

#if defined(__CUDACC__)
#define ADSURF_HDN __host__ __device__ __noinline__
#else
#define ADSURF_HDN inline
#endif

namespace adsurf {

struct DispModel {
  const double *d, *a, *b, *rho;
  int L;
  int ifunc;  // 1: Love waves (dltar1), 2: Rayleigh waves, Dunkin's matrix (dltar4) — the reference's ifunc
  // index of the topmost solid layer: 1 when layer 0 is water (Vs <= 0; the reference's llw = 0, :608), else 0
  ADSURF_HD int top() const { return b[0] <= 0.0 ? 1 : 0; }
};

// The normalised Rayleigh secular function (dltar4 with llw = -1, :217-297) in three pieces, so that the scalar
// evaluator (host check) and the warp-cooperative one (k_dispersion) run the same arithmetic:
//   per (period, layer):      k_alpha, k_beta, gammk                               (disp_period_consts)
//   per (evaluation, layer):  the layer's eigenfunction terms and matrix constants (disp_layer_terms: every sqrt,
//                             division, sincos and exp of the evaluation; no dependence between layers)
//   bottom-up:                E <- E . CA_m, rescaled                              (disp_layer_apply: 33 FMAs and an
//                             exact power-of-two rescale; the only sequential part), E[0] / max|E| at the top
constexpr int kDispTerms = 16;  // doubles per layer handed from disp_layer_terms to disp_layer_apply
constexpr int kDispPer = 6;     // doubles per layer kept per period: k_alpha, k_beta, gammk | warp evaluator: d, rho, 1/rho
enum { DT_CP = 0, DT_W, DT_X, DT_CQ, DT_Y, DT_Z, DT_A0, DT_GK, DT_G1, DT_AR, DT_BR, DT_KR, DT_IR, DT_TGK, DT_TG1 };

ADSURF_HD void disp_period_consts(double* __restrict__ per, double a, double b, double omega) {
  const double tt = b / omega;
  per[0] = omega / a;
  per[1] = omega / b;
  per[2] = 2.0 * tt * tt;
}

// sqrt(u) and 1/sqrt(u) from the rsqrt sequence of the pointwise kernels (u >= 0; 0 -> (0, inf), NaN -> NaN)
ADSURF_HD void disp_root(double u, double& r, double& ir) {
  ir = rsqrt_hd(u);
  r = u > 0.0 ? u * ir : u;
}

// eigenfunction terms of one wave type (var, :96-157): cos-like, sin-like / r, r * sin-like, exponent.
// Elementary functions: the kernels' inline exp / sincos (adsurf_math.cuh, <= 2 ulp), one code path per regime.
ADSURF_HD void disp_eig(double wvno, double xk, double r, double ir, double dm, double& cs, double& w, double& x,
                        double& ex) {
  const double p = r * dm;
  if (wvno < xk) {
    double sn;
    sincos_fast(p, &sn, &cs);
    w = sn * ir;
    x = -r * sn;
    ex = 0.0;
  } else if (wvno == xk) {
    cs = 1.0;
    w = dm;
    x = 0.0;
    ex = 0.0;
  } else {
    const double fac = p < 16.0 ? exp_neg(2.0 * p) : 0.0, sn = (1.0 - fac) * 0.5;
    cs = (1.0 + fac) * 0.5;
    w = sn * ir;
    x = r * sn;
    ex = p;
  }
}

// ir = 1 / rho of the layer
ADSURF_HD void disp_layer_terms(double* __restrict__ q, const double* __restrict__ per, double wvno, double k2, double t2,
                                double dm, double r1, double ir) {
  const double xka = per[0], xkb = per[1], gk = per[2];
  const double gam = gk * k2, g1 = gam - 1.0;
  double ra, ira, rb, irb;
  disp_root((wvno + xka) * fabs(wvno - xka), ra, ira);
  disp_root((wvno + xkb) * fabs(wvno - xkb), rb, irb);
  double pex, sex;
  disp_eig(wvno, xka, ra, ira, dm, q[DT_CP], q[DT_W], q[DT_X], pex);
  disp_eig(wvno, xkb, rb, irb, dm, q[DT_CQ], q[DT_Y], q[DT_Z], sex);
  const double exa = pex + sex;
  q[DT_A0] = exa < 60.0 ? exp_neg(exa) : 0.0;
  q[DT_GK] = gk;
  q[DT_G1] = g1;
  q[DT_AR] = gam * gk * r1;
  q[DT_BR] = g1 * g1 * r1;
  q[DT_KR] = k2 * ir;
  q[DT_IR] = ir;
  q[DT_TGK] = t2 * gk;
  q[DT_TG1] = t2 * g1;
}

// half-space vector (:231-244)
ADSURF_HD void disp_half_terms(double* __restrict__ q, const double* __restrict__ per, double wvno, double k2, double r1) {
  const double xka = per[0], xkb = per[1], gk = per[2];
  double ra, rb, unused;
  disp_root((wvno + xka) * fabs(wvno - xka), ra, unused);
  disp_root((wvno + xkb) * fabs(wvno - xkb), rb, unused);
  const double gam = gk * k2, g1 = gam - 1.0;
  q[0] = r1 * r1 * (g1 * g1 - gam * gk * ra * rb);
  q[1] = -r1 * ra;
  q[2] = r1 * (g1 - gk * ra * rb);
  q[3] = r1 * rb;
  q[4] = k2 - ra * rb;
}

// The reference divides E by max|E| after every layer (:278).  The layer step is linear, so in exact arithmetic the
// scales telescope: the normalised vector after layer m is the unnormalised one divided by ITS largest magnitude,
// whatever happened before.  Hence: an exact power-of-two rescale per layer (keeps the range, rounds nothing; integer
// work on the exponent fields instead of a max / reciprocal chain on the sequential path), and the reference's
// normalisation once, after layer 0 (disp_finish).  Values differ from the reference's by its own rounding noise.
ADSURF_HD double disp_pow2_scale(double n0, double n1, double n2, double n3, double n4) {
  auto hi = [](double v) { return (unsigned)((unsigned long long)dbl_bits(v) >> 32) & 0x7fffffffu; };
  unsigned h = hi(n0);
  h = h > hi(n1) ? h : hi(n1);
  h = h > hi(n2) ? h : hi(n2);
  h = h > hi(n3) ? h : hi(n3);
  h = h > hi(n4) ? h : hi(n4);
  const unsigned ex = h >> 20;                   // biased exponent of the largest magnitude
  if (ex == 0u || ex >= 2046u) return 1.0;       // zero / subnormal / Inf / NaN: leave alone
  return bits_dbl((long long)(2046u - ex) << 52);  // 2^(1023 - ex): the largest magnitude lands in [1, 2)
}

// one layer's terms as eight 16-byte pairs (the order of the DT_* enum)
struct DispTermRegs {
  Dbl2 v[kDispTerms / 2];
};
ADSURF_HD DispTermRegs disp_load_terms(const double* __restrict__ q) {  // q 16-byte aligned
  DispTermRegs r;
#pragma unroll
  for (int i = 0; i < kDispTerms / 2; ++i) r.v[i] = ld2(q + 2 * i);
  return r;
}

// E <- E . CA in the rank-structured form (adsurf_math.cuh), rescaled by a power of two
ADSURF_HD void disp_layer_apply(double& e0, double& e1, double& e2, double& e3, double& e4, const DispTermRegs& r) {
  const double cp = r.v[0].x, w = r.v[0].y, x = r.v[1].x, cq = r.v[1].y, y = r.v[2].x, z = r.v[2].y, a0 = r.v[3].x;
  const double gk = r.v[3].y, g1 = r.v[4].x, ar = r.v[4].y, br = r.v[5].x, kr = r.v[5].y, ir = r.v[6].x;
  const double tgk = r.v[6].y, tg1 = r.v[7].x;
  const double sa = e0 * ir + e2 * tgk + e4 * ar;
  const double sb = e0 * kr + e2 * tg1 + e4 * br;
  const double h1 = sb * y + e1 * cq, h2 = sa * cq + e3 * y;
  const double h3 = sa * z + e3 * cq, h4 = sb * cq + e1 * z;
  const double ca = cp * h4 - x * h3 - sb * a0;
  const double cb = cp * h2 - w * h1 - sa * a0;
  const double n0 = a0 * e0 + ar * ca + br * cb;
  const double n1 = cp * h1 - x * h2;
  const double n2 = a0 * e2 + gk * ca + g1 * cb;
  const double n3 = cp * h3 - w * h4;
  const double n4 = a0 * e4 + ir * ca + kr * cb;
  const double sc = disp_pow2_scale(n0, n1, n2, n3, n4);
  e0 = n0 * sc; e1 = n1 * sc; e2 = n2 * sc; e3 = n3 * sc; e4 = n4 * sc;
}

// the reference's normalised value after layer 0: E[0] / max|E| (an all-zero vector stays zero, :279)
ADSURF_HD double disp_finish(double e0, double e1, double e2, double e3, double e4) {
  const double big = fmax(fmax(fmax(fabs(e0), fabs(e1)), fmax(fabs(e2), fabs(e3))), fabs(e4));
  return big > 0.0 ? e0 / big : e0;
}
// a water layer on top (:284-295): the normalised vector of the topmost solid layer closed by the water column,
// cos-like(p) E[0] - rho_w sin-like(p)/r E[1] with the P terms of the water layer (q: DT_CP, DT_W of layer 0)
ADSURF_HD double disp_finish_water(double e0, double e1, double e2, double e3, double e4, const double* __restrict__ q,
                                   double rho_w) {
  const double big = fmax(fmax(fmax(fabs(e0), fabs(e1)), fmax(fabs(e2), fabs(e3))), fabs(e4));
  if (big > 0.0) {
    e0 = e0 / big;
    e1 = e1 / big;
  }
  return q[DT_CP] * e0 - rho_w * q[DT_W] * e1;
}
// P terms of the water layer
ADSURF_HD void disp_water_terms(double* __restrict__ q, double xka, double wvno, double dm) {
  double ra, ira, x, ex;
  disp_root((wvno + xka) * fabs(wvno - xka), ra, ira);
  disp_eig(wvno, xka, ra, ira, dm, q[DT_CP], q[DT_W], x, ex);
}

// Love waves (dltar1, :168-215): the SH stress-displacement pair (e1, e2) carried up from the half-space,
//   e1' = e1 cos + e2 mu z,  e2' = e1 y / mu + e2 cos,   mu = rho beta^2,
// with the same eigenfunction terms (disp_eig on k_beta), the same max-normalisation per layer in the reference and
// the same treatment here (power-of-two rescale, one normalisation at the top).
enum { LT_CQ = 0, LT_Y, LT_Z, LT_MU, LT_IMU };
ADSURF_HD void disp_love_terms(double* __restrict__ q, double xkb, double wvno, double dm, double xmu, double ixmu) {
  double rb, irb, ex;
  disp_root((wvno + xkb) * fabs(wvno - xkb), rb, irb);
  disp_eig(wvno, xkb, rb, irb, dm, q[LT_CQ], q[LT_Y], q[LT_Z], ex);
  q[LT_MU] = xmu;
  q[LT_IMU] = ixmu;
}
ADSURF_HD void disp_love_half(double* __restrict__ q, double xkb, double wvno, double rho1, double beta1) {
  double rb, irb;
  disp_root((wvno + xkb) * fabs(wvno - xkb), rb, irb);
  q[0] = rho1 * rb;
  q[1] = 1.0 / (beta1 * beta1);
}
ADSURF_HD void disp_love_apply(double& e1, double& e2, const double* __restrict__ q) {
  const double cq = q[LT_CQ], y = q[LT_Y], z = q[LT_Z];
  const double n1 = e1 * cq + (e2 * q[LT_MU]) * z;
  const double n2 = (e1 * y) * q[LT_IMU] + e2 * cq;
  const double sc = disp_pow2_scale(n1, n2, 0.0, 0.0, 0.0);
  e1 = n1 * sc;
  e2 = n2 * sc;
}
ADSURF_HD double disp_love_finish(double e1, double e2) {
  const double big = fmax(fabs(e1), fabs(e2));
  return big > 0.0 ? e1 / big : e1;
}

// One thread evaluates everything (host check; any number of layers).
struct DispScalarEval {
  DispModel M;
  double t;
  ADSURF_HD void set_period(double t_) { t = t_; }
  ADSURF_HDN double operator()(double c) const {
    double omega = kTwoPi / t;
    const double wvno = omega / c;
    omega = fmax(omega, 1.0e-4);
    const double k2 = wvno * wvno, t2 = -2.0 * k2;
    const int L = M.L, top = M.top();
    alignas(16) double per[kDispPer], q[kDispTerms];
    if (M.ifunc == 1) {
      disp_love_half(q, omega / M.b[L - 1], wvno, M.rho[L - 1], M.b[L - 1]);
      double e1 = q[0], e2 = q[1];
      for (int m = L - 2; m >= top; --m) {  // a water layer carries no SH motion
        const double xmu = M.rho[m] * M.b[m] * M.b[m];
        disp_love_terms(q, omega / M.b[m], wvno, M.d[m], xmu, 1.0 / xmu);
        disp_love_apply(e1, e2, q);
      }
      return L - 1 > top ? disp_love_finish(e1, e2) : e1;
    }
    disp_period_consts(per, M.a[L - 1], M.b[L - 1], omega);
    disp_half_terms(q, per, wvno, k2, M.rho[L - 1]);
    double e0 = q[0], e1 = q[1], e2 = q[2], e3 = q[3], e4 = q[4];
    for (int m = L - 2; m >= top; --m) {
      disp_period_consts(per, M.a[m], M.b[m], omega);
      disp_layer_terms(q, per, wvno, k2, t2, M.d[m], M.rho[m], 1.0 / M.rho[m]);
      disp_layer_apply(e0, e1, e2, e3, e4, disp_load_terms(q));
    }
    if (top == 1) {
      disp_water_terms(q, omega / M.a[0], wvno, M.d[0]);
      return disp_finish_water(e0, e1, e2, e3, e4, q, M.rho[0]);
    }
    return L > 1 ? disp_finish(e0, e1, e2, e3, e4) : e0;
  }
};

#if defined(__CUDACC__)
// One warp evaluates: lane m prepares layer m (lanes stride over the layers when there are more than 32), then every
// lane runs the same bottom-up product over the staged terms, the next layer's terms loaded while the current layer is
// worked on.  All lanes must call with the same arguments; all return the same value.  The warp's scratch —
// per [L][kDispPer], then lay [L][kDispTerms] — starts `off` doubles into the kernel's dynamic shared memory (named
// here, not passed as a pointer, so that the accesses compile to shared-memory instructions inside the non-inlined
// evaluator).
struct DispWarpEval {
  DispModel M;
  int off;
  int lane;
  double omega0;  // 2 pi / period, unclamped
  __device__ double* scratch() const { return adsurf_smem_head + off; }  // off includes the exp table (kSmemHead)
  __device__ void set_period(double t_) {
    omega0 = kTwoPi / t_;
    const double omega = fmax(omega0, 1.0e-4);
    double* per = scratch();
    __syncwarp();
    for (int m = lane; m < M.L; m += 32) {
      double* pm = per + m * kDispPer;
      if (M.ifunc == 1) {  // Love: beta, k_beta, mu, d, rho, 1/mu
        const double bm = M.b[m], xmu = M.rho[m] * bm * bm;
        pm[0] = bm;
        pm[1] = omega / bm;
        pm[2] = xmu;
        pm[5] = 1.0 / xmu;
      } else {
        disp_period_consts(pm, M.a[m], M.b[m], omega);
        pm[5] = 1.0 / M.rho[m];
      }
      pm[3] = M.d[m];
      pm[4] = M.rho[m];
    }
    __syncwarp();
  }
  __device__ __noinline__ double operator()(double c) const {
    const double wvno = omega0 / c;
    const double k2 = wvno * wvno, t2 = -2.0 * k2;
    const int L = M.L, top = M.top();
    const double* per = scratch();
    double* lay = scratch() + L * kDispPer;
    if (M.ifunc == 1) {
      for (int m = top + lane; m < L; m += 32) {
        const double* pm = per + m * kDispPer;
        if (m < L - 1) disp_love_terms(lay + m * kDispTerms, pm[1], wvno, pm[3], pm[2], pm[5]);
        else disp_love_half(lay + m * kDispTerms, pm[1], wvno, pm[4], pm[0]);
      }
      __syncwarp();
      const double* q = lay + (L - 1) * kDispTerms;
      double e1 = q[0], e2 = q[1];
      for (int m = L - 2; m >= top; --m) disp_love_apply(e1, e2, lay + m * kDispTerms);
      __syncwarp();
      return L - 1 > top ? disp_love_finish(e1, e2) : e1;
    }
    for (int m = lane; m < L; m += 32) {
      const double* pm = per + m * kDispPer;
      if (m < top) disp_water_terms(lay, pm[0], wvno, pm[3]);  // layer 0 is water
      else if (m < L - 1) disp_layer_terms(lay + m * kDispTerms, pm, wvno, k2, t2, pm[3], pm[4], pm[5]);
      else disp_half_terms(lay + m * kDispTerms, pm, wvno, k2, pm[4]);
    }
    __syncwarp();
    const double* q = lay + (L - 1) * kDispTerms;
    double e0 = q[0], e1 = q[1], e2 = q[2], e3 = q[3], e4 = q[4];
    if (L - 1 > top) {
      q -= kDispTerms;
      DispTermRegs cur = disp_load_terms(q);
      for (int m = L - 2; m >= top; --m) {
        if (m > top) q -= kDispTerms;  // the last layer re-reads its own terms: no predicated loads
        const DispTermRegs nxt = disp_load_terms(q);
        disp_layer_apply(e0, e1, e2, e3, e4, cur);
        cur = nxt;
      }
      e0 = top ? disp_finish_water(e0, e1, e2, e3, e4, lay, per[4]) : disp_finish(e0, e1, e2, e3, e4);
    }
    __syncwarp();  // the next evaluation overwrites the staged terms
    return e0;
  }
};
#endif

ADSURF_HD double np_sign(double v) { return v > 0.0 ? 1.0 : (v < 0.0 ? -1.0 : (v == 0.0 ? 0.0 : v)); }

// refine a bracketed root: Neville rational interpolation with bisection safeguards (:452-541)
template <class Eval>
ADSURF_HDN double disp_nevill(const Eval& ev, double c1, double c2, double del1, double del2) {
  double xs[20], ys[20];
  for (int i = 0; i < 20; ++i) xs[i] = ys[i] = 0.0;
  double c3 = 0.5 * (c1 + c2);
  double del3 = ev(c3);
  int nev = 1, m = 1;
  for (int it = 0; it < 98; ++it) {
    if (c3 < fmin(c1, c2) || c3 > fmax(c1, c2)) {  // estimate left the bracket: halve instead
      nev = 0;
      c3 = 0.5 * (c1 + c2);
      del3 = ev(c3);
    }
    const double s13 = del1 - del3, s32 = del3 - del2;
    if (np_sign(del3) * np_sign(del1) < 0.0) { c2 = c3; del2 = del3; }
    else { c1 = c3; del1 = del3; }
    if (fabs(c1 - c2) <= 1.0e-6 * c1) break;
    if (np_sign(s13) != np_sign(s32)) nev = 0;
    const double ss1 = fabs(del1), ss2 = fabs(del2);
    if (0.01 * ss1 > ss2 || 0.01 * ss2 > ss1 || nev == 0) {  // values too unequal for a polynomial fit
      c3 = 0.5 * (c1 + c2);
      del3 = ev(c3);
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
      del3 = ev(c3);
      nev = 2;
      m = m + 1 < 10 ? m + 1 : 10;
    } else {
      c3 = 0.5 * (c1 + c2);
      del3 = ev(c3);
      nev = 1;
      m = 1;
    }
  }
  return c3;
}

// walk the dc ladder until the secular function changes sign, then refine (:543-574); returns iret
template <class Eval>
ADSURF_HD bool disp_getsol(const Eval& ev, double& c1, double clow, double dc, double cm, double betmx, bool ifirst,
                            double& del1st) {
  double del1 = ev(c1);
  if (ifirst) del1st = del1;
  double idir = (!ifirst && np_sign(del1st) * np_sign(del1) < 0.0) ? -1.0 : 1.0;
  for (;;) {
    const double c2 = c1 + idir * dc;
    if (c2 <= clow) {
      idir = 1.0;
      c1 = clow;
      continue;
    }
    const double del2 = ev(c2);
    if (np_sign(del1) != np_sign(del2)) {
      c1 = disp_nevill(ev, c1, c2, del1, del2);
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
// ev evaluates the secular function of model M at the period given to set_period().
template <class Eval>
ADSURF_HD int disp_getc(Eval& ev, const DispModel& M, const double* ts, int NF, int nmodes, double dc, double* cout,
                         double* c) {
  const int L = M.L;
  double betmx = -1.0e20, betmn = 1.0e20;
  int jmn = 0;
  bool jsol = false;
  for (int i = 0; i < L; ++i) {
    const double v = M.b[i];
    if (v > 0.01 && v < betmn) { betmn = v; jmn = i; jsol = false; }
    else if (v < 0.01 && M.a[i] < betmn) { betmn = M.a[i]; jmn = i; jsol = true; }  // a fluid slower than the solids (:617-620)
    betmx = fmax(betmx, v);
  }
  // start a little below the slowest layer's Rayleigh velocity (the fluid's Vp if that is the slowest)
  const double cc = 0.9 * (jsol ? betmn : disp_gtsolh(M.a[jmn], M.b[jmn]));
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
      ev.set_period(ts[k]);
      const bool iret = disp_getsol(ev, c1, clow, dc, cm, betmx, ifirst, del1st);
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
