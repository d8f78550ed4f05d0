// adsurf_math.cuh — per-sample arithmetic of the Rayleigh secular function and its hand-derived
// adjoint.  Shared by the CUDA kernels (adsurf_kernels.cu) and by the host-side math check that
// tests/ builds from this same header (tests/host_math/host_math.cu) — the check compiles these
// functions for the CPU so the adjoint algebra can be verified against the oracle without a GPU.
// The product library never runs this code on the host.
//
// Specification: SURVEY.md Appendix A == /root/reference/ADsurf/_cps/_surf96_vectorAll_gpu.py
// :52-100 (Dunkin matrix), :102-192 (eigenfunction terms), :303-399 (recursion).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define ADSURF_HD __host__ __device__ __forceinline__
#else
#define ADSURF_HD inline
#endif

namespace adsurf {

constexpr double kTwoPi = 6.283185307179586;  // 2.0 * np.pi
constexpr double kLn10 = 2.302585092994046;
constexpr int kWarp = 32;
constexpr int kNumPar = 6;  // staged per-layer constants
enum { P_IA = 0, P_IB = 1, P_TB2 = 2, P_RHO = 3, P_IR = 4, P_D = 5 };

ADSURF_HD double rsqrt_hd(double u) {
#if defined(__CUDA_ARCH__)
  return rsqrt(u);
#else
  return 1.0 / sqrt(u);
#endif
}

// ------------------------------------------------------------------------------------------
// Per-layer quantities
// ------------------------------------------------------------------------------------------
struct Eig {
  double r, ir;    // sqrt((k+kx)|k-kx|) and its reciprocal
  double cs, sn;   // cos p / (1+e^-2p)/2   and   sin p / (1-e^-2p)/2
  double w, x;     // sn / r   and   -+ r sn
  double ex, fac;  // exponent carried out of the term (p when evanescent) and e^{-2p}
  int br;          // 0: oscillatory (k<kx), 1: evanescent (k>kx), 2: k==kx, 3: unordered (NaN)
};

// var_vector, reference :109-136 (P) / :140-167 (S)
ADSURF_HD Eig eig_terms(double k, double kx, double dm) {
  Eig e;
  const double u = (k + kx) * fabs(k - kx);
  e.ir = rsqrt_hd(u);
  e.r = u * e.ir;
  if (!(u > 0.0)) e.r = sqrt(u);  // u == 0 -> 0, NaN -> NaN
  const double p = e.r * dm;
  e.cs = 0.0; e.sn = 0.0; e.w = 0.0; e.x = 0.0; e.ex = 0.0; e.fac = 0.0; e.br = 3;
  if (k < kx) {
    e.br = 0;
    sincos(p, &e.sn, &e.cs);
    e.w = e.sn * e.ir;
    e.x = -e.r * e.sn;
  } else if (k > kx) {
    e.br = 1;
    e.ex = p;
    e.fac = (p < 16.0) ? exp(-2.0 * p) : 0.0;
    e.cs = (1.0 + e.fac) * 0.5;
    e.sn = (1.0 - e.fac) * 0.5;
    e.w = e.sn * e.ir;
    e.x = e.r * e.sn;
  } else if (k == kx) {
    e.br = 2;
    e.cs = 1.0;
    e.w = dm;
  }
  return e;
}

struct Layer {
  // inputs
  double ka, kb, gammk, rho, ir, dm;
  // derived scalars
  double gam, gamm1, twgm1, gmgmk, gmgm1, gm1sq, t;
  Eig P, S;
  double a0, a0pq, cpcq, cpy, cpz, cqw, cqx, xy, xz, wy, wz;
  // the 13 independent Dunkin entries (+ c22); c11 = cpcq
  double c00, c01, c02, c03, c04, c10, c12, c13, c30, c31, c32, c40, c42, c22;
};

// dnka_vector (reference :52-100) on top of var_vector (:170-182)
ADSURF_HD void layer_terms(Layer& y, double k, double k2) {
  y.P = eig_terms(k, y.ka, y.dm);
  y.S = eig_terms(k, y.kb, y.dm);
  const double exa = y.P.ex + y.S.ex;
  y.a0 = (exa < 60.0) ? exp(-exa) : 0.0;
  y.cpcq = y.P.cs * y.S.cs;
  y.cpy = y.P.cs * y.S.w;
  y.cpz = y.P.cs * y.S.x;
  y.cqw = y.S.cs * y.P.w;
  y.cqx = y.S.cs * y.P.x;
  y.xy = y.P.x * y.S.w;
  y.xz = y.P.x * y.S.x;
  y.wy = y.P.w * y.S.w;
  y.wz = y.P.w * y.S.x;

  y.gam = y.gammk * k2;
  y.gamm1 = y.gam - 1.0;
  y.twgm1 = y.gam + y.gamm1;
  y.gmgmk = y.gam * y.gammk;
  y.gmgm1 = y.gam * y.gamm1;
  y.gm1sq = y.gamm1 * y.gamm1;
  y.a0pq = y.a0 - y.cpcq;
  y.t = -2.0 * k2;
  const double rho = y.rho, ir = y.ir, rho2 = rho * rho, ir2 = ir * ir;

  y.c00 = y.cpcq - 2.0 * y.gmgm1 * y.a0pq - y.gmgmk * y.xz - k2 * y.gm1sq * y.wy;
  y.c01 = (k2 * y.cpy - y.cqx) * ir;
  y.c02 = -(y.twgm1 * y.a0pq + y.gammk * y.xz + k2 * y.gamm1 * y.wy) * ir;
  y.c03 = (y.cpz - k2 * y.cqw) * ir;
  y.c04 = -(2.0 * k2 * y.a0pq + y.xz + k2 * k2 * y.wy) * ir2;
  y.c10 = (y.gmgmk * y.cpz - y.gm1sq * y.cqw) * rho;
  y.c12 = y.gammk * y.cpz - y.gamm1 * y.cqw;
  y.c13 = -y.wz;
  y.c30 = (y.gm1sq * y.cpy - y.gmgmk * y.cqx) * rho;
  y.c31 = -y.xy;
  y.c32 = y.gamm1 * y.cpy - y.gammk * y.cqx;
  y.c40 = -(2.0 * y.gmgmk * y.gm1sq * y.a0pq + y.gmgmk * y.gmgmk * y.xz + y.gm1sq * y.gm1sq * y.wy) * rho2;
  y.c42 = -(y.gammk * y.gamm1 * y.twgm1 * y.a0pq + y.gam * y.gammk * y.gammk * y.xz + y.gamm1 * y.gm1sq * y.wy) * rho;
  y.c22 = y.a0 + 2.0 * (y.cpcq - y.c00);
}

// E <- E . CA   (row vector times Dunkin matrix, reference :376-383)
ADSURF_HD void row_apply(double (&e)[5], const Layer& y) {
  const double e0 = e[0], e1 = e[1], e2 = e[2], e3 = e[3], e4 = e[4];
  const double te2 = y.t * e2;
  e[0] = e0 * y.c00 + e1 * y.c10 + te2 * y.c42 + e3 * y.c30 + e4 * y.c40;
  e[1] = e0 * y.c01 + e1 * y.cpcq + te2 * y.c32 + e3 * y.c31 + e4 * y.c30;
  e[2] = e0 * y.c02 + e1 * y.c12 + e2 * y.c22 + e3 * y.c32 + e4 * y.c42;
  e[3] = e0 * y.c03 + e1 * y.c13 + te2 * y.c12 + e3 * y.cpcq + e4 * y.c10;
  e[4] = e0 * y.c04 + e1 * y.c03 + te2 * y.c02 + e3 * y.c01 + e4 * y.c00;
}

// s <- CA . s   (column vector of the layers above)
ADSURF_HD void col_apply(double (&s)[5], const Layer& y) {
  const double s0 = s[0], s1 = s[1], s2 = s[2], s3 = s[3], s4 = s[4];
  s[0] = y.c00 * s0 + y.c01 * s1 + y.c02 * s2 + y.c03 * s3 + y.c04 * s4;
  s[1] = y.c10 * s0 + y.cpcq * s1 + y.c12 * s2 + y.c13 * s3 + y.c03 * s4;
  s[2] = y.t * (y.c42 * s0 + y.c32 * s1 + y.c12 * s3 + y.c02 * s4) + y.c22 * s2;
  s[3] = y.c30 * s0 + y.c31 * s1 + y.c32 * s2 + y.cpcq * s3 + y.c01 * s4;
  s[4] = y.c40 * s0 + y.c30 * s1 + y.c42 * s2 + y.c10 * s3 + y.c00 * s4;
}

struct LayerGrad {
  double d, ka, kb, gammk, rho, k;  // dDelta/d(each input of the layer), k includes the k2 path
};

// adjoint of one eigenfunction block: (bcs, bw, bx, bex) -> (bd, bk, bkx)
ADSURF_HD void eig_adjoint(const Eig& e, double k, double kx, double dm, double bcs, double bw,
                                            double bx, double bex, double& bd, double& bk, double& bkx) {
  double bp = 0.0, br = 0.0, sgn = 0.0;
  if (e.br == 0) {
    const double bsn = bw * e.ir - e.r * bx;
    bp = -e.sn * bcs + e.cs * bsn;
    br = -e.w * e.ir * bw - e.sn * bx;
    sgn = -1.0;
  } else if (e.br == 1) {
    const double bsn = bw * e.ir + e.r * bx;
    const double bfac = 0.5 * (bcs - bsn);
    bp = -2.0 * e.fac * bfac + bex;
    br = -e.w * e.ir * bw + e.sn * bx;
    sgn = 1.0;
  } else if (e.br == 2) {
    bd += bw;  // w = d; the (measure-zero) dependence through r is dropped (finite one-sided value)
  }
  if (e.br < 2) {
    br += dm * bp;
    bd += e.r * bp;
    const double h = sgn * br * e.ir;  // d r / d u * 2 = 1/r
    bk += k * h;
    bkx -= kx * h;
  }
}

// Reverse sweep through one layer: given row vector e (= E_{m+1}) and column s (= s_m, already
// scaled by the upstream seed) returns dDelta/d(layer inputs).
ADSURF_HD LayerGrad layer_adjoint(const Layer& y, const double (&e)[5], const double (&s)[5],
                                                   double k, double k2) {
  const double e0 = e[0], e1 = e[1], e2 = e[2], e3 = e[3], e4 = e[4];
  const double s0 = s[0], s1 = s[1], s2 = s[2], s3 = s[3], s4 = s[4];
  const double e2s2 = e2 * s2;
  const double ts = y.t * e2;
  const double b00 = e0 * s0 + e4 * s4 - 2.0 * e2s2;
  const double b01 = e0 * s1 + e3 * s4;
  const double b02 = e0 * s2 + ts * s4;
  const double b03 = e0 * s3 + e1 * s4;
  const double b04 = e0 * s4;
  const double b10 = e1 * s0 + e4 * s3;
  const double b11 = e1 * s1 + e3 * s3;
  const double b12 = e1 * s2 + ts * s3;
  const double b13 = e1 * s3;
  const double b30 = e3 * s0 + e4 * s1;
  const double b31 = e3 * s1;
  const double b32 = e3 * s2 + ts * s1;
  const double b40 = e4 * s0;
  const double b42 = e4 * s2 + ts * s0;
  const double bt = e2 * (y.c42 * s0 + y.c32 * s1 + y.c12 * s3 + y.c02 * s4);

  const double rho = y.rho, ir = y.ir;
  const double B01 = ir * b01, B02 = ir * b02, B03 = ir * b03, B04 = ir * ir * b04;
  const double B10 = rho * b10, B30 = rho * b30, B40 = rho * rho * b40, B42 = rho * b42;
  const double gam = y.gam, gk = y.gammk, g1 = y.gamm1, tw = y.twgm1, gg = y.gmgmk, q1 = y.gm1sq;
  const double A = y.a0pq;

  // adjoints of the eigenfunction products
  const double bA = -2.0 * y.gmgm1 * b00 - tw * B02 - 2.0 * k2 * B04 - 2.0 * gg * q1 * B40 - gk * g1 * tw * B42;
  const double bxz = -gg * b00 - gk * B02 - B04 - gg * gg * B40 - gam * gk * gk * B42;
  const double bwy = -k2 * q1 * b00 - k2 * g1 * B02 - k2 * k2 * B04 - q1 * q1 * B40 - g1 * q1 * B42;
  const double bcpy = k2 * B01 + q1 * B30 + g1 * b32;
  const double bcqx = -B01 - gg * B30 - gk * b32;
  const double bcpz = B03 + gg * B10 + gk * b12;
  const double bcqw = -k2 * B03 - q1 * B10 - g1 * b12;
  const double bwz = -b13;
  const double bxy = -b31;
  const double bcpcq = b00 + b11 + 2.0 * e2s2 - bA;
  const double ba0 = bA + e2s2;

  // adjoints of the coefficient scalars
  const double bg1 = -2.0 * A * b00;
  const double bgg = -y.xz * b00 + y.cpz * B10 - y.cqx * B30 - 2.0 * (q1 * A + gg * y.xz) * B40;
  const double bq1 = -k2 * y.wy * b00 - y.cqw * B10 + y.cpy * B30 - 2.0 * (gg * A + q1 * y.wy) * B40;
  const double btw = -A * B02;
  const double bgm1 = -k2 * y.wy * B02 - y.cqw * b12 + y.cpy * b32;
  double bgk = -y.xz * B02 + y.cpz * b12 - y.cqx * b32 - (g1 * tw * A + 2.0 * gam * gk * y.xz) * B42;
  const double bgam = -(gk * (4.0 * gam - 3.0) * A + gk * gk * y.xz + 3.0 * q1 * y.wy) * B42 + tw * bg1 + gk * bgg +
                      2.0 * g1 * bq1 + 2.0 * btw + bgm1;
  bgk += gam * bgg + k2 * bgam;
  const double bk2 = -q1 * y.wy * b00 + y.cpy * B01 - g1 * y.wy * B02 - y.cqw * B03 - 2.0 * (A + k2 * y.wy) * B04 -
                     2.0 * bt + gk * bgam;
  const double brho = (-(y.c01 * b01 + y.c02 * b02 + y.c03 * b03) - 2.0 * y.c04 * b04 + y.c10 * b10 + y.c30 * b30 +
                       2.0 * y.c40 * b40 + y.c42 * b42) * ir;

  // products -> eigenfunction terms
  const Eig& P = y.P;
  const Eig& S = y.S;
  const double bcosp = S.cs * bcpcq + S.w * bcpy + S.x * bcpz;
  const double bcosq = P.cs * bcpcq + P.w * bcqw + P.x * bcqx;
  const double bw = S.cs * bcqw + S.w * bwy + S.x * bwz;
  const double bx = S.cs * bcqx + S.w * bxy + S.x * bxz;
  const double by = P.cs * bcpy + P.x * bxy + P.w * bwy;
  const double bz = P.cs * bcpz + P.x * bxz + P.w * bwz;
  const double bex = -y.a0 * ba0;

  LayerGrad g;
  g.d = 0.0; g.ka = 0.0; g.kb = 0.0;
  g.k = 2.0 * k * bk2;
  g.gammk = bgk;
  g.rho = brho;
  eig_adjoint(P, k, y.ka, y.dm, bcosp, bw, bx, bex, g.d, g.k, g.ka);
  eig_adjoint(S, k, y.kb, y.dm, bcosq, by, bz, bex, g.d, g.k, g.kb);
  return g;
}

// Half-space start vector (reference :328-347)
struct Half {
  double ka, kb, gammk, rho;
  double ra, ira, rb, irb, gam, gamm1, R;
  int bra, brb;  // 0 osc (k<kx), 1 evan (k>kx), 2 equal/unordered
};

ADSURF_HD void sqrt_pair(double k, double kx, double& r, double& ir, int& br) {
  const double u = (k + kx) * fabs(k - kx);
  ir = rsqrt_hd(u);
  r = u * ir;
  if (!(u > 0.0)) r = sqrt(u);
  br = (k < kx) ? 0 : ((k > kx) ? 1 : 2);
}

ADSURF_HD void half_space(Half& h, double k, double k2, double (&e)[5]) {
  sqrt_pair(k, h.ka, h.ra, h.ira, h.bra);
  sqrt_pair(k, h.kb, h.rb, h.irb, h.brb);
  h.gam = h.gammk * k2;
  h.gamm1 = h.gam - 1.0;
  h.R = h.ra * h.rb;
  const double rho = h.rho;
  e[0] = rho * rho * (h.gamm1 * h.gamm1 - h.gam * h.gammk * h.R);
  e[1] = -rho * h.ra;
  e[2] = rho * (h.gamm1 - h.gammk * h.R);
  e[3] = rho * h.rb;
  e[4] = k2 - h.R;
}

ADSURF_HD LayerGrad half_adjoint(const Half& h, const double (&e)[5], const double (&s)[5], double k,
                                                  double k2) {
  const double rho = h.rho, ir = 1.0 / rho;
  const double rho2s0 = rho * rho * s[0];
  const double bq1 = rho2s0;
  const double bgg = -rho2s0 * h.R;
  double bR = -rho2s0 * h.gam * h.gammk - rho * h.gammk * s[2] - s[4];
  const double bgm1 = rho * s[2];
  double bgk = -rho * h.R * s[2];
  double bra = -rho * s[1] + h.rb * bR;
  double brb = rho * s[3] + h.ra * bR;
  const double bgam = 2.0 * h.gamm1 * bq1 + h.gammk * bgg + bgm1;
  bgk += h.gam * bgg + k2 * bgam;
  const double bk2 = s[4] + h.gammk * bgam;
  LayerGrad g;
  g.d = 0.0;
  g.rho = (2.0 * e[0] * s[0] + e[1] * s[1] + e[2] * s[2] + e[3] * s[3]) * ir;
  g.gammk = bgk;
  g.k = 2.0 * k * bk2;
  g.ka = 0.0; g.kb = 0.0;
  if (h.bra < 2) {
    const double hh = (h.bra ? 1.0 : -1.0) * bra * h.ira;
    g.k += k * hh;
    g.ka -= h.ka * hh;
  }
  if (h.brb < 2) {
    const double hh = (h.brb ? 1.0 : -1.0) * brb * h.irb;
    g.k += k * hh;
    g.kb -= h.kb * hh;
  }
  return g;
}

// ------------------------------------------------------------------------------------------
// Sample-level drivers
// ------------------------------------------------------------------------------------------
struct Sample {
  double k, k2, omega, iw2;  // wavenumber, its square, clamped angular frequency, 1/omega^2
};

// omega = twopi / t as torch evaluates `python_float / tensor` (Tensor.__rtruediv__ ==
// reciprocal(t) * twopi) — the form every tensor-input caller of the reference goes through
// (_cps/_surf96_vectorAll_gpu.py:318, _matrixAll_gpu.py:256).  Matching the last bit matters only
// where a trial velocity coincides with the half-space velocity (see load_half).
ADSURF_HD double omega_of_period(double t) { return (1.0 / t) * kTwoPi; }

ADSURF_HD Sample make_sample(double omega0, double ic) {
  Sample z;
  z.k = ic * omega0;                          // (1/c) * omega   (reference :319, unclamped omega)
  z.omega = fmax(omega0, 1.0e-4);             // reference :326
  z.k2 = z.k * z.k;
  const double iw = 1.0 / z.omega;
  z.iw2 = iw * iw;
  return z;
}

ADSURF_HD void load_layer(Layer& y, const double* __restrict__ par, int L, int m, const Sample& z) {
  y.ka = z.omega * par[P_IA * L + m];
  y.kb = z.omega * par[P_IB * L + m];
  y.gammk = par[P_TB2 * L + m] * z.iw2;
  y.rho = par[P_RHO * L + m];
  y.ir = par[P_IR * L + m];
  y.dm = par[P_D * L + m];
}

// The half-space vector is linear in ra, rb = sqrt((k+kx)|k-kx|): at a trial velocity that coincides
// with the half-space velocity the result depends on the last bit of kx, so kx is formed with the
// reference's own correctly-rounded division omega/alpha (reference :328-329) from the raw
// velocities kept at par[kNumPar*L + 0/1].  Inside the layers the dependence on kx is smooth
// (even in r), and the cheaper omega * (1/alpha) is used.
ADSURF_HD void load_half(Half& h, const double* __restrict__ par, int L, const Sample& z) {
  const int m = L - 1;
  h.ka = z.omega / par[kNumPar * L + 0];
  h.kb = z.omega / par[kNumPar * L + 1];
  h.gammk = par[P_TB2 * L + m] * z.iw2;
  h.rho = par[P_RHO * L + m];
}

// forward only
ADSURF_HD double sample_forward(const double* __restrict__ par, int L, const Sample& z) {
  double e[5];
  Half h;
  load_half(h, par, L, z);
  half_space(h, z.k, z.k2, e);
  Layer y;
  for (int m = L - 2; m >= 0; --m) {
    load_layer(y, par, L, m, z);
    layer_terms(y, z.k, z.k2);
    row_apply(e, y);
  }
  return e[0];
}

// forward keeping E_{m+1} of every layer: slot m holds the row vector that multiplies CA_m.
// `est` points at this thread's lane; element (m, comp) is est[(m*5+comp)*kStride]
// (kStride = 32 in shared memory: lane-interleaved, conflict-free).
template <int kStride>
ADSURF_HD double sample_forward_store(const double* __restrict__ par, int L, const Sample& z,
                                      double* __restrict__ est) {
  double e[5];
  Half h;
  load_half(h, par, L, z);
  half_space(h, z.k, z.k2, e);
  Layer y;
  for (int m = L - 2; m >= 0; --m) {
#pragma unroll
    for (int i = 0; i < 5; ++i) est[(m * 5 + i) * kStride] = e[i];
    load_layer(y, par, L, m, z);
    layer_terms(y, z.k, z.k2);
    row_apply(e, y);
  }
  return e[0];
}

// Top-down adjoint sweep.  s_0 = seed * e_0; for every layer m (and finally the half-space, m = L-1)
// hands (m, dDelta/dd_m, /dalpha_m, /dbeta_m, /drho_m) * seed to `sink`.  Returns dDelta/dk * seed.
template <int kStride, typename Sink>
ADSURF_HD double sample_adjoint(const double* __restrict__ par, int L, const Sample& z, double seed,
                                const double* __restrict__ est, Sink& sink) {
  double s[5] = {seed, 0.0, 0.0, 0.0, 0.0};
  double e[5];
  double gk = 0.0;
  Layer y;
  for (int m = 0; m <= L - 2; ++m) {
#pragma unroll
    for (int i = 0; i < 5; ++i) e[i] = est[(m * 5 + i) * kStride];
    load_layer(y, par, L, m, z);
    layer_terms(y, z.k, z.k2);
    const LayerGrad g = layer_adjoint(y, e, s, z.k, z.k2);
    col_apply(s, y);
    gk += g.k;
    const double ia = par[P_IA * L + m], ib = par[P_IB * L + m];
    // k_alpha = omega/alpha, k_beta = omega/beta, gammk = 2 beta^2 / omega^2
    sink(m, g.d, -y.ka * ia * g.ka, -y.kb * ib * g.kb + 2.0 * y.gammk * ib * g.gammk, g.rho);
  }
  // half-space: dDelta/dE_hs = s_{L-1}
  Half h;
  load_half(h, par, L, z);
  half_space(h, z.k, z.k2, e);
  const LayerGrad g = half_adjoint(h, e, s, z.k, z.k2);
  gk += g.k;
  const int m = L - 1;
  const double ia = par[P_IA * L + m], ib = par[P_IB * L + m];
  sink(m, 0.0, -h.ka * ia * g.ka, -h.kb * ib * g.kb + 2.0 * h.gammk * ib * g.gammk, g.rho);
  return gk;
}

// per-model layer constants staged for the kernels: par[P_*][L]
ADSURF_HD void stage_one(double* __restrict__ par, int L, int m, double dm, double a, double b, double r) {
  par[P_IA * L + m] = 1.0 / a;
  par[P_IB * L + m] = 1.0 / b;
  par[P_TB2 * L + m] = 2.0 * b * b;
  par[P_RHO * L + m] = r;
  par[P_IR * L + m] = 1.0 / r;
  par[P_D * L + m] = dm;
  if (m == L - 1) {
    par[kNumPar * L + 0] = a;
    par[kNumPar * L + 1] = b;
  }
}

// doubles of staged constants per model
ADSURF_HD int par_doubles(int L) { return kNumPar * L + 2; }

}  // namespace adsurf
