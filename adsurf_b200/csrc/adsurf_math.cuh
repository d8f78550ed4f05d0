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
#include <string.h>

#if defined(__CUDACC__)
#define ADSURF_HD __host__ __device__ __forceinline__
#else
#define ADSURF_HD inline
#endif

namespace adsurf {

constexpr double kTwoPi = 6.283185307179586;  // 2.0 * np.pi
constexpr double kLn10 = 2.302585092994046;
constexpr int kWarp = 32;
constexpr int kNumPar = 10;  // staged per-layer record (80 bytes: five 16-byte loads)
// 1/alpha, 1/beta, 2 beta^2, rho, 1/rho, d | adjoint only: 1/alpha^3, 1/beta^3, 4 beta, (pad)
enum { P_IA = 0, P_IB = 1, P_TB2 = 2, P_RHO = 3, P_IR = 4, P_D = 5, P_IA3 = 6, P_IB3 = 7, P_CB = 8 };

// ---- inline elementary functions -------------------------------------------------------------
// Straight-line (call-free, branch-free on the common path) so that the P and S terms of a layer
// are scheduled together: the library versions cost about as many FP64 instructions but sit behind
// calls, slow-path branches and ~3x as many integer/uniform instructions, which compete with the
// half-rate FP64 pipe for issue slots.
ADSURF_HD long long dbl_bits(double v) {
#if defined(__CUDA_ARCH__)
  return __double_as_longlong(v);
#else
  long long b; memcpy(&b, &v, sizeof(b)); return b;
#endif
}
ADSURF_HD double bits_dbl(long long b) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double(b);
#else
  double v; memcpy(&v, &b, sizeof(v)); return v;
#endif
}

// 1/sqrt(u), u > 0 normal.  Device: 23-bit hardware seed (MUFU.RSQ64H) + one cubic Newton step
// (relative error ~ seed^3): about 1 ulp.  u == 0 / NaN callers test u themselves.
ADSURF_HD double rsqrt_hd(double u) {
#if defined(__CUDA_ARCH__)
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(u));
  const double t = u * y;
  const double e = fma(-t, y, 1.0);
  return fma(y * e, fma(0.375, e, 0.5), y);
#else
  return 1.0 / sqrt(u);
#endif
}

// "v < T" for a threshold T whose low 32 bits are zero (16, 60, 1e5) and v >= 0 (or NaN): compare the high
// words as integers — one ALU instruction instead of a DSETP that occupies the half-rate FP64 pipe.
// NaN (high word 0x7ff8....) compares false like the floating-point test.
ADSURF_HD int hi_word(double v) {
#if defined(__CUDA_ARCH__)
  return __double2hiint(v);
#else
  return (int)(dbl_bits(v) >> 32);
#endif
}
ADSURF_HD bool below16(double v) { return hi_word(v) < 0x40300000; }
ADSURF_HD bool below60(double v) { return hi_word(v) < 0x404E0000; }
ADSURF_HD bool below1e5(double v) { return hi_word(v) < 0x40F86A00; }

constexpr double kShifter = 6755399441055744.0;  // 1.5 * 2^52: adds round-to-nearest integer into the low mantissa bits

// Polynomial coefficients live in constant memory on the device so that DFMA reads them as
// c[bank][offset] operands instead of materialising each 64-bit literal with two UMOVs.
#define ADSURF_SIN_COEFS                                                                              \
  { 1.58969099521155010221e-10, -2.50507602534068634195e-08, 2.75573137070700676789e-06,               \
    -1.98412698298579493134e-04, 8.33333333332248946124e-03, -1.66666666666666324348e-01 }
#define ADSURF_COS_COEFS                                                                              \
  { -1.13596475577881948265e-11, 2.08757232129817482790e-09, -2.75573143513906633035e-07,              \
    2.48015872894767294178e-05, -1.38888888888741095749e-03, 4.16666666666666019037e-02 }
#define ADSURF_RED_CONSTS                                                                             \
  { 1.4426950408889634, -0.6931471805599453, -2.3190468138462996e-17, /* log2 e, -ln2 hi, -ln2 lo */    \
    0.6366197723675814, -1.5707963267948966, -6.123233995736766e-17, 1.4973849048591698e-33 /* 2/pi, -pi/2 x3 */ }
static const double kSinCH[6] = ADSURF_SIN_COEFS;
static const double kCosCH[6] = ADSURF_COS_COEFS;
static const double kRedCH[7] = ADSURF_RED_CONSTS;
#if defined(__CUDACC__)
__constant__ double kSinCD[6] = ADSURF_SIN_COEFS;
__constant__ double kCosCD[6] = ADSURF_COS_COEFS;
__constant__ double kRedCD[7] = ADSURF_RED_CONSTS;
#endif
#if defined(__CUDA_ARCH__)
#define ADSURF_K(tab, i) tab##D[i]
#else
#define ADSURF_K(tab, i) tab##H[i]
#endif

// 2^(j/32), j = 0..31 (correctly rounded).  On the device the table lives in the first kSmemHead
// doubles of every kernel's dynamic shared memory (per-lane indices: a constant-memory table would
// serialise); every kernel that evaluates exp_neg stages it with stage_exp_table().
constexpr int kSmemHead = 32;
#define ADSURF_EXP2_TABLE                                                                                     \
  { 1.0, 1.0218971486541166, 1.0442737824274138, 1.0671404006768237, 1.0905077326652577, 1.1143867425958924,    \
    1.1387886347566916, 1.1637248587775775, 1.189207115002721, 1.215247359980469, 1.241857812073484,            \
    1.2690509571917332, 1.2968395546510096, 1.3252366431597413, 1.3542555469368927, 1.383909881963832,          \
    1.4142135623730951, 1.4451808069770467, 1.4768261459394993, 1.5091644275934228, 1.5422108254079407,         \
    1.5759808451078865, 1.6104903319492543, 1.645755478153965, 1.681792830507429, 1.718619298122478,            \
    1.7562521603732995, 1.7947090750031072, 1.8340080864093424, 1.8741676341103, 1.9152065613971474,            \
    1.9571441241754002 }
static const double kExp2TabH[32] = ADSURF_EXP2_TABLE;
#if defined(__CUDACC__)
__constant__ double kExp2TabD[32] = ADSURF_EXP2_TABLE;
extern __shared__ double adsurf_smem_head[];
__device__ __forceinline__ void stage_exp_table() {
  for (int j = threadIdx.x; j < kSmemHead; j += blockDim.x) adsurf_smem_head[j] = kExp2TabD[j];
}
#endif
#if defined(__CUDA_ARCH__)
#define ADSURF_EXP2_TAB(j) adsurf_smem_head[j]
#else
#define ADSURF_EXP2_TAB(j) kExp2TabH[j]
#endif

// exp(-p) for 0 <= p < ~700: n = rint(-p 32/ln2), r = -p - n ln2/32 (two-term, |r| <= ln2/64),
// e^r by a degree-6 Taylor polynomial (truncation 3.5e-18), times 2^(j/32) from the table (j = n mod 32)
// scaled by 2^(n div 32) through the exponent field.  11 FP64 instructions, <= 2 ulp.
// For p >= ~700 the exponent field wraps and the value is meaningless — every caller discards it there
// (the e^{-2p} term is used only for p < 16, e^{-p} only for p < 60), so no clamp is spent on it.
ADSURF_HD double exp_neg(double p) {
  const double x = -p;
  const double t = fma(x, 46.16624130844683, kShifter);
  const double fn = t - kShifter;
  double r = fma(fn, -0.02166084939249829, x);
  r = fma(fn, -7.247021293269686e-19, r);
  const int n = (int)dbl_bits(t);  // low 32 bits hold the (two's complement) integer
  double q = fma(0.001388888888888889, r, 0.008333333333333333);
  q = fma(q, r, 0.041666666666666664);
  q = fma(q, r, 0.16666666666666666);
  q = fma(q, r, 0.5);
  q = fma(q, r, 1.0);
  q = fma(q, r, 1.0);
  const double tj = bits_dbl((long long)((unsigned long long)dbl_bits(ADSURF_EXP2_TAB(n & 31)) +
                                         ((unsigned long long)(long long)(n >> 5) << 52)));
  return tj * q;
}

// sin and cos of q for |q| < 1e5: Cody-Waite reduction by pi/2 in three FMA steps, then the fdlibm
// kernel polynomials on |r| <= pi/4 (coefficients of __kernel_sin/__kernel_cos, (C) 1993 Sun
// Microsystems, freely usable with this notice; < 1 ulp).  Larger arguments take the library path.
ADSURF_HD void sincos_core(double q, double* sn, double* cs);

ADSURF_HD void sincos_fast(double q, double* sn, double* cs) {
  if (!(fabs(q) < 1.0e5)) {
    sincos(q, sn, cs);
    return;
  }
  sincos_core(q, sn, cs);
}

// branch-free core, valid for |q| < 1e5
ADSURF_HD void sincos_core(double q, double* sn, double* cs) {
  const double t = fma(q, ADSURF_K(kRedC, 3), kShifter);
  const double fn = t - kShifter;
  double r = fma(fn, ADSURF_K(kRedC, 4), q);
  r = fma(fn, ADSURF_K(kRedC, 5), r);
  r = fma(fn, ADSURF_K(kRedC, 6), r);
  const double z = r * r;
  double ps = ADSURF_K(kSinC, 0);
#pragma unroll
  for (int i = 1; i < 6; ++i) ps = fma(ps, z, ADSURF_K(kSinC, i));
  const double s = fma(r * z, ps, r);
  double pc = ADSURF_K(kCosC, 0);
#pragma unroll
  for (int i = 1; i < 6; ++i) pc = fma(pc, z, ADSURF_K(kCosC, i));
  const double c = fma(z * z, pc, fma(-0.5, z, 1.0));
  const int n = (int)dbl_bits(t);
  const double a = (n & 1) ? c : s;
  const double b = (n & 1) ? s : c;
  *sn = (n & 2) ? -a : a;
  *cs = ((n + 1) & 2) ? -b : b;
}

// ------------------------------------------------------------------------------------------
// Per-layer quantities
// ------------------------------------------------------------------------------------------
// Structure used below (derived from the entries of Dunkin's matrix, reference :66-99).  With
//   a  = (gg*rho, gk, 1/rho),  b  = (q1*rho, g1, k2/rho)       acting on components (0,2,4)
//   a' = (1/rho, t*gk, gg*rho), b' = (k2/rho, t*g1, q1*rho)     (gk = gammk, g1 = gam-1, gg = gam*gk,
//                                                                q1 = g1^2, t = -2 k2, A = a0 - cp*cq)
// the compound matrix is
//   CA[024,024] = a0 I - a' (x) (A b + xz a) - b' (x) (A a + wy b)
//   CA[024,1] = cpy b' - cqx a',  CA[024,3] = cpz a' - cqw b',
//   CA[1,024] = cpz a  - cqw b,   CA[3,024] = cpy b  - cqx a,
//   CA[1,1] = CA[3,3] = cp cq,    CA[1,3] = -w z,   CA[3,1] = -x y
// i.e. two rank-one couplings around the Kronecker product of the P block [cp w; x cp] and the S block
// [cq y; z cq].  Row and column products therefore cost ~33 FMAs without ever forming the 13 entries
// or the 9 eigenfunction products, and the adjoint contracts against the same few scalars.
// regime flags of a (velocity, layer) pair: oscillatory (k < kx) and regular (k != kx) for P and S
enum { F_LTP = 1, F_POSP = 2, F_LTS = 4, F_POSS = 8, F_REGULAR = F_POSP | F_POSS };

struct Eig {
  double r, ir;     // sqrt((k+kx)|k-kx|) and its reciprocal
  double p;         // r * d
  double cs, sn;    // cos p / (1+e^-2p)/2   and   sin p / (1-e^-2p)/2
  double w, x;      // sn / r   and   -+ r sn
  double dcs, dsn;  // d cs / dp, d sn / dp
  double ex, eh;    // exponent carried out of the term (p when evanescent); e^{-p} (1 when oscillatory)
  double c0, c1;    // the transcendental results kept for the adjoint sweep: (sin, cos) or (e^{-p}, -)
  bool lt, pos;     // oscillatory branch (k < kx);  u > 0 (false: k == kx or unordered)
};

// var_vector, reference :109-136 (P) / :140-167 (S), split in three steps so that the adjoint sweep can
// re-use the transcendental values of the forward sweep instead of re-evaluating them:
//   eig_setup  — r, 1/r, p and the branch   (cheap; done in both sweeps)
//   eig_trans  — sin/cos or e^{-p}          (expensive; forward sweep only, results stored)
//   eig_finish — everything derived from them.
// e^{-2p} is formed as (e^{-p})^2 and a0 = e^{-(pex+sex)} as e^{-pex} e^{-sex} (thresholds kept).
ADSURF_HD void eig_setup(Eig& e, double k, double kx, double dm) {
  const double u = (k + kx) * fabs(k - kx);
  e.pos = u > 0.0;
  e.lt = k < kx;
  e.ir = rsqrt_hd(u);
  e.r = e.pos ? u * e.ir : u;  // u == 0 -> 0, NaN -> NaN
  e.p = e.r * dm;
}

ADSURF_HD void eig_trans(Eig& e) {
  if (e.lt) {  // warp-uniform in the dense kernels (depends on c and the layer, not on the period)
    sincos_fast(e.p, &e.c0, &e.c1);
  } else {
    e.c0 = exp_neg(e.p);
    e.c1 = 0.0;
  }
}

// branch-free bodies of the two regimes (oscillatory valid for |p| < 1e5)
ADSURF_HD void eig_osc(Eig& e) {
  sincos_core(e.p, &e.c0, &e.c1);
  e.sn = e.c0;
  e.cs = e.c1;
  e.dcs = -e.c0;
  e.dsn = e.c1;
  e.ex = 0.0;
  e.eh = 1.0;
  e.x = -e.r * e.sn;
}

ADSURF_HD void eig_evan(Eig& e) {
  e.c0 = exp_neg(e.p);
  e.c1 = 0.0;
  const double fac = below16(e.p) ? e.c0 * e.c0 : 0.0;
  e.cs = 0.5 + 0.5 * fac;
  e.sn = 0.5 - 0.5 * fac;
  e.dcs = -fac;
  e.dsn = fac;
  e.ex = e.p;  // p = 0 when k == kx
  e.eh = below60(e.p) ? e.c0 : 0.0;
  e.x = e.r * e.sn;
}

// forward sweep: transcendental + derived terms inside one (warp-uniform) branch, no selects
ADSURF_HD void eig_forward(Eig& e, double dm) {
  if (e.lt) {
    sincos_fast(e.p, &e.c0, &e.c1);
    e.sn = e.c0;
    e.cs = e.c1;
    e.dcs = -e.c0;
    e.dsn = e.c1;
    e.ex = 0.0;
    e.eh = 1.0;
    e.x = -e.r * e.sn;
  } else {
    e.c0 = exp_neg(e.p);
    e.c1 = 0.0;
    const double fac = below16(e.p) ? e.c0 * e.c0 : 0.0;
    e.cs = 0.5 + 0.5 * fac;
    e.sn = 0.5 - 0.5 * fac;
    e.dcs = -fac;
    e.dsn = fac;
    e.ex = e.p;  // p = 0 when k == kx
    e.eh = below60(e.p) ? e.c0 : 0.0;
    e.x = e.r * e.sn;
  }
  e.w = e.pos ? e.sn * e.ir : dm;  // k == kx: w = d (reference :135)
}

ADSURF_HD void eig_finish(Eig& e, double dm) {
  if (e.lt) {
    e.sn = e.c0;
    e.cs = e.c1;
    e.dcs = -e.c0;
    e.dsn = e.c1;
    e.ex = 0.0;
    e.eh = 1.0;
  } else {
    const double fac = below16(e.p) ? e.c0 * e.c0 : 0.0;
    e.cs = 0.5 + 0.5 * fac;
    e.sn = 0.5 - 0.5 * fac;
    e.dcs = -fac;
    e.dsn = fac;
    e.ex = e.p;  // p = 0 when k == kx
    e.eh = below60(e.p) ? e.c0 : 0.0;
  }
  e.w = e.pos ? e.sn * e.ir : dm;  // k == kx: w = d (reference :135)
  e.x = (e.lt ? -e.r : e.r) * e.sn;
}

struct Layer {
  // inputs
  double gk, rho, ir, dm;
  // coefficient scalars
  double gam, g1, gg, q1, ar, br, kr, tg, t1;
  Eig P, S;
  double a0;
  int fl;  // regime flags F_LTP | F_POSP | F_LTS | F_POSS (set by the layer source)
};

// forward sweep: evaluate the transcendentals (P, S set up by the layer source).  The four regime
// combinations get their own straight-line block so that the P and the S polynomial chains (each a
// ~16-deep dependent FP64 sequence, 8 cycles per link) are scheduled interleaved instead of one
// after the other behind separate branches.  The regime is warp-uniform in the dense kernels.
ADSURF_HD void layer_set_flags(Layer& y, int fl) {
  y.P.lt = fl & F_LTP;
  y.P.pos = fl & F_POSP;
  y.S.lt = fl & F_LTS;
  y.S.pos = fl & F_POSS;
}

ADSURF_HD void layer_eig_forward(Layer& y) {
  const bool small = below1e5(fabs(y.P.p)) && below1e5(fabs(y.S.p));
  const int fl = small ? y.fl : -1;  // one integer compare per regime instead of four extracted booleans
  if (fl == (F_REGULAR | F_LTS)) {          // c between the S and the P velocity of the layer: the common case
    layer_set_flags(y, F_REGULAR | F_LTS);
    eig_evan(y.P);
    eig_osc(y.S);
  } else if (fl == F_REGULAR) {             // slower than both: fully evanescent
    layer_set_flags(y, F_REGULAR);
    eig_evan(y.P);
    eig_evan(y.S);
  } else if (fl == (F_REGULAR | F_LTP | F_LTS)) {  // faster than both
    layer_set_flags(y, F_REGULAR | F_LTP | F_LTS);
    eig_osc(y.P);
    eig_osc(y.S);
  } else if (fl == (F_REGULAR | F_LTP)) {
    layer_set_flags(y, F_REGULAR | F_LTP);
    eig_osc(y.P);
    eig_evan(y.S);
  } else {  // huge phase, NaN, or c exactly on a layer velocity: generic path
    layer_set_flags(y, y.fl);
    eig_forward(y.P, y.dm);
    eig_forward(y.S, y.dm);
    y.a0 = below60(y.P.ex + y.S.ex) ? y.P.eh * y.S.eh : 0.0;
    return;
  }
  y.P.w = y.P.sn * y.P.ir;
  y.S.w = y.S.sn * y.S.ir;
  y.a0 = below60(y.P.ex + y.S.ex) ? y.P.eh * y.S.eh : 0.0;
}

// adjoint sweep: restore them from the store (c0, c1 already set by the caller)
ADSURF_HD void layer_eig_restore(Layer& y) {
  eig_finish(y.P, y.dm);
  eig_finish(y.S, y.dm);
  y.a0 = below60(y.P.ex + y.S.ex) ? y.P.eh * y.S.eh : 0.0;
}

// coefficient scalars from (gk, k2); t2 = -2 k2 (hoisted per sample)
ADSURF_HD void layer_coefs(Layer& y, double k2, double t2) {
  y.gam = y.gk * k2;
  y.g1 = y.gam - 1.0;
  y.gg = y.gam * y.gk;
  y.q1 = y.g1 * y.g1;
  y.ar = y.gg * y.rho;
  y.br = y.q1 * y.rho;
  y.kr = k2 * y.ir;
  y.tg = t2 * y.gk;
  y.t1 = t2 * y.g1;
}

// E <- E . CA   (row vector times Dunkin matrix, reference :376-383)
ADSURF_HD void row_apply(double (&e)[5], const Layer& y) {
  const double e0 = e[0], e1 = e[1], e2 = e[2], e3 = e[3], e4 = e[4];
  const double cp = y.P.cs, w = y.P.w, x = y.P.x, cq = y.S.cs, yy = y.S.w, z = y.S.x;
  const double sa = e0 * y.ir + e2 * y.tg + e4 * y.ar;
  const double sb = e0 * y.kr + e2 * y.t1 + e4 * y.br;
  const double h1 = sb * yy + e1 * cq;
  const double h2 = sa * cq + e3 * yy;
  const double h3 = sa * z + e3 * cq;
  const double h4 = sb * cq + e1 * z;
  const double Ca = cp * h4 - x * h3 - sb * y.a0;
  const double Cb = cp * h2 - w * h1 - sa * y.a0;
  e[1] = cp * h1 - x * h2;
  e[3] = cp * h3 - w * h4;
  e[0] = y.a0 * e0 + y.ar * Ca + y.br * Cb;
  e[2] = y.a0 * e2 + y.gk * Ca + y.g1 * Cb;
  e[4] = y.a0 * e4 + y.ir * Ca + y.kr * Cb;
}

// s <- CA . s   (column vector of the layers above)
ADSURF_HD void col_apply(double (&s)[5], const Layer& y) {
  const double s0 = s[0], s1 = s[1], s2 = s[2], s3 = s[3], s4 = s[4];
  const double cp = y.P.cs, w = y.P.w, x = y.P.x, cq = y.S.cs, yy = y.S.w, z = y.S.x;
  const double ta = y.ar * s0 + y.gk * s2 + y.ir * s4;
  const double tb = y.br * s0 + y.g1 * s2 + y.kr * s4;
  const double m1 = z * ta + cq * s1;
  const double m2 = cq * tb + z * s3;
  const double m3 = yy * tb + cq * s3;
  const double m4 = cq * ta + yy * s1;
  const double Da = cp * m2 - x * m1 - y.a0 * tb;
  const double Db = cp * m4 - w * m3 - y.a0 * ta;
  s[1] = cp * m1 - w * m2;
  s[3] = cp * m3 - x * m4;
  s[0] = y.a0 * s0 + y.ir * Da + y.kr * Db;
  s[2] = y.a0 * s2 + y.tg * Da + y.t1 * Db;
  s[4] = y.a0 * s4 + y.ar * Da + y.br * Db;
}

struct LayerGrad {
  // dDelta/d(thickness, gammk, rho, k) and, for each wave type, h = (dDelta/dr) / r * sign(k - kx):
  // dDelta/dk += k h, dDelta/dkx = -kx h  =>  dDelta/dalpha = omega^2 / alpha^3 * hp (kx = omega/alpha).
  double d, hp, hs, gammk, rho, k;
};

// adjoint of one eigenfunction block: (bcs, bw, bx, bex) -> bd += ..., bk += k h, returns h.  Both branches
// share one formula through dcs = d cs/dp, dsn = d sn/dp and the sign of (k - kx):
//   w = sn/r, x = sg r sn  =>  d/dr = (x bx - w bw)/r ;  d u/d k = 2 sg k, d r/d u = 1/(2 r).
ADSURF_HD double eig_adjoint(const Eig& e, double k, double dm, double bcs, double bw, double bx, double bex,
                             double& bd, double& bk) {
  const double bsn = bw * e.ir + (e.lt ? -e.r : e.r) * bx;
  const double bp = e.dcs * bcs + e.dsn * bsn + (e.lt ? 0.0 : bex);
  const double br = e.ir * (e.x * bx - e.w * bw) + dm * bp;
  const double h = (e.lt ? -e.ir : e.ir) * br;
  // k == kx: w = d and the (measure-zero) dependence through r is dropped (finite one-sided value)
  bd += e.pos ? e.r * bp : bw;
  bk += e.pos ? k * h : 0.0;
  return e.pos ? h : 0.0;
}

// Reverse sweep through one layer: given row vector e (= E_{m+1}) and column s (= s_m, already
// scaled by the upstream seed) returns dDelta/d(layer inputs), Delta = e . CA . s.
ADSURF_HD LayerGrad layer_adjoint(const Layer& y, const double (&e)[5], const double (&s)[5],
                                                   double k, double k2, double t2) {
  const double e0 = e[0], e1 = e[1], e2 = e[2], e3 = e[3], e4 = e[4];
  const double s0 = s[0], s1 = s[1], s2 = s[2], s3 = s[3], s4 = s[4];
  const double cp = y.P.cs, w = y.P.w, x = y.P.x, cq = y.S.cs, yy = y.S.w, z = y.S.x;
  const double a0 = y.a0;
  // the four contractions with the coefficient vectors
  const double sa = e0 * y.ir + e2 * y.tg + e4 * y.ar;
  const double sb = e0 * y.kr + e2 * y.t1 + e4 * y.br;
  const double ta = y.ar * s0 + y.gk * s2 + y.ir * s4;
  const double tb = y.br * s0 + y.g1 * s2 + y.kr * s4;
  // S-side combinations of the row vector, P-side combinations of the column vector
  const double h1 = sb * yy + e1 * cq;
  const double h2 = sa * cq + e3 * yy;
  const double h3 = sa * z + e3 * cq;
  const double h4 = sb * cq + e1 * z;
  const double g1_ = cp * s1 - w * tb;
  const double g2_ = cp * tb - x * s1;
  const double g3_ = cp * s3 - x * ta;
  const double g4_ = cp * ta - w * s3;
  // adjoints of the eigenfunction terms
  const double ba0 = e0 * s0 + e2 * s2 + e4 * s4 - sb * ta - sa * tb;
  const double bcp = h4 * ta + h2 * tb + h1 * s1 + h3 * s3;
  const double bx = -(h3 * ta + h2 * s1);
  const double bw = -(h1 * tb + h4 * s3);
  const double by = sb * g1_ + e3 * g2_;
  const double bz = sa * g3_ + e1 * g4_;
  const double bcq = e1 * g1_ + sa * g2_ + e3 * g3_ + sb * g4_;
  // adjoints of the contractions
  const double Dsa = cq * g2_ + z * g3_ - a0 * tb;
  const double Dsb = yy * g1_ + cq * g4_ - a0 * ta;
  const double Ca = cp * h4 - x * h3 - sb * a0;
  const double Cb = cp * h2 - w * h1 - sa * a0;
  // coefficient scalars
  const double b_ir = Dsa * e0 + Ca * s4;
  const double b_tg = Dsa * e2;
  const double b_ar = Dsa * e4 + Ca * s0;
  const double b_kr = Dsb * e0 + Cb * s4;
  const double b_t1 = Dsb * e2;
  const double b_br = Dsb * e4 + Cb * s0;
  const double b_gg = y.rho * b_ar;
  const double b_q1 = y.rho * b_br;
  const double b_g1 = Cb * s2 + t2 * b_t1 + 2.0 * y.g1 * b_q1;
  const double b_gam = b_g1 + y.gk * b_gg;
  const double b_gk = Ca * s2 + t2 * b_tg + y.gam * b_gg + k2 * b_gam;
  const double b_k2 = y.ir * b_kr + y.gk * b_gam - 2.0 * (y.gk * b_tg + y.g1 * b_t1);
  const double b_rho = y.gg * b_ar + y.q1 * b_br - y.ir * y.ir * (b_ir + k2 * b_kr);
  const double bex = -a0 * ba0;

  LayerGrad g;
  g.d = 0.0;
  g.k = 2.0 * k * b_k2;
  g.gammk = b_gk;
  g.rho = b_rho;
  g.hp = eig_adjoint(y.P, k, y.dm, bcp, bw, bx, bex, g.d, g.k);
  g.hs = eig_adjoint(y.S, k, y.dm, bcq, by, bz, bex, g.d, g.k);
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
  g.hp = 0.0; g.hs = 0.0;
  if (h.bra < 2) {
    g.hp = (h.bra ? 1.0 : -1.0) * bra * h.ira;
    g.k += k * g.hp;
  }
  if (h.brb < 2) {
    g.hs = (h.brb ? 1.0 : -1.0) * brb * h.irb;
    g.k += k * g.hs;
  }
  return g;
}

// ------------------------------------------------------------------------------------------
// Sample-level drivers
// ------------------------------------------------------------------------------------------
struct Sample {
  double k, k2, omega, iw, iw2, t2, w2;  // wavenumber, k^2, clamped angular frequency, 1/omega, 1/omega^2, -2 k^2, omega^2
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
  z.iw = 1.0 / z.omega;
  z.iw2 = z.iw * z.iw;
  z.t2 = -2.0 * z.k2;
  z.w2 = z.omega * z.omega;
  return z;
}

// ---- layer sources: fill the inputs, the eigenfunction set-up and the coefficient scalars of layer m
// DirectSource: everything per sample from (k, omega) — pointwise kernels, arbitrary (t, c) pairs.
struct DirectSource {
  static constexpr bool kUniformFlags = false;  // per-sample (t, c): regimes may differ between lanes
  const double* __restrict__ par;
  int L;
  ADSURF_HD void prep(Layer& y, int m, const Sample& z) const {
    y.gk = par[m * kNumPar + P_TB2] * z.iw2;
    y.rho = par[m * kNumPar + P_RHO];
    y.ir = par[m * kNumPar + P_IR];
    y.dm = par[m * kNumPar + P_D];
    eig_setup(y.P, z.k, z.omega * par[m * kNumPar + P_IA], y.dm);
    eig_setup(y.S, z.k, z.omega * par[m * kNumPar + P_IB], y.dm);
    y.fl = (y.P.lt ? F_LTP : 0) | (y.P.pos ? F_POSP : 0) | (y.S.lt ? F_LTS : 0) | (y.S.pos ? F_POSS : 0);
    layer_coefs(y, z.k2, z.t2);
  }
};

// TableSource: dense kernels.  Everything that depends on (trial velocity, layer) but not on the
// period is computed once per block into a shared-memory table and read as a broadcast:
//   r = sqrt(|k^2 - kx^2|) = omega * n,  n = sqrt(|1/c^2 - 1/v^2|);   p = omega * (n d);
//   gam = gk k^2 = 2 (beta/c)^2;  g1 = gam - 1;  br = g1^2 rho;   branch flags (c vs v).
// Per lane that leaves 6 multiplies for both eigenfunction set-ups (instead of 2 rsqrt sequences,
// the differences of squares and their selects) and 6 for the coefficient scalars.
constexpr int kTab = 10;  // 80-byte entries: five 16-byte loads; the last slot carries the flag bits
enum { T_PDA = 0, T_PDB = 1, T_NA = 2, T_NB = 3, T_INA = 4, T_INB = 5, T_GAM = 6, T_G1 = 7, T_BR = 8, T_FL = 9 };
struct alignas(16) Dbl2 {
  double x, y;
};
ADSURF_HD Dbl2 ld2(const double* p) { return *reinterpret_cast<const Dbl2*>(p); }

// one table entry: trial slowness ic = 1/c against layer m of the staged model
ADSURF_HD void table_entry(double* __restrict__ tab, const double* __restrict__ par, int L, int m, double ic) {
  const double sa = par[m * kNumPar + P_IA], sb = par[m * kNumPar + P_IB], dm = par[m * kNumPar + P_D];
  const double da = (ic + sa) * fabs(ic - sa), db = (ic + sb) * fabs(ic - sb);
  const double na = sqrt(da), nb = sqrt(db);
  tab[T_PDA] = na * dm;
  tab[T_PDB] = nb * dm;
  tab[T_NA] = na;
  tab[T_NB] = nb;
  tab[T_INA] = 1.0 / na;
  tab[T_INB] = 1.0 / nb;
  const double gam = par[m * kNumPar + P_TB2] * (ic * ic);
  const double g1 = gam - 1.0;
  tab[T_GAM] = gam;
  tab[T_G1] = g1;
  tab[T_BR] = g1 * g1 * par[m * kNumPar + P_RHO];
  const int fl = (ic < sa ? F_LTP : 0) | (da > 0.0 ? F_POSP : 0) | (ic < sb ? F_LTS : 0) | (db > 0.0 ? F_POSS : 0);
  tab[T_FL] = bits_dbl((long long)fl);  // raw bits, read back with dbl_bits()
}

struct TableSource {
  static constexpr bool kUniformFlags = true;   // flags come from the shared (velocity, layer) table
  const double* __restrict__ par;
  int L;
  const double* __restrict__ tab;  // [L-1][kTab] for the current trial velocity (16-byte aligned)
  ADSURF_HD void prep(Layer& y, int m, const Sample& z) const {
    // 16-byte broadcast loads: 2 for the layer record, 5 for the table entry
    const double* __restrict__ pr = par + m * kNumPar;
    const Dbl2 p23 = ld2(pr + P_TB2), p45 = ld2(pr + P_IR);
    const double* __restrict__ q = tab + m * kTab;
    const Dbl2 t01 = ld2(q + T_PDA), t23 = ld2(q + T_NA), t45 = ld2(q + T_INA), t67 = ld2(q + T_GAM),
               t89 = ld2(q + T_BR);
    const int fl = (int)dbl_bits(t89.y);
    y.rho = p23.y;
    y.ir = p45.x;
    y.dm = p45.y;
    y.gk = p23.x * z.iw2;
    y.gam = t67.x;
    y.g1 = t67.y;
    y.br = t89.x;
    y.gg = y.gam * y.gk;
    y.q1 = y.g1 * y.g1;      // adjoint only
    y.ar = y.gg * y.rho;
    y.kr = z.k2 * y.ir;
    y.tg = z.t2 * y.gk;
    y.t1 = z.t2 * y.g1;
    y.P.r = z.omega * t23.x;
    y.P.ir = z.iw * t45.x;
    y.P.p = z.omega * t01.x;
    y.S.r = z.omega * t23.y;
    y.S.ir = z.iw * t45.y;
    y.S.p = z.omega * t01.y;
    y.fl = fl;  // the per-eig booleans are set where a regime is selected (layer_set_flags)
  }
};

// The half-space vector is linear in ra, rb = sqrt((k+kx)|k-kx|): at a trial velocity that coincides
// with the half-space velocity the result depends on the last bit of kx, so kx is formed with the
// reference's own correctly-rounded division omega/alpha (reference :328-329) from the raw
// velocities kept at par[kNumPar*L + 0/1].  Inside the layers the dependence on kx is smooth
// (even in r), and the cheaper omega * (1/alpha) is used.
ADSURF_HD void load_half(Half& h, const double* __restrict__ par, int L, const Sample& z) {
  const int m = L - 1;
  h.ka = z.omega / par[kNumPar * L + 0];
  h.kb = z.omega / par[kNumPar * L + 1];
  h.gammk = par[m * kNumPar + P_TB2] * z.iw2;
  h.rho = par[m * kNumPar + P_RHO];
}

// forward only
template <typename Source>
ADSURF_HD double sample_forward(const Source& src, const Sample& z) {
  const double* __restrict__ par = src.par;
  const int L = src.L;
  double e[5];
  Half h;
  load_half(h, par, L, z);
  half_space(h, z.k, z.k2, e);
  Layer y;
  for (int m = L - 2; m >= 0; --m) {
    src.prep(y, m, z);
    layer_eig_forward(y);
    row_apply(e, y);
  }
  return e[0];
}

// Storage policy for what the forward sweep keeps for the adjoint sweep, kSlot = 9 doubles per layer:
// the row vector E_{m+1} (5) and the transcendental results of the P and S terms (2 + 2).  Element
// (m, j) of this thread lives at p[(m*kSlot+j)*kStride] (kStride = 32: lane-interleaved, i.e.
// conflict-free in shared memory and coalesced in global memory; kStride = 1 on the host).
constexpr int kSlot = 9;

template <int kStride>
struct PlainStore {
  double* p;
  typedef double* Slot;
  ADSURF_HD Slot slot(int m) const { return p + (size_t)m * (kSlot * kStride); }
  ADSURF_HD static double ld(Slot q, int j) { return q[j * kStride]; }
  ADSURF_HD static void st(Slot q, int j, double v) { q[j * kStride] = v; }
};

// forward keeping E_{m+1} of every layer: slot m holds the row vector that multiplies CA_m.
template <typename Source, typename Store>
ADSURF_HD double sample_forward_store(const Source& src, const Sample& z, const Store& est) {
  const double* __restrict__ par = src.par;
  const int L = src.L;
  double e[5];
  Half h;
  load_half(h, par, L, z);
  half_space(h, z.k, z.k2, e);
  Layer y;
  for (int m = L - 2; m >= 0; --m) {
    src.prep(y, m, z);
    layer_eig_forward(y);
    const typename Store::Slot q = est.slot(m);
#pragma unroll
    for (int i = 0; i < 5; ++i) Store::st(q, i, e[i]);
    Store::st(q, 5, y.P.c0);
    Store::st(q, 6, y.P.c1);
    Store::st(q, 7, y.S.c0);
    Store::st(q, 8, y.S.c1);
    row_apply(e, y);
  }
  return e[0];
}

// Top-down adjoint sweep.  s_0 = seed * e_0; for every layer m (and finally the half-space, m = L-1)
// hands (m, dDelta/dd_m, /dalpha_m, /dbeta_m, /drho_m) * seed to `sink`.  Returns dDelta/dk * seed.
// The stored values of the next layer are fetched one iteration ahead into the other of two
// register buffers (the loop is unrolled by two so the buffers swap without copies).
template <typename Source, typename Store, typename Sink>
ADSURF_HD double sample_adjoint(const Source& src, const Sample& z, double seed, const Store& est, Sink& sink) {
  const double* __restrict__ par = src.par;
  const int L = src.L;
  double s[5] = {seed, 0.0, 0.0, 0.0, 0.0};
  // software pipeline over the stored values: each group of registers is re-loaded for layer m+1 right
  // after its last use in layer m — the transcendental results once eig_finish has consumed them, the row
  // vector E once layer_adjoint is done — so the L2 latency of the slab hides behind the rest of the layer
  // (col_apply, the gradient sink, the next layer's set-up) without a second register buffer.
  double e[5], cv[4];
  double gk = 0.0;
  auto fetch_c = [&](int m) {
    const typename Store::Slot q = est.slot(m);
#pragma unroll
    for (int i = 0; i < 4; ++i) cv[i] = Store::ld(q, 5 + i);
  };
  auto fetch_e = [&](int m) {
    const typename Store::Slot q = est.slot(m);
#pragma unroll
    for (int i = 0; i < 5; ++i) e[i] = Store::ld(q, i);
  };
  auto layer = [&](int m) {
    Layer y;
    y.P.c0 = cv[0]; y.P.c1 = cv[1]; y.S.c0 = cv[2]; y.S.c1 = cv[3];
    src.prep(y, m, z);
    auto body = [&]() {
      layer_eig_restore(y);
      if (m + 1 <= L - 2) fetch_c(m + 1);
      const LayerGrad g = layer_adjoint(y, e, s, z.k, z.k2, z.t2);
      if (m + 1 <= L - 2) fetch_e(m + 1);
      col_apply(s, y);
      gk += g.k;
      // k_alpha = omega/alpha, k_beta = omega/beta, gammk = 2 beta^2 / omega^2:
      //   d/dalpha = omega^2/alpha^3 hp,  d/dbeta = omega^2/beta^3 hs + 4 beta/omega^2 dDelta/dgammk
      const double* __restrict__ pr = par + m * kNumPar;
      sink(m, g.d, (z.w2 * pr[P_IA3]) * g.hp, (z.w2 * pr[P_IB3]) * g.hs + (pr[P_CB] * z.iw2) * g.gammk, g.rho);
    };
    // Regime specialisation (dense kernels: the flags are warp-uniform table entries).  Re-stating the flags
    // as constants inside each branch lets the compiler fold every "lt ? a : b" / "pos ? a : b" select of
    // eig_finish / eig_adjoint away in the two regimes that cover ~95 % of (velocity, layer) pairs.
    if (Source::kUniformFlags && y.fl == (F_REGULAR | F_LTS)) {  // P evanescent, S oscillatory
      layer_set_flags(y, F_REGULAR | F_LTS);
      body();
    } else if (Source::kUniformFlags && y.fl == F_REGULAR) {     // both evanescent
      layer_set_flags(y, F_REGULAR);
      body();
    } else {
      layer_set_flags(y, y.fl);
      body();
    }
  };
  if (L > 1) {
    fetch_c(0);
    fetch_e(0);
  }
  for (int m = 0; m <= L - 2; ++m) layer(m);
  // half-space: dDelta/dE_hs = s_{L-1}
  Half h;
  load_half(h, par, L, z);
  half_space(h, z.k, z.k2, e);
  const LayerGrad g = half_adjoint(h, e, s, z.k, z.k2);
  gk += g.k;
  const int mh = L - 1;
  const double* __restrict__ prh = par + mh * kNumPar;
  sink(mh, 0.0, (z.w2 * prh[P_IA3]) * g.hp, (z.w2 * prh[P_IB3]) * g.hs + (prh[P_CB] * z.iw2) * g.gammk, g.rho);
  return gk;
}

// per-model layer constants staged for the kernels: par[L][kNumPar] (one 48-byte record per layer,
// so a layer needs one address computation), then the raw half-space velocities
ADSURF_HD void stage_one(double* __restrict__ par, int L, int m, double dm, double a, double b, double r) {
  par[m * kNumPar + P_IA] = 1.0 / a;
  par[m * kNumPar + P_IB] = 1.0 / b;
  par[m * kNumPar + P_TB2] = 2.0 * b * b;
  par[m * kNumPar + P_RHO] = r;
  par[m * kNumPar + P_IR] = 1.0 / r;
  par[m * kNumPar + P_D] = dm;
  par[m * kNumPar + P_IA3] = 1.0 / (a * a * a);
  par[m * kNumPar + P_IB3] = 1.0 / (b * b * b);
  par[m * kNumPar + P_CB] = 4.0 * b;
  par[m * kNumPar + 9] = 0.0;
  if (m == L - 1) {
    par[kNumPar * L + 0] = a;
    par[kNumPar * L + 1] = b;
  }
}

// doubles of staged constants per model
ADSURF_HD int par_doubles(int L) { return kNumPar * L + 2; }

}  // namespace adsurf
