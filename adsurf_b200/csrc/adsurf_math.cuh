// adsurf_math.cuh — per-sample arithmetic of the Rayleigh secular function and its hand-derived
// adjoint.  Shared by the CUDA kernels (adsurf_kernels.cu) and by the host-side math check that
// tests/ builds from this same header (tests/host_math/host_math.cpp) — the check compiles these
// functions for the CPU so the adjoint algebra can be verified against the oracle without a GPU.
// The product library never runs this code on the host.
//
// Specification: SURVEY.md Appendix A == /root/reference/ADsurf/_cps/_surf96_vectorAll_gpu.py
// :52-100 (Dunkin matrix), :102-192 (eigenfunction terms), :303-399 (recursion).
#pragma once
#include <math.h>
#include <string.h>

#include <type_traits>

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
// 1/alpha, 1/beta, 2 beta^2, rho, 1/rho, d | adjoint only: 1/alpha^3, 1/beta^3, 4 beta | 2 beta^2 rho
enum { P_IA = 0, P_IB = 1, P_TB2 = 2, P_RHO = 3, P_IR = 4, P_D = 5, P_IA3 = 6, P_IB3 = 7, P_CB = 8, P_GKR = 9 };

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
// v with its sign flipped when bit 31 of m is set
ADSURF_HD double flip_sign(double v, int m) {
#if defined(__CUDA_ARCH__)
  return __hiloint2double(__double2hiint(v) ^ (m & (int)0x80000000), __double2loint(v));
#else
  return bits_dbl(dbl_bits(v) ^ ((long long)(m & (int)0x80000000) << 32));
#endif
}
// v >= 0: min(v, T) for a threshold given by its high word (low word of the result is v's: |error| < 1 ulp(hi))
ADSURF_HD double clamp_hi(double v, int thi) {
#if defined(__CUDA_ARCH__)
  return __hiloint2double(min(__double2hiint(v), thi), __double2loint(v));
#else
  const long long b = dbl_bits(v);
  const int hi = (int)(b >> 32);
  return bits_dbl(((long long)(hi < thi ? hi : thi) << 32) | (b & 0xffffffffLL));
#endif
}
ADSURF_HD bool below16(double v) { return hi_word(v) < 0x40300000; }
ADSURF_HD bool below60(double v) { return hi_word(v) < 0x404E0000; }
ADSURF_HD bool below1e5(double v) { return hi_word(v) < 0x40F86A00; }

// v/2 and a+b as separate instructions: "0.5 + 0.5 v" contracted to one fma needs 0.5 twice, and an FP64
// instruction encodes only one immediate, so the second copy costs two moves per use
ADSURF_HD double half_of(double v) {
#if defined(__CUDA_ARCH__)
  return __dmul_rn(v, 0.5);
#else
  return v * 0.5;
#endif
}
ADSURF_HD double add_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dadd_rn(a, b);
#else
  return a + b;
#endif
}

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
#define ADSURF_EXP_CONSTS { 5.952044687368025e-11, 0.041666666666666664, 0.16666666666666666, 0.0 } /* -(ln2/32 tail), 1/24, 1/6 */
static const double kExpCH[4] = ADSURF_EXP_CONSTS;
static const double kSinCH[6] = ADSURF_SIN_COEFS;
static const double kCosCH[6] = ADSURF_COS_COEFS;
static const double kRedCH[7] = ADSURF_RED_CONSTS;
#if defined(__CUDACC__)
__constant__ double kExpCD[4] = ADSURF_EXP_CONSTS;
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
constexpr int kSmemHead = 40;  // 32 table entries + the constants of ADSURF_HC
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
// constants that share an FMA with another constant: an FP64 instruction takes one immediate / constant-bank operand,
// the second one would be rebuilt in a register pair (two moves) at every use; read from the shared-memory head it
// costs one load, or nothing where the compiler keeps it in registers across the layer loop
__constant__ double kHeadConstD[8] = {6755399441055744.0, 0.00138888880610466, -0.5, 0.0, 0.0, 0.0, 0.0, 0.0};
extern __shared__ double adsurf_smem_head[];
__device__ __forceinline__ void stage_exp_table() {
  for (int j = threadIdx.x; j < 32; j += blockDim.x) adsurf_smem_head[j] = kExp2TabD[j];
  if (threadIdx.x < 8) adsurf_smem_head[32 + threadIdx.x] = kHeadConstD[threadIdx.x];
}
#endif
#if defined(__CUDA_ARCH__)
#define ADSURF_EXP2_TAB(j) adsurf_smem_head[j]
#define ADSURF_HC(i, literal) adsurf_smem_head[32 + (i)]
#else
#define ADSURF_EXP2_TAB(j) kExp2TabH[j]
#define ADSURF_HC(i, literal) (literal)
#endif

// exp(-p) for 0 <= p < ~700: n = rint(-p 32/ln2), r = -p - n ln2/32 (two-term, |r| <= ln2/64),
// e^r by a degree-6 Taylor polynomial (truncation 3.5e-18), times 2^(j/32) from the table (j = n mod 32)
// scaled by 2^(n div 32) through the exponent field.  11 FP64 instructions, <= 2 ulp.
// Constants whose exact value does not matter are rounded to 21 significant bits, i.e. to doubles whose low
// word is zero, which an FP64 instruction encodes as an immediate (a full 64-bit literal costs two extra
// instructions): 32/ln2 (only picks n), the head of ln2/32 (the tail carries the rest: head + tail = ln2/32 to
// 2^-75) and the two highest Taylor coefficients (their rounding moves the result by < 1e-18).
// The reference sets e^{-(pP+pS)} to 0 beyond 60 (:170-178), a cut-off that removes a factor < 1e-26; here the
// exponential is simply evaluated (the argument is clamped at 700), which saves a compare and two selects per
// layer in both sweeps.  (Its other cut-off, e^{-2p} := 0 beyond p = 16, is kept: see evan_terms.)
ADSURF_HD double exp_neg(double p_in) {
  // p >= 0 (or NaN): doubles order like their high words, so an integer min on the high word clamps p to ~700
  // (e^-700 ~ 1e-304: its square and its products underflow to 0) in one ALU instruction; beyond ~708 the exponent arithmetic below would wrap.
  const double p = clamp_hi(p_in, 0x4085E000);
  const double x = -p;
  const double t = fma(x, 46.166229248046875, ADSURF_HC(0, kShifter));
  const double fn = t - kShifter;
  double r = fma(fn, -0.021660849452018738, x);
  r = fma(fn, ADSURF_K(kExpC, 0), r);
  const int n = (int)dbl_bits(t);  // low 32 bits hold the (two's complement) integer
  double q = fma(ADSURF_HC(1, 0.00138888880610466), r, 0.00833333283662796);
  q = fma(q, r, ADSURF_K(kExpC, 1));
  q = fma(q, r, ADSURF_K(kExpC, 2));
  q = fma(q, r, 0.5);
  q = fma(q, r, 1.0);
  q = fma(q, r, 1.0);
  // 2^(n div 32) goes into the exponent field: an add on the high word only (the 64-bit form costs a second add)
  const double tab = ADSURF_EXP2_TAB(n & 31);
#if defined(__CUDA_ARCH__)
  const double tj = __hiloint2double(__double2hiint(tab) + ((n >> 5) << 20), __double2loint(tab));
#else
  const double tj = bits_dbl((long long)((unsigned long long)dbl_bits(tab) + ((unsigned long long)(long long)(n >> 5) << 52)));
#endif
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
  const double t = fma(q, ADSURF_K(kRedC, 3), ADSURF_HC(0, kShifter));
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
  const double c = fma(z * z, pc, fma(ADSURF_HC(2, -0.5), z, 1.0));
  const int n = (int)dbl_bits(t);
  const double a = (n & 1) ? c : s;
  const double b = (n & 1) ? s : c;
  // quadrant signs: flip the sign bit with integer operations (a negation is an FP64-pipe instruction)
  *sn = flip_sign(a, n << 30);
  *cs = flip_sign(b, (n + 1) << 30);
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
//
// Frequency scaling.  Every wavenumber in CA is proportional to omega (k = omega/c, k_v = omega/v,
// r = omega n with n = sqrt(|1/c^2 - 1/v^2|), gammk = 2 beta^2/omega^2), so
//   CA(omega) = D^-1 CA~ D,   D = diag(1, omega, 1, omega, omega^2),
// where CA~ is the same matrix built from the frequency-free quantities (slowness 1/c for k, n for r,
// 2 beta^2 for gammk) and depends on omega only through the phases p = omega n d.  D is the same for all
// layers, so the product telescopes: the layer sweep runs on E~ = E D^-1 (and s~ = D s) with coefficient
// scalars that depend on (trial velocity, layer) only — in the dense kernels they come out of a shared
// table and a lane spends two multiplies (the two phases) on a layer's set-up instead of thirteen.  The
// half-space vector is formed as the reference does and converted once per sample.
// A second similarity, CA = R^-1 CA^ R with R = diag(rho^2, rho, rho, rho, 1) and CA^ free of rho, gives
// the density gradient of a layer from the vectors on either side of it:
//   rho dDelta/drho = (E'_0 s_0 - E_0 s'_0) - (E'_4 s_4 - E_4 s'_4),   E' = E.CA, s' = CA.s.
// regime flags of a (velocity, layer) pair: oscillatory (k < kx) and regular (k != kx) for P and S
enum { F_LTP = 1, F_POSP = 2, F_LTS = 4, F_POSS = 8, F_REGULAR = F_POSP | F_POSS,
       F_END = 16 };  // F_END: sentinel entry of the shared table (no layer), see TableSource

struct Eig {
  double r, ir;     // n = sqrt(|1/c^2 - 1/v^2|) and its reciprocal (frequency-free)
  double p;         // phase omega n d
  double cs, sn;    // cos p / (1+e^-2p)/2   and   sin p / (1-e^-2p)/2
  double w, x;      // sn / r   and   -+ r sn
  double dcs, dsn;  // d cs / dp, d sn / dp
  double c0, c1;    // the transcendental results kept for the adjoint sweep: (sin, cos) or (e^{-p}, -)
  bool lt, pos;     // oscillatory branch (k < kx);  u > 0 (false: k == kx or unordered)
};

// var_vector, reference :109-136 (P) / :140-167 (S).  The transcendental values (sin/cos or e^{-p}) are
// evaluated in the forward sweep only and stored; eig_finish re-derives everything else from them.
// e^{-2p} is formed as (e^{-p})^2 and a0 = e^{-(pex+sex)} as e^{-pex} e^{-sex} (thresholds kept).
// branch-free bodies of the two regimes (oscillatory valid for |p| < 1e5)
ADSURF_HD void eig_osc(Eig& e) {
  sincos_core(e.p, &e.c0, &e.c1);
  e.sn = e.c0;
  e.cs = e.c1;
  e.dcs = -e.c0;
  e.dsn = e.c1;
  e.x = -e.r * e.sn;
}

// (1 +- e^{-2p})/2 from e^{-p}; the e^{-2p} term is dropped beyond p = 16 as the reference does (:128-134)
ADSURF_HD void evan_terms(Eig& e) {
  const double fac = below16(e.p) ? e.c0 * e.c0 : 0.0;  // cut-off kept: it shapes d/d(thickness) at the 1e-8 level
  const double hf = half_of(fac);
  e.cs = add_rn(0.5, hf);
  e.sn = add_rn(0.5, -hf);
  e.dcs = -fac;
  e.dsn = fac;
}

ADSURF_HD void eig_evan(Eig& e) {
  e.c0 = exp_neg(e.p);
  e.c1 = 0.0;
  evan_terms(e);
  e.x = e.r * e.sn;
}

// a0 = e^{-(pex + sex)} (reference :170-178) from the stored e^{-p} of the evanescent terms
ADSURF_HD double layer_a0(const Eig& P, const Eig& S) { return (P.lt ? 1.0 : P.c0) * (S.lt ? 1.0 : S.c0); }

// generic forward: regime from the flags; wd = omega * thickness is the value of w when k == kx (reference :135)
ADSURF_HD void eig_forward(Eig& e, double wd) {
  if (e.lt) {
    sincos_fast(e.p, &e.c0, &e.c1);
    e.sn = e.c0;
    e.cs = e.c1;
    e.dcs = -e.c0;
    e.dsn = e.c1;
      e.x = -e.r * e.sn;
  } else {
    eig_evan(e);
  }
  e.w = e.pos ? e.sn * e.ir : wd;
}

// adjoint sweep: everything derived from the stored (c0, c1)
ADSURF_HD void eig_finish(Eig& e, double wd) {
  if (e.lt) {
    e.sn = e.c0;
    e.cs = e.c1;
    e.dcs = -e.c0;
    e.dsn = e.c1;
    } else {
    evan_terms(e);
  }
  e.w = e.pos ? e.sn * e.ir : wd;
  e.x = (e.lt ? -e.r : e.r) * e.sn;
}

struct Layer {
  double ir, tg, ar, kr, t1, br, gk, g1;  // coefficient scalars: a' = (ir, tg, ar), b' = (kr, t1, br), a = (ar, gk, ir), b = (br, g1, kr)
  double c1, c3, tqa, tqb, cb;            // adjoint chain constants (table_entry)
  double rho;                             // pointwise kernels only (d/dc)
  const double* pr;                       // the layer's staged record (thickness for the k == kx case)
  Eig P, S;
  double a0;
  int fl;  // regime flags F_LTP | F_POSP | F_LTS | F_POSS (set by the layer source)
};

ADSURF_HD void layer_set_flags(Layer& y, int fl) {
  y.P.lt = fl & F_LTP;
  y.P.pos = fl & F_POSP;
  y.S.lt = fl & F_LTS;
  y.S.pos = fl & F_POSS;
}

// adjoint sweep: restore them from the store (c0, c1 already set by the caller)
ADSURF_HD void layer_eig_restore(Layer& y, double omega) {
  const double wd = omega * y.pr[P_D];  // only read when k == kx
  eig_finish(y.P, wd);
  eig_finish(y.S, wd);
  y.a0 = layer_a0(y.P, y.S);
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

// what one layer contributes to the gradient: dDelta/d(thickness, alpha, beta, rho) and, for the pointwise
// kernels, dDelta/d(1/c) through this layer
struct LayerGrad {
  double d, a, b, rho, ic;
};

// adjoint of one eigenfunction block: (bcs, bw, bx, bex) -> bp = dDelta/dp and the return value
// t = r dDelta/dr = x bx - w bw + p bp   (w = sn/r, x = -+ r sn, p = omega r d)
ADSURF_HD double eig_adjoint(const Eig& e, double bcs, double bw, double bx, double bex, double& bp) {
  const double bsn = bw * e.ir + (e.lt ? -e.r : e.r) * bx;
  bp = fma(e.dsn, bsn, e.lt ? e.dcs * bcs : fma(e.dcs, bcs, bex));  // two FMAs when bex takes part
  return e.x * bx - e.w * bw + e.p * bp;
}

// Reverse sweep through one layer: row vector e (= E_{m+1}), column s (= s_m, already scaled by the upstream
// seed).  Returns dDelta/d(layer inputs), Delta = e . CA . s, and advances s <- CA . s on the way (the column
// product shares the contractions ta, tb, g1..g4, Da, Db with the adjoint: 13 extra FMAs instead of 33).
// tq carries T_m = E_m[0] s_m[0] - E_m[4] s_m[4] of the interface above the layer in and T_{m+1} out:
// rho dDelta/drho = T_m - T_{m+1} (the rho similarity above), T_0 = seed * determinant.
template <bool kWantIc>
ADSURF_HD LayerGrad layer_adjoint(const Layer& y, const double (&e)[5], double (&s)[5], double& tq, double omega,
                                  double ic, double ic2, double t2) {
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
  // adjoints of the contractions: Da = dDelta/dsa = the a-part of CA.s, Ca = dDelta/dta = the a-part of e.CA
  const double Da = cq * g2_ + z * g3_ - a0 * tb;
  const double Db = yy * g1_ + cq * g4_ - a0 * ta;
  const double Ca = cp * h4 - x * h3 - sb * a0;
  const double Cb = cp * h2 - w * h1 - sa * a0;
  // s <- CA . s
  s[1] = cq * g1_ + z * g4_;
  s[3] = cq * g3_ + yy * g2_;
  s[0] = a0 * s0 + y.ir * Da + y.kr * Db;
  s[2] = a0 * s2 + y.tg * Da + y.t1 * Db;
  s[4] = a0 * s4 + y.ar * Da + y.br * Db;
  // coefficient scalars -> 2 beta^2 (the only coefficient input besides rho and 1/c):
  //   d/dgk = s2 (Ca + ic2 Cb) + t2 e2 (Da + ic2 Db) + 2 gam rho b_ar + 2 g1 ic2 rho b_br
  const double b_ar = Da * e4 + Ca * s0;
  const double b_br = Db * e4 + Cb * s0;
  const double X1 = Ca + ic2 * Cb;
  const double X2 = Da + ic2 * Db;
  const double b_gk = s2 * X1 + (t2 * e2) * X2 + y.c1 * b_ar + y.c3 * b_br;
  const double bex = -a0 * ba0;
  double bpP, bpS;
  const double tP = eig_adjoint(y.P, bcp, bw, bx, bex, bpP);
  const double tS = eig_adjoint(y.S, bcq, by, bz, bex, bpS);
  LayerGrad g;
  // k == kx: w = omega d and the (measure-zero) dependence through r is dropped (finite one-sided value)
  g.d = omega * ((y.P.pos ? y.P.r * bpP : bw) + (y.S.pos ? y.S.r * bpS : by));
  g.a = y.P.pos ? y.tqa * tP : 0.0;
  g.b = (y.S.pos ? y.tqb * tS : 0.0) + y.cb * b_gk;
  const double tn = e0 * s[0] - e4 * s[4];
  g.rho = y.ir * (tq - tn);
  tq = tn;
  g.ic = 0.0;
  if (kWantIc) {
    // d/d(1/c): through n^2 = +-(ic2 - 1/v^2) and through ic2 in the coefficients
    const double b_kr = Db * e0 + Cb * s4;
    const double b_tg = Da * e2, b_t1 = Db * e2;
    const double gam = y.g1 + 1.0;
    const double b_ic2 = y.gk * (Cb * s2) + (y.gk * y.gk * y.rho) * b_ar + (2.0 * y.g1 * y.gk * y.rho) * b_br +
                         y.ir * b_kr - 2.0 * (y.gk * b_tg + (y.g1 + gam) * b_t1);
    const double hp = y.P.pos ? (y.P.lt ? -y.P.ir : y.P.ir) * y.P.ir * tP : 0.0;
    const double hs = y.S.pos ? (y.S.lt ? -y.S.ir : y.S.ir) * y.S.ir * tS : 0.0;
    g.ic = ic * (hp + hs + 2.0 * b_ic2);
  }
  return g;
}

// Half-space start vector (reference :328-347), in the reference's own (unscaled) variables
struct Half {
  double ka, kb, gammk, rho;
  double ra, ira, rb, irb, gam, gamm1, R;
  int bra, brb;  // 0 osc (k<kx), 1 evan (k>kx), 2 equal/unordered
};

struct HalfGrad {
  double hp, hs, gammk, rho, k;  // h = (dDelta/dr)/r * sign(k - kx):  dDelta/dalpha = omega^2/alpha^3 hp
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

ADSURF_HD HalfGrad half_adjoint(const Half& h, const double (&e)[5], const double (&s)[5], double k, double k2) {
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
  HalfGrad g;
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
  double k, k2, omega, iw, iw2;  // wavenumber, k^2, clamped angular frequency, 1/omega, 1/omega^2: half-space
  double ic, ic2, t2;            // k / omega (= 1/c unless omega is clamped), its square, -2 ic^2: the layers
};

// omega = twopi / t as torch evaluates `python_float / tensor` (Tensor.__rtruediv__ ==
// reciprocal(t) * twopi) — the form every tensor-input caller of the reference goes through
// (_cps/_surf96_vectorAll_gpu.py:318, _matrixAll_gpu.py:256).  Matching the last bit matters only
// where a trial velocity coincides with the half-space velocity (see load_half).
ADSURF_HD double omega_of_period(double t) { return (1.0 / t) * kTwoPi; }

// iw = 1 / max(omega0, 1e-4), precomputed by callers that revisit a period
ADSURF_HD Sample make_sample(double omega0, double ic, double iw) {
  Sample z;
  z.k = ic * omega0;                          // (1/c) * omega   (reference :319, unclamped omega)
  z.omega = fmax(omega0, 1.0e-4);             // reference :326
  z.k2 = z.k * z.k;
  z.iw = iw;
  z.iw2 = z.iw * z.iw;
  z.ic = (omega0 >= 1.0e-4) ? ic : z.k * z.iw;  // periods beyond 2 pi 1e4 s: k / omega differs from 1/c
  z.ic2 = z.ic * z.ic;
  z.t2 = -2.0 * z.ic2;
  return z;
}
ADSURF_HD Sample make_sample(double omega0, double ic) { return make_sample(omega0, ic, 1.0 / fmax(omega0, 1.0e-4)); }

// Dense kernels: the shared (velocity, layer) table is built for k/omega = 1/c, so a period beyond 2 pi 1e4 s
// (where the reference clamps omega but not k, making k/omega period-dependent) is evaluated with the clamp
// applied to both, i.e. at the clamped period.  The pointwise kernels follow the reference exactly.
ADSURF_HD Sample make_sample_dense(double omega0, double ic, double iw) {
  return make_sample(fmax(omega0, 1.0e-4), ic, iw);
}

// ---- layer sources: fill the frequency-free coefficient scalars, the phases and the regime of layer m.
// Everything that depends on (trial velocity, layer) but not on the period is one table entry:
//   n = sqrt(|1/c^2 - 1/v^2|), 1/n, n d (phase per unit omega) for P and S; the coefficient scalars
//   (1/rho, -2 gam, gam gk rho, ic2/rho, -2 ic2 g1, g1^2 rho, gk = 2 beta^2, g1 = gam - 1, gam = gk ic2);
//   the regime flags; and for the adjoint the constants that fold the chain rule to (alpha, beta):
//   1/(alpha^3 (ic2 - 1/alpha^2)), same for beta, 4 beta, 2 gam rho, 2 g1 ic2 rho.
// TableSource (dense kernels) reads entries that the block computed once per trial velocity into shared
// memory (16-byte broadcasts); DirectSource (pointwise kernels, arbitrary (t, c) pairs) computes the entry
// per sample.
constexpr int kTabFwd = 16, kTabAdj = 20;  // doubles per entry: forward sweep only / with the adjoint constants
enum {
  T_PDA = 0, T_PDB = 1, T_NA = 2, T_NB = 3, T_INA = 4, T_INB = 5, T_IR = 6, T_TG = 7, T_AR = 8, T_KR = 9,
  T_T1 = 10, T_BR = 11, T_GK = 12, T_G1 = 13, T_FL = 14, T_TQA = 15, T_TQB = 16, T_CB = 17, T_C1 = 18, T_C3 = 19
};
struct alignas(16) Dbl2 {
  double x, y;
};
ADSURF_HD Dbl2 ld2(const double* p) { return *reinterpret_cast<const Dbl2*>(p); }

// one table entry: trial slowness ic = 1/c against layer m of the staged model
// kFast (pointwise kernels: one entry per sample and layer): n and 1/n from the rsqrt sequence and the
// alpha/beta chain constants as +-(1/n)^2 / v^3 — no square root and no division; otherwise (dense kernels: one
// entry per (velocity, layer) and block) correctly rounded sqrt and divisions.
template <int kN, bool kFast = false>
ADSURF_HD void table_entry(double* __restrict__ tab, const double* __restrict__ par, int m, double ic) {
  const double* __restrict__ pr = par + m * kNumPar;
  const double sa = pr[P_IA], sb = pr[P_IB], dm = pr[P_D], rho = pr[P_RHO], ir = pr[P_IR], gk = pr[P_TB2];
  const double gkr = pr[P_GKR];  // gk rho, staged per (model, layer)
  const double da = (ic + sa) * fabs(ic - sa), db = (ic + sb) * fabs(ic - sb);
  double na, nb, ina, inb;
  if (kFast) {
    ina = rsqrt_hd(da);
    inb = rsqrt_hd(db);
    na = da > 0.0 ? da * ina : da;  // 0 -> 0, NaN -> NaN
    nb = db > 0.0 ? db * inb : db;
  } else {
    na = sqrt(da);
    nb = sqrt(db);
    ina = 1.0 / na;
    inb = 1.0 / nb;
  }
  const double ic2 = ic * ic;
  const double gam = gk * ic2;
  const double g1 = gam - 1.0;
  tab[T_PDA] = na * dm;
  tab[T_PDB] = nb * dm;
  tab[T_NA] = na;
  tab[T_NB] = nb;
  tab[T_INA] = ina;
  tab[T_INB] = inb;
  tab[T_IR] = ir;
  const double t1 = (-2.0 * ic2) * g1;  // (-2 ic2) does not depend on the layer: hoisted by the pointwise sweeps
  tab[T_TG] = -2.0 * gam;
  tab[T_AR] = gam * gkr;
  tab[T_KR] = ic2 * ir;
  tab[T_T1] = t1;
  tab[T_BR] = g1 * (g1 * rho);
  tab[T_GK] = gk;
  tab[T_G1] = g1;
  const int fl = (ic < sa ? F_LTP : 0) | (da > 0.0 ? F_POSP : 0) | (ic < sb ? F_LTS : 0) | (db > 0.0 ? F_POSS : 0);
  // raw bits, read back with dbl_bits(): regime flags in the low word, the layer index in the high word
  tab[T_FL] = bits_dbl(((long long)m << 32) | (long long)fl);
  tab[T_TQA] = kFast ? (ic < sa ? -pr[P_IA3] : pr[P_IA3]) * (ina * ina) : pr[P_IA3] / ((ic + sa) * (ic - sa));
  if (kN > kTabFwd) {
    tab[T_TQB] = kFast ? (ic < sb ? -pr[P_IB3] : pr[P_IB3]) * (inb * inb) : pr[P_IB3] / ((ic + sb) * (ic - sb));
    tab[T_CB] = pr[P_CB];
    tab[T_C1] = (2.0 * ic2) * gkr;  // 2 gam rho
    tab[T_C3] = -(t1 * rho);        // 2 g1 ic2 rho
  }
}

// fill a Layer from a table entry (kN = kTabFwd: the adjoint constants are not there and not read)
template <int kN>
ADSURF_HD void layer_from_entry(Layer& y, const double* __restrict__ q, const double* __restrict__ pr, double omega,
                                const Dbl2 t01) {
  const Dbl2 t23 = ld2(q + T_NA), t45 = ld2(q + T_INA), t67 = ld2(q + T_IR),
             t89 = ld2(q + T_AR), tab_ = ld2(q + T_T1), tcd = ld2(q + T_GK), tef = ld2(q + T_FL);
  y.pr = pr;
  y.P.p = omega * t01.x;
  y.S.p = omega * t01.y;
  y.P.r = t23.x;
  y.S.r = t23.y;
  y.P.ir = t45.x;
  y.S.ir = t45.y;
  y.ir = t67.x;
  y.tg = t67.y;
  y.ar = t89.x;
  y.kr = t89.y;
  y.t1 = tab_.x;
  y.br = tab_.y;
  y.gk = tcd.x;
  y.g1 = tcd.y;
  y.fl = (int)dbl_bits(tef.x);  // the per-eig booleans are set where a regime is selected (layer_set_flags)
  if (kN > kTabFwd) {
    const Dbl2 tgh = ld2(q + T_TQB), tij = ld2(q + T_C1);
    y.tqa = tef.y;
    y.tqb = tgh.x;
    y.cb = tgh.y;
    y.c1 = tij.x;
    y.c3 = tij.y;
  }
}

template <bool kIc>
struct DirectSourceT {
  static constexpr bool kWantIc = kIc;          // dF/dc wanted (gradient and Jacobian entries; not the DMF step)
  const double* __restrict__ par;
  int L;
  ADSURF_HD const double* half_k() const { return nullptr; }
  ADSURF_HD void prep(Layer& y, int m, const Sample& z) const {
    alignas(16) double tab[kTabAdj];
    table_entry<kTabAdj, true>(tab, par, m, z.ic);
    layer_from_entry<kTabAdj>(y, tab, par + m * kNumPar, z.omega, Dbl2{tab[T_PDA], tab[T_PDB]});
    y.rho = par[m * kNumPar + P_RHO];
  }
};
typedef DirectSourceT<true> DirectSource;

// Dense kernels: the entries of one trial velocity in the block's shared table.  Layout of the table (built by
// the kernels' fill_table): per velocity one SENTINEL entry (flags = F_END, phases 0) followed by the L-1 layer
// entries, and one more sentinel after the last velocity — so the entry before layer 0 and the entry after layer
// L-2 are both sentinels.  The sweeps walk the entries with a running pointer and stop at a sentinel: no layer
// index, no bounds tests, no predicated loads in the layer loops.
template <int kN>
struct TableSource {
  static constexpr bool kWantIc = false;
  const double* __restrict__ par;
  int L;
  const double* __restrict__ tab;  // entry of layer 0 for the current trial velocity (16-byte aligned)
  const double* hkq;               // this lane's staged (omega/alpha_hs, omega/beta_hs), see load_half
  ADSURF_HD const double* half_k() const { return hkq; }
  ADSURF_HD static int flags(const double* q) { return (int)dbl_bits(q[T_FL]); }
  // every phase of the sample below 1e5 (the inline sincos's range): p = omega n d, n <= max(1/c, 1/beta)
  ADSURF_HD bool small_phases(const Sample& z) const {
    return z.omega * fmax(z.ic, par[kNumPar * L + 2]) * par[kNumPar * L + 3] < 1.0e5;
  }
  // the layer's staged record (read only where c coincides with a layer velocity): its index rides in the entry
  ADSURF_HD const double* record(const double* q) const { return par + (int)(dbl_bits(q[T_FL]) >> 32) * kNumPar; }
};
// doubles of one velocity's entries in the table (sentinel + layers) and of a table of tc velocities
ADSURF_HD int table_stride(int L, int kN) { return L * kN; }
ADSURF_HD int table_doubles(int L, int tc, int kN) { return (tc * L + 1) * kN; }
// sentinel entry
ADSURF_HD void table_sentinel(double* __restrict__ tab, int kN) {
  for (int i = 0; i < kN; ++i) tab[i] = 0.0;
  tab[T_FL] = bits_dbl((long long)F_END);
}

// The half-space vector is linear in ra, rb = sqrt((k+kx)|k-kx|): at a trial velocity that coincides
// with the half-space velocity the result depends on the last bit of kx, so kx is formed with the
// reference's own correctly-rounded division omega/alpha (reference :328-329) from the raw
// velocities kept at par[kNumPar*L + 0/1].
// hk: the sample's (omega/alpha_hs, omega/beta_hs) if the caller staged them per period (dense kernels: two
// correctly rounded divisions per period and block instead of four per sample), else null.
ADSURF_HD void load_half(Half& h, const double* __restrict__ par, int L, const Sample& z,
                         const double* __restrict__ hk) {
  const int m = L - 1;
  if (hk) {
    const Dbl2 k = ld2(hk);
    h.ka = k.x;
    h.kb = k.y;
  } else {
    h.ka = z.omega / par[kNumPar * L + 0];
    h.kb = z.omega / par[kNumPar * L + 1];
  }
  h.gammk = par[m * kNumPar + P_TB2] * z.iw2;
  h.rho = par[m * kNumPar + P_RHO];
}
// what a caller stages per period for load_half (omega0 unclamped)
ADSURF_HD void stage_half_k(double* __restrict__ hk, const double* __restrict__ par, int L, double omega0) {
  const double w = fmax(omega0, 1.0e-4);
  hk[0] = w / par[kNumPar * L + 0];
  hk[1] = w / par[kNumPar * L + 1];
}

// half-space vector in the scaled variables of the layer sweep: E~ = E D^-1, D = diag(1, w, 1, w, w^2)
ADSURF_HD void half_start(const double* __restrict__ par, int L, const Sample& z, const double* __restrict__ hk,
                          double (&e)[5]) {
  Half h;
  load_half(h, par, L, z, hk);
  half_space(h, z.k, z.k2, e);
  e[1] *= z.iw;
  e[3] *= z.iw;
  e[4] *= z.iw2;
}

// Storage policy for what the forward sweep keeps for the adjoint sweep, kSlot = 9 doubles per layer:
// the row vector E_{m+1} (5) and the transcendental results of the P and S terms (2 + 2).  Element
// (m, j) of this thread lives at p[(m*kSlot+j)*kStride] (kStride = 32: lane-interleaved, i.e.
// conflict-free in shared memory and coalesced in global memory; kStride = 1 on the host).
constexpr int kSlot = 9;
// Pair stores (Store::kPairs, the dense adjoint kernel's global slab): the same values as 16-byte pairs, one vector
// access per pair — (E0,E1) (E2,E3) (E4,cP0) (cS0,cS1) and, only under an oscillatory P term, (cP1,-) — in slots of
// kPairSlot doubles: 4 accesses per layer and sweep instead of 7-8 (a slab access costs an LSU instruction besides
// its bytes, profiles/r2_variants.txt).  st2 / ld2 address a pair by its index.
constexpr int kPairSlot = 10;

template <int kStride>
struct PlainStore {
  static constexpr bool kPairs = false;
  double* p;
  typedef double* Slot;
  ADSURF_HD Slot slot(int m) const { return p + (size_t)m * (kSlot * kStride); }
  ADSURF_HD static Slot step(Slot q, int n) { return q + n * (kSlot * kStride); }  // n slots further
  ADSURF_HD static double ld(Slot q, int j) { return q[j * kStride]; }
  ADSURF_HD static void st(Slot q, int j, double v) { q[j * kStride] = v; }
};
struct NoStore {
  static constexpr bool kPairs = false;
  typedef int Slot;
  ADSURF_HD Slot slot(int) const { return 0; }
  ADSURF_HD static Slot step(Slot q, int) { return q; }
  ADSURF_HD static void st(Slot, int, double) {}
};

constexpr int kEO = F_REGULAR | F_LTS;          // P evanescent, S oscillatory: c between the S and P velocity
constexpr int kEE = F_REGULAR;                  // both evanescent: c below both
constexpr int kOO = F_REGULAR | F_LTP | F_LTS;  // both oscillatory: c above both
constexpr int kOE = F_REGULAR | F_LTP;

// Pointwise sweeps: lanes of a warp are different (period, velocity) picks, so regimes may differ between lanes —
// but neighbouring picks mostly share them.  Returns the regime flags if every lane of the warp is in the same
// regular regime with phases inside the inline sincos's range (then the branch-free specialised layer bodies
// apply), else -1.  All 32 lanes must be converged here.
ADSURF_HD int uniform_regime(const Layer& y) {
  const bool ok = (y.fl & F_REGULAR) == F_REGULAR && (!(y.fl & F_LTP) || below1e5(fabs(y.P.p))) &&
                  (!(y.fl & F_LTS) || below1e5(fabs(y.S.p)));
#if defined(__CUDA_ARCH__)
  const int fl0 = __shfl_sync(0xffffffffu, y.fl, 0);
  return __all_sync(0xffffffffu, ok && y.fl == fl0) ? fl0 : -1;
#else
  return ok ? y.fl : -1;
#endif
}

// one layer of the bottom-up sweep: eigenfunction terms of the regime kFl (>= 0: regular, |phase| < 1e5, branch-free
// bodies; -1: huge phase, NaN, or c exactly on a layer velocity), optional store, E <- E . CA
template <int kFl, bool kStore, typename Store>
ADSURF_HD void forward_layer(Layer& y, double (&e)[5], double omega, typename Store::Slot q) {
  if (kFl >= 0) {
    layer_set_flags(y, kFl);
    if (kFl & F_LTP) eig_osc(y.P); else eig_evan(y.P);
    if (kFl & F_LTS) eig_osc(y.S); else eig_evan(y.S);
    y.P.w = y.P.sn * y.P.ir;
    y.S.w = y.S.sn * y.S.ir;
  } else {
    layer_set_flags(y, y.fl);
    const double wd = omega * y.pr[P_D];
    eig_forward(y.P, wd);
    eig_forward(y.S, wd);
  }
  y.a0 = layer_a0(y.P, y.S);
  if constexpr (kStore && Store::kPairs) {
    // e is still E_{m+1} here (row_apply below turns it into E_m): the whole slot in 4 (5) vector stores
    Store::st2(q, 0, e[0], e[1]);
    Store::st2(q, 1, e[2], e[3]);
    Store::st2(q, 2, e[4], y.P.c0);
    Store::st2(q, 3, y.S.c0, y.S.c1);
    if (kFl >= 0 ? (kFl & F_LTP) != 0 : y.P.lt) Store::st2(q, 4, y.P.c1, y.P.c1);
  } else if (kStore) {
    // the second value (cos) exists only for an oscillatory term; the adjoint sweep loads it on the same rule
    Store::st(q, 5, y.P.c0);
    if (kFl >= 0 ? (kFl & F_LTP) != 0 : y.P.lt) Store::st(q, 6, y.P.c1);
    Store::st(q, 7, y.S.c0);
    if (kFl >= 0 ? (kFl & F_LTS) != 0 : y.S.lt) Store::st(q, 8, y.S.c1);
  }
  row_apply(e, y);
}

// Bottom-up sweep E <- E . CA_m, m = L-2 .. 0, from the half-space vector; kStore keeps E_{m+1} and the
// transcendental results of every layer in slot m for the adjoint sweep.  Pointwise sources (one table entry
// computed per sample and layer, regimes differ between lanes): the generic layer body.
template <bool kStore, bool kIc, typename Store>
ADSURF_HD double forward_sweep(const DirectSourceT<kIc>& src, const Sample& z, const Store& est) {
  const int L = src.L;
  double e[5];
  half_start(src.par, L, z, nullptr, e);
  for (int m = L - 2; m >= 0; --m) {
    const typename Store::Slot q = est.slot(m);
    if (kStore) {  // E_{m+1} multiplies CA_m
#pragma unroll
      for (int i = 0; i < 5; ++i) Store::st(q, i, e[i]);
    }
    Layer y;
    src.prep(y, m, z);
    const int ur = uniform_regime(y);
    if (ur == kEO) forward_layer<kEO, kStore, Store>(y, e, z.omega, q);
    else if (ur == kEE) forward_layer<kEE, kStore, Store>(y, e, z.omega, q);
    else if (ur == kOO) forward_layer<kOO, kStore, Store>(y, e, z.omega, q);  // shallow layers under long-period picks
    else forward_layer<-1, kStore, Store>(y, e, z.omega, q);
  }
  return e[0];
}

// Dense kernels (warp-uniform regime flags from the shared table): consecutive layers in the same regime run in a
// straight-line loop of their own — no selects, the P and S polynomial chains (each a ~16-deep dependent FP64
// sequence, 8 cycles per link) interleaved by the scheduler.  The table entries and the store slots are walked
// with running pointers; the next layer's flags and phase pair (which heads the layer's longest dependent chain:
// phase -> exp / sincos -> ...) are read one layer ahead, so the branch that closes a run never waits for shared
// memory; a sentinel entry ends the sweep.
template <bool kStore, int kN, typename Store>
ADSURF_HD double forward_sweep(const TableSource<kN>& src, const Sample& z, const Store& est) {
  const int L = src.L;
  double e[5];
  half_start(src.par, L, z, src.half_k(), e);
  const double* __restrict__ q = src.tab + (L - 2) * kN;  // layer L-2 (the leading sentinel when there is no layer)
  typename Store::Slot sq = est.slot(L - 2);
  int fl = TableSource<kN>::flags(q);
  Dbl2 ph = ld2(q + T_PDA);
  if (kStore && !Store::kPairs && !(fl & F_END)) {
#pragma unroll
    for (int i = 0; i < 5; ++i) Store::st(sq, i, e[i]);
  }
  const bool small = src.small_phases(z);
  auto run = [&](auto tag) {
    constexpr int kFl = decltype(tag)::value;
    do {
      const int fln = TableSource<kN>::flags(q - kN);
      const Dbl2 phn = ld2(q - kN + T_PDA);
      Layer y;
      layer_from_entry<kN>(y, q, kFl < 0 ? src.record(q) : src.par, z.omega, ph);
      forward_layer<kFl, kStore, Store>(y, e, z.omega, sq);
      q -= kN;
      sq = Store::step(sq, -1);
      if (kStore && !Store::kPairs && !(fln & F_END)) {  // E_m multiplies CA_{m-1}: stored right after it is formed, a whole layer
#pragma unroll                          // before its registers are overwritten (no store/overwrite hazard)
        for (int i = 0; i < 5; ++i) Store::st(sq, i, e[i]);
      }
      fl = fln;
      ph = phn;
    } while (kFl >= 0 && fl == kFl);
  };
  while (!(fl & F_END)) {
    if (!small) run(std::integral_constant<int, -1>{});
    else if (fl == kEO) run(std::integral_constant<int, kEO>{});
    else if (fl == kEE) run(std::integral_constant<int, kEE>{});
    else if (fl == kOO) run(std::integral_constant<int, kOO>{});
    else if (fl == kOE) run(std::integral_constant<int, kOE>{});
    else run(std::integral_constant<int, -1>{});
  }
  return e[0];
}

template <typename Source>
ADSURF_HD double sample_forward(const Source& src, const Sample& z) {
  return forward_sweep<false>(src, z, NoStore{});
}
template <typename Source, typename Store>
ADSURF_HD double sample_forward_store(const Source& src, const Sample& z, const Store& est) {
  return forward_sweep<true>(src, z, est);
}

// half-space end of the adjoint sweep: dDelta/dE_hs = D^-1 s~_{L-1}, in the reference's variables; returns
// omega * dDelta/dk and the gradients of the half-space "layer" L-1
ADSURF_HD double adjoint_half_space(const double* __restrict__ par, int L, const Sample& z, const double* hk,
                                    const double (&s)[5], LayerGrad& g) {
  double e[5];
  Half h;
  load_half(h, par, L, z, hk);
  half_space(h, z.k, z.k2, e);
  const double su[5] = {s[0], s[1] * z.iw, s[2], s[3] * z.iw, s[4] * z.iw2};
  const HalfGrad hg = half_adjoint(h, e, su, z.k, z.k2);
  const double* __restrict__ prh = par + (L - 1) * kNumPar;
  const double w2 = z.omega * z.omega;
  // k_alpha = omega/alpha, k_beta = omega/beta, gammk = 2 beta^2 / omega^2:
  //   d/dalpha = omega^2/alpha^3 hp,  d/dbeta = omega^2/beta^3 hs + 4 beta/omega^2 dDelta/dgammk
  g.d = 0.0;
  g.a = (w2 * prh[P_IA3]) * hg.hp;
  g.b = (w2 * prh[P_IB3]) * hg.hs + (prh[P_CB] * z.iw2) * hg.gammk;
  g.rho = hg.rho;
  return z.omega * hg.k;  // k = omega * (k/omega)
}

// Top-down adjoint sweep.  s_0 = seed * e_0; for every layer m (and finally the half-space, m = L-1)
// hands (m, dDelta/dd_m, /dalpha_m, /dbeta_m, /drho_m) * seed to `sink`.  det is the forward sweep's result.
// Returns seed * dDelta/d(k/omega).  Pointwise sources: generic layer body (regimes differ between lanes).
// Software pipeline over the stored values: each group of registers is re-loaded for layer m+1 right
// after its last use in layer m — the transcendental results once eig_finish has consumed them, the row
// vector E once layer_adjoint is done — so the latency of the slab hides behind the rest of the layer
// without a second register buffer.
template <bool kIc, typename Store, typename Sink>
ADSURF_HD double sample_adjoint(const DirectSourceT<kIc>& src, const Sample& z, double seed, double det,
                                const Store& est, Sink& sink) {
  const int L = src.L;
  double s[5] = {seed, 0.0, 0.0, 0.0, 0.0};
  double tq = seed * det;  // T_0
  double gic = 0.0;
  double e[5], cv[4];
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
  // the gradient of layer m goes to the sink while layer m+1 is being worked on (see the dense sweep)
  LayerGrad pend;
  pend.d = pend.a = pend.b = pend.rho = 0.0;
  const int last = L - 2;
  if (L > 1) {
    fetch_c(0);
    fetch_e(0);
  }
  for (int m = 0; m <= last; ++m) {
    Layer y;
    y.P.c0 = cv[0]; y.P.c1 = cv[1]; y.S.c0 = cv[2]; y.S.c1 = cv[3];
    src.prep(y, m, z);
    const int mn = m + 1 <= last ? m + 1 : m;  // the last layer re-reads its own slot: no predicated loads
    auto body = [&](auto tag) {
      constexpr int kFl = decltype(tag)::value;
      layer_set_flags(y, kFl >= 0 ? kFl : y.fl);
      layer_eig_restore(y, z.omega);
      sink(m > 0 ? m - 1 : 0, pend.d, pend.a, pend.b, pend.rho);  // layer m-1 (zeros before layer 0)
      fetch_c(mn);
      const LayerGrad g = layer_adjoint<kIc>(y, e, s, tq, z.omega, z.ic, z.ic2, z.t2);
      fetch_e(mn);
      gic += g.ic;
      pend = g;
    };
    const int ur = uniform_regime(y);
    if (ur == kEO) body(std::integral_constant<int, kEO>{});
    else if (ur == kEE) body(std::integral_constant<int, kEE>{});
    else if (ur == kOO) body(std::integral_constant<int, kOO>{});
    else body(std::integral_constant<int, -1>{});
  }
  if (L > 1) sink(last, pend.d, pend.a, pend.b, pend.rho);
  LayerGrad gh;
  gic += adjoint_half_space(src.par, L, z, nullptr, s, gh);
  sink(L - 1, gh.d, gh.a, gh.b, gh.rho);
  return gic;
}

// Dense kernels.  The sink is a WALKING sink: put() deposits one layer's four gradients and moves on to the next
// layer's accumulators; the sweep calls it L + 1 times — first with zeros (the gradient of layer m goes to the
// sink while layer m+1 is being worked on: the sink's exchange and shared-memory round trips, short-scoreboard
// latencies in a dependent chain, then overlap FP64 work instead of stalling the warp at the end of every
// layer), then layers 0 .. L-2, then the half-space L-1.  Table entries, store slots and accumulators are walked
// with running pointers; the store has one slot more than layers, so the read-ahead of the last layer needs no
// bounds test; kFl >= 0 states the regime flags as a compile-time constant, which folds every "lt ? a : b" /
// "pos ? a : b" select of eig_finish / eig_adjoint away; consecutive layers in the same regime (the usual case:
// oscillatory S near the surface, evanescent below) run in a loop of their own.
template <int kN, typename Store, typename Sink>
ADSURF_HD void sample_adjoint(const TableSource<kN>& src, const Sample& z, double seed, double det, const Store& est,
                              Sink& sink) {
  const int L = src.L;
  double s[5] = {seed, 0.0, 0.0, 0.0, 0.0};
  double tq = seed * det;  // T_0
  double e[5], cv[4] = {0.0, 0.0, 0.0, 0.0};
  auto fetch_c = [&](typename Store::Slot q, int flm) {  // flm: regime flags of the layer in slot q
    if constexpr (!Store::kPairs) {
      cv[0] = Store::ld(q, 5);
      if (flm & F_LTP) cv[1] = Store::ld(q, 6);
      cv[2] = Store::ld(q, 7);
      if (flm & F_LTS) cv[3] = Store::ld(q, 8);
    }
  };
  auto fetch_e = [&](typename Store::Slot q) {
    if constexpr (!Store::kPairs) {
#pragma unroll
      for (int i = 0; i < 5; ++i) e[i] = Store::ld(q, i);
    }
  };
  // pair stores: the whole slot (row vector into en, transcendental values into cv) in 4 (5) vector loads
  auto fetch_pairs = [&](typename Store::Slot q, int flm, double (&en)[5]) {
    if constexpr (Store::kPairs) {
      const Dbl2 p0 = Store::ld2(q, 0), p1 = Store::ld2(q, 1), p2 = Store::ld2(q, 2), p3 = Store::ld2(q, 3);
      en[0] = p0.x; en[1] = p0.y; en[2] = p1.x; en[3] = p1.y; en[4] = p2.x;
      cv[0] = p2.y; cv[2] = p3.x; cv[3] = p3.y;
      if (flm & F_LTP) cv[1] = Store::ld2(q, 4).x;
    }
  };
  LayerGrad pend;
  pend.d = pend.a = pend.b = pend.rho = 0.0;
  const double* __restrict__ q = src.tab;  // layer 0 (a sentinel when there is no layer)
  typename Store::Slot sq = est.slot(0);
  int fl = TableSource<kN>::flags(q);
  if (!(fl & F_END)) {
    fetch_c(sq, fl);
    fetch_e(sq);
    fetch_pairs(sq, fl, e);
  }
  auto layer = [&](auto tag) {
    constexpr int kFl = decltype(tag)::value;
    const int fln = TableSource<kN>::flags(q + kN);
    Layer y;
    y.P.c0 = cv[0]; y.P.c1 = cv[1]; y.S.c0 = cv[2]; y.S.c1 = cv[3];
    layer_from_entry<kN>(y, q, kFl < 0 ? src.record(q) : src.par, z.omega, ld2(q + T_PDA));
    layer_set_flags(y, kFl >= 0 ? kFl : y.fl);
    sq = Store::step(sq, 1);
    layer_eig_restore(y, z.omega);
    sink.put(pend.d, pend.a, pend.b, pend.rho);
    fetch_c(sq, fln);  // after the last layer: the slot behind it (values unused)
    // the next layer's row vector is requested a whole layer ahead, into registers of its own: issued only after
    // the last use of e it arrives late and the top of the next layer waits for L2 (long-scoreboard stalls;
    // measured 2.776 vs 2.711 G samples/s, profiles/r2_variants.txt)
    double en[5];
    if constexpr (Store::kPairs) {
      fetch_pairs(sq, fln, en);
    } else {
#pragma unroll
      for (int i = 0; i < 5; ++i) en[i] = Store::ld(sq, i);
    }
    const LayerGrad g = layer_adjoint<false>(y, e, s, tq, z.omega, z.ic, z.ic2, z.t2);
#pragma unroll
    for (int i = 0; i < 5; ++i) e[i] = en[i];
    pend = g;
    q += kN;
    fl = fln;
  };
  auto run = [&](auto tag) {
    constexpr int kFl = decltype(tag)::value;
    do {
      layer(tag);
    } while (fl == kFl);
  };
  // (config 5: 55 % of the (velocity, layer) pairs are EE, 39 % EO, 6 % OO — trial velocities above a shallow
  // layer's Vp —, OE never occurs: alpha > beta)
  while (!(fl & F_END)) {
    if (fl == kEO) run(std::integral_constant<int, kEO>{});
    else if (fl == kEE) run(std::integral_constant<int, kEE>{});
    else if (fl == kOO) run(std::integral_constant<int, kOO>{});
    else layer(std::integral_constant<int, -1>{});
  }
  sink.put(pend.d, pend.a, pend.b, pend.rho);
  LayerGrad gh;
  adjoint_half_space(src.par, L, z, src.half_k(), s, gh);
  sink.put(gh.d, gh.a, gh.b, gh.rho);
}

// adapter: a sink indexed by the layer number as a walking sink (host check, shared-memory reduce sink)
template <typename Sink>
struct WalkingSink {
  Sink& sink;
  int m;
  ADSURF_HD explicit WalkingSink(Sink& s) : sink(s), m(-1) {}
  ADSURF_HD void put(double gd, double ga, double gb, double gr) {
    if (m >= 0) sink(m, gd, ga, gb, gr);
    ++m;
  }
};

// per-model layer constants staged for the kernels: par[L][kNumPar] (one 80-byte record per layer),
// then the raw half-space velocities
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
  par[m * kNumPar + P_GKR] = (2.0 * b * b) * r;
  if (m == L - 1) {
    par[kNumPar * L + 0] = a;
    par[kNumPar * L + 1] = b;
  }
}

// bounds used by TableSource::small_phases: max_m 1/beta_m and max_m d_m over the layers above the half-space
ADSURF_HD void stage_bounds(double* __restrict__ par, int L, const double* __restrict__ d,
                            const double* __restrict__ beta) {
  double smax = 0.0, dmax = 0.0;
  for (int m = 0; m < L - 1; ++m) {
    smax = fmax(smax, 1.0 / beta[m]);
    dmax = fmax(dmax, d[m]);
  }
  par[kNumPar * L + 2] = smax;
  par[kNumPar * L + 3] = dmax;
}

// doubles of staged constants per model
ADSURF_HD int par_doubles(int L) { return kNumPar * L + 4; }

}  // namespace adsurf
