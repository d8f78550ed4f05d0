// adsurf_kernels.cu — hand-written sm_100a FP64 kernels for the Rayleigh secular function
// (Dunkin compound-matrix recursion) and its reverse-mode gradient, plus the C ABI
// declared in include/adsurf_b200.h.
//
// What it computes is specified in SURVEY.md Appendix A, which restates
// /root/reference/ADsurf/_cps/_surf96_vectorAll_gpu.py:52-100 (Dunkin matrix), :102-192
// (exponent-scaled eigenfunction terms) and :303-399 (layer recursion, no normalisation).
// How it computes it is new:
//   * one thread per (model, period, trial velocity) sample; lanes of a warp run over PERIODS at a fixed trial
//     velocity, so det[s,k,f] is written coalesced and the three-way branch "c <,==,> layer velocity" is
//     warp-uniform (it does not depend on the period);
//   * a block = (model, velocity chunk, group of period tiles) chosen by a host-side planner that follows the
//     device's SM count (plan_grid).  Per-model layer records (1/alpha, 1/beta,
//     2 beta^2, rho, 1/rho, d, ...) are staged once per block in shared memory.  The layers are swept in
//     frequency-free variables (CA(omega) = D^-1 CA~ D, D = diag(1, w, 1, w, w^2)): everything about a layer
//     except its two phases depends on (velocity, layer) only — sqrt(|1/c^2 - 1/v^2|), its reciprocal, the
//     phase per unit omega, the eight coefficient scalars, the regime flags, the adjoint's chain constants —
//     and is computed once per <= 8 velocities into a shared table read by every lane as 16-byte broadcasts
//     (TableSource); a lane spends two multiplies on a layer's set-up; sentinel entries bracket every velocity's
//     layers, so the sweeps walk the table (and the store slots, and the accumulators) with running pointers;
//   * the 5-component compound state lives in registers; the 5x5 Dunkin matrix is never formed: the layer
//     product uses its rank structure (two rank-one couplings around the Kronecker product of the P and S
//     blocks), 33 FMAs per product; elementary functions are inline (table-driven exp, Cody-Waite sincos);
//   * the adjoint is a hand-derived reverse sweep: pass 1 (bottom-up) stores, per layer, the row vector
//     E_{m+1} and the transcendental results (up to 9 doubles per thread) in a per-resident-warp slab of global
//     memory that stays in L2 (evict_last accesses; adsurf_init optionally reserves the persisting carve-out);
//     pass 2 (top-down) carries the column vector s_m = CA_{m-1}...CA_0 e_0, requests the stored values a whole
//     layer ahead into registers of their own, and contracts
//     dDelta = E_{m+1} . dCA_m . s_m analytically into (d, alpha, beta, rho) while advancing s in the same
//     expressions — no transcendental is evaluated twice; the kernel is persistent (3 blocks x 4 warps per SM:
//     the register file allows 3 warps per scheduler up to 170 registers).  In both passes consecutive layers of
//     one regime run in a straight-line loop, with the values that head a dependent chain (regime flags, phases,
//     stored transcendentals) read one layer ahead and the gradient sink deferred by one layer;
//   * per-model gradient sums: one warp fold + per-lane shared accumulators -> one partial per block ->
//     deterministic second-stage reduce (no atomics);
//   * the dense forward pass fuses max/min over the trial velocities (what the DMF needs: combined across blocks
//     with ordered-key 64-bit atomics) and an optional sign bitmap, so det never has to be materialised;
//   * one iteration of an inversion loop is a replayable launch sequence (adsurf_inversion_step): the iteration
//     index, StepLR and the Adam bias corrections come from a device counter, the optimiser step (velocity chain,
//     Tikhonov terms, projection, Adam, history) runs one thread per (model, layer);
//   * dispersion curves: the reference's sequential root search, one thread per model (adsurf_dispersion.cuh);
//     Monte-Carlo ensembles: Philox4x32-10 normals on the device.
// No tensor cores (not a contraction), no library calls on the path.

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include "adsurf_b200.h"
#include "adsurf_math.cuh"
#include "adsurf_dispersion.cuh"

namespace {

using namespace adsurf;


// Programmatic dependent launch (the kernels of one inversion iteration form a dependent chain in one stream /
// CUDA graph): every kernel of the chain first waits for the grid before it (griddepcontrol.wait: completion and
// memory visibility; a no-op for a plain launch) and then lets the grid after it start being scheduled
// (griddepcontrol.launch_dependents: that grid's blocks become resident as slots free up and wait at their own
// griddepcontrol.wait), so launch latency and the ramp of the next kernel overlap the tail of this one.
__device__ __forceinline__ void chain_wait_then_release() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;");
}
// host side: launches issued while t_chain_pdl is set carry the programmatic-serialisation attribute
thread_local bool t_chain_pdl = false;
template <typename... KArgs, typename... Args>
inline void launch_chain(void (*kern)(KArgs...), unsigned grid, unsigned threads, size_t smem, cudaStream_t st,
                         Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid, 1, 1);
  cfg.blockDim = dim3(threads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = t_chain_pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// gradient sinks for sample_adjoint
// Warp-reduce the 4 gradients of a layer with a transposed butterfly: after the xor-16 and xor-8
// exchanges each lane carries one of the four values (which one = lane bits 4,3), three more
// exchanges finish the sums; lanes 0, 8, 16, 24 then add theirs into the per-warp shared
// accumulators [4][L].  12 SHFL + 6 DADD instead of 40 SHFL + 20 DADD for four plain butterflies.
struct ReduceSink {
  double* acc;
  int L, lane;
  __device__ __forceinline__ void operator()(int m, double gd, double ga, double gb, double gr) const {
    const bool hi16 = lane & 16, hi8 = lane & 8;
    // xor 16: lower half keeps (gd, ga), upper half keeps (gb, gr)
    const double k0 = hi16 ? gb : gd, k1 = hi16 ? gr : ga;
    const double x0 = hi16 ? gd : gb, x1 = hi16 ? ga : gr;
    const double a0 = k0 + __shfl_xor_sync(0xffffffffu, x0, 16);
    const double a1 = k1 + __shfl_xor_sync(0xffffffffu, x1, 16);
    // xor 8: bit 3 clear keeps the first of the pair, set keeps the second
    const double kk = hi8 ? a1 : a0, xx = hi8 ? a0 : a1;
    double v = kk + __shfl_xor_sync(0xffffffffu, xx, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    if ((lane & 7) == 0) acc[(lane >> 3) * L + m] += v;  // lane>>3: 0 gd, 1 ga, 2 gb, 3 grho
  }
};

// Adjoint with the slabs in global memory (shared memory is free): fold the warp once (lane l with lane l^16)
// and keep 16 per-lane partial sums per (layer, gradient) in shared memory: acc[row][4][16], row = layer + 1 (row 0
// takes the zeros a walking sweep deposits first).  2 exchanges + 2 read-modify-writes per layer, no dependent
// shuffle chain; the 16 lanes x warps are summed once per virtual block.
constexpr int kAccRow = 64;  // doubles per (warp, layer): [4][16]
struct FoldSink {
  double* p;  // this lane's cell of the current row: row + ((lane >> 4) << 5) + (lane & 15)
  bool hi;    // upper half-warp: owns (gb, grho); the lower half owns (gd, ga)
  __device__ __forceinline__ FoldSink(double* acc, int lane) : p(acc + ((lane >> 4) << 5) + (lane & 15)), hi(lane & 16) {}
  __device__ __forceinline__ void fold(double* q, double gd, double ga, double gb, double gr) const {
    const double k0 = hi ? gb : gd, k1 = hi ? gr : ga;
    const double x0 = hi ? gd : gb, x1 = hi ? ga : gr;
    const double a0 = k0 + __shfl_xor_sync(0xffffffffu, x0, 16);
    const double a1 = k1 + __shfl_xor_sync(0xffffffffu, x1, 16);
    q[0] += a0;
    q[16] += a1;
  }
  // walking form (dense sweep): rows in the order the sweep visits them
  __device__ __forceinline__ void put(double gd, double ga, double gb, double gr) {
    fold(p, gd, ga, gb, gr);
    p += kAccRow;
  }
  // indexed form (pointwise sweep): layer m lives in row m + 1
  __device__ __forceinline__ void operator()(int m, double gd, double ga, double gb, double gr) const {
    fold(p + (m + 1) * kAccRow, gd, ga, gb, gr);
  }
};
// sum of the 16 lane cells of (warp w, layer m, gradient which) in accs[wpb][L+1][4][16]
__device__ __forceinline__ double fold_total(const double* __restrict__ accs, int wpb, int L, int m, int which) {
  double v = 0.0;
  for (int w = 0; w < wpb; ++w) {
    const double* q = accs + ((size_t)w * (L + 1) + (m + 1)) * kAccRow + (which << 4);
#pragma unroll
    for (int l = 0; l < 16; ++l) v += q[l];
  }
  return v;
}

struct JacSink {  // write this sample's Jacobian rows
  double *jd, *ja, *jb, *jr;
  bool active;
  __device__ __forceinline__ void operator()(int m, double gd, double ga, double gb, double gr) const {
    if (!active) return;
    if (jd) jd[m] = gd;
    if (ja) ja[m] = ga;
    if (jb) jb[m] = gb;
    if (jr) jr[m] = gr;
  }
};

// stage the per-model layer constants; kBounds: also the phase bounds the dense sweeps test once per sample
// (TableSource::small_phases) — a serial loop over the layers on one thread, which the pointwise kernels (one
// sample per thread and staging) skip
template <bool kBounds>
__device__ __forceinline__ void stage_params(double* __restrict__ par, const double* __restrict__ d,
                                             const double* __restrict__ alpha, const double* __restrict__ beta,
                                             const double* __restrict__ rho, int s, int L) {
  stage_exp_table();
  for (int m = threadIdx.x; m < L; m += blockDim.x) {
    stage_one(par, L, m, d[(size_t)s * L + m], alpha[(size_t)s * L + m], beta[(size_t)s * L + m],
              rho[(size_t)s * L + m]);
  }
  if (kBounds && threadIdx.x == blockDim.x - 1) stage_bounds(par, L, d + (size_t)s * L, beta + (size_t)s * L);
}

// ------------------------------------------------------------------------------------------
// Kernels
// ------------------------------------------------------------------------------------------
struct GridPlan {
  int nft;     // period tiles of 32 per model
  int wpb;     // warps per block = period tiles handled by one block
  int nfg;     // period-tile groups per model = ceil(nft / wpb)
  int nchunk;  // trial-velocity chunks per model
  int chunk;   // trial velocities per chunk
  int bpm;     // blocks per model = nchunk * nfg
  int tc;      // trial velocities per shared-memory table fill
};

// Block = (model s, velocity chunk ch, period group fg); warp w owns period tile fg*wpb + w (32 periods,
// one per lane) and walks the chunk's trial velocities.  Every tc velocities the block refills the
// shared table of per-(velocity, layer) quantities (TableSource in adsurf_math.cuh).
struct BlockCoord {
  int s, ch, fg, blk;
};
__device__ __forceinline__ BlockCoord block_coord(int vb, const GridPlan& gp) {
  BlockCoord bc;
  bc.s = vb / gp.bpm;
  bc.blk = vb - bc.s * gp.bpm;
  bc.ch = bc.blk / gp.nfg;
  bc.fg = bc.blk - bc.ch * gp.nfg;
  return bc;
}

// Table of n trial velocities: per velocity a sentinel entry followed by the L-1 layer entries, one more sentinel
// at the end (TableSource in adsurf_math.cuh); ics[cc] = 1 / c.
template <int kTab>
__device__ __forceinline__ void fill_table(double* __restrict__ tab, double* __restrict__ ics,
                                           const double* __restrict__ par, int L, const double* __restrict__ cs,
                                           int n) {
  for (int e = threadIdx.x; e < n * L; e += blockDim.x) {
    const int cc = e / L, j = e - cc * L;
    double* __restrict__ q = tab + (size_t)e * kTab;
    if (j == 0) {
      ics[cc] = 1.0 / cs[cc];
      table_sentinel(q, kTab);
    } else {
      table_entry<kTab>(q, par, j - 1, 1.0 / cs[cc]);
    }
  }
  if (threadIdx.x == blockDim.x - 1) table_sentinel(tab + (size_t)n * L * kTab, kTab);
}

// Order-preserving map double -> uint64 (monotonic for non-NaN values; NaN maps to the largest key) so that a
// maximum over blocks is one native 64-bit atomicMax per lane instead of a partials array and a reduction loop;
// key 0 is below every value: a zeroed buffer is the identity.  The minimum travels as the maximum of -v.
__device__ __forceinline__ unsigned long long order_key(double v) {
  if (v != v) return ~0ull;
  const long long b = __double_as_longlong(v);
  return b < 0 ? ~(unsigned long long)b : (unsigned long long)b | 0x8000000000000000ull;
}
__device__ __forceinline__ double order_value(unsigned long long k) {
  if (k == ~0ull) return NAN;
  return __longlong_as_double((long long)((k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k));
}

// Dense grid, forward only: det / fused max,min over k / sign bitmap.
template <int kMaxRegs>
__global__ void __maxnreg__(kMaxRegs) k_grid_fwd(const double* __restrict__ d, const double* __restrict__ alpha,
                                                  const double* __restrict__ beta, const double* __restrict__ rho,
                                                  const double* __restrict__ t, const double* __restrict__ c, int L,
                                                  int NF, int K, GridPlan gp, double* __restrict__ det,
                                                  double* __restrict__ pmax, double* __restrict__ pmin, int keyed,
                                                  uint32_t* __restrict__ signbits) {
  chain_wait_then_release();
  extern __shared__ double smem[];
  double* par = smem + kSmemHead;
  constexpr int kTab = kTabFwd;
  double* tab = par + par_doubles(L);                   // [tc][sentinel + L-1 layers][kTab] + sentinel
  double* ics = tab + table_doubles(L, gp.tc, kTab);    // [tc]
  double* hks = ics + ((gp.tc + 1) & ~1);               // [wpb*32][2]: omega/alpha_hs, omega/beta_hs per lane
  const BlockCoord bc = block_coord(blockIdx.x, gp);
  const int s = bc.s;
  stage_params<true>(par, d, alpha, beta, rho, s, L);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ft = bc.fg * gp.wpb + warp;
  const bool live = ft < gp.nft;  // warps beyond the last period tile only help filling the table
  const int f = ft * kWarp + lane;
  const bool active = live && f < NF;
  const int fc = active ? f : NF - 1;
  const double omega0 = omega_of_period(t[(size_t)s * NF + fc]);
  const double iw0 = 1.0 / fmax(omega0, 1.0e-4);
  double* hk = hks + 2 * threadIdx.x;
  __syncthreads();  // staged parameters visible
  stage_half_k(hk, par, L, omega0);
  const int k0 = bc.ch * gp.chunk, k1 = min(K, k0 + gp.chunk);
  double vmax = -INFINITY, vmin = INFINITY;
  bool anynan = false;
  const int nw = (NF + 31) / 32;
  for (int kb = k0; kb < k1; kb += gp.tc) {
    const int n = min(gp.tc, k1 - kb);
    __syncthreads();  // staged parameters visible / previous table consumed
    fill_table<kTab>(tab, ics, par, L, c + (size_t)s * K + kb, n);
    __syncthreads();
    if (!live) continue;
    for (int cc = 0; cc < n; ++cc) {
      const int k = kb + cc;
      const Sample z = make_sample_dense(omega0, ics[cc], iw0);
      const TableSource<kTab> src{par, L, tab + (size_t)(cc * L + 1) * kTab, hk};
      const double v = sample_forward(src, z);
      if (det && active) det[((size_t)s * K + k) * NF + f] = v;
      vmax = fmax(vmax, v);
      vmin = fmin(vmin, v);
      anynan |= (v != v);
      if (signbits) {
        const uint32_t bits = __ballot_sync(0xffffffffu, active && (v < 0.0));
        if (lane == 0) signbits[((size_t)s * K + k) * nw + ft] = bits;
      }
    }
  }
  if (pmax && active) {
    // torch.max/min propagate NaN; fmax/fmin do not
    if (anynan) { vmax = NAN; vmin = NAN; }
    if (keyed) {  // [S,NF] ordered keys, combined across the velocity chunks with atomicMax
      atomicMax(reinterpret_cast<unsigned long long*>(pmax) + (size_t)s * NF + f, order_key(vmax));
      atomicMax(reinterpret_cast<unsigned long long*>(pmin) + (size_t)s * NF + f, order_key(-vmin));
    } else {      // [S,nchunk,NF] partials for k_minmax_reduce
      pmax[((size_t)s * gp.nchunk + bc.ch) * NF + f] = vmax;
      pmin[((size_t)s * gp.nchunk + bc.ch) * NF + f] = vmin;
    }
  }
}

__global__ void k_minmax_reduce(const double* __restrict__ pmax, const double* __restrict__ pmin, int NF, int nchunk,
                                double* __restrict__ cmax, double* __restrict__ cmin, size_t total) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const size_t s = i / NF, f = i % NF;
  double vmax = -INFINITY, vmin = INFINITY;
  bool anynan = false;
  for (int ch = 0; ch < nchunk; ++ch) {
    const double a = pmax[(s * nchunk + ch) * NF + f], b = pmin[(s * nchunk + ch) * NF + f];
    anynan |= (a != a) | (b != b);
    vmax = fmax(vmax, a);
    vmin = fmin(vmin, b);
  }
  if (anynan) { vmax = NAN; vmin = NAN; }
  if (cmax) cmax[i] = vmax;
  if (cmin) cmin[i] = vmin;
}

// E-store in global memory (L2-resident working set), lane-interleaved.  Slab accesses carry an L2 evict_last
// policy so the write-back cache keeps them resident (inside the persisting carve-out that adsurf_init reserves)
// while det streams through with evict-first stores.
__device__ __forceinline__ unsigned long long l2_evict_last_policy() {
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
struct GlobalStore {
  static constexpr bool kPairs = false;
  double* p;
  struct Slot {
    double* q;
    unsigned long long pol;
  };
  unsigned long long pol;
  __device__ __forceinline__ Slot slot(int m) const { return Slot{p + (ptrdiff_t)m * (kSlot * kWarp), pol}; }
  __device__ __forceinline__ static Slot step(Slot s, int n) { return Slot{s.q + n * (kSlot * kWarp), s.pol}; }
  __device__ __forceinline__ static double ld(Slot s, int j) {
    double v;
    asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(s.q + j * kWarp), "l"(s.pol));
    return v;
  }
  __device__ __forceinline__ static void st(Slot s, int j, double v) {
    asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(s.q + j * kWarp), "d"(v), "l"(s.pol) : "memory");
  }
};

// the same slab as 16-byte pairs (adsurf_math.cuh, kPairSlot): p = slab base + 2 * lane
struct PairStore {
  static constexpr bool kPairs = true;
  double* p;
  struct Slot {
    double* q;
    unsigned long long pol;
  };
  unsigned long long pol;
  __device__ __forceinline__ Slot slot(int m) const { return Slot{p + (ptrdiff_t)m * (kPairSlot * kWarp), pol}; }
  __device__ __forceinline__ static Slot step(Slot s, int n) { return Slot{s.q + n * (kPairSlot * kWarp), s.pol}; }
  __device__ __forceinline__ static Dbl2 ld2(Slot s, int pr) {
    Dbl2 v;
    asm volatile("ld.global.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;"
                 : "=d"(v.x), "=d"(v.y)
                 : "l"(s.q + pr * (2 * kWarp)), "l"(s.pol));
    return v;
  }
  __device__ __forceinline__ static void st2(Slot s, int pr, double a, double b) {
    asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(s.q + pr * (2 * kWarp)), "d"(a), "d"(b),
                 "l"(s.pol)
                 : "memory");
  }
  // scalar interface (unused by the pair paths; keeps the templates that name it well-formed)
  __device__ __forceinline__ static double ld(Slot, int) { return 0.0; }
  __device__ __forceinline__ static void st(Slot, int, double) {}
};

// Dense grid, fused forward + adjoint with per-model gradient reduction.  Persistent: each block
// walks the virtual blocks vb = blockIdx.x, blockIdx.x + gridDim.x, ... (vb -> model, velocity chunk, period
// group).  The per-layer row vectors E_{m+1} and transcendental results live in a per-resident-warp slab of
// global memory that stays in L2 (148 SMs x 12 warps x 43.8 KB = 78 MB of the 126 MB L2 at 20 layers), so
// residency is bounded by registers only.
template <int kThreads, int kMaxRegs>
__global__ void __maxnreg__(kMaxRegs)
k_grid_bwd(const double* __restrict__ d, const double* __restrict__ alpha, const double* __restrict__ beta,
           const double* __restrict__ rho, const double* __restrict__ t, const double* __restrict__ c,
           const double* __restrict__ gdet, int L, int NF, int K, GridPlan gp, int total_vblocks,
           double* __restrict__ det, double* __restrict__ part, double* __restrict__ estore) {
  extern __shared__ double smem[];
  constexpr int kTab = kTabAdj;
  double* par = smem + kSmemHead;                       // staged layer constants (after the exp table)
  double* accs = par + par_doubles(L);                  // [wpb][L+1][4][16]
  double* tab = accs + (size_t)gp.wpb * kAccRow * (L + 1);  // [tc][sentinel + L-1 layers][kTab] + sentinel
  double* ics = tab + table_doubles(L, gp.tc, kTab);    // [tc]
  double* hks = ics + ((gp.tc + 1) & ~1);               // [wpb*32][2]: omega/alpha_hs, omega/beta_hs per lane
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t slab = (size_t)(L - 1) * kPairSlot * kWarp;  // (the read-ahead of the last layer touches the next slab)
  const PairStore st{estore + ((size_t)blockIdx.x * gp.wpb + warp) * slab + 2 * lane, l2_evict_last_policy()};
  double* acc = accs + (size_t)warp * kAccRow * (L + 1);
  for (int vb = blockIdx.x; vb < total_vblocks; vb += gridDim.x) {
    const BlockCoord bc = block_coord(vb, gp);
    const int s = bc.s;
    __syncthreads();  // previous virtual block has flushed its accumulators
    stage_params<true>(par, d, alpha, beta, rho, s, L);
    for (int i = threadIdx.x; i < gp.wpb * kAccRow * (L + 1); i += blockDim.x) accs[i] = 0.0;
    const int ft = bc.fg * gp.wpb + warp;
    const bool live = ft < gp.nft;
    const int f = ft * kWarp + lane;
    const bool active = live && f < NF;
    const int fc = active ? f : NF - 1;
    const double omega0 = omega_of_period(t[(size_t)s * NF + fc]);
    const double iw0 = 1.0 / fmax(omega0, 1.0e-4);
    double* hk = hks + 2 * threadIdx.x;
    __syncthreads();  // staged parameters visible
    stage_half_k(hk, par, L, omega0);
    const int k0 = bc.ch * gp.chunk, k1 = min(K, k0 + gp.chunk);
    for (int kb = k0; kb < k1; kb += gp.tc) {
      const int n = min(gp.tc, k1 - kb);
      __syncthreads();
      fill_table<kTab>(tab, ics, par, L, c + (size_t)s * K + kb, n);
      __syncthreads();
      if (!live) continue;
      for (int cc = 0; cc < n; ++cc) {
        const int k = kb + cc;
        const Sample z = make_sample_dense(omega0, ics[cc], iw0);
        const TableSource<kTab> src{par, L, tab + (size_t)(cc * L + 1) * kTab, hk};
        const size_t idx = ((size_t)s * K + k) * NF + fc;
        // inactive lanes (period tile overhang) replay a valid sample with a zero seed
        const double seed = active ? (gdet ? gdet[idx] : 1.0) : 0.0;
        FoldSink sink{acc, lane};
        const double v = sample_forward_store(src, z, st);
        if (det && active) __stcs(det + idx, v);  // streaming store: keep the store slabs in L2
        sample_adjoint(src, z, seed, v, st, sink);
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 4 * L; i += blockDim.x) {
      const int which = i / L, m = i - which * L;  // output order [4][L]
      part[((size_t)s * gp.bpm + bc.blk) * 4 * L + i] = fold_total(accs, gp.wpb, L, m, which);
    }
  }
}

// second stage: sum block partials [S][bpm][4][L] -> gd, ga, gb, grho [S,L]
__global__ void k_grad_reduce(const double* __restrict__ part, int L, int bpm, double* __restrict__ gd,
                              double* __restrict__ ga, double* __restrict__ gb, double* __restrict__ grho) {
  const int s = blockIdx.x;
  for (int i = threadIdx.x; i < 4 * L; i += blockDim.x) {
    double v = 0.0;
    for (int b = 0; b < bpm; ++b) v += part[((size_t)s * bpm + b) * 4 * L + i];
    const int which = i / L, m = i - which * L;
    double* out = which == 0 ? gd : (which == 1 ? ga : (which == 2 ? gb : grho));
    if (out) out[(size_t)s * L + m] = v;
  }
}

// ---- ensemble optimiser step (N3) -----------------------------------------------------------------------
__device__ __forceinline__ double sgn(double v) { return v > 0.0 ? 1.0 : (v < 0.0 ? -1.0 : (v == 0.0 ? 0.0 : NAN)); }
__device__ __forceinline__ double brocher_vp(double b) {
  return 0.9409 + 2.0947 * b - 0.8206 * b * b + 0.2683 * b * b * b - 0.0251 * b * b * b * b;
}
__device__ __forceinline__ double brocher_dvp(double b) {
  return 2.0947 - 2.0 * 0.8206 * b + 3.0 * 0.2683 * b * b - 4.0 * 0.0251 * b * b * b;
}
__device__ __forceinline__ double brocher_rho(double a) {
  return 1.6612 * a - 0.4721 * a * a + 0.0671 * a * a * a - 0.0043 * a * a * a * a + 0.000106 * a * a * a * a * a;
}
__device__ __forceinline__ double brocher_drho(double a) {
  return 1.6612 - 2.0 * 0.4721 * a + 3.0 * 0.0671 * a * a - 4.0 * 0.0043 * a * a * a + 5.0 * 0.000106 * a * a * a * a;
}
// second-difference operator tridiag(-1, 2, -1) with corner entries 1 applied along an axis of length n
__device__ __forceinline__ double lap(double xm, double x0, double xp, int i, int n) {
  if (n == 1) return x0;
  if (i == 0) return x0 - xp;
  if (i == n - 1) return x0 - xm;
  return 2.0 * x0 - xm - xp;
}

// One optimiser step of the inversion loop = two launches with one thread per (model, layer):
//   k_ens_prepare  reads the CURRENT (not yet projected) b — the point where the misfit was evaluated — and forms the
//                  gradient of the data term w.r.t. b through the velocity chain (summing the adjoint kernel's block
//                  partials on the way when they are handed over unreduced), the sign fields of the two Tikhonov terms
//                  and their contribution to the misfit; snapshots b and d so that phase 2 can update in place;
//   k_ens_adam     projection (Vs clip, half-space rule, depth clip: the serial recurrences over the layers are a few
//                  cheap operations per layer and are simply recomputed by every thread up to its own layer),
//                  iterate / misfit history, Tikhonov subgradients, Adam, next Vp / rho, iteration counter.
// The iteration index, learning-rate schedule and Adam bias corrections come either from the host (lr, bc1, bc2s) or
// from a device counter, which makes the launch sequence of one iteration a replayable CUDA graph.
struct EnsArgs {
  double *b, *d, *alpha, *rho, *mb, *vb, *md, *vd;
  const double *gd, *ga, *gb, *gr;  // reduced gradients [S,L] ...
  const double *part, *loss_part;   // ... or the adjoint kernel's block partials [S][bpm][4][L], [S][bpm]
  int bpm, npicks;
  double* loss;                     // [S]: data misfit in (reduced-gradient form) / total misfit out
  const double *b0, *a0, *r0, *lob, *upb, *mind, *maxd;
  double *gt, *gdr, *sv, *sh, *tik, *bo, *dold, *shh;  // scratch: 7 x [S,L] + [2,L]
  const double* halo;               // [4,L]: Vs rows row0-2, row0-1, row0+S, row0+S+1 of a sharded ensemble (or NULL)
  int row0, S_all;                  // global index of local row 0 and global ensemble size (horizontal damping)
  int S, L, vp_mode, invert_thick, layering, proj_flags;
  double ratio, dv, dh, lr, gamma, beta1, beta2, eps, bc1, bc2s;
  int step_size;
  int* counter;                     // device: [0] iteration index j (0-based), [1] block ticket; NULL: host scalars
  int T;                            // history slots (counter form)
  double *loss_hist, *iter_b, *iter_d;  // counter form: bases of [T,S], [T,S,L]; host form: this iteration's slot
};

// what phase 1 leaves for phase 2 of the SAME (model, layer) cell: kept in registers by the one-launch form
struct EnsCell {
  double gt, gdr, x0, d0;
};
__device__ __forceinline__ EnsCell ens_prepare_cell(const EnsArgs& A, int s, int l) {
  const int L = A.L, S = A.S;
  const int i = s * L + l;
  const double x0 = A.b[i];
  double gd, ga, gb, gr;
  if (A.part) {
    gd = ga = gb = gr = 0.0;
    for (int k = 0; k < A.bpm; ++k) {
      const double* q = A.part + ((size_t)s * A.bpm + k) * 4 * L + l;
      gd += q[0];
      ga += q[L];
      gb += q[2 * L];
      gr += q[3 * L];
    }
  } else {
    gd = A.gd ? A.gd[i] : 0.0;
    ga = A.ga[i];
    gb = A.gb[i];
    gr = A.gr[i];
  }
  double g = gb;
  if (A.vp_mode == 1 || (A.vp_mode == 3 && x0 <= 4.6)) {
    g += (ga + gr * brocher_drho(A.alpha[i])) * brocher_dvp(x0);
  } else if (A.vp_mode == 2) {
    g += ga * A.ratio;
  }
  A.gt[i] = g;
  A.gdr[i] = gd;
  double tik = 0.0;
  if (A.dv > 0.0) {
    const double mv = A.dv * lap(l > 0 ? A.b[i - 1] : 0.0, x0, l < L - 1 ? A.b[i + 1] : 0.0, l, L) / L;
    tik += fabs(mv);
    A.sv[i] = sgn(mv);
  }
  if (A.dh > 0.0) {
    const int r = A.row0 + s, Sa = A.S_all;
    const double* h = A.halo;
    const double xm = s > 0 ? A.b[i - L] : (h ? h[L + l] : 0.0);
    const double xp = s < S - 1 ? A.b[i + L] : (h ? h[2 * L + l] : 0.0);
    const double mh = A.dh * lap(xm, x0, xp, r, Sa) / L;
    tik += fabs(mh);
    A.sh[i] = sgn(mh);
    if (h) {  // sign field of the two rows next to the shard (owned by the neighbouring ranks)
      if (s == 0 && r > 0) A.shh[l] = sgn(A.dh * lap(h[l], h[L + l], x0, r - 1, Sa) / L);
      if (s == S - 1 && r + 1 < Sa) A.shh[L + l] = sgn(A.dh * lap(x0, h[2 * L + l], h[3 * L + l], r + 1, Sa) / L);
    }
  }
  const double d0 = A.d[i];
  A.tik[i] = tik;
  A.bo[i] = x0;
  A.dold[i] = d0;
  return EnsCell{g, gd, x0, d0};
}

__global__ void k_ens_prepare(EnsArgs A) {
  chain_wait_then_release();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < A.S * A.L) ens_prepare_cell(A, i / A.L, i % A.L);
}

__device__ __forceinline__ double adam(double p, double g, double& m, double& v, const EnsArgs& A, double lr, double bc1,
                                       double bc2s) {
  m = m + (1.0 - A.beta1) * (g - m);              // exp_avg.lerp_(grad, 1 - beta1)
  v = A.beta2 * v + (1.0 - A.beta2) * g * g;      // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
  const double denom = sqrt(v) / bc2s + A.eps;    // bias_correction2_sqrt
  return p - (lr / bc1) * (m / denom);
}

// iteration index, learning rate (StepLR as torch chains it: lr <- lr * gamma every step_size steps) and Adam bias
// corrections of step j + 1, from the device counter or the host scalars
struct EnsStep {
  int j;
  double lr, bc1, bc2s;
};
__device__ __forceinline__ EnsStep ens_step_of(const EnsArgs& A) {
  EnsStep t{0, A.lr, A.bc1, A.bc2s};
  if (A.counter) {
    t.j = A.counter[0];
    for (int n = A.step_size > 0 ? t.j / A.step_size : 0; n > 0; --n) t.lr *= A.gamma;
    t.bc1 = 1.0 - pow(A.beta1, (double)(t.j + 1));
    t.bc2s = sqrt(1.0 - pow(A.beta2, (double)(t.j + 1)));
  }
  return t;
}

template <bool kOwn>
__device__ __forceinline__ void ens_adam_cell(const EnsArgs& A, const EnsStep& T, int s, int l, const EnsCell& own) {
  const int L = A.L, S = A.S, j = T.j;
  const int i = s * L + l;
  const size_t o = (size_t)s * L;
  const bool rec = !A.counter || j < A.T;
  const size_t ho = A.counter ? (size_t)j * S * L : 0;
  // --- projection of Vs (before the optimiser step, as the reference orders it); torch.clamp keeps NaN
  auto proj_b = [&](int ll) {
    double x = (kOwn && ll == l) ? own.x0 : A.bo[o + ll];
    if (x != x && (A.proj_flags & 2)) x = A.b0[o + ll];            // NaN repair of the ensemble vs-only loop (:624-626)
    if (x == x) x = fmin(fmax(x, A.lob[o + ll]), A.upb[o + ll]);   // clip_(min, max): max wins when min > max
    return x;
  };
  double x = proj_b(l);
  if ((A.proj_flags & 1) && l == L - 1 && L > 1) {                 // half-space >= layer above (:129/:599/:631)
    const double prev = proj_b(L - 2);
    if (x == x && prev == prev) x = fmax(x, prev);
  }
  if (A.iter_b && rec) A.iter_b[ho + i] = x;
  // --- projection of the depths: cumulative depth clipped per layer (LR) or against the previous projected
  // depth (LN), back to thicknesses >= min thickness (:133-147, :603-620)
  double th = kOwn ? own.d0 : A.dold[i];
  if (A.invert_thick) {
    const double floor_t = A.mind[0];
    double depth = 0.0, pdepth = 0.0, z = 0.0;
    for (int ll = 0; ll <= l; ++ll) {
      pdepth = z;
      depth += A.dold[o + ll];
      z = depth;
      if (A.layering == 1) {
        z = fmin(fmax(z, A.mind[o + ll]), A.maxd[o + ll]);
      } else if (A.layering == 2) {
        const double lo = ll == 0 ? A.mind[o + ll] : pdepth + A.mind[o + ll];
        z = fmin(fmax(z, lo), A.maxd[o + ll]);
      }
    }
    th = fmax(z - pdepth, floor_t);
    if (A.iter_d && rec) A.iter_d[ho + i] = th;
  }
  // --- Tikhonov subgradients on top of the chain gradient of phase 1, Adam, next alpha / rho
  double g = kOwn ? own.gt : A.gt[i];
  if (A.dv > 0.0) {
    const double c = lap(l > 0 ? A.sv[i - 1] : 0.0, A.sv[i], l < L - 1 ? A.sv[i + 1] : 0.0, l, L);
    g += A.dv / L * c;
  }
  if (A.dh > 0.0) {
    const int r = A.row0 + s, Sa = A.S_all;
    const double sm = s > 0 ? A.sh[i - L] : (A.halo && r > 0 ? A.shh[l] : 0.0);
    const double sp = s < S - 1 ? A.sh[i + L] : (A.halo && r + 1 < Sa ? A.shh[L + l] : 0.0);
    g += A.dh / L * lap(sm, A.sh[i], sp, r, Sa);
  }
  double m = A.mb[i], v = A.vb[i];
  const double xn = adam(x, g, m, v, A, T.lr, T.bc1, T.bc2s);
  A.mb[i] = m;
  A.vb[i] = v;
  A.b[i] = xn;
  if (A.vp_mode == 1 || (A.vp_mode == 3 && xn <= 4.6)) {
    // USArray (3): entries above 4.6 km/s keep their current Vp / rho, as the reference's masked in-place
    // update leaves them (_model.py:440-446; its AK135 loop writes into temporaries)
    const double an = brocher_vp(xn);
    A.alpha[i] = an;
    A.rho[i] = brocher_rho(an);
  } else if (A.vp_mode == 2) {
    A.alpha[i] = xn * A.ratio;
  }
  if (A.invert_thick) {
    double m2 = A.md[i], v2 = A.vd[i];
    A.d[i] = adam(th, kOwn ? own.gdr : A.gdr[i], m2, v2, A, T.lr, T.bc1, T.bc2s);
    A.md[i] = m2;
    A.vd[i] = v2;
  }
  if (l == 0) {  // total misfit of model s
    double tot;
    if (A.loss_part) {
      tot = 0.0;
      for (int k = 0; k < A.bpm; ++k) tot += A.loss_part[(size_t)s * A.bpm + k];
      tot /= (double)A.npicks;
    } else {
      tot = A.loss[s];
    }
    double add = 0.0;
    for (int ll = 0; ll < L; ++ll) add += A.tik[o + ll];
    tot += add;
    A.loss[s] = tot;
    if (A.loss_hist && rec) A.loss_hist[(A.counter ? (size_t)j * S : 0) + s] = tot;
  }
}

// the last block to finish advances the iteration index (every block has read it by then)
__device__ __forceinline__ void ens_advance(const EnsArgs& A, int j) {
  if (!A.counter) return;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&A.counter[1], 1) == (int)gridDim.x - 1) {
      A.counter[1] = 0;
      A.counter[0] = j + 1;
    }
  }
}

__global__ void k_ens_adam(EnsArgs A) {
  chain_wait_then_release();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const EnsStep T = ens_step_of(A);
  if (i < A.S * A.L) ens_adam_cell<false>(A, T, i / A.L, i % A.L, EnsCell{});
  ens_advance(A, T.j);
}

// both phases in one launch when no model needs another block's data (damp_horizontal == 0): a block owns
// blockDim / L whole models, a barrier separates the phases
__global__ void k_ens_fused(EnsArgs A) {
  chain_wait_then_release();
  const int L = A.L, mpb = blockDim.x / L;
  const int local = threadIdx.x / L, l = threadIdx.x - local * L;
  const int s = blockIdx.x * mpb + local;
  const bool mine = local < mpb && s < A.S;
  const EnsStep T = ens_step_of(A);
  EnsCell own{};
  if (mine) own = ens_prepare_cell(A, s, l);
  __syncthreads();
  if (mine) ens_adam_cell<true>(A, T, s, l, own);
  ens_advance(A, T.j);
}

struct PointsPlan {
  int units;   // ceil(N/32) warp units per model
  int wpb;
  int bpm;
  int global;  // 1: persistent blocks with the store slabs in the global (L2) slab
};

// Pointwise forward at the observed picks.
__global__ void __launch_bounds__(128) k_points_fwd(const double* __restrict__ d, const double* __restrict__ alpha,
                                                    const double* __restrict__ beta, const double* __restrict__ rho,
                                                    const double* __restrict__ t, const double* __restrict__ c, int L,
                                                    int N, PointsPlan pp, double* __restrict__ F) {
  extern __shared__ double smem[];
  double* par = smem + kSmemHead;
  const int s = blockIdx.x / pp.bpm;
  const int blk = blockIdx.x - s * pp.bpm;
  stage_params<false>(par, d, alpha, beta, rho, s, L);
  __syncthreads();
  const int i = blk * blockDim.x + threadIdx.x;
  if (blk * blockDim.x + (threadIdx.x & ~31) >= N) return;  // whole warps only: the sweep votes across the warp
  const size_t idx = (size_t)s * N + (i < N ? i : N - 1);   // overhang lanes replay the last pick
  const double omega0 = omega_of_period(t[idx]);
  const Sample z = make_sample(omega0, 1.0 / c[idx]);
  const double v = sample_forward(DirectSource{par, L}, z);
  if (i < N) F[idx] = v;
}

// DMF compression: returns |G| and d|G|/dF (reference _model.py:455-464)
__device__ __forceinline__ void dmf_compress(int mode, double F, double n, double invN, double& absG, double& seed) {
  if (mode == ADSURF_COMPRESS_NONE) {
    absG = fabs(F);
    seed = (F > 0.0 ? 1.0 : (F < 0.0 ? -1.0 : (F == 0.0 ? 0.0 : NAN))) * invN;
    return;
  }
  const bool norm = (mode == ADSURF_COMPRESS_EXP) || (mode == ADSURF_COMPRESS_LOG) || (mode == ADSURF_COMPRESS_NORM);
  const double in = norm ? 1.0 / n : 1.0;
  const double u = F * in;
  const double au = fabs(u);
  const double sg = (u > 0.0 ? 1.0 : (u < 0.0 ? -1.0 : (u == 0.0 ? 0.0 : NAN)));
  if (mode == ADSURF_COMPRESS_NORM) {
    absG = au;
    seed = sg * in * invN;
  } else if (mode == ADSURF_COMPRESS_LOG) {
    absG = log(au + 1.0);
    seed = sg / (au + 1.0) * in * invN;
  } else {
    const double pw = pow(0.1, au);
    absG = 1.0 - pw;  // |0.1^|u| - 1|
    // G = pw - 1 <= 0; d|G|/du = -dG/du = ln(10) pw sign(u); torch's abs'(0) = 0 when G == 0 (u == 0)
    seed = kLn10 * pw * sg * in * invN;
  }
}

// Pointwise forward + adjoint.  kind 0: upstream gF given; kind 1: Jacobian rows; kind 2: fused DMF
// (normalisation read from the dense pass partials, loss partial written per block).
template <int kKind, bool kGlobal>
__global__ void __maxnreg__(kGlobal ? 168 : 232)
k_points_bwd(const double* __restrict__ d, const double* __restrict__ alpha, const double* __restrict__ beta,
             const double* __restrict__ rho, const double* __restrict__ t, const double* __restrict__ c,
             const double* __restrict__ gF, int L, int N, PointsPlan pp, int total_vblocks, double* __restrict__ F,
             double* __restrict__ gc, double* __restrict__ part, double* __restrict__ estore,
             // Jacobian outputs
             double* __restrict__ Jd, double* __restrict__ Ja, double* __restrict__ Jb, double* __restrict__ Jrho,
             // DMF
             double* __restrict__ pmax, double* __restrict__ pmin, int nchunk, int compress, int with_grad,
             double* __restrict__ norm, double* __restrict__ loss_part,
             // fused optimiser step (DMF kind, no horizontal damping): the block that completes a model's partials runs it
             int* __restrict__ model_done, EnsArgs E) {
  chain_wait_then_release();
  // kGlobal: persistent blocks, store slabs in the L2-resident global slab, per-lane fold accumulators (large
  // ensembles); otherwise one block per (model, 32*wpb picks) with the slabs in shared memory (small problems).
  extern __shared__ double smem[];
  const int acc_per_warp = kGlobal ? kAccRow * (L + 1) : 4 * L;
  double* par = smem + kSmemHead;
  double* accs = par + par_doubles(L);                  // [wpb][L+1][4][16] or [wpb][4][L]
  double* lossw = accs + (size_t)pp.wpb * acc_per_warp; // [wpb]
  double* ests = lossw + pp.wpb;                        // [wpb][L-1][kSlot][32] (shared variant only)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t slab = (size_t)(L - 1) * kSlot * kWarp;
  double* acc = accs + (size_t)warp * acc_per_warp;
  const unsigned long long slab_policy = kGlobal ? l2_evict_last_policy() : 0ull;
  double* estp = kGlobal ? estore + ((size_t)blockIdx.x * pp.wpb + warp) * slab + lane : ests + (size_t)warp * slab + lane;
  for (int vb = blockIdx.x; vb < total_vblocks; vb += gridDim.x) {
    const int s = vb / pp.bpm;
    const int blk = vb - s * pp.bpm;
    __syncthreads();
    stage_params<false>(par, d, alpha, beta, rho, s, L);
    for (int i = threadIdx.x; i < pp.wpb * acc_per_warp + pp.wpb; i += blockDim.x) accs[i] = 0.0;
    __syncthreads();
    const int unit = blk * pp.wpb + warp;
    if (unit < pp.units) {
      const int i = unit * kWarp + lane;
      const bool active = i < N;
      const size_t idx = (size_t)s * N + (active ? i : N - 1);
      const double cc = c[idx];
      const double ic = 1.0 / cc;
      const Sample z = make_sample(omega_of_period(t[idx]), ic);
      const DirectSourceT<kKind != 2> src{par, L};  // the DMF step does not return dF/dc
      // max / min over the trial velocities from the dense pass: ordered keys [S,N], read once per pick (before the
      // sweep: their latency hides behind it) and reset to the identity for the next evaluation
      const bool normalised = kKind == 2 && (compress == ADSURF_COMPRESS_EXP || compress == ADSURF_COMPRESS_LOG ||
                                             compress == ADSURF_COMPRESS_NORM);
      unsigned long long key_max = 0ull, key_min = 0ull;
      if (normalised) {
        unsigned long long* kmax = reinterpret_cast<unsigned long long*>(pmax) + idx;
        unsigned long long* kmin = reinterpret_cast<unsigned long long*>(pmin) + idx;
        key_max = *kmax;
        key_min = *kmin;
        if (active) {
          *kmax = 0ull;
          *kmin = 0ull;
        }
      }
      typename std::conditional<kGlobal, GlobalStore, PlainStore<kWarp>>::type store;
      store.p = estp;
      if constexpr (kGlobal) store.pol = slab_policy;
      const double v = sample_forward_store(src, z, store);
      if (F && active) F[idx] = v;
      double seed = 1.0;
      if (kKind == 0) seed = gF[idx];
      if (kKind == 2) {
        double n = 1.0;
        if (normalised) {
          n = order_value(key_max) + order_value(key_min);  // max - min (the minimum travels negated)
          if (norm && active) norm[idx] = n;
        }
        double absG;
        dmf_compress(compress, v, n, 1.0 / (double)N, absG, seed);
        double lsum = warp_sum(active ? absG : 0.0);
        if (lane == 0) lossw[warp] = lsum;
      }
      if (kKind != 2 || with_grad) {
        double gk;
        if (kKind == 1) {
          JacSink sink{Jd ? Jd + idx * L : nullptr, Ja ? Ja + idx * L : nullptr, Jb ? Jb + idx * L : nullptr,
                       Jrho ? Jrho + idx * L : nullptr, active};
          gk = sample_adjoint(src, z, seed, v, store, sink);
        } else if (kGlobal) {
          FoldSink sink{acc, lane};
          gk = sample_adjoint(src, z, active ? seed : 0.0, v, store, sink);  // overhang lanes: zero seed
        } else {
          ReduceSink sink{acc, L, lane};
          gk = sample_adjoint(src, z, active ? seed : 0.0, v, store, sink);
        }
        if (gc && active) gc[idx] = -ic * z.ic * gk;  // gk = dDelta/d(k/omega), k/omega = 1/c
      }
    }
    __syncthreads();
    if (kKind != 1) {
      for (int i = threadIdx.x; i < 4 * L; i += blockDim.x) {
        double v = 0.0;
        if (kGlobal) {
          v = fold_total(accs, pp.wpb, L, i % L, i / L);
        } else {
          for (int w = 0; w < pp.wpb; ++w) v += accs[(size_t)w * 4 * L + i];
        }
        part[((size_t)s * pp.bpm + blk) * 4 * L + i] = v;
      }
      if (kKind == 2 && threadIdx.x == 0) {
        double v = 0.0;
        for (int w = 0; w < pp.wpb; ++w) v += lossw[w];
        loss_part[(size_t)s * pp.bpm + blk] = v;
      }
    }
    if (kKind == 2 && model_done) {
      // Optimiser step of model s by the block that delivers its last partials (threadFenceReduction pattern):
      // every block publishes its partials, takes a ticket, and the holder of the last ticket sees them all.
      __shared__ int last_block;
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) {
        const int ticket = atomicAdd(&model_done[s], 1);
        last_block = ticket == pp.bpm - 1;
        if (last_block) model_done[s] = 0;  // ready for the next iteration
      }
      __syncthreads();
      if (last_block) {
        __threadfence();
        const EnsStep T = ens_step_of(E);
        for (int l = threadIdx.x; l < L; l += blockDim.x) ens_prepare_cell(E, s, l);
        __syncthreads();
        for (int l = threadIdx.x; l < L; l += blockDim.x) ens_adam_cell<false>(E, T, s, l, EnsCell{});
        if (threadIdx.x == 0) {  // the last MODEL to finish advances the iteration index
          __threadfence();
          if (atomicAdd(&E.counter[1], 1) == E.S - 1) {
            E.counter[1] = 0;
            E.counter[0] = T.j + 1;
          }
        }
      }
    }
  }
}

__global__ void k_loss_reduce(const double* __restrict__ loss_part, int bpm, int N, double* __restrict__ loss, int S) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  double v = 0.0;
  for (int b = 0; b < bpm; ++b) v += loss_part[(size_t)s * bpm + b];
  loss[s] = v / (double)N;
}

// FP64 peak probe: 8 independent DFMA chains per thread
__global__ void __launch_bounds__(256) k_dfma_probe(int iters, double* __restrict__ out, double a, double b) {
  double x0 = threadIdx.x * 1e-9, x1 = x0 + 1.0, x2 = x0 + 2.0, x3 = x0 + 3.0, x4 = x0 + 4.0, x5 = x0 + 5.0,
         x6 = x0 + 6.0, x7 = x0 + 7.0;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  const double r = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (r == 12345.678) out[0] = r;  // never true; keeps the chains alive
}

// modal dispersion curves (N1): the reference's sequential root search (adsurf_dispersion.cuh), one model per WARP.
// The search is data-dependent control flow all the way down (ladder walk, Neville / bisection choices): 32 models
// sharing a warp serialise on their divergence (measured 24 x slower per warp than one model alone), so a warp carries
// one model, runs the control flow redundantly in every lane, and shares out the per-layer preparation of each
// secular-function evaluation over its lanes (DispWarpEval); the ensemble spreads over the warps.
constexpr int kDispThreads = 128;
__host__ __device__ inline size_t disp_warp_doubles(int L) { return (size_t)L * (kDispPer + kDispTerms); }
__global__ void __launch_bounds__(kDispThreads)
k_dispersion(const double* __restrict__ d, const double* __restrict__ a, const double* __restrict__ b,
             const double* __restrict__ rho, const double* __restrict__ t, int S, int L, int NF, int nmodes, double dc,
             int ifunc, double* __restrict__ cout, double* __restrict__ work, int* __restrict__ status) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s = blockIdx.x * (kDispThreads / kWarp) + warp;
  stage_exp_table();
  __syncthreads();
  if (s >= S) return;  // whole warps: the evaluator synchronises the warp only
  const DispModel M{d + (size_t)s * L, a + (size_t)s * L, b + (size_t)s * L, rho + (size_t)s * L, L, ifunc};
  DispWarpEval ev{M, kSmemHead + warp * (int)disp_warp_doubles(L), lane, 0.0};
  const int rc = disp_getc(ev, M, t + (size_t)s * NF, NF, nmodes, dc, cout + (size_t)s * nmodes * NF, work + (size_t)s * NF);
  if (lane == 0) status[s] = rc;
}

// ---- Monte-Carlo ensemble generation on the device (N2) -------------------------------------------------
// Philox4x32-10 counter-based generator (Salmon et al. 2011): counter = (model, layer, draw kind, 0), key = seed, so
// every (model, layer) draw is independent of the launch geometry and reproducible from the seed alone.
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// one standard normal per (model, layer, kind): Box-Muller on two 53-bit-spaced uniforms in (0, 1)
__device__ __forceinline__ double philox_normal(unsigned long long seed, uint32_t s, uint32_t l, uint32_t kind) {
  uint32_t r[4];
  philox4x32_10(s, l, kind, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
  const double u1 = ((double)r[0] * 4294967296.0 + (double)r[1] + 0.5) * (1.0 / 18446744073709551616.0);
  const double u2 = ((double)r[2] * 4294967296.0 + (double)r[3] + 0.5) * (1.0 / 18446744073709551616.0);
  double sn, cs;
  sincospi(2.0 * u2, &sn, &cs);
  return sqrt(-2.0 * log(u1)) * cs;
}

struct McArgs {
  const double *d, *a, *b, *rho;              // base model [L]
  const double *vs_lo, *vs_hi;                // [L]
  const double *dep_lo, *dep_hi;              // [L] depth bounds (leading zero already dropped) or NULL
  const double* ak135;                        // [nak,3] (USArray) or NULL
  double *od, *oa, *ob, *orho;                // [S,L]
  int S, L, vel_method, layering, nak;        // vel_method 1 Brocher, 2 Constant, 3 USArray
  double ratio, sigma_vs, sigma_thick;
  unsigned long long seed;
};

// one thread per model: Vs ~ N(vs_l, ((hi-lo)/sigma_vs)^2) clipped to [lo, hi]; optional depth perturbation (LR: depth
// clipped per layer; LN: against the previous perturbed depth), thickness >= dep_lo[1], last = second to last; then
// Vp / rho from the velocity rule.  Row 0 is the unperturbed model (reference _MonteCarlo.py:81-169).
__global__ void k_mc_generate(McArgs A) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= A.S) return;
  const int L = A.L;
  const size_t o = (size_t)s * L;
  double pdepth = 0.0, depth0 = 0.0;
  for (int l = 0; l < L; ++l) {
    double vs = A.b[l], th = A.d[l];
    if (s > 0) {
      const double z = philox_normal(A.seed, (uint32_t)s, (uint32_t)l, 0u);
      vs = fmin(fmax(vs + (A.vs_hi[l] - A.vs_lo[l]) / A.sigma_vs * z, A.vs_lo[l]), A.vs_hi[l]);
    }
    depth0 += A.d[l];
    if (A.dep_lo && A.layering && A.sigma_thick > 0.0 && s > 0) {
      const double z = philox_normal(A.seed, (uint32_t)s, (uint32_t)l, 1u);
      double dep = depth0 + (A.dep_hi[l] - A.dep_lo[l]) / A.sigma_thick * z;
      const double lo = (A.layering == 2 && l > 0) ? pdepth + A.dep_lo[l] : A.dep_lo[l];
      dep = fmin(fmax(dep, lo), A.dep_hi[l]);  // np.clip: the upper bound wins when lo > hi
      th = fmax(dep - pdepth, A.dep_lo[1 < L ? 1 : 0]);
      pdepth = dep;
    }
    A.ob[o + l] = vs;
    A.od[o + l] = th;
    double vp = A.a[l], rr = A.rho[l];
    if (s > 0) {
      if (A.vel_method == 1 || A.vel_method == 3) {
        vp = brocher_vp(vs);
        rr = brocher_rho(vp);
        if (A.vel_method == 3 && vs > 4.6 && A.ak135) {  // nearest AK135 entry (columns: depth, rho, vp)
          double bv = INFINITY, br = INFINITY, nvp = vp, nrr = rr;
          for (int j = 0; j < A.nak; ++j) {
            const double dv = fabs(A.ak135[3 * j + 2] - vp), dr = fabs(A.ak135[3 * j + 1] - rr);
            if (dv < bv) { bv = dv; nvp = A.ak135[3 * j + 2]; }
            if (dr < br) { br = dr; nrr = A.ak135[3 * j + 1]; }
          }
          vp = nvp;
          rr = nrr;
        }
      } else if (A.vel_method == 2) {
        vp = vs * A.ratio;
      }
    }
    A.oa[o + l] = vp;
    A.orho[o + l] = rr;
  }
  if (A.dep_lo && A.layering && A.sigma_thick > 0.0 && L > 1) A.od[o + L - 1] = A.od[o + L - 2];  // row 0 too (:119-123)
}

__global__ void k_math_probe(const double* __restrict__ x, int n, double* __restrict__ rs, double* __restrict__ ex,
                             double* __restrict__ sn, double* __restrict__ cs) {
  stage_exp_table();
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  rs[i] = rsqrt_hd(x[i]);
  ex[i] = exp_neg(x[i]);
  sincos_fast(x[i], sn + i, cs + i);
}

// ------------------------------------------------------------------------------------------
// Host-side planning
// ------------------------------------------------------------------------------------------
constexpr size_t kSmemBudget = 100 * 1024;   // per block, so that two blocks fit an SM
constexpr size_t kSmemMax = 227 * 1024;

// SM count of the current device (grids, slab sizing and work splitting follow it; nothing assumes 148)
inline int sm_count() {
  int dev = 0, nsm = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
      nsm <= 0) {
    cudaGetLastError();
    nsm = 148;
  }
  return nsm;
}

inline size_t par_bytes(int L) { return (size_t)(kSmemHead + par_doubles(L)) * sizeof(double); }

// store slab of one warp: a slot per layer above the half-space.  The dense sweep's read-ahead of the last layer
// touches the slot after it — the next warp's first slot (values unused) — so the whole store ends with one spare
// slot (slab_bytes); a spare slot per warp would push the live slabs past the persisting-L2 capacity.
inline size_t slot_bytes() { return (size_t)kSlot * kWarp * sizeof(double); }
inline size_t pair_slot_bytes() { return (size_t)kPairSlot * kWarp * sizeof(double); }  // global slabs: the larger layout
inline size_t est_bytes_per_warp(int L) { return (size_t)(L > 1 ? L - 1 : 0) * slot_bytes(); }

// warps per block for the pointwise adjoint with the slabs in shared memory; 0 if even one warp does not fit
inline int adj_wpb(int L) {
  const size_t fixed = par_bytes(L) + 64;
  const size_t per_warp = est_bytes_per_warp(L) + (size_t)4 * L * sizeof(double) + sizeof(double);
  for (int w = 4; w >= 1; w >>= 1) {
    const size_t need = fixed + (size_t)w * per_warp;
    if (need <= (w > 1 ? kSmemBudget : kSmemMax)) return w;
  }
  return 0;
}

inline size_t adj_smem(int L, int wpb) {
  return par_bytes(L) + (size_t)wpb * ((size_t)4 * L * sizeof(double) + sizeof(double)) +
         (size_t)wpb * est_bytes_per_warp(L);
}

// The adjoint kernels with global (L2-resident) slabs run 3 blocks x 4 warps per SM at <= 168 registers: the
// register file is split per scheduler (16384 registers each), so residency goes in steps of one warp per
// scheduler — 3 warps up to 170 registers per thread, 4 only at <= 128.
constexpr int kBwdWarps = 4, kBwdBlocksPerSM = 3, kFwdWarpsPerSM = 16;
inline int max_resident_adj_warps(int nsm) { return nsm * kBwdBlocksPerSM * kBwdWarps; }
inline size_t slab_bytes(int L, int nsm) {
  return (size_t)max_resident_adj_warps(nsm) * (size_t)(L > 1 ? L - 1 : 0) * pair_slot_bytes() + pair_slot_bytes();
}

inline size_t table_bytes(int L, int tc, bool adjoint) {
  return (size_t)table_doubles(L, tc, adjoint ? kTabAdj : kTabFwd) * sizeof(double) +
         (size_t)((tc + 1) & ~1) * sizeof(double);
}

inline size_t grid_fwd_smem(int L, int tc, int wpb) {
  return par_bytes(L) + table_bytes(L, tc, false) + (size_t)wpb * kWarp * 2 * sizeof(double);
}
inline size_t grid_bwd_smem(int L, int tc, int wpb) {
  return par_bytes(L) + (size_t)wpb * kAccRow * (L + 1) * sizeof(double) + table_bytes(L, tc, true) +
         (size_t)wpb * kWarp * 2 * sizeof(double);
}

// Split of the dense grid into blocks = (model, velocity chunk, period group of wpb tiles).
// A thread walks its chunk's velocities one after the other, so the run time is
//   (waves of resident blocks) x (chunk + table fills),
// which for small problems (one model x a few hundred periods: BASELINE configs 1-3) is dominated by how the
// block count rounds against the machine: the planner enumerates (warps per block, chunks per model) and keeps
// the cheapest; for large problems every split costs the same and the tie-break is ~16 waves.
inline GridPlan plan_grid_search(int S, int L, int NF, int K, bool adjoint, int nsm) {
  GridPlan best;
  best.nft = (NF + kWarp - 1) / kWarp;
  best.wpb = 0;
  double best_cost = 1e300;
  const int nft = best.nft;
  const int wmax = adjoint ? kBwdWarps : 8;
  const int per = L > 1 ? L - 1 : 1;
  // one sample of L layers in units of a 20-layer one (measured: 2 590 instructions at L = 20, 390 at L = 3)
  const double unit = 0.05 + 0.95 * per / 19.0;
  int tcmax = 8;  // table fills of up to 8 velocities, fewer for very deep models (keeps the table under 26 KB)
  while (tcmax > 1 && table_bytes(L, tcmax, adjoint) > 26 * 1024) tcmax >>= 1;
  // experiments: ADSURF_PLAN_GRID_FWD="warps per block,velocities per chunk,velocities per table fill" overrides the
  // search for the forward-only kernel (ignored when it does not fit the kernel's limits)
  if (!adjoint) {
    int w = 0, chunk = 0, tc = 0;
    const char* env = getenv("ADSURF_PLAN_GRID_FWD");
    if (env && sscanf(env, "%d,%d,%d", &w, &chunk, &tc) == 3 && w >= 1 && w <= 4 && chunk >= 1 && tc >= 1 &&
        table_bytes(L, tc < chunk ? tc : chunk, false) <= 26 * 1024) {
      if (chunk > K) chunk = K;
      best.wpb = w;
      best.nfg = (nft + w - 1) / w;
      best.chunk = chunk;
      best.nchunk = (K + chunk - 1) / chunk;
      best.tc = tc < chunk ? tc : chunk;
      best.bpm = best.nchunk * best.nfg;
      return best;
    }
  }
  for (int w = wmax < nft ? wmax : nft; w >= 1; w = (w > 4 ? 4 : w - 1)) {
    if (adjoint && w != (wmax < nft ? wmax : nft)) break;  // the adjoint kernel is compiled for 4-warp blocks
    const int nfg = (nft + w - 1) / w;
    int bps = adjoint ? kBwdBlocksPerSM : kFwdWarpsPerSM / w;
    // forward kernel, shallow models: as many velocities per fill as the table budget holds (one fill per block)
    int tcw = tcmax;
    while (!adjoint && tcw < 64 && table_bytes(L, 2 * tcw, adjoint) <= 26 * 1024) tcw <<= 1;
    const size_t smem = adjoint ? grid_bwd_smem(L, tcw, w) : grid_fwd_smem(L, tcw, w);
    const int bsm = (int)(kSmemMax / (smem + 1024));
    if (bsm < bps) bps = bsm;
    if (bps < 1) bps = 1;
    if (bps > 32) bps = 32;
    const double R = (double)bps * nsm;  // resident blocks
    const double live = (double)nft / ((double)nfg * w);  // fraction of a block's warps that own a period tile
    int prev_chunk = 0;
    for (int nchunk = 1; nchunk <= K; nchunk += (nchunk < 64 ? 1 : nchunk / 16)) {
      const int chunk = (K + nchunk - 1) / nchunk;
      if (chunk == prev_chunk) continue;
      prev_chunk = chunk;
      const int nch = (K + chunk - 1) / chunk;
      const int tc = chunk < tcw ? chunk : tcw;
      const double B = (double)S * nfg * nch;
      const double full = floor(B / R), rem = B - full * R;
      // a last wave that does not fill the machine: blocks are dealt round-robin to the SMs, so the busiest SM holds
      // ceil(rem / nsm) of them; a warp then runs faster (less contention for the FP64 pipe) but not faster than
      // its own dependent-issue latency allows (~0.55 of the time it takes at full residency)
      const double cap = adjoint ? (double)(kBwdBlocksPerSM * kBwdWarps) : (double)kFwdWarpsPerSM;
      const double waves = full + (rem > 0 ? fmin(1.0, fmax(0.55, ceil(rem / nsm) * w / cap)) : 0.0);
      const int fills = (chunk + tc - 1) / tc;
      // Two views of the same launch, in units of one 20-layer sample per warp slot:
      //  * latency (at most one wave of blocks: config 2's single model): a thread's own timeline, its chunk of
      //    samples + per fill the entries it computes (~0.04 each round) and two barriers (~0.05) + staging (~0.3);
      //  * throughput (several waves: ensembles): the instruction work of a block — its live warps' samples, per
      //    fill a round of table entries per 32 entries (~0.1) and two barriers per warp, staging once — divided
      //    over its warps.  The set-up is paid per block, not per warp: for shallow models it is a third of all
      //    instructions with one-warp blocks (config 3, measured: 2 warps x 12-20 velocities with one fill run 7 %
      //    faster than 1 x 20; larger chunks are slower, they leave the last wave emptier), and fewer resident
      //    warps issue less (6-warp blocks: 12 of 16 warp slots, measured 9 % slower).
      const double t_lat = chunk * unit + fills * (0.05 + 0.04 * ceil((double)tc * per / (w * kWarp))) + 0.3;
      const double block_work = (double)w * live * chunk * unit + fills * (0.05 * w + 0.1 * ceil((double)tc * L / kWarp)) + 0.3;
      const double occ = fmin(1.0, (double)bps * w * live / cap);  // warps that issue
      const double t_thr = block_work / w / pow(occ, 0.6);
      const double mix = fmin(1.0, fmax(0.0, B / R - 1.0));  // 0: one wave or less, 1: two waves or more
      double cost = waves * ((1.0 - mix) * t_lat + mix * t_thr);
      if (!(mix > 0.0)) cost *= 1.0 + 0.02 * (1.0 - live);  // dead warps only help with the fills
      cost *= 1.0 + 1e-3 * fabs(log2((B / R + 1e-9) / 16.0));  // tie-break: about 16 waves
      if (cost < best_cost) {
        best_cost = cost;
        best.wpb = w;
        best.nfg = nfg;
        best.tc = tc;
        best.chunk = chunk;
        best.nchunk = nch;
        best.bpm = nch * nfg;
      }
    }
  }
  return best;
}

// the search costs tens of microseconds and an inversion loop asks for the same plan every iteration
inline GridPlan plan_grid(int S, int L, int NF, int K, bool adjoint, int nsm) {
  struct Key {
    int S, L, NF, K, adjoint, nsm;
  };
  thread_local Key key[2] = {{0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 0, 0}};
  thread_local GridPlan plan[2];
  const int slot = adjoint ? 1 : 0;
  const Key k{S, L, NF, K, slot, nsm};
  if (memcmp(&k, &key[slot], sizeof(Key)) != 0) {
    plan[slot] = plan_grid_search(S, L, NF, K, adjoint, nsm);
    key[slot] = k;
  }
  return plan[slot];
}

inline PointsPlan plan_points(int S, int L, int N, bool allow_global, int nsm) {
  PointsPlan pp;
  pp.units = (N + kWarp - 1) / kWarp;
  // large ensembles: 12 resident warps per SM with the slabs in L2; small problems: slabs in shared memory
  pp.global = allow_global && (long long)S * pp.units >= (long long)nsm * 8;
  pp.wpb = pp.global ? kBwdWarps : adj_wpb(L);
  if (pp.global) {
    // a block owns one model: pick the warps per block (4, 3 or 2; always 12 resident warps per SM) that leaves the
    // fewest idle warps in a model's last block (300 picks = 10 warp units: 2 warps per block, not 4 + 4 + 2)
    int best_pad = 1 << 30;
    for (int w = kBwdWarps; w >= 2; --w) {
      const int pad = (pp.units + w - 1) / w * w - pp.units;
      if (pad < best_pad) {
        best_pad = pad;
        pp.wpb = w;
      }
    }
  }
  if (pp.wpb > pp.units) {
    int w = 1;
    while (w < pp.units) w <<= 1;
    pp.wpb = w < pp.wpb ? w : pp.wpb;
  }
  pp.bpm = pp.wpb > 0 ? (pp.units + pp.wpb - 1) / pp.wpb : 0;
  return pp;
}

inline size_t points_estore_bytes(const PointsPlan& pp, int L, int nsm) { return pp.global ? slab_bytes(L, nsm) : 0; }

inline size_t points_smem(const PointsPlan& pp, int L) {
  return pp.global ? par_bytes(L) + (size_t)pp.wpb * (kAccRow * (L + 1) + 1) * sizeof(double) : adj_smem(L, pp.wpb);
}

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

inline int check_launch() {
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? ADSURF_OK : ADSURF_E_CUDA;
}

template <typename Kern>
inline int set_smem(Kern kern, size_t bytes) {
  if (bytes > 48 * 1024) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) {
      cudaGetLastError();
      return ADSURF_E_CUDA;
    }
  }
  return ADSURF_OK;
}

// launch k_points_bwd<kKind, global?> (persistent when the slabs live in global memory)
template <int kKind>
inline int launch_points_bwd(const PointsPlan& pp, int S, int L, int N, const double* d, const double* alpha,
                             const double* beta, const double* rho, const double* t, const double* c,
                             const double* gF, double* F, double* gc, double* part, double* estore, double* Jd,
                             double* Ja, double* Jb, double* Jrho, double* pmax, double* pmin, int nchunk,
                             int compress, int with_grad, double* norm, double* loss_part, int nsm, cudaStream_t st,
                             int* model_done = nullptr, const EnsArgs* fused = nullptr) {
  EnsArgs E;
  memset(&E, 0, sizeof(E));
  if (fused) E = *fused;
  if (!fused) model_done = nullptr;
  const size_t smem = points_smem(pp, L);
  const long long total_vb = (long long)S * pp.bpm;
  if (total_vb > 0x7fffffffLL) return ADSURF_E_SHAPE;
  const int threads = pp.wpb * kWarp;
  if (pp.global) {
    auto kern = k_points_bwd<kKind, true>;
    if (set_smem(kern, smem)) return ADSURF_E_CUDA;
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem) != cudaSuccess || occ < 1) {
      cudaGetLastError();
      return ADSURF_E_CUDA;
    }
    const int cap = kBwdBlocksPerSM * kBwdWarps / pp.wpb;  // the slab is sized for 12 warps per SM
    if (occ > cap) occ = cap;
    long long grid = (long long)occ * nsm;
    if (grid > total_vb) grid = total_vb;
    launch_chain(kern, (unsigned)grid, threads, smem, st, d, alpha, beta, rho, t, c, gF, L, N, pp, (int)total_vb, F, gc,
                 part, estore, Jd, Ja, Jb, Jrho, pmax, pmin, nchunk, compress, with_grad, norm, loss_part, model_done, E);
  } else {
    auto kern = k_points_bwd<kKind, false>;
    if (set_smem(kern, smem)) return ADSURF_E_CUDA;
    launch_chain(kern, (unsigned)total_vb, threads, smem, st, d, alpha, beta, rho, t, c, gF, L, N, pp, (int)total_vb, F,
                 gc, part, (double*)nullptr, Jd, Ja, Jb, Jrho, pmax, pmin, nchunk, compress, with_grad, norm, loss_part,
                 model_done, E);
  }
  return check_launch();
}

}  // namespace

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" {

const char* adsurf_version(void) { return "adsurf_b200 0.2 (sm_100a, fp64)"; }

const char* adsurf_error_string(int code) {
  switch (code) {
    case ADSURF_OK: return "ok";
    case ADSURF_E_NULL: return "required pointer is NULL";
    case ADSURF_E_SHAPE: return "bad shape (non-positive or unsupported dimension)";
    case ADSURF_E_LAYERS: return "too many layers for the shared-memory adjoint state";
    case ADSURF_E_WORKSPACE: return "workspace missing or too small";
    case ADSURF_E_CUDA: return "CUDA runtime error";
    case ADSURF_E_ARG: return "bad argument";
    default: return "unknown error";
  }
}

int adsurf_init(int device, unsigned flags) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    cudaGetLastError();
    return ADSURF_E_ARG;
  }
  if (flags & ~(unsigned)ADSURF_INIT_PERSIST_L2) return ADSURF_E_ARG;
  if (flags & ADSURF_INIT_PERSIST_L2) {
    int cur = 0, maxb = 0;
    if (cudaGetDevice(&cur) != cudaSuccess) return ADSURF_E_CUDA;
    if (cur != device && cudaSetDevice(device) != cudaSuccess) return ADSURF_E_CUDA;
    cudaError_t e = cudaDeviceGetAttribute(&maxb, cudaDevAttrMaxPersistingL2CacheSize, device);
    if (e == cudaSuccess && maxb > 0) e = cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)maxb);
    if (cur != device) cudaSetDevice(cur);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return ADSURF_E_CUDA;
    }
  }
  return ADSURF_OK;
}

size_t adsurf_workspace_bytes(int op, int S, int L, int n, int K) {
  if (S <= 0 || L <= 0 || n <= 0) return 0;
  const int nsm = sm_count();
  switch (op) {
    case ADSURF_OP_POINTS_FWD:
    case ADSURF_OP_POINTS_JAC:
      return 0;
    case ADSURF_OP_POINTS_BWD: {
      const PointsPlan pp = plan_points(S, L, n, true, nsm);
      return align256((size_t)S * pp.bpm * 4 * L * sizeof(double)) + align256(points_estore_bytes(pp, L, nsm));
    }
    case ADSURF_OP_GRID_FWD: {
      if (K <= 0) return 0;
      const GridPlan gp = plan_grid(S, L, n, K, false, nsm);
      return 2 * align256((size_t)S * gp.nchunk * n * sizeof(double));
    }
    case ADSURF_OP_GRID_BWD: {
      if (K <= 0) return 0;
      const GridPlan gp = plan_grid(S, L, n, K, true, nsm);
      return align256((size_t)S * gp.bpm * 4 * L * sizeof(double)) + align256(slab_bytes(L, nsm));
    }
    case ADSURF_OP_DMF_STEP: {
      const PointsPlan pp = plan_points(S, L, n, true, nsm);
      size_t b = align256((size_t)S * pp.bpm * 4 * L * sizeof(double)) + align256((size_t)S * pp.bpm * sizeof(double)) +
                 align256(points_estore_bytes(pp, L, nsm));
      if (K > 0) b += 2 * align256((size_t)S * n * sizeof(double));  // ordered max / min keys of the dense pass
      return b;
    }
    case ADSURF_OP_ENSEMBLE_STEP:
      return align256((size_t)(7 * (size_t)S * L + 2 * (size_t)L) * sizeof(double));
    case ADSURF_OP_DISPERSION:
      return align256((size_t)S * n * sizeof(double));
    default:
      return 0;
  }
}

int adsurf_plan_describe(int op, int S, int L, int n, int K, int* out) {
  if (!out) return ADSURF_E_NULL;
  if (S <= 0 || L <= 0 || n <= 0) return ADSURF_E_SHAPE;
  const int nsm = sm_count();
  memset(out, 0, 8 * sizeof(int));
  out[7] = nsm;
  if (op == ADSURF_OP_GRID_FWD || op == ADSURF_OP_GRID_BWD) {
    if (K <= 0) return ADSURF_E_SHAPE;
    const GridPlan gp = plan_grid(S, L, n, K, op == ADSURF_OP_GRID_BWD, nsm);
    out[0] = gp.wpb; out[1] = gp.nfg; out[2] = gp.nchunk; out[3] = gp.chunk; out[4] = gp.tc; out[5] = gp.bpm;
    out[6] = gp.nft;
    return ADSURF_OK;
  }
  if (op == ADSURF_OP_POINTS_BWD || op == ADSURF_OP_DMF_STEP || op == ADSURF_OP_POINTS_JAC) {
    const PointsPlan pp = plan_points(S, L, n, op != ADSURF_OP_POINTS_JAC, nsm);
    out[0] = pp.wpb; out[1] = pp.units; out[5] = pp.bpm; out[6] = pp.global;
    return ADSURF_OK;
  }
  return ADSURF_E_ARG;
}

int adsurf_det_points_fwd(const double* d, const double* alpha, const double* beta, const double* rho,
                          const double* t, const double* c, int S, int L, int N, double* F, void* stream) {
  if (!d || !alpha || !beta || !rho || !t || !c || !F) return ADSURF_E_NULL;
  if (S <= 0 || L <= 0 || N <= 0 || L > ADSURF_MAX_LAYERS) return ADSURF_E_SHAPE;
  PointsPlan pp;
  pp.units = (N + kWarp - 1) / kWarp;
  pp.wpb = 4;
  pp.bpm = (N + 127) / 128;
  pp.global = 0;
  const size_t smem = par_bytes(L);
  if (set_smem(k_points_fwd, smem)) return ADSURF_E_CUDA;
  k_points_fwd<<<(unsigned)((size_t)S * pp.bpm), 128, smem, (cudaStream_t)stream>>>(d, alpha, beta, rho, t, c, L, N,
                                                                                      pp, F);
  return check_launch();
}

static int points_adjoint_common(int kind, const double* d, const double* alpha, const double* beta,
                                 const double* rho, const double* t, const double* c, const double* gF, int S, int L,
                                 int N, double* F, double* gd, double* ga, double* gb, double* grho, double* gc,
                                 double* Jd, double* Ja, double* Jb, double* Jrho, void* ws, size_t ws_bytes,
                                 cudaStream_t st) {
  if (!d || !alpha || !beta || !rho || !t || !c) return ADSURF_E_NULL;
  if (S <= 0 || L <= 0 || N <= 0) return ADSURF_E_SHAPE;
  if (L - 1 > ADSURF_MAX_ADJ_LAYERS) return ADSURF_E_LAYERS;
  const int nsm = sm_count();
  const PointsPlan pp = plan_points(S, L, N, kind == 0, nsm);  // the Jacobian entry has no workspace: shared slabs
  if (pp.wpb <= 0) return ADSURF_E_LAYERS;
  if (kind == 0) {
    const size_t part_bytes = align256((size_t)S * pp.bpm * 4 * L * sizeof(double));
    if (!ws || ws_bytes < part_bytes + align256(points_estore_bytes(pp, L, nsm))) return ADSURF_E_WORKSPACE;
    double* part = (double*)ws;
    double* estore = pp.global ? (double*)((char*)ws + part_bytes) : nullptr;
    const int rc = launch_points_bwd<0>(pp, S, L, N, d, alpha, beta, rho, t, c, gF, F, gc, part, estore, nullptr,
                                        nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, 1, nullptr, nullptr, nsm, st);
    if (rc) return rc;
    k_grad_reduce<<<S, 128, 0, st>>>(part, L, pp.bpm, gd, ga, gb, grho);
    return check_launch();
  }
  return launch_points_bwd<1>(pp, S, L, N, d, alpha, beta, rho, t, c, nullptr, F, gc, nullptr, nullptr, Jd, Ja, Jb,
                              Jrho, nullptr, nullptr, 0, 0, 1, nullptr, nullptr, nsm, st);
}

int adsurf_det_points_bwd(const double* d, const double* alpha, const double* beta, const double* rho,
                          const double* t, const double* c, const double* gF, int S, int L, int N, double* F,
                          double* gd, double* galpha, double* gbeta, double* grho, double* gc, void* ws,
                          size_t ws_bytes, void* stream) {
  if (!gF) return ADSURF_E_NULL;
  return points_adjoint_common(0, d, alpha, beta, rho, t, c, gF, S, L, N, F, gd, galpha, gbeta, grho, gc, nullptr,
                               nullptr, nullptr, nullptr, ws, ws_bytes, (cudaStream_t)stream);
}

int adsurf_det_points_jac(const double* d, const double* alpha, const double* beta, const double* rho,
                          const double* t, const double* c, int S, int L, int N, double* F, double* Jd,
                          double* Jalpha, double* Jbeta, double* Jrho, double* Jc, void* stream) {
  return points_adjoint_common(1, d, alpha, beta, rho, t, c, nullptr, S, L, N, F, nullptr, nullptr, nullptr, nullptr,
                               Jc, Jd, Jalpha, Jbeta, Jrho, nullptr, 0, (cudaStream_t)stream);
}

static int grid_fwd_launch(const double* d, const double* alpha, const double* beta, const double* rho,
                           const double* t, const double* c, int S, int L, int NF, int K, double* det, double* pmax,
                           double* pmin, int keyed, uint32_t* signbits, const GridPlan& gp, cudaStream_t st) {
  const size_t smem = grid_fwd_smem(L, gp.tc, gp.wpb);
  if (smem > kSmemMax) return ADSURF_E_SHAPE;
  if (set_smem(k_grid_fwd<128>, smem)) return ADSURF_E_CUDA;
  launch_chain(k_grid_fwd<128>, (unsigned)((size_t)S * gp.bpm), gp.wpb * kWarp, smem, st, d, alpha, beta, rho, t, c, L,
               NF, K, gp, det, pmax, pmin, keyed, signbits);
  return check_launch();
}

int adsurf_det_grid_fwd(const double* d, const double* alpha, const double* beta, const double* rho,
                        const double* t, const double* c, int S, int L, int NF, int K, double* det, double* cmax,
                        double* cmin, uint32_t* signbits, void* ws, size_t ws_bytes, void* stream) {
  if (!d || !alpha || !beta || !rho || !t || !c) return ADSURF_E_NULL;
  if (S <= 0 || L <= 0 || NF <= 0 || K <= 0 || L > ADSURF_MAX_LAYERS) return ADSURF_E_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  const GridPlan gp = plan_grid(S, L, NF, K, false, sm_count());
  if ((size_t)S * (size_t)gp.bpm > 0x7fffffffull) return ADSURF_E_SHAPE;
  double *pmax = nullptr, *pmin = nullptr;
  if (cmax || cmin) {
    const size_t one = align256((size_t)S * gp.nchunk * NF * sizeof(double));
    if (!ws || ws_bytes < 2 * one) return ADSURF_E_WORKSPACE;
    pmax = (double*)ws;
    pmin = (double*)((char*)ws + one);
  }
  const int rc = grid_fwd_launch(d, alpha, beta, rho, t, c, S, L, NF, K, det, pmax, pmin, 0, signbits, gp, st);
  if (rc) return rc;
  if (pmax) {
    const size_t total = (size_t)S * NF;
    k_minmax_reduce<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(pmax, pmin, NF, gp.nchunk, cmax, cmin, total);
    return check_launch();
  }
  return ADSURF_OK;
}

int adsurf_det_grid_bwd(const double* d, const double* alpha, const double* beta, const double* rho,
                        const double* t, const double* c, const double* gdet, int S, int L, int NF, int K,
                        double* det, double* gd, double* galpha, double* gbeta, double* grho, void* ws,
                        size_t ws_bytes, void* stream) {
  if (!d || !alpha || !beta || !rho || !t || !c) return ADSURF_E_NULL;
  if (S <= 0 || L <= 0 || NF <= 0 || K <= 0) return ADSURF_E_SHAPE;
  if (L - 1 > ADSURF_MAX_ADJ_LAYERS) return ADSURF_E_LAYERS;
  const int nsm = sm_count();
  const GridPlan gp = plan_grid(S, L, NF, K, true, nsm);
  if (gp.wpb <= 0) return ADSURF_E_LAYERS;
  const size_t part_bytes = align256((size_t)S * gp.bpm * 4 * L * sizeof(double));
  const size_t need = part_bytes + align256(slab_bytes(L, nsm));
  if (!ws || ws_bytes < need) return ADSURF_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = grid_bwd_smem(L, gp.tc, gp.wpb);
  if (smem > kSmemMax) return ADSURF_E_LAYERS;
  double* estore = (double*)((char*)ws + part_bytes);
  const long long total_vb = (long long)S * gp.bpm;
  if (total_vb > 0x7fffffffLL) return ADSURF_E_SHAPE;
  const int threads = gp.wpb * kWarp;
  auto kern = k_grid_bwd<kBwdWarps * kWarp, 168>;  // 3 x 4 = 12 warps / SM
  if (set_smem(kern, smem)) return ADSURF_E_CUDA;
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem) != cudaSuccess || occ < 1) {
    cudaGetLastError();
    return ADSURF_E_CUDA;
  }
  if (occ > kBwdBlocksPerSM) occ = kBwdBlocksPerSM;  // the slab is sized for 3 x 4 warps per SM
  long long grid = (long long)occ * nsm;
  if (grid > total_vb) grid = total_vb;
  kern<<<(unsigned)grid, threads, smem, st>>>(d, alpha, beta, rho, t, c, gdet, L, NF, K, gp, (int)total_vb, det,
                                              (double*)ws, estore);
  if (check_launch()) return ADSURF_E_CUDA;
  k_grad_reduce<<<S, 128, 0, st>>>((const double*)ws, L, gp.bpm, gd, galpha, gbeta, grho);
  return check_launch();
}

// dense pass (normalised modes) + pointwise forward/adjoint with the DMF seed; leaves the block partials in ws
struct DmfLaunch {
  PointsPlan pp;
  double *part, *loss_part;
};
static int dmf_launch(const double* d, const double* alpha, const double* beta, const double* rho, const double* t,
                      const double* c, const double* C, int S, int L, int N, int K, int compress, int with_grad,
                      double* F, double* norm, void* ws, size_t ws_bytes, bool zero_keys, cudaStream_t st,
                      DmfLaunch& out, int* model_done = nullptr, EnsArgs* fused = nullptr) {
  if (!d || !alpha || !beta || !rho || !t || !c) return ADSURF_E_NULL;
  if (compress < ADSURF_COMPRESS_NONE || compress > ADSURF_COMPRESS_NORM) return ADSURF_E_ARG;
  const bool normalised =
      (compress == ADSURF_COMPRESS_EXP || compress == ADSURF_COMPRESS_LOG || compress == ADSURF_COMPRESS_NORM);
  if (normalised && !C) return ADSURF_E_NULL;
  if (S <= 0 || L <= 0 || N <= 0 || (normalised && K <= 0)) return ADSURF_E_SHAPE;
  if (L - 1 > ADSURF_MAX_ADJ_LAYERS) return ADSURF_E_LAYERS;
  const int nsm = sm_count();
  const PointsPlan pp = plan_points(S, L, N, true, nsm);
  if (pp.wpb <= 0) return ADSURF_E_LAYERS;
  const size_t need = adsurf_workspace_bytes(ADSURF_OP_DMF_STEP, S, L, N, normalised ? K : 0);
  if (!ws || ws_bytes < need) return ADSURF_E_WORKSPACE;
  char* w = (char*)ws;
  double* part = (double*)w;
  w += align256((size_t)S * pp.bpm * 4 * L * sizeof(double));
  double* loss_part = (double*)w;
  w += align256((size_t)S * pp.bpm * sizeof(double));
  double* estore = pp.global ? (double*)w : nullptr;
  w += align256(points_estore_bytes(pp, L, nsm));
  double *pmax = nullptr, *pmin = nullptr;
  if (normalised) {
    // dense pass: max/min over the trial velocities at every observed period (constant for autograd), combined
    // across blocks as ordered keys [S,N] x 2; the pointwise kernel reads each key once and resets it to the identity
    const GridPlan gp = plan_grid(S, L, N, K, false, nsm);
    if ((size_t)S * (size_t)gp.bpm > 0x7fffffffull) return ADSURF_E_SHAPE;
    const size_t one = align256((size_t)S * N * sizeof(double));
    pmax = (double*)w;
    pmin = (double*)(w + one);
    if (zero_keys && cudaMemsetAsync(w, 0, 2 * one, st) != cudaSuccess) {
      cudaGetLastError();
      return ADSURF_E_CUDA;
    }
    const int rc = grid_fwd_launch(d, alpha, beta, rho, t, C, S, L, N, K, nullptr, pmax, pmin, 1, nullptr, gp, st);
    if (rc) return rc;
  }
  out.pp = pp;
  out.part = part;
  out.loss_part = loss_part;
  if (fused) {  // the optimiser step reads the block partials of this launch
    fused->part = part;
    fused->loss_part = loss_part;
    fused->bpm = pp.bpm;
  }
  return launch_points_bwd<2>(pp, S, L, N, d, alpha, beta, rho, t, c, nullptr, F, nullptr, part, estore, nullptr,
                              nullptr, nullptr, nullptr, pmax, pmin, 0, compress, with_grad, norm, loss_part,
                              nsm, st, model_done, fused);
}

int adsurf_dmf_step(const double* d, const double* alpha, const double* beta, const double* rho, const double* t,
                    const double* c, const double* C, int S, int L, int N, int K, int compress, int with_grad,
                    double* loss, double* F, double* norm, double* gd, double* galpha, double* gbeta, double* grho,
                    void* ws, size_t ws_bytes, void* stream) {
  if (!loss) return ADSURF_E_NULL;
  cudaStream_t st = (cudaStream_t)stream;
  DmfLaunch dl;
  const int rc = dmf_launch(d, alpha, beta, rho, t, c, C, S, L, N, K, compress, with_grad, F, norm, ws, ws_bytes, true, st,
                            dl);
  if (rc) return rc;
  k_loss_reduce<<<(S + 127) / 128, 128, 0, st>>>(dl.loss_part, dl.pp.bpm, N, loss, S);
  if (check_launch()) return ADSURF_E_CUDA;
  if (with_grad) {
    k_grad_reduce<<<S, 128, 0, st>>>(dl.part, L, dl.pp.bpm, gd, galpha, gbeta, grho);
    return check_launch();
  }
  return ADSURF_OK;
}

static int ens_scratch(EnsArgs& A, void* ws, size_t ws_bytes) {
  const size_t SL = (size_t)A.S * A.L;
  if (SL > 0x7fffffffull) return ADSURF_E_SHAPE;
  if (!ws || ws_bytes < adsurf_workspace_bytes(ADSURF_OP_ENSEMBLE_STEP, A.S, A.L, 1, 0)) return ADSURF_E_WORKSPACE;
  double* w = (double*)ws;
  A.gt = w;
  A.gdr = w + SL;
  A.sv = w + 2 * SL;
  A.sh = w + 3 * SL;
  A.tik = w + 4 * SL;
  A.bo = w + 5 * SL;
  A.dold = w + 6 * SL;
  A.shh = w + 7 * SL;
  return ADSURF_OK;
}

static int ens_launch(EnsArgs& A, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int rc = ens_scratch(A, ws, ws_bytes);
  if (rc) return rc;
  const size_t SL = (size_t)A.S * A.L;
  const int threads = 128;
  if (!(A.dh > 0.0) && A.L <= threads) {
    const int mpb = threads / A.L;
    launch_chain(k_ens_fused, (unsigned)((A.S + mpb - 1) / mpb), threads, 0, st, A);
    return check_launch();
  }
  const int blocks = (int)((SL + threads - 1) / threads);
  launch_chain(k_ens_prepare, (unsigned)blocks, threads, 0, st, A);
  if (check_launch()) return ADSURF_E_CUDA;
  launch_chain(k_ens_adam, (unsigned)blocks, threads, 0, st, A);
  return check_launch();
}

int adsurf_ensemble_adam_step(double* b, double* d, double* alpha, double* rho, const double* gd,
                              const double* galpha, const double* gbeta, const double* grho, double* loss, double* mb,
                              double* vb, double* md, double* vd, const double* b0, const double* alpha0,
                              const double* rho0, const double* lower_b, const double* upper_b,
                              const double* layer_mindepth, const double* layer_maxdepth, int S, int L, int vp_mode,
                              double vp_vs_ratio, int invert_thick, int layering, int proj_flags,
                              double damp_vertical, double damp_horizontal, double lr, double beta1, double beta2, double eps, int step,
                              double* iter_b, double* iter_d, void* ws, size_t ws_bytes, void* stream) {
  if (!b || !d || !alpha || !rho || !galpha || !gbeta || !grho || !loss || !mb || !vb || !b0 || !alpha0 || !rho0 ||
      !lower_b || !upper_b)
    return ADSURF_E_NULL;
  if (invert_thick && (!gd || !md || !vd || !layer_mindepth || !layer_maxdepth)) return ADSURF_E_NULL;
  if (S <= 0 || L <= 0 || step < 1) return ADSURF_E_SHAPE;
  if (vp_mode < 0 || vp_mode > 3 || layering < 0 || layering > 2) return ADSURF_E_ARG;
  EnsArgs A;
  memset(&A, 0, sizeof(A));
  A.b = b; A.d = d; A.alpha = alpha; A.rho = rho; A.mb = mb; A.vb = vb; A.md = md; A.vd = vd;
  A.gd = gd; A.ga = galpha; A.gb = gbeta; A.gr = grho; A.loss = loss;
  A.b0 = b0; A.a0 = alpha0; A.r0 = rho0; A.lob = lower_b; A.upb = upper_b; A.mind = layer_mindepth; A.maxd = layer_maxdepth;
  A.row0 = 0; A.S_all = S;
  A.S = S; A.L = L; A.vp_mode = vp_mode; A.invert_thick = invert_thick; A.layering = layering; A.proj_flags = proj_flags;
  A.ratio = vp_vs_ratio; A.dv = damp_vertical; A.dh = damp_horizontal; A.lr = lr; A.beta1 = beta1; A.beta2 = beta2;
  A.eps = eps;
  A.bc1 = 1.0 - pow(beta1, (double)step);
  A.bc2s = sqrt(1.0 - pow(beta2, (double)step));
  A.iter_b = iter_b; A.iter_d = iter_d;
  return ens_launch(A, ws, ws_bytes, (cudaStream_t)stream);
}

size_t adsurf_inversion_workspace_bytes(const adsurf_inversion* p) {
  if (!p || p->S <= 0 || p->L <= 0 || p->N <= 0) return 0;
  const bool normalised = (p->compress == ADSURF_COMPRESS_EXP || p->compress == ADSURF_COMPRESS_LOG ||
                           p->compress == ADSURF_COMPRESS_NORM);
  return align256(adsurf_workspace_bytes(ADSURF_OP_DMF_STEP, p->S, p->L, p->N, normalised ? p->K : 0)) +
         adsurf_workspace_bytes(ADSURF_OP_ENSEMBLE_STEP, p->S, p->L, 1, 0) + align256((size_t)p->S * sizeof(int));
}

int adsurf_inversion_step(const adsurf_inversion* p, void* ws, size_t ws_bytes, void* stream) {
  if (!p) return ADSURF_E_NULL;
  if (!p->b || !p->d || !p->alpha || !p->rho || !p->mb || !p->vb || !p->b0 || !p->alpha0 || !p->rho0 || !p->lower_b ||
      !p->upper_b || !p->counter || !p->loss)
    return ADSURF_E_NULL;
  if (p->invert_thick && (!p->md || !p->vd || !p->layer_mindepth || !p->layer_maxdepth)) return ADSURF_E_NULL;
  if (p->vp_mode < 0 || p->vp_mode > 3 || p->layering < 0 || p->layering > 2 || p->T < 0) return ADSURF_E_ARG;
  if (p->S_all < p->S || p->row0 < 0 || p->row0 + p->S > p->S_all) return ADSURF_E_ARG;
  if (p->damp_horizontal > 0.0 && p->S_all > p->S && (!p->halo || p->S < 2)) return ADSURF_E_ARG;
  const size_t need = adsurf_inversion_workspace_bytes(p);
  if (!ws || ws_bytes < need || need == 0) return ADSURF_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t ens_bytes = adsurf_workspace_bytes(ADSURF_OP_ENSEMBLE_STEP, p->S, p->L, 1, 0);
  const size_t done_bytes = align256((size_t)p->S * sizeof(int));
  const size_t dmf_bytes = need - ens_bytes - done_bytes;
  EnsArgs A;
  memset(&A, 0, sizeof(A));
  A.b = p->b; A.d = p->d; A.alpha = p->alpha; A.rho = p->rho; A.mb = p->mb; A.vb = p->vb; A.md = p->md; A.vd = p->vd;
  A.npicks = p->N; A.loss = p->loss;
  A.b0 = p->b0; A.a0 = p->alpha0; A.r0 = p->rho0; A.lob = p->lower_b; A.upb = p->upper_b;
  A.mind = p->layer_mindepth; A.maxd = p->layer_maxdepth;
  A.halo = p->S_all > p->S ? p->halo : nullptr; A.row0 = p->row0; A.S_all = p->S_all;
  A.S = p->S; A.L = p->L; A.vp_mode = p->vp_mode; A.invert_thick = p->invert_thick; A.layering = p->layering;
  A.proj_flags = p->proj_flags;
  A.ratio = p->vp_vs_ratio; A.dv = p->damp_vertical; A.dh = p->damp_horizontal; A.lr = p->lr; A.gamma = p->gamma;
  A.beta1 = p->beta1; A.beta2 = p->beta2; A.eps = p->eps; A.step_size = p->step_size;
  A.counter = p->counter; A.T = p->T; A.loss_hist = p->loss_hist; A.iter_b = p->iter_b_hist; A.iter_d = p->iter_d_hist;
  const int rs = ens_scratch(A, (char*)ws + dmf_bytes, ens_bytes);
  if (rs) return rs;
  // Small problems (single-model inversions: fewer blocks than SMs) are latency-bound, and there a launch is worth
  // saving: the block of the pointwise kernel that delivers a model's last partials runs that model's optimiser
  // step in the same launch (ticket counters model_done[S], zero between calls).  For ensembles the optimiser step is
  // a launch of its own, parallel over all (model, layer) cells — run inside the pointwise kernel it would
  // serialise ~10 us of dependent loads per model behind that model's block (measured: 1024 models 3 870 instead of
  // 4 560 it/s) — and with horizontal damping two launches (the neighbours' sign field is a grid-wide dependency).
  // (The order keys of the dense pass and the counters are reset by their readers: ws zero-filled before the first
  // call of a loop.)
  const PointsPlan ppq = plan_points(p->S, p->L, p->N, true, sm_count());
  const bool in_kernel = !(p->damp_horizontal > 0.0) && (long long)p->S * ppq.bpm <= sm_count();
  int* model_done = (int*)((char*)ws + dmf_bytes + ens_bytes);
  DmfLaunch dl;
  // the launches of an iteration are a dependent chain: programmatic dependent launch when ADSURF_PDL=1
  static const bool pdl = []() { const char* e = getenv("ADSURF_PDL"); return e && e[0] == '1'; }();
  struct Scope {
    explicit Scope(bool on) { t_chain_pdl = on; }
    ~Scope() { t_chain_pdl = false; }
  } scope(pdl);
  const int rc = dmf_launch(p->d, p->alpha, p->b, p->rho, p->t, p->c, p->C, p->S, p->L, p->N, p->K, p->compress, 1,
                            nullptr, nullptr, ws, dmf_bytes, false, st, dl, in_kernel ? model_done : nullptr,
                            in_kernel ? &A : nullptr);
  if (rc || in_kernel) return rc;
  A.part = dl.part; A.loss_part = dl.loss_part; A.bpm = dl.pp.bpm;
  return ens_launch(A, (char*)ws + dmf_bytes, ens_bytes, st);
}

int adsurf_dispersion_curves(const double* d, const double* alpha, const double* beta, const double* rho,
                             const double* t, int S, int L, int NF, int nmodes, double dc, int ifunc, double* c,
                             int* status, void* ws, size_t ws_bytes, void* stream) {
  if (!d || !alpha || !beta || !rho || !t || !c || !status) return ADSURF_E_NULL;
  if (S <= 0 || L <= 0 || NF <= 0 || nmodes <= 0) return ADSURF_E_SHAPE;
  if (!(dc > 0.0) || (ifunc != 1 && ifunc != 2)) return ADSURF_E_ARG;
  if (!ws || ws_bytes < (size_t)S * NF * sizeof(double)) return ADSURF_E_WORKSPACE;
  const int wpb = kDispThreads / kWarp;  // one model per warp
  const long long blocks = ((long long)S + wpb - 1) / wpb;
  if (blocks > 0x7fffffffLL) return ADSURF_E_SHAPE;
  const size_t smem = (kSmemHead + wpb * disp_warp_doubles(L)) * sizeof(double);
  if (smem > 227 * 1024) return ADSURF_E_SHAPE;  // more than 360 layers
  if (const int rc = set_smem(k_dispersion, smem)) return rc;
  k_dispersion<<<(unsigned)blocks, kDispThreads, smem, (cudaStream_t)stream>>>(d, alpha, beta, rho, t, S, L, NF, nmodes,
                                                                              dc, ifunc, c, (double*)ws, status);
  return check_launch();
}

int adsurf_mc_generate(const double* d, const double* alpha, const double* beta, const double* rho, const double* vs_lower,
                       const double* vs_upper, const double* depth_lower, const double* depth_upper,
                       const double* ak135, int n_ak135, int S, int L, int vel_method, double vp_vs_ratio,
                       double sigma_vs, double sigma_thick, int layering, unsigned long long seed, double* out_d,
                       double* out_alpha, double* out_beta, double* out_rho, void* stream) {
  if (!d || !alpha || !beta || !rho || !vs_lower || !vs_upper || !out_d || !out_alpha || !out_beta || !out_rho)
    return ADSURF_E_NULL;
  if (S <= 0 || L <= 0) return ADSURF_E_SHAPE;
  if (vel_method < 1 || vel_method > 3 || layering < 0 || layering > 2 || !(sigma_vs > 0.0)) return ADSURF_E_ARG;
  if ((depth_lower == nullptr) != (depth_upper == nullptr)) return ADSURF_E_NULL;
  McArgs A;
  A.d = d; A.a = alpha; A.b = beta; A.rho = rho; A.vs_lo = vs_lower; A.vs_hi = vs_upper;
  A.dep_lo = depth_lower; A.dep_hi = depth_upper; A.ak135 = ak135; A.nak = ak135 ? n_ak135 : 0;
  A.od = out_d; A.oa = out_alpha; A.ob = out_beta; A.orho = out_rho;
  A.S = S; A.L = L; A.vel_method = vel_method; A.layering = layering;
  A.ratio = vp_vs_ratio; A.sigma_vs = sigma_vs; A.sigma_thick = sigma_thick; A.seed = seed;
  k_mc_generate<<<(S + 127) / 128, 128, 0, (cudaStream_t)stream>>>(A);
  return check_launch();
}

int adsurf_math_probe(const double* x, int n, double* out_rsqrt, double* out_expneg, double* out_sin,
                      double* out_cos, void* stream) {
  if (!x || !out_rsqrt || !out_expneg || !out_sin || !out_cos) return ADSURF_E_NULL;
  if (n <= 0) return ADSURF_E_SHAPE;
  k_math_probe<<<(n + 255) / 256, 256, kSmemHead * sizeof(double), (cudaStream_t)stream>>>(x, n, out_rsqrt, out_expneg, out_sin, out_cos);
  return check_launch();
}

int adsurf_fp64_peak_probe(int iters, int repeats, double* tflops, double* ms, void* stream) {
  if (!tflops) return ADSURF_E_NULL;
  if (iters <= 0 || repeats <= 0) return ADSURF_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  double* out = nullptr;
  if (cudaMalloc(&out, sizeof(double)) != cudaSuccess) return ADSURF_E_CUDA;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int blocks = sm_count() * 8, threads = 256;
  k_dfma_probe<<<blocks, threads, 0, st>>>(iters, out, 0.999999, 1e-9);  // warm-up
  float best = 1e30f;
  for (int r = 0; r < repeats; ++r) {
    cudaEventRecord(e0, st);
    k_dfma_probe<<<blocks, threads, 0, st>>>(iters, out, 0.999999, 1e-9);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float t = 0.f;
    cudaEventElapsedTime(&t, e0, e1);
    if (t < best) best = t;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  if (cudaGetLastError() != cudaSuccess) return ADSURF_E_CUDA;
  const double flops = 2.0 * 8.0 * 16.0 * (double)iters * (double)blocks * (double)threads;
  *tflops = flops / ((double)best * 1e-3) / 1e12;
  if (ms) *ms = best;
  return ADSURF_OK;
}

}  // extern "C"
