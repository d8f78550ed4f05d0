/*
 * adsurf_b200.h — C ABI of the B200-native Rayleigh secular-function ("determinant") path.
 *
 * Drop-in boundary for the hot path of liufeng2317/ADsurf.  The reference has no
 * FFI: its boundary is a set of Python module functions.  Each entry point below names the
 * reference function (file:line under /root/reference/ADsurf) whose arithmetic it replaces; the
 * Python shims in adsurf_b200/_cps/ keep the reference signatures and call these through ctypes
 * (INTEGRATION.md shows the binding).
 *
 * Conventions
 *   - every array pointer is a DEVICE pointer to float64, row-major, contiguous, caller-owned;
 *   - S = models, L = layers (last = half-space), N = observed picks, NF = periods, K = trial
 *     phase velocities;  model arrays d, alpha, beta, rho are [S,L];
 *   - no water layer (reference llw = -1) and Rayleigh/Dunkin (reference ifunc = 2) only — the
 *     only combination the reference's torch path can execute;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, no host sync;
 *   - `ws` is scratch of at least adsurf_workspace_bytes(...) bytes (may be NULL when that is 0);
 *   - outputs are overwritten, never accumulated; the library keeps no state between calls and is re-entrant
 *     (calls on distinct streams may run from distinct host threads).  The ONLY device-wide effect is opt-in:
 *     adsurf_init(device, ADSURF_INIT_PERSIST_L2), see below;
 *   - grids, slab sizes and work splits follow the SM count of the CURRENT device at call time (query and
 *     launch on the same device);
 *   - return value: ADSURF_OK (0) or a negative ADSURF_E_* code; never throws, NaN/Inf inputs
 *     propagate to the outputs (the reference's host loop repairs them, _model.py:624-626).
 */
#ifndef ADSURF_B200_H
#define ADSURF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ADSURF_OK 0
#define ADSURF_E_NULL -1     /* required pointer is NULL */
#define ADSURF_E_SHAPE -2    /* non-positive or unsupported dimension */
#define ADSURF_E_LAYERS -3   /* L too large for the shared-memory adjoint state (L-1 > ADSURF_MAX_ADJ_LAYERS) */
#define ADSURF_E_WORKSPACE -4 /* ws too small / NULL */
#define ADSURF_E_CUDA -5     /* CUDA runtime error (launch failure, bad device pointer...) */
#define ADSURF_E_ARG -6      /* bad enum / flag */

#define ADSURF_MAX_LAYERS 4096   /* forward kernels: model parameters must fit shared memory */
#define ADSURF_MAX_ADJ_LAYERS 96 /* adjoint kernels: 72 B of state per (thread, layer), one warp per block minimum */

/* operation ids for adsurf_workspace_bytes */
enum adsurf_op {
  ADSURF_OP_POINTS_FWD = 0,
  ADSURF_OP_POINTS_BWD = 1,
  ADSURF_OP_POINTS_JAC = 2,
  ADSURF_OP_GRID_FWD = 3,
  ADSURF_OP_GRID_BWD = 4,
  ADSURF_OP_DMF_STEP = 5,
  ADSURF_OP_ENSEMBLE_STEP = 6, /* adsurf_ensemble_adam_step: n, K ignored */
  ADSURF_OP_DISPERSION = 7     /* adsurf_dispersion_curves: n = NF */
};

/* DMF compression (reference _model.py:450-462) */
enum adsurf_compress {
  ADSURF_COMPRESS_NONE = 0,       /* compress=False: G = F (no normalisation) */
  ADSURF_COMPRESS_EXP = 1,        /* G = 0.1^|F/n| - 1,  n = max_k det - min_k det */
  ADSURF_COMPRESS_LOG = 2,        /* G = log(|F/n| + 1) */
  ADSURF_COMPRESS_EXP_NONORM = 3, /* compress=True, normalized=False: G = 0.1^|F| - 1 */
  ADSURF_COMPRESS_NORM = 4        /* G = F/n only (Monte-Carlo pre-screen, _MonteCarlo.py:188-189) */
};

const char* adsurf_version(void);
const char* adsurf_error_string(int code);

/* Optional, explicit per-device set-up.  flags:
 *   ADSURF_INIT_PERSIST_L2  raise cudaLimitPersistingL2CacheSize of `device` to the maximum the part allows.  The
 *     adjoint kernels keep their per-warp layer slabs (43.8 KB per resident warp at 20 layers, 78 MB per GPU) in L2
 *     with evict_last accesses; L2 honours that policy only inside the persisting carve-out, which is 0 bytes by
 *     default.  The limit is CONTEXT-WIDE (it is shared with every other library in the process, though it changes
 *     nothing for code that issues no persisting accesses), hence opt-in: without it everything still works, part
 *     of the slabs is merely written back to HBM (measured on B200: ~10x the DRAM traffic of adsurf_det_grid_bwd at
 *     the same run time within 0.2 %: the path is FP64-bound, not bandwidth-bound).
 * Nothing else in this library touches device-wide state.  No reference counterpart. */
#define ADSURF_INIT_PERSIST_L2 1u
int adsurf_init(int device, unsigned flags);

/* Scratch bytes needed by an operation.  n = N (points ops) or NF (grid ops); K ignored by points ops
 * except DMF_STEP. */
size_t adsurf_workspace_bytes(int op, int S, int L, int n, int K);

/* How an operation would be split over the current device (diagnostics, tests, documentation): out[8] =
 * grid ops:   {warps per block, period groups, velocity chunks, velocities per chunk, velocities per table fill,
 *              blocks per model, period tiles, SM count};
 * points ops: {warps per block, warp units per model, 0, 0, 0, blocks per model, slabs in L2 (1) or shared (0), SMs}. */
int adsurf_plan_describe(int op, int S, int L, int n, int K, int* out);

/* F[s,i] = Delta_R(model s; t[s,i], c[s,i]).
 * Replaces dltar_vector / dltar4_vector: _cps/_surf96_vectorAll_gpu.py:194,303-399 (batched),
 * _cps/_surf96_vector_gpu.py:194,303-399 and _cps/_surf96_vector.py (S = 1). */
int adsurf_det_points_fwd(const double* d, const double* alpha, const double* beta, const double* rho,
                          const double* t, const double* c, int S, int L, int N,
                          double* F, void* stream);

/* Reverse mode of the above: given gF[S,N] returns gd, galpha, gbeta, grho [S,L] (reduced over the
 * picks of each model) and gc[S,N] (may be NULL).  Also returns F when F != NULL.
 * Replaces torch.autograd over dltar4_vector (caller: loss.backward(), _model.py:592). */
int adsurf_det_points_bwd(const double* d, const double* alpha, const double* beta, const double* rho,
                          const double* t, const double* c, const double* gF, int S, int L, int N,
                          double* F, double* gd, double* galpha, double* gbeta, double* grho, double* gc,
                          void* ws, size_t ws_bytes, void* stream);

/* Per-sample Jacobian rows: Jd, Jalpha, Jbeta, Jrho [S,N,L], Jc [S,N] (= dF/dc), F [S,N] (any may be
 * NULL).  Replaces autograd.grad(F, x, grad_outputs=eye, is_grads_batched=True) of the sensitivity
 * notebook (examples/02_2 cell 12). */
int adsurf_det_points_jac(const double* d, const double* alpha, const double* beta, const double* rho,
                          const double* t, const double* c, int S, int L, int N,
                          double* F, double* Jd, double* Jalpha, double* Jbeta, double* Jrho, double* Jc,
                          void* stream);

/* det[s,k,f] = Delta_R(model s; t[s,f], c[s,k])  (c is dim 1, f fastest — the reference layout).
 * det, cmax/cmin [S,NF] (max/min over k, fused) and signbits may each be NULL.
 * signbits: uint32 [S, K, ceil(NF/32)], bit (f%32) of word [s,k,f/32] set when det < 0.
 * Replaces dltar_matrix / dltar4_matrix: _cps/_surf96_matrixAll_gpu.py:183,242-353 (batched),
 * _cps/_surf96_matrix_gpu.py:186,245-342 and _cps/_surf96_matrix.py (S = 1, output [K,NF]),
 * plus torch.max/min over dim 1 at _model.py:455. */
int adsurf_det_grid_fwd(const double* d, const double* alpha, const double* beta, const double* rho,
                        const double* t, const double* c, int S, int L, int NF, int K,
                        double* det, double* cmax, double* cmin, uint32_t* signbits,
                        void* ws, size_t ws_bytes, void* stream);

/* Fused forward + adjoint on the dense grid: gd, galpha, gbeta, grho [S,L] = sum over (k,f) of
 * gdet[s,k,f] * dDelta/dparam.  gdet == NULL means gdet = 1 everywhere.  det (may be NULL) receives
 * the forward values.  Replaces dltar4_matrix(...).backward(gdet). */
int adsurf_det_grid_bwd(const double* d, const double* alpha, const double* beta, const double* rho,
                        const double* t, const double* c, const double* gdet, int S, int L, int NF, int K,
                        double* det, double* gd, double* galpha, double* gbeta, double* grho,
                        void* ws, size_t ws_bytes, void* stream);

/* One fused Determinant-Misfit-Function evaluation with gradient (a1+a4+a6 of SURVEY.md §8a):
 *   n[s,i]  = max_k Delta(s; t[s,i], C[s,k]) - min_k ...         (dense pass, constant for autograd)
 *   G[s,i]  = compress(Delta(s; t[s,i], c[s,i]) / n[s,i])
 *   loss[s] = (1/N) sum_i |G[s,i]|                                (data term only)
 * and d loss[s] / d(d, alpha, beta, rho)[s,:] for backward seed 1 per model (_model.py:592).
 * F (may be NULL) receives the raw pointwise determinants, norm (may be NULL) n[s,i].
 * with_grad = 0 skips the adjoint (Monte-Carlo pre-screen, _MonteCarlo.py:184-190).
 * Replaces multi_cal_grad.forward + backward (_model.py:448-464,592), iter_cal_grad (:320-337). */
int adsurf_dmf_step(const double* d, const double* alpha, const double* beta, const double* rho,
                    const double* t, const double* c, const double* C, int S, int L, int N, int K,
                    int compress, int with_grad,
                    double* loss, double* F, double* norm,
                    double* gd, double* galpha, double* gbeta, double* grho,
                    void* ws, size_t ws_bytes, void* stream);

/* Register-resident DFMA microbenchmark: runs `iters` dependent-chain iterations on every SM and
 * returns the achieved FP64 rate in TFLOP/s (2 flops per DFMA) in *tflops.  Synchronises the stream.
 * Used by bench.py for the FP64 roofline denominator (MEASURED_PEAKS.json has no FP64 entry). */
int adsurf_fp64_peak_probe(int iters, int repeats, double* tflops, double* ms, void* stream);

/* Velocity-model chain and optimiser step of the ensemble inversion loop, on the device.
 * One call finishes iteration j after adsurf_dmf_step produced the data misfit `loss` [S] and its gradients
 * gd, galpha, gbeta, grho [S,L] at the current (d, alpha, beta, rho):
 *   1. total misfit: loss[s] += damp_v sum_l |(L_v b_s)_l| / L + damp_h sum_l |(L_h b)_{s,l}| / L   (_model.py:466-477)
 *   2. dloss/db through the velocity chain (vp_mode 1 Brocher: alpha = P4(b), rho = P5(alpha); 2 Constant:
 *      alpha = ratio b; 3 USArray: Brocher where b <= 4.6; 0: alpha, rho fixed) plus the Tikhonov subgradients
 *   3. parameter projection done by the reference BEFORE the optimiser step (_model.py:593-632, :122-159): clip b
 *      to [lower_b, upper_b]; proj_flags bit 0: b[:,L-1] >= b[:,L-2] (ensemble loops and the single-model
 *      vs-and-thick loop), bit 1: NaN entries reset to b0 first (ensemble vs-only loop); with invert_thick the cumulative depths
 *      are clipped (layering 1 = LR, 2 = LN, 0 = none) and turned back into thicknesses >= min_thick;
 *      the projected iterates are recorded in iter_b / iter_d (may be NULL)
 *   4. Adam (torch.optim.Adam semantics, no weight decay/amsgrad) with learning rate `lr` and step count
 *      `step` (1-based) on b, and on d when invert_thick
 *   5. alpha, rho for the next iteration from the updated b.
 * b, d, alpha, rho and the Adam moments mb, vb, md, vd [S,L] are updated in place; b0, alpha0, rho0 are the
 * initial model (NaN repair / fixed values); ws >= adsurf_workspace_bytes(ADSURF_OP_ENSEMBLE_STEP, S, L, 1, 0).
 * vp_mode 3 (USArray): entries whose updated Vs exceeds 4.6 km/s keep their current Vp / rho (the reference's
 * masked in-place update, _model.py:440-446).  Replaces the torch autograd chain, torch.optim.Adam and the numpy
 * clipping round trip of multi_inversion (_model.py:589-636). */
int adsurf_ensemble_adam_step(double* b, double* d, double* alpha, double* rho,
                              const double* gd, const double* galpha, const double* gbeta, const double* grho,
                              double* loss, double* mb, double* vb, double* md, double* vd,
                              const double* b0, const double* alpha0, const double* rho0,
                              const double* lower_b, const double* upper_b,
                              const double* layer_mindepth, const double* layer_maxdepth,
                              int S, int L, int vp_mode, double vp_vs_ratio, int invert_thick, int layering,
                              int proj_flags, double damp_vertical, double damp_horizontal,
                              double lr, double beta1, double beta2, double eps, int step,
                              double* iter_b, double* iter_d, void* ws, size_t ws_bytes, void* stream);

/* One WHOLE inversion iteration, replayable: adsurf_dmf_step (without its two reduction launches) followed by the
 * optimiser step above, with the iteration index j read from DEVICE memory — so the launch sequence can be captured
 * once in a CUDA graph and replayed `T` times with no host work in between.  Launches per iteration: the dense pass
 * (normalised modes only) + the pointwise forward/adjoint kernel + the optimiser step (one launch; two with
 * damp_horizontal > 0: the sign field of the neighbouring models is a grid-wide dependency; none for problems with
 * fewer blocks than SMs, where the block that delivers a model's last partial sums runs its optimiser step).
 *   counter[0] (device int) = j, the 0-based index of the iteration about to run; the last kernel increments it.
 *               counter[1] is scratch (must be 0 at the first call).
 *   learning rate of iteration j: lr * gamma^(j / step_size) (torch StepLR stepped after every optimiser step);
 *   Adam bias corrections use step j + 1;
 *   loss[S] receives the total misfit of iteration j; loss_hist [T,S], iter_b_hist / iter_d_hist [T,S,L] (each may
 *   be NULL) receive row j while j < T (what the reference appends to its Python lists, _model.py:597-636);
 *   sharded ensembles (one process per GPU, this rank owns global rows row0 .. row0+S-1 of S_all): with
 *   damp_horizontal > 0 `halo` [4,L] holds the CURRENT Vs rows row0-2, row0-1, row0+S, row0+S+1 (rows outside
 *   [0, S_all) are ignored) — the only data a rank needs from its neighbours (reference :471-473); S >= 2 then.
 *   ws (adsurf_inversion_workspace_bytes) must be ZERO-FILLED before the first call of a loop and left alone
 *   between calls: the dense pass combines its per-block max / min in it with atomics and each call leaves it reset.
 * Replaces one trip of the loops at _model.py:118-186 (iter_inversion, S = 1) and :589-636 (multi_inversion). */
typedef struct adsurf_inversion {
  int S, L, N, K, T;
  const double *t, *c, *C;                           /* picks [S,N] (period, velocity), trial velocities [S,K] */
  double *b, *d, *alpha, *rho, *mb, *vb, *md, *vd;   /* state [S,L], updated in place */
  const double *b0, *alpha0, *rho0, *lower_b, *upper_b, *layer_mindepth, *layer_maxdepth; /* [S,L] */
  double *loss, *loss_hist, *iter_b_hist, *iter_d_hist;
  int* counter;
  const double* halo;
  int row0, S_all;
  int compress, vp_mode, invert_thick, layering, proj_flags, step_size;
  double vp_vs_ratio, damp_vertical, damp_horizontal, lr, gamma, beta1, beta2, eps;
} adsurf_inversion;
size_t adsurf_inversion_workspace_bytes(const adsurf_inversion* p);
int adsurf_inversion_step(const adsurf_inversion* p, void* ws, size_t ws_bytes, void* stream);

/* Modal Rayleigh phase-velocity curves by the classical root search, for S models at once:
 * c[s,n,k] = phase velocity of mode n (0 = fundamental .. nmodes-1) of model s at period t[s,k], 0 where the mode does
 * not exist.  One warp per model (the search is divergent control flow throughout; the lanes share the per-layer
 * work of every secular-function evaluation) runs the reference's own sequential algorithm — start 10 % below the slowest
 * layer's half-space Rayleigh velocity, walk a `dc` ladder from the previous period's root (and above the previous
 * mode's root) to the first sign change of the normalised secular function, refine by Neville interpolation with
 * bisection safeguards to 1e-6 relative — with identical bookkeeping (clow, one = 1e-2, onea = 1.5), so mode indices
 * and the zeros for absent modes are the reference's and the roots agree with it to its own iterates (1e-6 km/s).
 * Periods are searched in the order given (the reference passes them ascending).  status[s] = 1 when the fundamental
 * mode cannot be bracketed (the reference raises DispersionError), else 0.  ws >= S*NF doubles.
 * Replaces _surf96.surf96 / getc / getsol / nevill / gtsolh (_surf96.py:452-760) behind cal_dispersion_curve
 * (_ADsurf.py:232-310), phase velocity (itype = 0); group velocity is two calls at t/(1 +- dt) on the host.
 * ifunc as in the reference (_utils.py:20-23): 2 = Rayleigh waves, Dunkin's matrix (dltar4; what cal_dispersion_curve
 * uses), 1 = Love waves (dltar1, _surf96.py:168-215: the same search on the SH secular function; alpha only enters the
 * starting velocity); 3 (fast delta matrix) is ADSURF_E_ARG.  beta[s,0] <= 0 marks layer 0 of model s as water (the
 * reference's llw = 0, :608, :284-295): the solid stack ends above it and the water column closes the equation. */
int adsurf_dispersion_curves(const double* d, const double* alpha, const double* beta, const double* rho,
                             const double* t, int S, int L, int NF, int nmodes, double dc, int ifunc,
                             double* c, int* status, void* ws, size_t ws_bytes, void* stream);

/* Monte-Carlo ensemble of initial models, generated on the device: row 0 = the base model
 * (d, alpha, beta, rho [L]); rows 1..S-1: Vs_l ~ Normal(beta_l, ((vs_upper_l - vs_lower_l) / sigma_vs)^2) clipped to
 * [vs_lower_l, vs_upper_l]; with depth bounds (depth_lower/upper [L], leading zero already dropped as the reference
 * does) and sigma_thick > 0 the cumulative depths are perturbed the same way and clipped (layering 1 = LR: per layer;
 * 2 = LN: at least depth_lower_l below the previous perturbed depth), thickness >= depth_lower[1], last = second to
 * last; Vp / rho by vel_method (1 Brocher, 2 Constant ratio with rho kept, 3 USArray: Brocher, nearest AK135 entry
 * [n_ak135,3] = (depth, rho, vp) above 4.6 km/s).  Random numbers: Philox4x32-10 keyed by `seed`, counter = (model,
 * layer, draw kind) — reproducible and independent of the launch geometry, but NOT the numpy Mersenne-Twister stream
 * of the reference: ensembles agree with the reference's in distribution, not draw by draw (the host generator
 * `ADsurf_MonteCarlo(..., device=None)` keeps numpy's stream for seeded parity).
 * Replaces the generation half of ADsurf_MonteCarlo (_MonteCarlo.py:81-169). */
int adsurf_mc_generate(const double* d, const double* alpha, const double* beta, const double* rho,
                       const double* vs_lower, const double* vs_upper,
                       const double* depth_lower, const double* depth_upper,
                       const double* ak135, int n_ak135,
                       int S, int L, int vel_method, double vp_vs_ratio, double sigma_vs, double sigma_thick,
                       int layering, unsigned long long seed,
                       double* out_d, double* out_alpha, double* out_beta, double* out_rho, void* stream);

/* Validation hook: evaluates the kernels' inline elementary functions on the device,
 * out_rsqrt[i] = 1/sqrt(x[i]), out_expneg[i] = exp(-x[i]), out_sin/out_cos[i] = sin/cos(x[i]).
 * Used by the GPU tests to bound their error against libm (no reference counterpart). */
int adsurf_math_probe(const double* x, int n, double* out_rsqrt, double* out_expneg, double* out_sin,
                      double* out_cos, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ADSURF_B200_H */
