// One GaussianFit.fit() for one spot, written once for two executors:
//   * WarpExec (fit_kernels.cu): the 32 lanes of a warp; a batch of 32 voxels is one voxel per lane,
//     the normal-equation sums are FP64 tensor-core tiles (pass_fused_mma).  The fit engine drives lmder
//     with the register / shuffle algebra of lm_warp.h; the standalone GaussianFit kernel
//     (k_generic_fit) uses run_lm below with the generic shared-memory algebra of lm_core.h;
//   * SerialExec (tests/hostsim): one "lane" on the CPU, used to debug the numerics in a
//     container without a GPU.  It is test infrastructure, never a product fallback.
// Follows GaussianFit.__init__ / fit / to_natural_paramaters of External/Fitting_v4.py:166-393
// (and Fitting_v3.py:51-254).
#pragma once
#include "gauss_model.h"
#include "lm_core.h"

namespace ia3 {

// Per-spot scratch that must be visible to every lane (shared memory on the device).
constexpr int GRAM_PITCH = 36;    // doubles per column: lanes write consecutive words, fragment reads hit 16 distinct banks

template <typename T>
struct SpotShared {
  LMState st;
  VoxConsts<T> vc;
  double Ag[NTRI + NP];   // J^T J (packed upper triangle) followed by J^T f
  double etab[NEXP];      // exp table of the parameter transforms (one slot per lane)
  double scal[NSCAL + 3]; // the division / square-root scalars of the parameter transforms (one per lane, lm_warp.h)
  double x0[NP];
  double small10[10], large10[10];
  double gram[(NP + 1) * GRAM_PITCH];   // device: [J | f] of the current batch of 32 voxels, column-major (pass_fused_mma)
};

// Executor interface (WarpExec in fit_kernels.cu, SerialExec in tests/hostsim):
//   static constexpr int W           lanes that share one spot
//   int lane(); void sync();
//   double allsum(double), int allsum_int(int), double allmax(double)   -> same value on every lane
//   void argmin(double& v, int& k)    -> smallest v (lowest k on ties) on every lane
//   void reduce_store<N>(double (&v)[N], double* out)   out[i] = sum over lanes of v[i] (v is clobbered;
//                                                        visible to all lanes after the next sync())

// ---- initial guess (GaussianFit.__init__, Fitting_v4.py:174-185) ---------------------------
// dv[0..m) holds the float64 voxel values in voxel order.  Selects the 10 smallest (or largest)
// values in ascending (descending) order; equal values are taken in voxel order, like a stable
// sort.  Lane l owns voxels l, l+W, l+2W, ...; `taken` is its private bitmask (m <= 32*W... the
// host executor (W = 1) keeps the mask in a caller-provided array instead).
template <typename Exec>
IA3_HD void select10(Exec& ex, const double* dv, int m, bool largest, double* out10) {
  typename Exec::TakenMask taken;
  taken.clear();
  for (int r = 0; r < 10; ++r) {
    double best = INFINITY;
    int bk = 0x7fffffff;
    int slot = 0;
    for (int k = ex.lane(); k < m; k += Exec::W, ++slot) {
      if (taken.test(slot)) continue;
      const double v = largest ? -dv[k] : dv[k];
      if (v < best) { best = v; bk = k; }   // ascending k inside a lane: first minimum kept
    }
    ex.argmin(best, bk);
    if (bk != 0x7fffffff && (bk % Exec::W) == ex.lane()) taken.set(bk / Exec::W);
    if (ex.lane() == 0) out10[r] = largest ? -best : best;
  }
  ex.sync();
}

// Same selection for windows of any size, with the "taken" marks in a scratch array of m doubles
// (standalone GaussianFit on arbitrary voxel lists).
template <typename Exec>
IA3_HD void select10_scratch(Exec& ex, const double* dv, double* tmp, int m, bool largest, double* out10) {
  for (int k = ex.lane(); k < m; k += Exec::W) tmp[k] = largest ? -dv[k] : dv[k];
  ex.sync();
  for (int r = 0; r < 10; ++r) {
    double best = INFINITY;
    int bk = 0x7fffffff;
    for (int k = ex.lane(); k < m; k += Exec::W) {
      const double v = tmp[k];
      if (v < best) { best = v; bk = k; }
    }
    ex.argmin(best, bk);
    if (ex.lane() == 0) { out10[r] = largest ? -best : best; if (bk != 0x7fffffff) tmp[bk] = INFINITY; }
    ex.sync();
  }
}

// numpy's float64 pairwise sum for n = 10 (8-wide unrolled block, then the tail), / 10
IA3_HD double mean10_numpy(const double* a) {
  double res = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
  res += a[8];
  res += a[9];
  return res / 10.0;
}

// Builds x0 (as float32-rounded values, :185).  init_w_raw[3]: v4 -> the three entries are the
// same scalar init_w; v3 -> per-axis init_w.  Also fills fp.init_wt for v3.
IA3_HDN void initial_guess(FitParams& fp, const double* small10_asc, const double* large10_desc,
                          const double* init_w_raw, double* x0) {
  const double eps = exp(-10.0);
  double asc[10];
  for (int i = 0; i < 10; ++i) asc[i] = large10_desc[9 - i];
  double mb = mean10_numpy(small10_asc), mh = mean10_numpy(asc);
  double bk = log(mb > eps ? mb : eps), hh = log(mh > eps ? mh : eps);   // np.max([mean, eps])
  if (mb != mb) bk = mb;  // np.max propagates NaN
  if (mh != mh) hh = mh;
  double wg[3];
  if (fp.personality == 4) {
    for (int i = 0; i < 3; ++i) {
      double wsq = init_w_raw[i] * init_w_raw[i];
      wg[i] = log((fp.max_w2 - wsq) / (wsq - fp.min_w2));
    }
  } else {
    // Fitting_v3.py:70-75 -- the guard compares w^2 with the *unsquared* bounds and the
    // replacement value is 1.5**2 (which is then squared again)
    const double max_w = sqrt(fp.max_w2), min_w = sqrt(fp.min_w2);
    for (int i = 0; i < 3; ++i) {
      double iw = init_w_raw[i];
      if (iw * iw > max_w || iw * iw < min_w) iw = 1.5 * 1.5;
      wg[i] = log((fp.max_w2 - iw * iw) / (iw * iw - fp.min_w2));
      fp.init_wt[i] = wg[i];
    }
  }
  const double raw[NP] = {bk, hh, 0, 0, 0, wg[0], wg[1], wg[2], 0, 0};
  for (int i = 0; i < NP; ++i) x0[i] = (double)(float)raw[i];
}

// ---- voxel passes --------------------------------------------------------------------------
// Vox provides: int m; void get(int k, T& X0, T& X1, T& X2, T& data) with coordinates relative
// to the spot's integer origin.

// |f| with MINPACK enorm's semantics.  enorm keeps three accumulators (large / mid / small
// components) so that it neither overflows nor underflows; what matters for parity is how it
// behaves when the model blows up (v3 has no overflow guards and regularly proposes bk ~ 1e6):
//   * one +-inf residual           -> inf          (x1max = inf, s1 = 1)
//   * two or more inf residuals    -> NaN          ((inf/inf)^2)
//   * finite but > rgiant/m        -> finite norm, where a plain sum of squares gives inf
// lmder's step-bound update then takes different branches for inf / NaN / finite (the test
// "0.1*fnorm1 >= fnorm" is false for NaN), so the trust region shrinks by 0.1 or by ~0.25.
// Mid-range components use the plain sum of squares; the (very rare) large ones need a second
// sweep once the largest magnitude is known.
template <typename T, typename Exec, typename Vox>
IA3_HD double pass_residual(Exec& ex, const VoxConsts<T>& vc, const Vox& vox, double* sum_abs) {
  const double agiant = 1.304e19 / (double)vox.m;
  double s2 = 0.0, sa = 0.0, big = 0.0;
  int ninf = 0, nlarge = 0, nnan = 0;
  for (int k = ex.lane(); k < vox.m; k += Exec::W) {
    T X0, X1, X2, d;
    vox.get(k, X0, X1, X2, d);
    const double r = (double)eval_res<T>(vc, X0, X1, X2, d);
    const double a = fabs(r);
    sa += a;
    if (a < agiant) s2 += r * r;
    else if (a != a) nnan += 1;
    else { nlarge += 1; if (a > DBL_MAX) ninf += 1; else big = fmax(big, a); }
  }
  s2 = ex.allsum(s2);
  if (sum_abs) *sum_abs = ex.allsum(sa);
  nlarge = ex.allsum_int(nlarge);
  nnan = ex.allsum_int(nnan);
  if (nlarge == 0) return nnan ? NAN : sqrt(s2);   // sqrt taken here so callers get the norm
  ninf = ex.allsum_int(ninf);
  if (ninf == 1) return INFINITY;
  if (ninf >= 2) return NAN;
  big = ex.allmax(big);
  double s1 = 0.0;
  for (int k = ex.lane(); k < vox.m; k += Exec::W) {
    T X0, X1, X2, d;
    vox.get(k, X0, X1, X2, d);
    const double a = fabs((double)eval_res<T>(vc, X0, X1, X2, d));
    if (a >= agiant) { const double q = a / big; s1 += q * q; }
  }
  s1 = ex.allsum(s1);
  return big * sqrt(s1 + (s2 / big) / big);
}

// Ag_out[0..55) = J^T J (packed upper triangle), Ag_out[55..65) = J^T f
template <typename T, typename Exec, typename Vox>
IA3_HD void pass_jacobian(Exec& ex, const VoxConsts<T>& vc, const Vox& vox, double* Ag_out) {
  double acc[NTRI + NP];
#pragma unroll
  for (int i = 0; i < NTRI + NP; ++i) acc[i] = 0.0;
  for (int k = ex.lane(); k < vox.m; k += Exec::W) {
    T X0, X1, X2, d, res;
    float J[NP];
    vox.get(k, X0, X1, X2, d);
    eval_jac<T>(vc, X0, X1, X2, d, res, J);
    const double r = (double)res;
    int idx = 0;
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      const double ji = (double)J[i];
      acc[NTRI + i] += ji * r;
#pragma unroll
      for (int j = i; j < NP; ++j) { acc[idx] += ji * (double)J[j]; ++idx; }
    }
  }
  ex.template reduce_store<NTRI + NP>(acc, Ag_out);
}

// Residual norm AND normal equations at the same point in one sweep over the voxels.  lmder evaluates
// f(x + p) and, if the step is accepted, J(x + p): same point, same exp per voxel.  Doing both at once
// saves a voxel sweep per accepted step (most steps are) at the price of a discarded Jacobian per
// rejected one.  Residuals that are inf / NaN / huge go through pass_residual's enorm-faithful path
// (the step is then rejected and the sums are never used).
#if defined(__CUDACC__)
#ifndef IA3_FIT_MMA
#define IA3_FIT_MMA 1
#endif
// C(8x8) += A(8x4) B(4x8) in FP64 on the tensor cores.  Lane l holds A[l / 4][l % 4], B[l % 4][l / 4] and
// C[l / 4][2 (l % 4) + {0, 1}].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// pass_fused on the warp: the normal-equation sums as G^T G with G = [J | f] (m x 11).  The 65 sums
// used to be 65 FP64 accumulators per lane plus a 65-entry butterfly reduction; here the lanes stage
// the 11 values of their voxel in shared memory (32 voxels per batch) and three 8x8 FP64 tensor-core
// tiles accumulate G^T G over the batch's eight 4-voxel slices: rows/columns 0-7 (T00), 0-7 x 8-15
// (T01) and 8-15 x 8-15 (T11) of the 16x16 padded product.  6 accumulator registers per lane instead
// of 130, and the cross-lane reduction is part of the MMA.  Same products as the scalar path (float32
// Jacobian entries widened to FP64), different summation order.
// One batch of 32 voxels: the lanes stage [J | f] of their voxel in `gt`, then three 8x8 FP64 tensor-core
// tiles accumulate G^T G over the batch's eight 4-voxel slices, STARTING FROM ZERO.  t[0..5] = this lane's
// entries of T00 (rows/columns 0-7), T01 (0-7 x 8-15), T11 (8-15 x 8-15).  The running sums are formed
// by adding whole batch tiles in batch order (pass_fused_mma below, pass_cta in fit_kernels.cu): the
// result does not depend on how many warps share the batches of one spot.
template <typename T, typename Vox>
__device__ __forceinline__ void mma_batch_tile(const VoxConsts<T>& vc, const Vox& vox, int k0, int lane, double* gt, double (&t)[6],
                                               int& nbad, double agiant) {
  const int g = lane >> 2, q = lane & 3;
  const int k = k0 + lane;
  float J[NP];
  double r = 0.0;
  if (k < vox.m) {
    T X0, X1, X2, d, res;
    vox.get(k, X0, X1, X2, d);
    eval_jac<T>(vc, X0, X1, X2, d, res, J);
    r = (double)res;
    if (!(fabs(r) < agiant)) nbad += 1;
  } else {
#pragma unroll
    for (int i = 0; i < NP; ++i) J[i] = 0.f;
  }
  __syncwarp();                                  // the previous batch's fragments have been read
#pragma unroll
  for (int i = 0; i < NP; ++i) gt[i * GRAM_PITCH + lane] = (double)J[i];
  gt[NP * GRAM_PITCH + lane] = r;
  __syncwarp();
#pragma unroll
  for (int e = 0; e < 6; ++e) t[e] = 0.0;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const double f0 = gt[g * GRAM_PITCH + 4 * c + q];                           // column g, voxel 4c + q of the batch
    const double f1 = (g < 3) ? gt[(8 + g) * GRAM_PITCH + 4 * c + q] : 0.0;     // columns 8, 9 and f; 11-15 are padding
    dmma884(t[0], t[1], f0, f0);
    dmma884(t[2], t[3], f0, f1);
    dmma884(t[4], t[5], f1, f1);
  }
}

// scatter the upper triangle / the J^T f column of the three accumulated tiles to the packed layout
// lm_factor reads; returns (f, f) = |f|^2 on every lane
__device__ __forceinline__ double mma_scatter(const double (&a)[6], int lane, double* Ag_out) {
  static_assert(NP == 10, "tile bookkeeping below assumes 10 parameters + the residual column");
  const int g = lane >> 2, q = lane & 3;
  const int j0 = 2 * q, j1 = j0 + 1;
  if (g <= j0) Ag_out[tri(g, j0)] = a[0];
  if (g <= j1) Ag_out[tri(g, j1)] = a[1];
  const int ja = 8 + j0, jb = ja + 1;
  if (ja < NP) Ag_out[tri(g, ja)] = a[2]; else if (ja == NP) Ag_out[NTRI + g] = a[2];
  if (jb < NP) Ag_out[tri(g, jb)] = a[3]; else if (jb == NP) Ag_out[NTRI + g] = a[3];
  const int i = 8 + g;
  if (i < NP) {
    if (ja < NP) { if (i <= ja) Ag_out[tri(i, ja)] = a[4]; } else if (ja == NP) Ag_out[NTRI + i] = a[4];
    if (jb < NP) { if (i <= jb) Ag_out[tri(i, jb)] = a[5]; } else if (jb == NP) Ag_out[NTRI + i] = a[5];
  }
  return __shfl_sync(0xffffffffu, a[4], 9);      // (f, f): lane 9 = row 8 + 2, column 8 + 2
}

// pass_fused on one warp: the normal-equation sums as G^T G with G = [J | f] (m x 11).  The 65 sums
// used to be 65 FP64 accumulators per lane plus a 65-entry butterfly reduction; here three 8x8 FP64
// tensor-core tiles hold them (6 accumulator registers per lane) and the cross-lane reduction is part of
// the MMA.  Same products as the scalar path (float32 Jacobian entries widened to FP64), different
// summation order.
template <typename T, typename Exec, typename Vox>
__device__ __forceinline__ double pass_fused_mma(Exec& ex, const VoxConsts<T>& vc, const Vox& vox, double* Ag_out, double* gt) {
  const double agiant = 1.304e19 / (double)vox.m;
  const int lane = ex.lane();
  double acc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  int nbad = 0;
  for (int k0 = 0; k0 < vox.m; k0 += 32) {
    double t[6];
    mma_batch_tile<T>(vc, vox, k0, lane, gt, t, nbad, agiant);
#pragma unroll
    for (int e = 0; e < 6; ++e) acc[e] += t[e];
  }
  const double s2 = mma_scatter(acc, lane, Ag_out);
  nbad = ex.allsum_int(nbad);
  if (nbad == 0) return sqrt(s2);
  return pass_residual<T>(ex, vc, vox, (double*)0);
}
#endif

template <typename T, typename Exec, typename Vox>
IA3_HD double pass_fused(Exec& ex, const VoxConsts<T>& vc, const Vox& vox, double* Ag_out, double* gram) {
#if defined(__CUDA_ARCH__) && IA3_FIT_MMA
  if constexpr (Exec::W == 32) return pass_fused_mma<T>(ex, vc, vox, Ag_out, gram);
#endif
  (void)gram;
  const double agiant = 1.304e19 / (double)vox.m;
  double acc[NTRI + NP];
#pragma unroll
  for (int i = 0; i < NTRI + NP; ++i) acc[i] = 0.0;
  double s2 = 0.0;
  int nbad = 0;
  for (int k = ex.lane(); k < vox.m; k += Exec::W) {
    T X0, X1, X2, d, res;
    float J[NP];
    vox.get(k, X0, X1, X2, d);
    eval_jac<T>(vc, X0, X1, X2, d, res, J);
    const double r = (double)res;
    if (fabs(r) < agiant) s2 += r * r; else nbad += 1;
    int idx = 0;
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      const double ji = (double)J[i];
      acc[NTRI + i] += ji * r;
#pragma unroll
      for (int j = i; j < NP; ++j) { acc[idx] += ji * (double)J[j]; ++idx; }
    }
  }
  ex.template reduce_store<NTRI + NP>(acc, Ag_out);
  s2 = ex.allsum(s2);
  nbad = ex.allsum_int(nbad);
  if (nbad == 0) return sqrt(s2);
  return pass_residual<T>(ex, vc, vox, (double*)0);
}

struct FitResult {
  double p_raw[NP];
  float ps[NOUT];
  int nfev, njev, info;
};

// Per-voxel constants at x, by the whole warp: the 19 FP64 exps one per lane, the rest (divisions,
// square roots, the 36 quadratic-form coefficients) by lane 0.  Always with the Jacobian constants:
// an accepted trial point is the next Jacobian point, so lmder's pair "f(x + p), then J(x + p)" costs
// one call instead of two.
template <typename T, typename Exec>
IA3_HD void build_consts_par(Exec& ex, const FitParams& fp, const double* cen_est, const double* origin,
                             const double* x, SpotShared<T>& sh) {
  for (int i = ex.lane(); i < NEXP; i += Exec::W) sh.etab[i] = (i == 1) ? 0.0 : exp(exp_slot_arg(fp, x, i));
  ex.sync();
  if (ex.lane() == 0) finish_consts<T>(fp, cen_est, origin, x, sh.etab, true, sh.vc);
  ex.sync();
}

// Everything that is live at the top of lmder's outer loop: a run can be suspended there and resumed
// later (by another kernel launch, by another number of warps) without changing a single bit of its
// trajectory.  lm_outer recomputes the factorisation (R, ipvt, acn, qtf, B0, rq ...) from the sums, so
// only the iterate, the scaling, four scalars, the counters and the sums themselves are kept: 728 bytes.
struct LMLive {
  double x[NP], diag[NP];
  double Ag[NTRI + NP];
  double fnorm, xnorm, delta, par;
  int iter, nfev, njev, info;
};
// i-th 8-byte word of the live state, read from (st, Ag) / written back to them
constexpr int LMLIVE_WORDS = (int)(sizeof(LMLive) / 8);
IA3_HD void lm_live_save(const LMState& st, const double* Ag, LMLive& out) {
  for (int i = 0; i < NP; ++i) { out.x[i] = st.x[i]; out.diag[i] = st.diag[i]; }
  for (int i = 0; i < NTRI + NP; ++i) out.Ag[i] = Ag[i];
  out.fnorm = st.fnorm; out.xnorm = st.xnorm; out.delta = st.delta; out.par = st.par;
  out.iter = st.iter; out.nfev = st.nfev; out.njev = st.njev; out.info = st.info;
}
IA3_HD void lm_live_restore(const LMLive& in, LMState& st, double* Ag) {
  for (int i = 0; i < NP; ++i) { st.x[i] = in.x[i]; st.diag[i] = in.diag[i]; }
  for (int i = 0; i < NTRI + NP; ++i) Ag[i] = in.Ag[i];
  st.fnorm = in.fnorm; st.xnorm = in.xnorm; st.delta = in.delta; st.par = in.par;
  st.iter = in.iter; st.nfev = in.nfev; st.njev = in.njev; st.info = in.info;
}
enum { LM_START_FRESH = 0, LM_START_CONTINUE = 1 };

// Runs leastsq from sh.x0 (LM_START_FRESH) or continues the run whose state is in sh.st / sh.Ag
// (LM_START_CONTINUE).  Every lane must call this; results are in sh.st (shared) after return.
// cap > 0: return true ("suspended", state consistent in sh.st / sh.Ag) once cap function evaluations
// have been spent in this call and the run is not finished; false = finished.
template <typename T, typename Exec, typename Vox>
IA3_HD bool run_lm(Exec& ex, const FitParams& fp, const LMConfig& cfg, const double* cen_est,
                   const double* origin, const Vox& vox, SpotShared<T>& sh, int cap = 0, int start = LM_START_FRESH) {
  LMState& st = sh.st;
  if (start == LM_START_FRESH) {
    build_consts_par<T>(ex, fp, cen_est, origin, sh.x0, sh);
    const double fn0 = pass_fused<T>(ex, sh.vc, vox, sh.Ag, sh.gram);       // f(x0) and J(x0)
    lm_init(ex, st, sh.x0, fn0);
  }
  ex.sync();
  const int nfev_entry = st.nfev;
  for (;;) {
    // sh.Ag holds J^T J, J^T f at st.x (x0, or the trial point that was just accepted)
    ex.sync();
    if (cap > 0 && st.nfev - nfev_entry >= cap) return true;
    if (!lm_outer(ex, st, cfg, sh.Ag, sh.Ag + NTRI)) break;
    int action;
    for (;;) {
      lm_propose(ex, st);
      build_consts_par<T>(ex, fp, cen_est, origin, st.xt, sh);
      const double fn1 = pass_fused<T>(ex, sh.vc, vox, sh.Ag, sh.gram);   // lm_outer has consumed the old sums
      action = lm_judge(ex, st, cfg, fn1);
      if (action != LM_RETRY) break;
    }
    if (action == LM_DONE) break;
  }
  ex.sync();
  return false;
}

// to_natural_paramaters() with the final parameters (Fitting_v4.py:244-258): ps[0..9] natural
// parameters, ps[10] = mean |residual| over the fitted voxels; all cast to float32.
template <typename T, typename Exec, typename Vox>
IA3_HD void finish_fit(Exec& ex, const FitParams& fp, const double* cen_est, const double* origin,
                       const Vox& vox, SpotShared<T>& sh, FitResult* out /*lane 0 writes*/) {
  if (ex.lane() == 0) {
    build_consts<T>(fp, cen_est, origin, sh.st.x, false, sh.vc);
  }
  ex.sync();
  double sa = 0.0;
  pass_residual<T>(ex, sh.vc, vox, &sa);
  if (ex.lane() == 0) {
    double nat[10];
    natural_params(fp, cen_est, sh.st.x, nat);
    for (int i = 0; i < 10; ++i) out->ps[i] = (float)nat[i];
    out->ps[10] = (float)(sa / (double)vox.m);
    for (int i = 0; i < NP; ++i) out->p_raw[i] = sh.st.x[i];
    out->nfev = sh.st.nfev;
    out->njev = sh.st.njev;
    out->info = sh.st.info;
  }
  ex.sync();
}

}  // namespace ia3
