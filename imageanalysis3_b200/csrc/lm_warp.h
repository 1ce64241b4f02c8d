// lmder's 10x10 algebra for ONE WARP, written for latency (device only).
//
// lm_core.h states the algorithm (MINPACK lmder / lmpar driven from J^T J, J^T f, |f|^2) for a generic
// executor, with every matrix in shared memory and a barrier after each dependent step; measured on
// B200 that costs ~47 k cycles per function evaluation (ncu / clock64: lm_outer 17.5 k, lmpar 21.6 k,
// parameter constants 5.3 k, lm_judge 2.7 k) against 5.8 k for the voxel pass of an 8-warp team -- the
// serial algebra, not the model, bounds a fit.  Here the same steps run out of registers:
//   * lane l (l < 10, mirrored in both half-warps so that every scalar is uniform) owns COLUMN l of the
//     Jacobian: its entries of R / S / R^T R stay in registers, indexed by compile-time row numbers
//     (every loop is unrolled); a step's pivot column reaches the other lanes by shuffles, never through
//     shared memory, so a step costs no barrier;
//   * the small bookkeeping (remaining column norms, pivot order) is replicated in every lane, by
//     position, so pivoting is a handful of predicated moves with qrfac's rule (largest remaining
//     norm, first position wins, NaN never wins);
//   * 1/sqrt(d) comes from one rsqrt, the diagonal entry is d * rsqrt(d); divisions by pivots are
//     multiplications by the stored reciprocals.
// Same algorithm, same pivoting, same stopping logic as lm_core.h; sums are formed in a different
// order and pivots carry <= 2 ulp instead of <= 1, i.e. the two agree to rounding (the parity tests
// compare both with scipy's MINPACK).  The factorisation's products are parked in the LMState arrays of
// the spot (shared memory) between calls in a lane-major layout -- st.R[i][l] = R[i][position of
// column l] -- so nothing has to stay in registers across the voxel pass.
// Control flow: every branch that has a shuffle behind it is taken on a VOTE result (__all_sync /
// __any_sync), which ptxas knows to be warp-uniform; otherwise it wraps each shuffle in a
// WARPSYNC.COLLECTIVE / ENDCOLLECTIVE pair (measured: 3x the cycles of the shuffles themselves).
// Each function starts with __syncwarp() for the same reason (it is called under `if (warp == 0)`).
#pragma once
#include "gauss_model.h"
#include "lm_core.h"

#if defined(__CUDACC__)
namespace ia3 {
namespace lw {

constexpr unsigned FULLM = 0xffffffffu;
#ifdef IA3_FIT_PROF
__device__ unsigned long long g_lw_prof[16];     // cycles: factor init / steps / B0+store; outer rest; propose GN / parl / damped (count in [7]) / tail; judge
#define LWP_T(v) const long long v = clock64()
#define LWP_ADD(i, a, b) do { if ((threadIdx.x & 31) == 0) atomicAdd(&g_lw_prof[i], (unsigned long long)((b) - (a))); } while (0)
#else
#define LWP_T(v) do {} while (0)
#define LWP_ADD(i, a, b) do {} while (0)
#endif
__device__ __forceinline__ double bc(double v, int src) { return __shfl_sync(FULLM, v, src, 16); }
__device__ __forceinline__ double sum16(double v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(FULLM, v, o, 16);
  return v;
}
__device__ __forceinline__ double max16(double v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULLM, v, o, 16));
  return v;
}

// butterfly reduce-scatter step over the 16-lane halves: N per-lane partial sums -> ceil(N / 2), the other half goes
// to the partner lane (xor O); after the steps 8, 4, 2, 1 every total sits in exactly one lane of each half
template <int N, int O>
__device__ __forceinline__ void rs_step16(double* v) {
  const bool up = (threadIdx.x & O) != 0;
#pragma unroll
  for (int i = 0; i < (N + 1) / 2; ++i) {
    const double a = v[2 * i];
    const double b = (2 * i + 1 < N) ? v[2 * i + 1] : 0.0;
    const double send = up ? a : b;
    const double keep = up ? b : a;
    v[i] = keep + __shfl_xor_sync(FULLM, send, O, 16);
  }
}

// index of A[i][j] (either order) in the packed upper triangle
__device__ __forceinline__ int tri_sym(int i, int j) { return i < j ? tri(i, j) : tri(j, i); }

// lmder's outer-iteration prologue after a Jacobian pass at st.x (A = J^T J, g = J^T f): pivoted Cholesky of
// A with qrfac's pivot rule (largest remaining column norm; among exactly equal norms the lowest column --
// qrfac takes the first in its current order, which differs only when two POSITIVE remaining norms are
// bit-identical), qtf = R^-T P^T g, the scaling and gradient tests of lmder.
// Leaves in st: R[i][l], B0[a][l] (P^T A P = R^T R up to rounding, column of lane l), rq[l], acn[l], pos[l],
// ipvt[p], qtf[p], rinv[p], nsing.  Returns false if lmder stops here (gnorm <= gtol).
__device__ __noinline__ bool outer(LMState& st, const LMConfig& cfg, const double* A, const double* g) {
  __syncwarp();
  LWP_T(c0);
  const int l = threadIdx.x & 15;
  const bool act = l < NP;
  const int lane0 = (threadIdx.x & 31) == 0;
  const int c = act ? l : 0;
  double Rc[NP], qv[NP], rinvU[NP];
  int colP[NP];
  double rdl = act ? A[tri(c, c)] : -1.0;      // remaining squared norm of my column
  const double acn = act ? sqrt(rdl) : 0.0;
  double w = act ? g[c] : 0.0;                 // forward substitution R^T qtf = P^T g, by own column
  bool done = !act;                            // my column has its position
  int myPos = NP, nsing = NP;
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    // arg max of the remaining norms with two 32-bit warp reductions on an order-preserving key
    const unsigned long long bits = (unsigned long long)__double_as_longlong(rdl);
    unsigned long long key = (bits >> 63) ? ~bits : (bits | 0x8000000000000000ull);
    if (done || rdl != rdl) key = 0ull;        // placed columns never; NaN only when nothing else is left
    const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
    const unsigned mh = __reduce_max_sync(FULLM, hi);
    const unsigned ml = __reduce_max_sync(FULLM, hi == mh ? lo : 0u);
    const unsigned bal = __ballot_sync(FULLM, !done && hi == mh && lo == ml);
    const int pc = __ffs(bal & 0xffffu) - 1;
    const double dd = bc(rdl, pc);
    // dd <= 0 (or NaN): exactly dependent / zero column -- row j of R stays zero (qrfac: rdiag = 0)
    const bool pos_ok = dd > 0.0;
    if (!pos_ok && nsing == NP) nsing = j;
    const double inv = pos_ok ? rsqrt(dd) : 0.0;
    const double rjj = pos_ok ? dd * inv : 0.0;
    const bool is_p = act && l == pc;
    const bool later = !done && !is_p;
    double s = A[tri_sym(pc, c)];
#pragma unroll
    for (int i = 0; i < j; ++i) s -= bc(Rc[i], pc) * Rc[i];
    const double r = (later && pos_ok) ? s * inv : (is_p ? rjj : 0.0);
    Rc[j] = r;
    if (later) rdl -= r * r;
    const double wq = bc(w, pc);
    const double q = pos_ok ? wq * inv : 0.0;
    qv[j] = q;
    rinvU[j] = inv;
    if (later && pos_ok) w -= r * q;
    if (is_p) { done = true; myPos = j; }
    colP[j] = pc;
  }
  LWP_T(c1);
  LWP_ADD(0, c0, c1);
  double rq = 0.0;
#pragma unroll
  for (int i = 0; i < NP; ++i) rq += Rc[i] * qv[i];
  if (act) {
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      st.R[i][c] = Rc[i];
      st.B0[i][c] = (i <= myPos) ? A[tri_sym(colP[i], c)] : 0.0;
    }
    st.rq[c] = rq;
    st.acn[c] = acn;
    st.pos[c] = myPos;
  }
#pragma unroll
  for (int p = 0; p < NP; ++p)
    if (l == p) { st.ipvt[p] = colP[p]; st.qtf[p] = qv[p]; st.rinv[p] = rinvU[p]; }
  if (l == 0) st.nsing = nsing;
  LWP_T(c2);
  LWP_ADD(1, c1, c2);
  LWP_ADD(8, 0, 1);
  // scaling (first iteration), gradient norm test, diag = max(diag, acnorm)
  double diagl = act ? st.diag[c] : 0.0;
  const double xl = act ? st.x[c] : 0.0;
  if (__any_sync(FULLM, st.iter == 1)) {
    diagl = act ? ((acn == 0.0) ? 1.0 : acn) : 0.0;
    const double dx = diagl * xl;
    const double xnorm = sqrt(sum16(dx * dx));
    double delta = cfg.factor * xnorm;
    if (delta == 0.0) delta = cfg.factor;
    if (lane0) { st.xnorm = xnorm; st.delta = delta; }
  }
  double gnorm = 0.0;
  const double fnorm = st.fnorm;
  if (fnorm != 0.0) {
    const double rf = 1.0 / fnorm;
    double sum = 0.0;
#pragma unroll
    for (int i = 0; i < NP; ++i) sum += Rc[i] * (qv[i] * rf);
    if (act && acn != 0.0) gnorm = fabs(sum / acn);
  }
  gnorm = max16(gnorm);
  if (lane0) { st.njev += 1; st.gnorm = gnorm; }
  const bool stop = __all_sync(FULLM, gnorm <= cfg.gtol);
  if (stop) { if (lane0) st.info = 4; }
  else if (act) st.diag[c] = fmax(diagl, acn);
  __syncwarp();
  LWP_T(c3);
  LWP_ADD(2, c2, c3);
  return !stop;
}

// forward substitution T^T w = v for a column-distributed upper-triangular T (Tc = this lane's column, entries
// by row), v by position held by the lane at that position; returns |w|^2 (uniform)
__device__ __forceinline__ double fwd_norm2(const double (&Tc)[NP], const double (&tinv)[NP], const int (&colP)[NP], int myPos, double vl) {
  double n2 = 0.0;
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    const double t = bc(vl, colP[j]) * tinv[j];
    n2 += t * t;
    if (myPos > j) vl -= Tc[j] * t;
  }
  return n2;
}

// MINPACK lmpar + the trial point: on return st.par is the LM parameter, st.p the step (sign flipped, as
// lmder uses it), st.xt = x + p, st.pnorm = |D p|.
__device__ __noinline__ void propose(LMState& st) {
  __syncwarp();
  const double dwarf = DBL_MIN;
  const int l = threadIdx.x & 15;
  const bool act = l < NP;
  const int lane0 = (threadIdx.x & 31) == 0;
  const int c = act ? l : 0;
  double Rc[NP], rinvU[NP], wa[NP];
  int colP[NP];
  const int nsing = st.nsing;
  const int myPos = act ? st.pos[c] : NP;
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    Rc[i] = act ? st.R[i][c] : 0.0;
    rinvU[i] = st.rinv[i];
    colP[i] = st.ipvt[i];
    wa[i] = (i < nsing) ? st.qtf[i] : 0.0;
  }
  const double diagl = act ? st.diag[c] : 0.0;
  const double rql = act ? st.rq[c] : 0.0;
  const double delta = st.delta;
  // Gauss-Newton direction: back substitution with the non-singular leading block of R (axpy form)
  LWP_T(c0);
  double xl = 0.0;
#pragma unroll
  for (int j = NP - 1; j >= 0; --j) {     // rows >= nsing: wa = 0 and rinv = 0, i.e. x = 0 and no update
    const double temp = wa[j] * rinvU[j];
    if (myPos == j) xl = temp;
#pragma unroll
    for (int i = 0; i < j; ++i) wa[i] -= bc(Rc[i], colP[j]) * temp;
  }
  double wa2 = diagl * xl;
  double dxnorm = sqrt(sum16(wa2 * wa2));
  double fp = dxnorm - delta;
  double par = 0.0;
  LWP_T(c1);
  LWP_ADD(3, c0, c1);
  LWP_ADD(9, 0, 1);
  if (!__all_sync(FULLM, fp <= 0.1 * delta)) {
    double parl = 0.0;
    if (__all_sync(FULLM, nsing >= NP)) {
      const double n2 = fwd_norm2(Rc, rinvU, colP, myPos, act ? diagl * (wa2 / dxnorm) : 0.0);
      const double temp = sqrt(n2);
      parl = ((fp / delta) / temp) / temp;
    }
    const double gv = act ? rql / diagl : 0.0;
    const double gnorm = sqrt(sum16(gv * gv));
    double paru = gnorm / delta;
    if (paru == 0.0) paru = dwarf / fmin(delta, 0.1);
    par = st.par;
    par = fmax(par, parl);
    par = fmin(par, paru);
    if (par == 0.0) par = gnorm / dxnorm;
    double B0c[NP];
#pragma unroll
    for (int a = 0; a < NP; ++a) B0c[a] = act ? st.B0[a][c] : 0.0;
    const double d2 = diagl * diagl;
    LWP_T(c2);
    LWP_ADD(4, c1, c2);
    for (int iter = 1;; ++iter) {
      LWP_T(c3);
      if (par == 0.0) par = fmax(dwarf, 0.001 * paru);
      // S = chol(R^T R + par D^2) (qrsolv's triangular factor up to row signs), left-looking, no pivoting
      double Sc[NP], sinvU[NP], yU[NP];
#pragma unroll
      for (int j = 0; j < NP; ++j) {
        const int pc = colP[j];
        double v = (myPos >= j) ? B0c[j] : 0.0;
        if (myPos == j) v += par * d2;
#pragma unroll
        for (int i = 0; i < j; ++i) v -= bc(Sc[i], pc) * Sc[i];
        const double vp = bc(v, pc);
        const double inv = (vp > 0.0) ? rsqrt(vp) : 0.0;
        const double sjj = (vp > 0.0) ? vp * inv : 0.0;
        sinvU[j] = inv;
        Sc[j] = (myPos == j) ? sjj : ((myPos > j && act) ? v * inv : 0.0);
      }
      LWP_T(c5);
      LWP_ADD(6, c3, c5);
      // forward: S^T y = R^T qtf (axpy form), backward: S z = y; x[ipvt[j]] = z[j]
      double wl = rql;
#pragma unroll
      for (int j = 0; j < NP; ++j) {
        const double y = bc(wl, colP[j]) * sinvU[j];
        yU[j] = y;
        if (myPos > j) wl -= Sc[j] * y;
      }
#pragma unroll
      for (int j = NP - 1; j >= 0; --j) {
        const double z = yU[j] * sinvU[j];
        if (myPos == j) xl = z;
#pragma unroll
        for (int i = 0; i < j; ++i) yU[i] -= bc(Sc[i], colP[j]) * z;
      }
      LWP_T(c4);
      LWP_ADD(5, c3, c4);
      LWP_ADD(7, 0, 1);
      wa2 = diagl * xl;
      dxnorm = sqrt(sum16(wa2 * wa2));
      const double prev = fp;
      fp = dxnorm - delta;
      if (__any_sync(FULLM, fabs(fp) <= 0.1 * delta || (parl == 0.0 && fp <= prev && prev < 0.0) || iter == 10)) break;
      const double n2 = fwd_norm2(Sc, sinvU, colP, myPos, act ? diagl * (wa2 / dxnorm) : 0.0);
      const double temp = sqrt(n2);
      const double parc = ((fp / delta) / temp) / temp;
      if (fp > 0.0) parl = fmax(parl, par);
      if (fp < 0.0) paru = fmin(paru, par);
      par = fmax(parl, par + parc);
    }
  }
  // trial point
  const double pl = -xl;
  const double dp = diagl * pl;
  const double pnorm = sqrt(sum16(dp * dp));
  const int iter1 = st.iter == 1;
  __syncwarp();
  if (act) { st.p[c] = pl; st.xt[c] = st.x[c] + pl; }
  if (lane0) {
    st.par = par;
    st.pnorm = pnorm;
    if (iter1) st.delta = fmin(delta, pnorm);
  }
  __syncwarp();
}

// Given fnorm1 = |f(st.xt)|: ratio test, trust-region update, convergence tests (lm_core.h: lm_judge).
__device__ __noinline__ int judge(LMState& st, const LMConfig& cfg, double fnorm1) {
  __syncwarp();
  const int l = threadIdx.x & 15;
  const bool act = l < NP;
  const int lane0 = (threadIdx.x & 31) == 0;
  const int c = act ? l : 0;
  const double fnorm = st.fnorm;
  const double pnorm = st.pnorm;
  double delta = st.delta, par = st.par, xnorm = st.xnorm;
  const double gnorm = st.gnorm;
  const int nfev = st.nfev + 1;
  double actred = -1.0;
  if (0.1 * fnorm1 < fnorm) { const double q = fnorm1 / fnorm; actred = 1.0 - q * q; }
  // |R p[ipvt]|: row i = sum over the columns (lanes) of R[i][column] * p[column]
  const double pl = act ? st.p[c] : 0.0;
  double rowp[NP];
#pragma unroll
  for (int i = 0; i < NP; ++i) rowp[i] = (act ? st.R[i][c] : 0.0) * pl;
  rs_step16<NP, 8>(rowp);
  rs_step16<5, 4>(rowp);
  rs_step16<3, 2>(rowp);
  rs_step16<2, 1>(rowp);
  const double rp2 = sum16(rowp[0] * rowp[0]);
  const double temp1 = sqrt(rp2) / fnorm;
  const double temp2 = (sqrt(par) * pnorm) / fnorm;
  const double prered = temp1 * temp1 + temp2 * temp2 / 0.5;
  const double dirder = -(temp1 * temp1 + temp2 * temp2);
  double ratio = 0.0;
  if (prered != 0.0) ratio = actred / prered;
  if (ratio <= 0.25) {
    double temp = 0.5;
    if (actred < 0.0) temp = 0.5 * dirder / (dirder + 0.5 * actred);
    if (0.1 * fnorm1 >= fnorm || temp < 0.1) temp = 0.1;
    delta = temp * fmin(delta, pnorm / 0.1);
    par = par / temp;
  } else if (par == 0.0 || ratio >= 0.75) {
    delta = pnorm / 0.5;
    par = 0.5 * par;
  }
  const bool accepted = __all_sync(FULLM, ratio >= 1.0e-4);
  if (accepted) {
    const double v = act ? st.xt[c] : 0.0;
    const double dx = (act ? st.diag[c] : 0.0) * v;
    xnorm = sqrt(sum16(dx * dx));
    if (act) st.x[c] = v;
  }
  int info = 0;
  const bool small = fabs(actred) <= cfg.ftol && prered <= cfg.ftol && 0.5 * ratio <= 1.0;
  if (small) info = 1;
  if (delta <= cfg.xtol * xnorm) info = 2;
  if (small && info == 2) info = 3;
  if (info == 0) {
    if (nfev >= cfg.maxfev) info = 5;
    if (fabs(actred) <= DBL_EPSILON && prered <= DBL_EPSILON && 0.5 * ratio <= 1.0) info = 6;
    if (delta <= DBL_EPSILON * xnorm) info = 7;
    if (gnorm <= DBL_EPSILON) info = 8;
  }
  __syncwarp();
  if (lane0) {
    st.nfev = nfev;
    st.fnorm1 = fnorm1;
    st.delta = delta;
    st.par = par;
    st.info = info;
    if (accepted) { st.xnorm = xnorm; st.fnorm = fnorm1; st.iter += 1; }
  }
  __syncwarp();
  if (info != 0) return LM_DONE;
  return accepted ? LM_ACCEPTED : LM_RETRY;
}

// Per-voxel constants at x by the whole warp: the 19 FP64 exps one per lane; then the 16 + 5 scalars that cost a
// division or a square root (gauss_model.h: consts_scalar), one per lane -- every lane prepares its own numerator,
// denominator and offsets with a few cheap selects and ALL lanes then run ONE division (a switch over the formulas
// would execute sixteen divisions one after the other: divergence); then the products by lane 0.  Same operations
// per scalar as the serial build_consts.  `etab`, `scal`: shared scratch.
template <typename T>
__device__ __forceinline__ void build_consts(const FitParams& fp, const double* cen_est, const double* origin, const double* x,
                                             double* etab, double* scal, VoxConsts<T>& vc) {
  const int lane = threadIdx.x & 31;
  const bool v4 = (fp.personality == 4);
  const double LOGMAX = 709.782712893384;
  __syncwarp();
  if (lane < NEXP) etab[lane] = (lane == 1) ? 0.0 : exp(exp_slot_arg(fp, x, lane));
  __syncwarp();
  {
    const double d = fp.delta, minw = fp.min_w2, maxw = fp.max_w2, dws = maxw - minw;
    double numer = 0.0, denom = 1.0, off1 = 0.0, off2 = 0.0, raw = 0.0, lo_val = 0.0, hi_val = 0.0;
    bool guard = false;
    if (lane < 2) {                       // t (x[9]), p (x[8]): 2 / (1 + e) - 1
      const int j = 9 - lane;
      numer = 2.0; denom = 1.0 + etab[j]; off1 = -1.0; raw = x[j]; lo_val = -1.0; hi_val = 1.0; guard = v4;
    } else if (lane < 5) {                // ws_i: dws / (1 + e) + minw
      const int j = 5 + (lane - 2);
      numer = dws; denom = 1.0 + etab[j]; off1 = minw; raw = x[j]; lo_val = minw; hi_val = dws + minw; guard = v4;
    } else if (lane < 8) {                // centre: 2 d / (1 + e) - d + c_est   (v3: 2 d e' / (1 + e), Fitting_v3.py:86)
      const int i = lane - 5;
      const double ce = cen_est[i];
      numer = v4 ? 2.0 * d : 2.0 * d * ((i == 2) ? etab[3] : etab[2 + i]);
      denom = 1.0 + etab[2 + i]; off1 = -d; off2 = ce; raw = x[2 + i]; lo_val = -d + ce; hi_val = d + ce; guard = v4;
    } else if (lane < 11) {               // norm_xp..: -d e / ((1 + e) (1 + e))
      const double ex = etab[11 + (lane - 8)];
      numer = -d * ex; denom = (1 + ex) * (1 + ex);
    } else if (lane < 14) {               // norm_w_i (Fitting_v4.py:369-375)
      const double w = x[5 + (lane - 11)], e = etab[14 + (lane - 11)];
      const double dd = (w > 0) ? maxw * e + minw : minw * e + maxw;
      numer = 0.5 * (maxw - minw) * e; denom = dd * dd;
    } else if (lane < 16) {               // norm_p, norm_t: e / (1 + e e)
      const double e = etab[17 + (lane - 14)];
      numer = e; denom = 1 + e * e;
    }
    double r = numer / denom;
    r = r + off1;
    r = r + off2;
    if (guard && raw >= LOGMAX) r = lo_val;
    if (guard && raw <= -LOGMAX) r = hi_val;
    if (lane < NSCAL_FIRST) scal[lane] = r;
  }
  __syncwarp();
  if (lane < 3) scal[16 + lane] = 1.0 / scal[2 + lane];
  else if (lane < 5) scal[19 + (lane - 3)] = sqrt(1 - scal[lane - 3] * scal[lane - 3]);
  __syncwarp();
  if (lane == 0) finish_consts_scalars<T>(fp, origin, x, etab, scal, true, vc);
  __syncwarp();
}

}  // namespace lw
}  // namespace ia3
#endif
