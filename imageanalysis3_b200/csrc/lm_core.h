// MINPACK-lmder-faithful Levenberg-Marquardt driver, driven from the normal-equation sums
// (J^T J, J^T f, |f|^2) that a warp accumulates per spot.
//
// The reference fits every spot with scipy.optimize.leastsq(calc_eps, p0, Dfun=calc_jac)
// (External/Fitting_v4.py:388, External/Fitting_v3.py:250), i.e. MINPACK lmder with
// ftol = xtol = 1.49012e-8, gtol = 0, factor = 100, mode = 1.  SURVEY.md App. G / B.8 shows
// that the result is only reproducible if the *trust-region logic* is lmder's, so this file
// restates lmder / lmpar (public-domain MINPACK, Argonne 1980) step by step.  Two deliberate
// changes, both in how the small linear algebra is carried out, not in what is computed:
//   * lmder's Householder QR of the m x 10 Jacobian (qrfac with column pivoting) is replaced
//     by a pivoted Cholesky factorisation of the FP64 10x10 matrix J^T J with qrfac's pivot
//     rule (largest remaining column norm, first index wins ties).  R^T R = P^T J^T J P, so R
//     agrees with qrfac's R up to row signs, which cancel everywhere R is used together with
//     qtf = R^-T P^T J^T f.
//   * lmpar's qrsolv (55 Givens rotations eliminating sqrt(par)*D against R) is replaced by a
//     Cholesky factorisation of R^T R + par*D^2: the same upper-triangular S up to row signs
//     (S^T S = R^T R + par D^2), the same step x = (S^T S)^-1 R^T qtf, a tenth of the
//     sequential depth.
//
// Everything here is FP64 on 10-vectors / 10x10 matrices held in the spot's shared-memory
// state.  The routines are written for an executor `Ex` (fit_spot.h): the O(n^2)/O(n^3) loops
// are strided over the lanes of the warp that owns the spot (`for (k = ex.lane(); ...; k +=
// Ex::W)`), short reductions over 10 numbers are done redundantly by every lane from shared
// memory (so all lanes hold bit-identical scalars and take the same branches), and
// `ex.sync()` separates dependent phases.  With the one-lane host executor (tests/hostsim)
// the same code runs serially.  No per-voxel work here.
#pragma once
#include "ia3_common.h"

namespace ia3 {

struct LMConfig {
  double ftol, xtol, gtol, factor;
  int maxfev;
};

struct LMState {
  double x[NP];        // accepted raw parameters
  double xt[NP];       // trial point x + p                     (lmder wa2)
  double p[NP];        // step                                   (lmder wa1, sign already flipped)
  double diag[NP];     // variable scaling
  double acn[NP];      // column norms of J                      (qrfac acnorm / lmder wa2)
  double qtf[NP];      // first n entries of Q^T f
  double R[NP][NP];    // upper triangle: R (pivoted order)
  double B0[NP][NP];   // R^T R, upper triangle (pivoted order)
  double S[NP][NP];    // Cholesky factor of R^T R + par D^2, upper triangle
  double rinv[NP];     // 1 / R[j][j] (0 where the pivot is 0)
  double sinv[NP];     // 1 / S[j][j]
  double rq[NP];       // R^T qtf (pivoted order)
  double rd[NP], gp[NP], y[NP];
  double w1[NP], w2[NP], w3[NP];
  int ipvt[NP];
  double fnorm, fnorm1, xnorm, delta, par, gnorm, pnorm;
  int iter, nfev, njev, info;
  int pos[NP];         // lm_warp.h: position of column l in the pivot order (inverse of ipvt)
  int nsing;           // lm_warp.h: first position with a zero pivot (NP if none)
  int pad_;
};

IA3_HD double enorm_n(const double* v, int n) {
  double s = 0.0;
  for (int i = 0; i < n; ++i) s += v[i] * v[i];
  return sqrt(s);
}

// B0 = R^T R, rq = R^T qtf (both in pivoted order) and rinv from st.R / st.qtf
template <typename Ex>
IA3_HDN void lm_post_factor(Ex& ex, LMState& st) {
  for (int e = ex.lane(); e < NP * NP + NP; e += Ex::W) {
    if (e < NP * NP) {
      const int a = e / NP, b = e - a * NP;
      if (b < a) continue;                          // upper triangle only
      double s = 0.0;
      for (int i = 0; i <= a; ++i) s += st.R[i][a] * st.R[i][b];
      st.B0[a][b] = s;
    } else {
      const int j = e - NP * NP;
      double s = 0.0;
      for (int i = 0; i <= j; ++i) s += st.R[i][j] * st.qtf[i];
      st.rq[j] = s;
      const double rjj = st.R[j][j];
      st.rinv[j] = (rjj != 0.0) ? 1.0 / rjj : 0.0;
    }
  }
  ex.sync();
}

// Pivoted Cholesky of the packed symmetric A (55 entries) with MINPACK qrfac's pivoting.
// Outputs st.R (upper), st.ipvt, st.acn, st.qtf = R^-T (g permuted), and lm_post_factor's products.
// Step j: lane k - j forms the entry (j, k) of the Schur complement (lane 0 the pivot itself), the
// pivot's reciprocal square root is broadcast, one barrier per step.
template <typename Ex>
IA3_HDN void lm_factor(Ex& ex, LMState& st, const double* A, const double* g) {
  const int ln = ex.lane();
  for (int i = ln; i < NP; i += Ex::W) {
    const double a = A[tri(i, i)];
    st.acn[i] = sqrt(a);
    st.rd[i] = a;              // remaining squared norms (Schur complement diagonal)
    st.ipvt[i] = i;
    st.gp[i] = g[i];
    for (int j = 0; j < NP; ++j) st.R[i][j] = 0.0;
  }
  ex.sync();
#pragma unroll 1
  for (int j = 0; j < NP; ++j) {
    int kmax = j;
    for (int k = j + 1; k < NP; ++k) if (st.rd[k] > st.rd[kmax]) kmax = k;
    if (kmax != j) {
      ex.sync();                                 // everybody has read rd[] before it is permuted
      for (int i = ln; i < j; i += Ex::W) { const double t = st.R[i][j]; st.R[i][j] = st.R[i][kmax]; st.R[i][kmax] = t; }
      if (ln == 0) {
        { const double t = st.rd[j]; st.rd[j] = st.rd[kmax]; st.rd[kmax] = t; }
        { const double t = st.gp[j]; st.gp[j] = st.gp[kmax]; st.gp[kmax] = t; }
        { const int t = st.ipvt[j]; st.ipvt[j] = st.ipvt[kmax]; st.ipvt[kmax] = t; }
      }
      ex.sync();
    }
    const double d = st.rd[j];
    if (!(d > 0.0)) {
      // exactly dependent / zero column: qrfac leaves rdiag = 0 and lmpar then treats this and
      // all later columns as singular (row j of R stays zero).
      continue;
    }
    const double rjj = sqrt(d);
    const double inv = 1.0 / rjj;
    const int pj = st.ipvt[j];
    for (int k = j + ln; k < NP; k += Ex::W) {
      if (k == j) { st.R[j][j] = rjj; continue; }
      const int pk = st.ipvt[k];
      double s = A[pj < pk ? tri(pj, pk) : tri(pk, pj)];
      for (int i = 0; i < j; ++i) s -= st.R[i][j] * st.R[i][k];
      const double r = s * inv;
      st.R[j][k] = r;
      st.rd[k] -= r * r;
    }
    ex.sync();
  }
  // qtf = R^-T gp (forward substitution in axpy form; rows with zero pivot give 0)
  for (int k = ln; k < NP; k += Ex::W) st.w1[k] = st.gp[k];
  ex.sync();
#pragma unroll 1
  for (int j = 0; j < NP; ++j) {
    const double rjj = st.R[j][j];
    const double q = (rjj != 0.0) ? st.w1[j] / rjj : 0.0;
    if (ln == 0) st.qtf[j] = q;
    for (int k = j + 1 + ln; k < NP; k += Ex::W) st.w1[k] -= st.R[j][k] * q;
    ex.sync();
  }
  lm_post_factor(ex, st);
}

// Solve min |R P^T x - qtf|^2 + par |D x|^2 (MINPACK qrsolv's job): S = chol(R^T R + par D^2),
// z = S^-1 S^-T rq, x[ipvt[j]] = z[j].  Leaves S / sinv in st; x in xout (unpermuted); uses st.w3, st.y.
template <typename Ex>
IA3_HDN void lm_damped_solve(Ex& ex, LMState& st, double par, double* xout) {
  const int ln = ex.lane();
  // left-looking Cholesky: at step j lane k - j forms entry (j, k) (lane 0 the pivot), the reciprocal
  // of the pivot's square root is broadcast from lane 0
#pragma unroll 1
  for (int j = 0; j < NP; ++j) {
    const int k0 = j + ln;
    double v = 0.0;
    if (k0 < NP) {
      v = st.B0[j][k0];
      if (ln == 0) { const double dj = st.diag[st.ipvt[j]]; v += par * (dj * dj); }
      for (int i = 0; i < j; ++i) v -= st.S[i][j] * st.S[i][k0];
    }
    double sjj = 0.0, inv = 0.0;
    if (ln == 0) { sjj = (v > 0.0) ? sqrt(v) : 0.0; inv = (sjj != 0.0) ? 1.0 / sjj : 0.0; }
    inv = ex.bcast(inv, 0);
    if (k0 < NP) {
      if (ln == 0) { st.S[j][j] = sjj; st.sinv[j] = inv; }
      else st.S[j][k0] = v * inv;
    }
    for (int k = k0 + Ex::W; k < NP; k += Ex::W) {        // executors narrower than NP lanes (host: W = 1)
      double v2 = st.B0[j][k];
      for (int i = 0; i < j; ++i) v2 -= st.S[i][j] * st.S[i][k];
      st.S[j][k] = v2 * inv;
    }
    ex.sync();
  }
  // forward: S^T y = rq  (axpy form)
  for (int k = ln; k < NP; k += Ex::W) st.w3[k] = st.rq[k];
  ex.sync();
#pragma unroll 1
  for (int j = 0; j < NP; ++j) {
    const double y = st.w3[j] * st.sinv[j];
    if (ln == 0) st.y[j] = y;
    for (int k = j + 1 + ln; k < NP; k += Ex::W) st.w3[k] -= st.S[j][k] * y;
    ex.sync();
  }
  // backward: S z = y
#pragma unroll 1
  for (int j = NP - 1; j >= 0; --j) {
    const double z = st.y[j] * st.sinv[j];
    if (ln == 0) xout[st.ipvt[j]] = z;
    for (int i = ln; i < j; i += Ex::W) st.y[i] -= st.S[i][j] * z;
    ex.sync();
  }
}

// MINPACK lmpar: on return st.par is the LM parameter and xout the (positive-sign) step.
template <typename Ex>
IA3_HDN void lm_lmpar(Ex& ex, LMState& st, double* xout) {
  const double dwarf = DBL_MIN;
  const int ln = ex.lane();
  double* wa1 = st.w1;
  double* wa2 = st.w2;
  double (*r)[NP] = st.R;
  const double delta = st.delta;
  int nsing = NP;
  for (int j = 0; j < NP; ++j) if (r[j][j] == 0.0 && nsing == NP) nsing = j;
  for (int j = ln; j < NP; j += Ex::W) {
    wa1[j] = (j < nsing) ? st.qtf[j] : 0.0;
    if (j >= nsing) xout[st.ipvt[j]] = 0.0;
  }
  ex.sync();
  // Gauss-Newton direction: back substitution with the non-singular leading block of R
#pragma unroll 1
  for (int k = 0; k < nsing; ++k) {
    const int j = nsing - 1 - k;
    const double temp = wa1[j] * st.rinv[j];
    if (ln == 0) xout[st.ipvt[j]] = temp;
    for (int i = ln; i < j; i += Ex::W) wa1[i] -= r[i][j] * temp;
    ex.sync();
  }
  for (int j = ln; j < NP; j += Ex::W) wa2[j] = st.diag[j] * xout[j];
  ex.sync();
  int iter = 0;
  double dxnorm = enorm_n(wa2, NP);
  double fp = dxnorm - delta;
  if (fp <= 0.1 * delta) { if (ln == 0) st.par = 0.0; ex.sync(); return; }

  double parl = 0.0;
  if (nsing >= NP) {
    for (int j = ln; j < NP; j += Ex::W) { const int l = st.ipvt[j]; wa1[j] = st.diag[l] * (wa2[l] / dxnorm); }
    ex.sync();
    double n2 = 0.0;
#pragma unroll 1
    for (int j = 0; j < NP; ++j) {               // R^T w = wa1, axpy form; only |w| is needed
      const double t = wa1[j] * st.rinv[j];
      n2 += t * t;
      for (int k = j + 1 + ln; k < NP; k += Ex::W) wa1[k] -= r[j][k] * t;
      ex.sync();
    }
    const double temp = sqrt(n2);
    parl = ((fp / delta) / temp) / temp;
  }
  double g2 = 0.0;
  for (int j = 0; j < NP; ++j) { const double v = st.rq[j] / st.diag[st.ipvt[j]]; g2 += v * v; }
  const double gnorm = sqrt(g2);
  double paru = gnorm / delta;
  if (paru == 0.0) paru = dwarf / fmin(delta, 0.1);

  double par = st.par;
  par = fmax(par, parl);
  par = fmin(par, paru);
  if (par == 0.0) par = gnorm / dxnorm;

  for (;;) {
    ++iter;
    if (par == 0.0) par = fmax(dwarf, 0.001 * paru);
    ex.sync();
    lm_damped_solve(ex, st, par, xout);
    for (int j = ln; j < NP; j += Ex::W) wa2[j] = st.diag[j] * xout[j];
    ex.sync();
    dxnorm = enorm_n(wa2, NP);
    double temp = fp;
    fp = dxnorm - delta;
    if (fabs(fp) <= 0.1 * delta || (parl == 0.0 && fp <= temp && temp < 0.0) || iter == 10) break;
    for (int j = ln; j < NP; j += Ex::W) { const int l = st.ipvt[j]; wa1[j] = st.diag[l] * (wa2[l] / dxnorm); }
    ex.sync();
    double n2 = 0.0;
#pragma unroll 1
    for (int j = 0; j < NP; ++j) {               // S^T w = wa1, axpy form; only |w| is needed
      const double t = wa1[j] * st.sinv[j];
      n2 += t * t;
      for (int k = j + 1 + ln; k < NP; k += Ex::W) wa1[k] -= st.S[j][k] * t;
      ex.sync();
    }
    temp = sqrt(n2);
    const double parc = ((fp / delta) / temp) / temp;
    if (fp > 0.0) parl = fmax(parl, par);
    if (fp < 0.0) paru = fmin(paru, par);
    par = fmax(parl, par + parc);
  }
  ex.sync();
  if (ln == 0) st.par = par;
  ex.sync();
}

// ---- lmder split into the three places where the warp has to evaluate the model ----------
// Scalars of LMState are written by every lane with the same value (benign), arrays by the
// lanes that own the elements.

template <typename Ex>
IA3_HD void lm_init(Ex& ex, LMState& st, const double* x0, double fnorm0) {
  for (int j = ex.lane(); j < NP; j += Ex::W) st.x[j] = x0[j];
  if (ex.lane() == 0) {
    st.fnorm = fnorm0;
    st.par = 0.0;
    st.iter = 1;
    st.nfev = 1;
    st.njev = 0;
    st.info = 0;
    st.xnorm = 0.0;
    st.delta = 0.0;
  }
  ex.sync();
}

// After a Jacobian pass at st.x (A = J^T J, g = J^T f).  Returns false if lmder stops here.
template <typename Ex>
IA3_HDN bool lm_outer(Ex& ex, LMState& st, const LMConfig& cfg, const double* A, const double* g) {
  const int ln = ex.lane();
  lm_factor(ex, st, A, g);
  const int iter = st.iter;
  if (iter == 1) {
    for (int j = ln; j < NP; j += Ex::W) {
      const double d = (st.acn[j] == 0.0) ? 1.0 : st.acn[j];
      st.diag[j] = d;
      st.w3[j] = d * st.x[j];
    }
    ex.sync();
    const double xnorm = enorm_n(st.w3, NP);
    double delta = cfg.factor * xnorm;
    if (delta == 0.0) delta = cfg.factor;
    ex.sync();
    if (ln == 0) { st.xnorm = xnorm; st.delta = delta; }
  }
  double gnorm = 0.0;
  const double fnorm = st.fnorm;
  if (fnorm != 0.0) {
    for (int j = ln; j < NP; j += Ex::W) st.y[j] = st.qtf[j] / fnorm;     // lmder divides every term
    ex.sync();
    for (int j = ln; j < NP; j += Ex::W) {
      const int l = st.ipvt[j];
      if (st.acn[l] != 0.0) {
        double sum = 0.0;
        for (int i = 0; i <= j; ++i) sum += st.R[i][j] * st.y[i];
        gnorm = fmax(gnorm, fabs(sum / st.acn[l]));
      }
    }
  }
  gnorm = ex.allmax(gnorm);
  ex.sync();
  if (ln == 0) { st.njev += 1; st.gnorm = gnorm; }
  if (gnorm <= cfg.gtol) { if (ln == 0) st.info = 4; ex.sync(); return false; }
  for (int j = ln; j < NP; j += Ex::W) st.diag[j] = fmax(st.diag[j], st.acn[j]);
  ex.sync();
  return true;
}

// Compute the LM step and the trial point st.xt (model must then be evaluated at st.xt).
template <typename Ex>
IA3_HDN void lm_propose(Ex& ex, LMState& st) {
  const int ln = ex.lane();
  lm_lmpar(ex, st, st.p);
  ex.sync();
  for (int j = ln; j < NP; j += Ex::W) {
    const double pj = -st.p[j];
    st.p[j] = pj;
    st.xt[j] = st.x[j] + pj;
    st.w3[j] = st.diag[j] * pj;
  }
  ex.sync();
  const double pnorm = enorm_n(st.w3, NP);
  const double delta = st.delta;
  const int iter = st.iter;
  ex.sync();
  if (ln == 0) {
    st.pnorm = pnorm;
    if (iter == 1) st.delta = fmin(delta, pnorm);
  }
  ex.sync();
}

enum { LM_RETRY = 0, LM_ACCEPTED = 1, LM_DONE = 2 };

// Given fnorm1 = |f(st.xt)|: ratio test, trust-region update, convergence tests.  Every lane
// computes the (identical) decision; lane-strided / lane-0 writes.
template <typename Ex>
IA3_HDN int lm_judge(Ex& ex, LMState& st, const LMConfig& cfg, double fnorm1) {
  const int ln = ex.lane();
  const double fnorm = st.fnorm;
  const double pnorm = st.pnorm;
  double delta = st.delta, par = st.par, xnorm = st.xnorm;
  const double gnorm = st.gnorm;
  const int nfev = st.nfev + 1;
  double actred = -1.0;
  if (0.1 * fnorm1 < fnorm) { const double q = fnorm1 / fnorm; actred = 1.0 - q * q; }
  // w3 = R * p[ipvt]  (row i: sum_{j>=i} R[i][j] p[ipvt[j]], j ascending as lmder accumulates)
  for (int i = ln; i < NP; i += Ex::W) {
    double s = 0.0;
    for (int j = i; j < NP; ++j) s += st.R[i][j] * st.p[st.ipvt[j]];
    st.w3[i] = s;
  }
  ex.sync();
  const double temp1 = enorm_n(st.w3, NP) / fnorm;
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
  const bool accepted = ratio >= 1.0e-4;
  ex.sync();
  if (accepted) {
    for (int j = ln; j < NP; j += Ex::W) { const double v = st.xt[j]; st.x[j] = v; st.w3[j] = st.diag[j] * v; }
    ex.sync();
    xnorm = enorm_n(st.w3, NP);
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
  ex.sync();
  if (ln == 0) {
    st.nfev = nfev;
    st.fnorm1 = fnorm1;
    st.delta = delta;
    st.par = par;
    st.info = info;
    if (accepted) { st.xnorm = xnorm; st.fnorm = fnorm1; st.iter += 1; }
  }
  ex.sync();
  if (info != 0) return LM_DONE;
  return accepted ? LM_ACCEPTED : LM_RETRY;
}

}  // namespace ia3
