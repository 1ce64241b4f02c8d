// MINPACK-lmder-faithful Levenberg-Marquardt driver, driven from the normal-equation sums
// (J^T J, J^T f, |f|^2) that a warp accumulates per spot.
//
// The reference fits every spot with scipy.optimize.leastsq(calc_eps, p0, Dfun=calc_jac)
// (External/Fitting_v4.py:388, External/Fitting_v3.py:250), i.e. MINPACK lmder with
// ftol = xtol = 1.49012e-8, gtol = 0, factor = 100, mode = 1.  SURVEY.md App. G / B.8 shows
// that the result is only reproducible if the *trust-region logic* is lmder's, so this file
// restates lmder / lmpar / qrsolv (public-domain MINPACK, Argonne 1980) step by step.  The one
// deliberate change: lmder's Householder QR of the m x 10 Jacobian (qrfac with column
// pivoting) is replaced by a pivoted Cholesky factorisation of the FP64 10x10 matrix J^T J
// with qrfac's pivot rule (largest remaining column norm, first index wins ties).  R^T R =
// P^T J^T J P, so R agrees with qrfac's R up to row signs, which cancel everywhere R is used
// together with qtf = R^-T P^T J^T f.
//
// Everything here is scalar FP64 on a 10-vector / 10x10 matrix and is executed by ONE lane of
// the warp that owns the spot (state lives in shared memory).  No per-voxel work.
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
  double R[NP][NP];    // upper triangle: R; strict lower triangle: scratch for qrsolv's S^T
  double sdiag[NP];
  double w1[NP], w2[NP], w3[NP];
  int ipvt[NP];
  double fnorm, fnorm1, xnorm, delta, par, gnorm, pnorm;
  int iter, nfev, njev, info;
};

IA3_HD double enorm_n(const double* v, int n) {
  double s = 0.0;
  for (int i = 0; i < n; ++i) s += v[i] * v[i];
  return sqrt(s);
}

// Pivoted Cholesky of the packed symmetric A (55 entries) with MINPACK qrfac's pivoting.
// Outputs st.R (upper), st.ipvt, st.acn and st.qtf = R^-T (g permuted).
IA3_HDN void lm_factor(LMState& st, const double* A, const double* g) {
  double M[NP][NP];
  double rd[NP];  // remaining squared norms (Schur complement diagonal)
  double gp[NP];
  for (int i = 0; i < NP; ++i) {
    for (int j = i; j < NP; ++j) { double a = A[tri(i, j)]; M[i][j] = a; M[j][i] = a; }
    st.acn[i] = sqrt(A[tri(i, i)]);
    rd[i] = A[tri(i, i)];
    st.ipvt[i] = i;
    gp[i] = g[i];
    for (int j = 0; j < NP; ++j) st.R[i][j] = 0.0;
  }
  // M is permuted symmetrically in place as pivots are chosen.
  for (int j = 0; j < NP; ++j) {
    int kmax = j;
    for (int k = j + 1; k < NP; ++k) if (rd[k] > rd[kmax]) kmax = k;
    if (kmax != j) {
      for (int i = 0; i < NP; ++i) { double t = M[i][j]; M[i][j] = M[i][kmax]; M[i][kmax] = t; }
      for (int i = 0; i < NP; ++i) { double t = M[j][i]; M[j][i] = M[kmax][i]; M[kmax][i] = t; }
      for (int i = 0; i < j; ++i) { double t = st.R[i][j]; st.R[i][j] = st.R[i][kmax]; st.R[i][kmax] = t; }
      { double t = rd[j]; rd[j] = rd[kmax]; rd[kmax] = t; }
      { double t = gp[j]; gp[j] = gp[kmax]; gp[kmax] = t; }
      { int t = st.ipvt[j]; st.ipvt[j] = st.ipvt[kmax]; st.ipvt[kmax] = t; }
    }
    double d = rd[j];
    if (!(d > 0.0)) {
      // exactly dependent / zero column: qrfac leaves rdiag = 0 and lmpar then treats this and
      // all later columns as singular.
      st.R[j][j] = 0.0;
      for (int k = j + 1; k < NP; ++k) st.R[j][k] = 0.0;
      continue;
    }
    double rjj = sqrt(d);
    st.R[j][j] = rjj;
    for (int k = j + 1; k < NP; ++k) {
      double s = M[j][k];
      for (int i = 0; i < j; ++i) s -= st.R[i][j] * st.R[i][k];
      double r = s / rjj;
      st.R[j][k] = r;
      rd[k] -= r * r;
    }
  }
  // qtf = R^-T gp  (forward substitution; rows with zero pivot give 0)
  for (int j = 0; j < NP; ++j) {
    double s = gp[j];
    for (int i = 0; i < j; ++i) s -= st.R[i][j] * st.qtf[i];
    st.qtf[j] = (st.R[j][j] != 0.0) ? s / st.R[j][j] : 0.0;
  }
}

// MINPACK qrsolv: solve min |R P^T x - qtb|^2 + |D x|^2 given d = sqrt(par)*diag.
IA3_HDN void lm_qrsolv(LMState& st, const double* d, double* x /*out, unpermuted*/, double* wa) {
  double (*r)[NP] = st.R;
  double xsave[NP];
  for (int j = 0; j < NP; ++j) {
    for (int i = j; i < NP; ++i) r[i][j] = r[j][i];
    xsave[j] = r[j][j];
    wa[j] = st.qtf[j];
  }
  for (int j = 0; j < NP; ++j) {
    int l = st.ipvt[j];
    if (d[l] != 0.0) {
      for (int k = j; k < NP; ++k) st.sdiag[k] = 0.0;
      st.sdiag[j] = d[l];
      double qtbpj = 0.0;
      for (int k = j; k < NP; ++k) {
        if (st.sdiag[k] == 0.0) continue;
        double c, s;
        if (fabs(r[k][k]) < fabs(st.sdiag[k])) {
          double cotan = r[k][k] / st.sdiag[k];
          s = 0.5 / sqrt(0.25 + 0.25 * (cotan * cotan));
          c = s * cotan;
        } else {
          double tn = st.sdiag[k] / r[k][k];
          c = 0.5 / sqrt(0.25 + 0.25 * (tn * tn));
          s = c * tn;
        }
        r[k][k] = c * r[k][k] + s * st.sdiag[k];
        double temp = c * wa[k] + s * qtbpj;
        qtbpj = -s * wa[k] + c * qtbpj;
        wa[k] = temp;
        for (int i = k + 1; i < NP; ++i) {
          double t2 = c * r[i][k] + s * st.sdiag[i];
          st.sdiag[i] = -s * r[i][k] + c * st.sdiag[i];
          r[i][k] = t2;
        }
      }
    }
    st.sdiag[j] = r[j][j];
    r[j][j] = xsave[j];
  }
  int nsing = NP;
  for (int j = 0; j < NP; ++j) {
    if (st.sdiag[j] == 0.0 && nsing == NP) nsing = j;
    if (nsing < NP) wa[j] = 0.0;
  }
  for (int k = 0; k < nsing; ++k) {
    int j = nsing - 1 - k;
    double sum = 0.0;
    for (int i = j + 1; i < nsing; ++i) sum += r[i][j] * wa[i];
    wa[j] = (wa[j] - sum) / st.sdiag[j];
  }
  for (int j = 0; j < NP; ++j) x[st.ipvt[j]] = wa[j];
}

// MINPACK lmpar: on return st.par is the LM parameter and xout the (positive-sign) step.
IA3_HDN void lm_lmpar(LMState& st, double* xout) {
  const double dwarf = DBL_MIN;
  double* wa1 = st.w1;
  double* wa2 = st.w2;
  double (*r)[NP] = st.R;
  const double delta = st.delta;
  int nsing = NP;
  for (int j = 0; j < NP; ++j) {
    wa1[j] = st.qtf[j];
    if (r[j][j] == 0.0 && nsing == NP) nsing = j;
    if (nsing < NP) wa1[j] = 0.0;
  }
  for (int k = 0; k < nsing; ++k) {
    int j = nsing - 1 - k;
    wa1[j] = wa1[j] / r[j][j];
    double temp = wa1[j];
    for (int i = 0; i < j; ++i) wa1[i] -= r[i][j] * temp;
  }
  for (int j = 0; j < NP; ++j) xout[st.ipvt[j]] = wa1[j];

  int iter = 0;
  for (int j = 0; j < NP; ++j) wa2[j] = st.diag[j] * xout[j];
  double dxnorm = enorm_n(wa2, NP);
  double fp = dxnorm - delta;
  if (fp <= 0.1 * delta) { st.par = 0.0; return; }

  double parl = 0.0;
  if (nsing >= NP) {
    for (int j = 0; j < NP; ++j) { int l = st.ipvt[j]; wa1[j] = st.diag[l] * (wa2[l] / dxnorm); }
    for (int j = 0; j < NP; ++j) {
      double sum = 0.0;
      for (int i = 0; i < j; ++i) sum += r[i][j] * wa1[i];
      wa1[j] = (wa1[j] - sum) / r[j][j];
    }
    double temp = enorm_n(wa1, NP);
    parl = ((fp / delta) / temp) / temp;
  }
  for (int j = 0; j < NP; ++j) {
    double sum = 0.0;
    for (int i = 0; i <= j; ++i) sum += r[i][j] * st.qtf[i];
    wa1[j] = sum / st.diag[st.ipvt[j]];
  }
  double gnorm = enorm_n(wa1, NP);
  double paru = gnorm / delta;
  if (paru == 0.0) paru = dwarf / fmin(delta, 0.1);

  double par = st.par;
  par = fmax(par, parl);
  par = fmin(par, paru);
  if (par == 0.0) par = gnorm / dxnorm;

  for (;;) {
    ++iter;
    if (par == 0.0) par = fmax(dwarf, 0.001 * paru);
    double temp = sqrt(par);
    for (int j = 0; j < NP; ++j) wa1[j] = temp * st.diag[j];
    lm_qrsolv(st, wa1, xout, st.w3);
    for (int j = 0; j < NP; ++j) wa2[j] = st.diag[j] * xout[j];
    dxnorm = enorm_n(wa2, NP);
    temp = fp;
    fp = dxnorm - delta;
    if (fabs(fp) <= 0.1 * delta || (parl == 0.0 && fp <= temp && temp < 0.0) || iter == 10) break;
    for (int j = 0; j < NP; ++j) { int l = st.ipvt[j]; wa1[j] = st.diag[l] * (wa2[l] / dxnorm); }
    for (int j = 0; j < NP; ++j) {
      wa1[j] = wa1[j] / st.sdiag[j];
      double t2 = wa1[j];
      for (int i = j + 1; i < NP; ++i) wa1[i] -= r[i][j] * t2;
    }
    temp = enorm_n(wa1, NP);
    double parc = ((fp / delta) / temp) / temp;
    if (fp > 0.0) parl = fmax(parl, par);
    if (fp < 0.0) paru = fmin(paru, par);
    par = fmax(parl, par + parc);
  }
  st.par = par;
}

// ---- lmder split into the three places where the warp has to evaluate the model ----------

IA3_HD void lm_init(LMState& st, const double* x0, double fnorm0) {
  for (int j = 0; j < NP; ++j) st.x[j] = x0[j];
  st.fnorm = fnorm0;
  st.par = 0.0;
  st.iter = 1;
  st.nfev = 1;
  st.njev = 0;
  st.info = 0;
  st.xnorm = 0.0;
  st.delta = 0.0;
}

// After a Jacobian pass at st.x (A = J^T J, g = J^T f).  Returns false if lmder stops here.
IA3_HDN bool lm_outer(LMState& st, const LMConfig& cfg, const double* A, const double* g) {
  st.njev += 1;
  lm_factor(st, A, g);
  if (st.iter == 1) {
    for (int j = 0; j < NP; ++j) { st.diag[j] = st.acn[j]; if (st.acn[j] == 0.0) st.diag[j] = 1.0; }
    for (int j = 0; j < NP; ++j) st.w3[j] = st.diag[j] * st.x[j];
    st.xnorm = enorm_n(st.w3, NP);
    st.delta = cfg.factor * st.xnorm;
    if (st.delta == 0.0) st.delta = cfg.factor;
  }
  double gnorm = 0.0;
  if (st.fnorm != 0.0) {
    for (int j = 0; j < NP; ++j) {
      int l = st.ipvt[j];
      if (st.acn[l] != 0.0) {
        double sum = 0.0;
        for (int i = 0; i <= j; ++i) sum += st.R[i][j] * (st.qtf[i] / st.fnorm);
        gnorm = fmax(gnorm, fabs(sum / st.acn[l]));
      }
    }
  }
  st.gnorm = gnorm;
  if (gnorm <= cfg.gtol) { st.info = 4; return false; }
  for (int j = 0; j < NP; ++j) st.diag[j] = fmax(st.diag[j], st.acn[j]);
  return true;
}

// Compute the LM step and the trial point st.xt (model must then be evaluated at st.xt).
IA3_HDN void lm_propose(LMState& st) {
  lm_lmpar(st, st.p);
  for (int j = 0; j < NP; ++j) {
    st.p[j] = -st.p[j];
    st.xt[j] = st.x[j] + st.p[j];
    st.w3[j] = st.diag[j] * st.p[j];
  }
  st.pnorm = enorm_n(st.w3, NP);
  if (st.iter == 1) st.delta = fmin(st.delta, st.pnorm);
}

enum { LM_RETRY = 0, LM_ACCEPTED = 1, LM_DONE = 2 };

// Given fnorm1 = |f(st.xt)|: ratio test, trust-region update, convergence tests.
IA3_HDN int lm_judge(LMState& st, const LMConfig& cfg, double fnorm1) {
  st.nfev += 1;
  st.fnorm1 = fnorm1;
  const double fnorm = st.fnorm;
  double actred = -1.0;
  if (0.1 * fnorm1 < fnorm) { double q = fnorm1 / fnorm; actred = 1.0 - q * q; }
  for (int j = 0; j < NP; ++j) st.w3[j] = 0.0;
  for (int j = 0; j < NP; ++j) {
    double temp = st.p[st.ipvt[j]];
    for (int i = 0; i <= j; ++i) st.w3[i] += st.R[i][j] * temp;
  }
  double temp1 = enorm_n(st.w3, NP) / fnorm;
  double temp2 = (sqrt(st.par) * st.pnorm) / fnorm;
  double prered = temp1 * temp1 + temp2 * temp2 / 0.5;
  double dirder = -(temp1 * temp1 + temp2 * temp2);
  double ratio = 0.0;
  if (prered != 0.0) ratio = actred / prered;
  if (ratio <= 0.25) {
    double temp = 0.5;
    if (actred < 0.0) temp = 0.5 * dirder / (dirder + 0.5 * actred);
    if (0.1 * fnorm1 >= fnorm || temp < 0.1) temp = 0.1;
    st.delta = temp * fmin(st.delta, st.pnorm / 0.1);
    st.par = st.par / temp;
  } else if (st.par == 0.0 || ratio >= 0.75) {
    st.delta = st.pnorm / 0.5;
    st.par = 0.5 * st.par;
  }
  bool accepted = false;
  if (ratio >= 1.0e-4) {
    for (int j = 0; j < NP; ++j) { st.x[j] = st.xt[j]; st.w3[j] = st.diag[j] * st.x[j]; }
    st.xnorm = enorm_n(st.w3, NP);
    st.fnorm = fnorm1;
    st.iter += 1;
    accepted = true;
  }
  const bool small = fabs(actred) <= cfg.ftol && prered <= cfg.ftol && 0.5 * ratio <= 1.0;
  if (small) st.info = 1;
  if (st.delta <= cfg.xtol * st.xnorm) st.info = 2;
  if (small && st.info == 2) st.info = 3;
  if (st.info != 0) return LM_DONE;
  if (st.nfev >= cfg.maxfev) st.info = 5;
  if (fabs(actred) <= DBL_EPSILON && prered <= DBL_EPSILON && 0.5 * ratio <= 1.0) st.info = 6;
  if (st.delta <= DBL_EPSILON * st.xnorm) st.info = 7;
  if (st.gnorm <= DBL_EPSILON) st.info = 8;
  if (st.info != 0) return LM_DONE;
  return accepted ? LM_ACCEPTED : LM_RETRY;
}

}  // namespace ia3
