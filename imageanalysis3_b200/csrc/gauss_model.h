// The constrained 10-parameter rotated 3D Gaussian of the reference's GaussianFit, in its two
// "personalities":
//   V4: External/Fitting_v4.py:165-396  (overflow guards, bk clip in calc_f only)
//   V3: External/Fitting_v3.py:50-257   (no guards, to_center quirk at :86 -- c2 uses exp(-c1_)
//       in its numerator -- and the optional width prior weight_sigma, :173-177 / :216-221)
// Raw parameter order (both): p_ = [bk, h, xp, yp, zp, w1, w2, w3, pp, tp]; the model axes
// "x,y,z" are the image axes 0,1,2 (z, x, y of the stack) because the callers pass
// X = [z_keep, x_keep, y_keep] (Fitting_v4.py:614, :173).
//
// Two layers:
//   ModelConsts  -- everything that depends on the parameters only, computed once per
//                   evaluation in FP64 (scalar, a few dozen transcendental calls);
//   eval_res / eval_jac -- the per-voxel part, templated on the arithmetic type T
//                   (double = reference precision under numpy>=2, SURVEY App. B.5;
//                    float = the FP32-FMA fast mode).
// The Jacobian the reference hands to MINPACK is cast to float32 (Fitting_v4.py:365); eval_jac
// reproduces that rounding, so J^T J is accumulated from float32-rounded entries.
#pragma once
#include "ia3_common.h"

namespace ia3 {

struct FitParams {       // per-launch constants of a GaussianFit family
  double min_w2, max_w2; // min_w^2, max_w^2               (Fitting_v4.py:168-169)
  double delta;          // delta_center
  double init_wt[3];     // v3: transformed init widths (Fitting_v3.py:75); v4: unused
  double weight_sigma;   // v3 width prior (0 = off)
  int personality;       // 3 or 4
};

struct ModelConsts {
  // residual
  double ebk_f;          // exp(bk) as used in calc_f  (v4: bk clipped to +-709.78)
  double h;              // log-height
  double c[3];           // centre
  double q[6];           // x2c, y2c, z2c, xyc, xzc, yzc
  double pen;            // v3: weight_sigma * |init_w - w|  (added to every residual)
  // jacobian
  double ebk_j;          // exp(bk) unclipped (calc_jac f1)
  double ncen[3];        // norm_xp, norm_yp, norm_zp
  double a6[6], a7[6], a8[6], a9[6], a10[6];  // quadratic-form coefficients of f6..f10 over
                         // (xt2, yt2, zt2, xtyt, xtzt, ytzt), with norm_w / norm_p / norm_t folded in
  double jpen[3];        // v3: +-weight_sigma added to f6..f8
};

// One out-of-line copy of the FP64 exp for the per-parameter (scalar) code: model_consts alone has
// ~20 call sites and an inlined exp is ~60 instructions.
IA3_HDN double exp_s(double v) { return exp(v); }

// ---- the transcendental part of the parameter transforms, one exp per slot -------------------
// All FP64 exps a model_consts call needs, as a table so that a warp evaluates them in parallel
// (one slot per lane, fit_spot.h: build_consts_par) instead of one lane evaluating 19 in a row:
//   0 exp(bk) as used by calc_f (v4: bk clipped)   2..4 exp(xp|yp|zp) (v3: exp(-xp) ...)
//   5..7 exp(w1..w3)   8 exp(pp)   9 exp(tp)        10 exp(bk) unclipped (calc_jac f1)
//   11..13 exp(-|xp|..)   14..16 exp(-|w1|..)   17 exp(-|pp|/2)   18 exp(-|tp|/2)       (1 unused)
constexpr int NEXP = 19;

IA3_HD double exp_slot_arg(const FitParams& fp, const double* x, int i) {
  const bool v4 = (fp.personality == 4);
  if (i == 0) {
    const double bk = x[0];
    if (!v4) return bk;
    return bk < -709.78 ? -709.78 : (bk > 709.78 ? 709.78 : bk);      // np.clip, Fitting_v4.py:287 (NaN passes)
  }
  if (i < 10) return (!v4 && i >= 2 && i <= 4) ? -x[i] : x[i];
  if (i == 10) return x[0];
  const double a = -fabs(x[i - 9]);
  return (i >= 17) ? a / 2 : a;
}

IA3_HD double v4_sigmoid_guarded(double a, double ea, double lo_val, double hi_val, double num, double off) {
  // value = num/(1+exp(a)) + off with the reference's saturation guards at |a| >= log(DBL_MAX); ea = exp(a)
  const double LOGMAX = 709.782712893384;  // np.log(np.finfo(np.float64).max)
  if (a >= LOGMAX) return lo_val;
  if (a <= -LOGMAX) return hi_val;
  return num / (1.0 + ea) + off;
}

IA3_HD double norm_w_fn(double w, double e /* exp(-|w|) */, double minw, double maxw) {
  // Fitting_v4.py:369-375 / Fitting_v3.py:232-238
  const double d = (w > 0) ? maxw * e + minw : minw * e + maxw;
  return 0.5 * (maxw - minw) * e / (d * d);
}

// The parameter-only part of the model in two layers, so that a warp can spread the expensive bit over its lanes
// (lm_warp.h: build_consts) while the serial form below evaluates exactly the same expressions:
//   consts_scalar(k)   the 21 scalars that cost a division or a square root each -- k = 0 t, 1 p, 2..4 ws1..3,
//                      5..7 centre, 8..10 norm_xp.., 11..13 norm_w1..3, 14 norm_p, 15 norm_t (from x and the exp
//                      table only), 16..18 s_i = 1 / ws_i, 19 tc, 20 pc (from scalars 0..4)
//   model_consts_tail  everything else: products and sums of those scalars
constexpr int NSCAL = 21;
constexpr int NSCAL_FIRST = 16;         // scalars 0..15 are independent of each other

IA3_HD double consts_scalar(const FitParams& fp, const double* cen_est, const double* x, const double* e, const double* S, int k) {
  const bool v4 = (fp.personality == 4);
  const double d = fp.delta, minw = fp.min_w2, maxw = fp.max_w2, dws = maxw - minw;
  const double LOGMAX = 709.782712893384;
  switch (k) {
    case 0: return v4 ? v4_sigmoid_guarded(x[9], e[9], -1.0, 1.0, 2.0, -1.0) : 2.0 / (1.0 + e[9]) - 1.0;            // t
    case 1: return v4 ? v4_sigmoid_guarded(x[8], e[8], -1.0, 1.0, 2.0, -1.0) : 2.0 / (1.0 + e[8]) - 1.0;            // p
    case 2: case 3: case 4: {
      const int i = k - 2;
      return v4 ? v4_sigmoid_guarded(x[5 + i], e[5 + i], minw, dws + minw, dws, minw) : dws / (1.0 + e[5 + i]) + minw;
    }
    case 5: case 6: case 7: {
      const int i = k - 5;
      if (v4) {      // c = 2*delta/(1+exp(c_)) - delta + center_est   (left-to-right as written, :198)
        const double raw = x[2 + i];
        if (raw >= LOGMAX) return -d + cen_est[i];
        if (raw <= -LOGMAX) return d + cen_est[i];
        return 2.0 * d / (1.0 + e[2 + i]) - d + cen_est[i];
      }
      // v3: e[2..4] = exp(-xp), exp(-yp), exp(-zp); the third centre mixes in exp(-yp)   (Fitting_v3.py:86, sic)
      const double num = (i == 2) ? e[3] : e[2 + i];
      return 2.0 * d * num / (1.0 + e[2 + i]) - d + cen_est[i];
    }
    case 8: case 9: case 10: { const double ex = e[11 + (k - 8)]; return -d * ex / ((1 + ex) * (1 + ex)); }
    case 11: case 12: case 13: return norm_w_fn(x[5 + (k - 11)], e[14 + (k - 11)], minw, maxw);
    case 14: { const double e_p = e[17]; return e_p / (1 + e_p * e_p); }
    case 15: { const double e_t = e[18]; return e_t / (1 + e_t * e_t); }
    case 16: case 17: case 18: return 1.0 / S[2 + (k - 16)];
    case 19: return sqrt(1 - S[0] * S[0]);
    default: return sqrt(1 - S[1] * S[1]);
  }
}

// want_jac = false skips the Jacobian-only constants.  e[NEXP]: exp table at x (exp_slot_arg); S: consts_scalar(0..20).
IA3_HD void model_consts_tail(const FitParams& fp, const double* x, const double* e, const double* S, bool want_jac,
                              ModelConsts& mc) {
  const double w1 = x[5], w2 = x[6], w3 = x[7];
  const bool v4 = (fp.personality == 4);
  const double t = S[0], p = S[1];
  mc.c[0] = S[5]; mc.c[1] = S[6]; mc.c[2] = S[7];
  const double p2 = p * p, t2 = t * t, tc2 = 1 - t2, pc2 = 1 - p2;
  const double tc = S[19], pc = S[20];
  const double s1 = S[16], s2 = S[17], s3 = S[18];
  mc.q[0] = pc2 * tc2 * s1 + t2 * s2 + p2 * tc2 * s3;
  mc.q[1] = pc2 * t2 * s1 + tc2 * s2 + p2 * t2 * s3;
  mc.q[2] = p2 * s1 + pc2 * s3;
  mc.q[3] = 2 * tc * t * (pc2 * s1 - s2 + p2 * s3);
  mc.q[4] = 2 * p * pc * tc * (s3 - s1);
  mc.q[5] = 2 * p * pc * t * (s3 - s1);
  mc.h = x[1];
  mc.ebk_f = e[0];
  mc.pen = 0.0;
  if (!v4 && fp.weight_sigma > 0) {
    const double d0 = fp.init_wt[0] - w1, d1 = fp.init_wt[1] - w2, d2 = fp.init_wt[2] - w3;
    mc.pen = fp.weight_sigma * sqrt(d0 * d0 + d1 * d1 + d2 * d2);
  }
  if (!want_jac) return;
  mc.ebk_j = e[10];
  mc.ncen[0] = S[8]; mc.ncen[1] = S[9]; mc.ncen[2] = S[10];
  const double nw1 = S[11], nw2 = S[12], nw3 = S[13];
  const double norm_p = S[14], norm_t = S[15];
  // order of the six monomials: xt2, yt2, zt2, xtyt, xtzt, ytzt
  // f6 (Fitting_v4.py:355)
  mc.a6[0] = -pc2 * tc2 * nw1; mc.a6[1] = -pc2 * t2 * nw1; mc.a6[2] = -p2 * nw1;
  mc.a6[3] = -2 * pc2 * t * tc * nw1; mc.a6[4] = 2 * p * pc * tc * nw1; mc.a6[5] = 2 * p * pc * t * nw1;
  // f7 (:356)
  mc.a7[0] = -t2 * nw2; mc.a7[1] = -tc2 * nw2; mc.a7[2] = 0.0;
  mc.a7[3] = 2 * t * tc * nw2; mc.a7[4] = 0.0; mc.a7[5] = 0.0;
  // f8 (:357)
  mc.a8[0] = -p2 * tc2 * nw3; mc.a8[1] = -p2 * t2 * nw3; mc.a8[2] = -pc2 * nw3;
  mc.a8[3] = -2 * p2 * t * tc * nw3; mc.a8[4] = -2 * p * pc * tc * nw3; mc.a8[5] = -2 * p * pc * t * nw3;
  // f9 (:360) = f2*(s3-s1)*((2pc2-1)(tc xtzt + t ytzt) + p pc (tc2 xt2 + 2 t tc xtyt + t2 yt2 - zt2))*norm_p
  {
    const double k = (s3 - s1) * norm_p, a = 2 * pc2 - 1.0, b = p * pc;
    mc.a9[0] = k * b * tc2; mc.a9[1] = k * b * t2; mc.a9[2] = -k * b;
    mc.a9[3] = k * b * 2 * t * tc; mc.a9[4] = k * a * tc; mc.a9[5] = k * a * t;
  }
  // f10 (:363) = f2*((pc2 s1 - s2 + p2 s3)(t tc (yt2 - xt2) - (t2 - tc2) xtyt) + p pc (s1 - s3)(t xtzt - tc ytzt))*norm_t
  {
    const double u = (pc2 * s1 - s2 + p2 * s3) * norm_t, v = p * pc * (s1 - s3) * norm_t;
    mc.a10[0] = -u * t * tc; mc.a10[1] = u * t * tc; mc.a10[2] = 0.0;
    mc.a10[3] = -u * (t2 - tc2); mc.a10[4] = v * t; mc.a10[5] = -v * tc;
  }
  mc.jpen[0] = mc.jpen[1] = mc.jpen[2] = 0.0;
  if (!v4 && fp.weight_sigma != 0) {
    // int(init_w>w)*ws - int(init_w<w)*ws   (Fitting_v3.py:217-221)
    const double wv[3] = {w1, w2, w3};
    for (int i = 0; i < 3; ++i)
      mc.jpen[i] = (fp.init_wt[i] > wv[i] ? fp.weight_sigma : 0.0) - (fp.init_wt[i] < wv[i] ? fp.weight_sigma : 0.0);
  }
}

IA3_HD void model_consts_e(const FitParams& fp, const double* cen_est, const double* x, const double* e, bool want_jac,
                           ModelConsts& mc) {
  double S[NSCAL];
  for (int k = 0; k < NSCAL; ++k) {
    // the Jacobian-only scalars (8..15) read exp slots that are not filled when want_jac is false
    S[k] = (!want_jac && k >= 8 && k < NSCAL_FIRST) ? 0.0 : consts_scalar(fp, cen_est, x, e, S, k);
  }
  model_consts_tail(fp, x, e, S, want_jac, mc);
}


// serial convenience form (every exp evaluated here)
IA3_HD void model_consts(const FitParams& fp, const double* cen_est, const double* x, bool want_jac, ModelConsts& mc) {
  double e[NEXP];
  const int n = want_jac ? NEXP : 10;
  for (int i = 0; i < n; ++i) e[i] = (i == 1) ? 0.0 : exp_s(exp_slot_arg(fp, x, i));
  model_consts_e(fp, cen_est, x, e, want_jac, mc);
}

// Per-voxel constants narrowed to the evaluation type, with the centre expressed relative to
// an integer origin so that the float mode keeps sub-pixel accuracy at x,y ~ 2000.
template <typename T>
struct VoxConsts {
  T ebk_f, h, c[3], q[6], pen;
  T ebk_j, ncen[3], a6[6], a7[6], a8[6], a9[6], a10[6], jpen[3];
};

template <typename T>
IA3_HD void narrow_consts(const ModelConsts& mc, const double* origin, bool want_jac, VoxConsts<T>& vc) {
  vc.ebk_f = (T)mc.ebk_f; vc.h = (T)mc.h; vc.pen = (T)mc.pen;
  for (int i = 0; i < 3; ++i) vc.c[i] = (T)(mc.c[i] - origin[i]);
  for (int i = 0; i < 6; ++i) vc.q[i] = (T)mc.q[i];
  if (!want_jac) return;
  vc.ebk_j = (T)mc.ebk_j;
  for (int i = 0; i < 3; ++i) { vc.ncen[i] = (T)mc.ncen[i]; vc.jpen[i] = (T)mc.jpen[i]; }
  for (int i = 0; i < 6; ++i) {
    vc.a6[i] = (T)mc.a6[i]; vc.a7[i] = (T)mc.a7[i]; vc.a8[i] = (T)mc.a8[i];
    vc.a9[i] = (T)mc.a9[i]; vc.a10[i] = (T)mc.a10[i];
  }
}

// model_consts + narrow_consts as ONE out-of-line routine writing straight into the (shared-memory)
// per-voxel constants: the intermediate ModelConsts then lives in registers, not on the stack.
template <typename T>
IA3_HDN void build_consts(const FitParams& fp, const double* cen_est, const double* origin, const double* x,
                          bool want_jac, VoxConsts<T>& vc) {
  ModelConsts mc;
  model_consts(fp, cen_est, x, want_jac, mc);
  narrow_consts<T>(mc, origin, want_jac, vc);
}

// the non-transcendental rest of build_consts, given the exp table (shared memory) of a warp
template <typename T>
IA3_HDN void finish_consts(const FitParams& fp, const double* cen_est, const double* origin, const double* x,
                           const double* e, bool want_jac, VoxConsts<T>& vc) {
  ModelConsts mc;
  model_consts_e(fp, cen_est, x, e, want_jac, mc);
  narrow_consts<T>(mc, origin, want_jac, vc);
}

// the cheap rest of build_consts given the exp table and the 21 scalars (lm_warp.h computes those over a warp's lanes)
template <typename T>
IA3_HDN void finish_consts_scalars(const FitParams& fp, const double* origin, const double* x, const double* e,
                                   const double* S, bool want_jac, VoxConsts<T>& vc) {
  ModelConsts mc;
  model_consts_tail(fp, x, e, S, want_jac, mc);
  narrow_consts<T>(mc, origin, want_jac, vc);
}

template <typename T> IA3_HD T exp_t(T v);
template <> IA3_HD double exp_t<double>(double v) { return exp(v); }
template <> IA3_HD float exp_t<float>(float v) { return expf(v); }

// Gaussian part f0 = exp(h - xsigmax/2) at voxel (X0,X1,X2) given relative to the origin.
template <typename T>
IA3_HD T eval_f0(const VoxConsts<T>& vc, T X0, T X1, T X2) {
  const T xt = X0 - vc.c[0], yt = X1 - vc.c[1], zt = X2 - vc.c[2];
  const T xs = vc.q[0] * xt * xt + vc.q[1] * yt * yt + vc.q[2] * zt * zt + vc.q[3] * xt * yt + vc.q[4] * xt * zt +
               vc.q[5] * yt * zt;
  return exp_t<T>(vc.h - (T)0.5 * xs);
}

// residual  f - im (+ v3 penalty)                        (calc_eps, Fitting_v4.py:320)
template <typename T>
IA3_HD T eval_res(const VoxConsts<T>& vc, T X0, T X1, T X2, T data) {
  return (vc.ebk_f + eval_f0<T>(vc, X0, X1, X2)) - data + vc.pen;
}

// residual and float32-rounded Jacobian row            (calc_jac, Fitting_v4.py:321-367)
template <typename T>
IA3_HD void eval_jac(const VoxConsts<T>& vc, T X0, T X1, T X2, T data, T& res, float* J) {
  const T xt = X0 - vc.c[0], yt = X1 - vc.c[1], zt = X2 - vc.c[2];
  const T xs = vc.q[0] * xt * xt + vc.q[1] * yt * yt + vc.q[2] * zt * zt + vc.q[3] * xt * yt + vc.q[4] * xt * zt +
               vc.q[5] * yt * zt;
  const T f2 = exp_t<T>(vc.h - (T)0.5 * xs);
  res = (vc.ebk_f + f2) - data + vc.pen;
  const T m0 = xt * xt, m1 = yt * yt, m2 = zt * zt, m3 = xt * yt, m4 = xt * zt, m5 = yt * zt;
  J[0] = (float)vc.ebk_j;
  J[1] = (float)f2;
  J[2] = (float)((f2 * ((T)2 * vc.q[0] * xt + vc.q[3] * yt + vc.q[4] * zt)) * vc.ncen[0]);
  J[3] = (float)((f2 * (xt * vc.q[3] + (T)2 * vc.q[1] * yt + vc.q[5] * zt)) * vc.ncen[1]);
  J[4] = (float)((f2 * (xt * vc.q[4] + yt * vc.q[5] + (T)2 * vc.q[2] * zt)) * vc.ncen[2]);
  J[5] = (float)(f2 * (vc.a6[0] * m0 + vc.a6[1] * m1 + vc.a6[2] * m2 + vc.a6[3] * m3 + vc.a6[4] * m4 + vc.a6[5] * m5) + vc.jpen[0]);
  J[6] = (float)(f2 * (vc.a7[0] * m0 + vc.a7[1] * m1 + vc.a7[3] * m3) + vc.jpen[1]);
  J[7] = (float)(f2 * (vc.a8[0] * m0 + vc.a8[1] * m1 + vc.a8[2] * m2 + vc.a8[3] * m3 + vc.a8[4] * m4 + vc.a8[5] * m5) + vc.jpen[2]);
  J[8] = (float)(f2 * (vc.a9[0] * m0 + vc.a9[1] * m1 + vc.a9[2] * m2 + vc.a9[3] * m3 + vc.a9[4] * m4 + vc.a9[5] * m5));
  J[9] = (float)(f2 * (vc.a10[0] * m0 + vc.a10[1] * m1 + vc.a10[3] * m3 + vc.a10[4] * m4 + vc.a10[5] * m5));
}

// natural parameters [hf, c0, c1, c2, bkf, w1f, w2f, w3f, t, p] (to_natural_paramaters, :244-258)
IA3_HDN void natural_params(const FitParams& fp, const double* cen_est, const double* x, double* out10) {
  ModelConsts mc;
  model_consts(fp, cen_est, x, false, mc);
  const bool v4 = (fp.personality == 4);
  const double minw = fp.min_w2, dws = fp.max_w2 - fp.min_w2;
  double ws[3], t, p;
  if (v4) {
    for (int i = 0; i < 3; ++i) ws[i] = v4_sigmoid_guarded(x[5 + i], exp_s(x[5 + i]), minw, dws + minw, dws, minw);
    t = v4_sigmoid_guarded(x[9], exp_s(x[9]), -1.0, 1.0, 2.0, -1.0);
    p = v4_sigmoid_guarded(x[8], exp_s(x[8]), -1.0, 1.0, 2.0, -1.0);
  } else {
    for (int i = 0; i < 3; ++i) ws[i] = dws / (1.0 + exp_s(x[5 + i])) + minw;
    t = 2.0 / (1.0 + exp_s(x[9])) - 1.0;
    p = 2.0 / (1.0 + exp_s(x[8])) - 1.0;
  }
  out10[0] = exp_s(x[1]);
  out10[1] = mc.c[0]; out10[2] = mc.c[1]; out10[3] = mc.c[2];
  out10[4] = exp_s(x[0]);
  out10[5] = sqrt(ws[0]); out10[6] = sqrt(ws[1]); out10[7] = sqrt(ws[2]);
  out10[8] = t; out10[9] = p;
}

}  // namespace ia3
