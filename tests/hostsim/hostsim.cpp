// TEST INFRASTRUCTURE: compiles the numerical core of the fit kernel (csrc/fit_spot.h,
// lm_core.h, gauss_model.h) with g++ and runs it with a one-lane executor, so the
// lmder-faithful driver and the Gaussian model can be checked against scipy on a machine
// without a GPU.  Not linked into the product library and not a fallback for it.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../imageanalysis3_b200/csrc/fit_spot.h"

using namespace ia3;

struct SerialExec {
  static constexpr int W = 1;
  int lane() const { return 0; }
  void sync() const {}
  double allsum(double v) const { return v; }
  int allsum_int(int v) const { return v; }
  double allmax(double v) const { return v; }
  double bcast(double v, int) const { return v; }
  void argmin(double&, int&) const {}
  struct TakenMask {
    std::vector<bool> bits;
    void clear() { bits.assign(1 << 15, false); }
    bool test(int i) const { return bits[i]; }
    void set(int i) { bits[i] = true; }
  };
  template <int N>
  void reduce_store(double (&v)[N], double* out) const { for (int i = 0; i < N; ++i) out[i] = v[i]; }
};

template <typename T>
struct ArrayVox {
  int m;
  const int* rel;      // m x 3 relative coordinates
  const float* data;   // m float32 data
  void get(int k, T& X0, T& X1, T& X2, T& d) const {
    X0 = (T)rel[3 * k]; X1 = (T)rel[3 * k + 1]; X2 = (T)rel[3 * k + 2]; d = (T)data[k];
  }
};

template <typename T>
static int run(int personality, const double* values, const int* coords, int m, const double* cen,
               double delta, double min_w, double max_w, const double* init_w, double weight_sigma,
               int maxfev, double* p_raw, float* ps, int* stats, int cap = 0) {
  if (m < NP) return 1;
  FitParams fp;
  fp.min_w2 = min_w * min_w; fp.max_w2 = max_w * max_w; fp.delta = delta;
  fp.weight_sigma = weight_sigma; fp.personality = personality;
  LMConfig cfg{1.49012e-8, 1.49012e-8, 0.0, 100.0, maxfev};
  double origin[3];
  std::vector<int> rel(3 * m);
  std::vector<float> data(m);
  for (int i = 0; i < 3; ++i) origin[i] = (double)(long long)cen[i];
  for (int k = 0; k < m; ++k) {
    for (int i = 0; i < 3; ++i) rel[3 * k + i] = coords[3 * k + i] - (int)origin[i];
    data[k] = (float)values[k];
  }
  SerialExec ex;
  static SpotShared<T> sh;
  select10(ex, values, m, false, sh.small10);
  select10(ex, values, m, true, sh.large10);
  initial_guess(fp, sh.small10, sh.large10, init_w, sh.x0);
  ArrayVox<T> vox{m, rel.data(), data.data()};
  const FitParams fp_in = fp;          // initial_guess fills the v3 width prior in fp
  bool suspended = run_lm<T>(ex, fp, cfg, cen, origin, vox, sh, cap, LM_START_FRESH);
  int rounds = 0;
  while (suspended) {
    // what the device does between two launches (fit_kernels.cu: fit_one): park the live part of LMState + the normal
    // equation sums (LMLive), forget everything else, rebuild the prologue, continue
    static LMLive park;
    lm_live_save(sh.st, sh.Ag, park);
    memset(static_cast<void*>(&sh), 0xA5, sizeof(sh));
    fp = fp_in;
    select10(ex, values, m, false, sh.small10);
    select10(ex, values, m, true, sh.large10);
    initial_guess(fp, sh.small10, sh.large10, init_w, sh.x0);
    lm_live_restore(park, sh.st, sh.Ag);
    suspended = run_lm<T>(ex, fp, cfg, cen, origin, vox, sh, cap, LM_START_CONTINUE);
    ++rounds;
  }
  if (cap > 0) stats[3] = rounds;
  FitResult res;
  finish_fit<T>(ex, fp, cen, origin, vox, sh, &res);
  memcpy(p_raw, res.p_raw, sizeof(double) * NP);
  memcpy(ps, res.ps, sizeof(float) * NOUT);
  stats[0] = res.nfev; stats[1] = res.njev; stats[2] = res.info;
  return 0;
}

extern "C" int hostsim_fit(int personality, int use_float, const double* values, const int* coords, int m,
                           const double* cen, double delta, double min_w, double max_w, const double* init_w,
                           double weight_sigma, int maxfev, double* p_raw, float* ps, int* stats) {
  if (use_float)
    return run<float>(personality, values, coords, m, cen, delta, min_w, max_w, init_w, weight_sigma, maxfev, p_raw, ps, stats);
  return run<double>(personality, values, coords, m, cen, delta, min_w, max_w, init_w, weight_sigma, maxfev, p_raw, ps, stats);
}

// the same fit, suspended every `cap` function evaluations and resumed from the parked state only
// (stats[3] = number of suspensions): must be bit-identical to hostsim_fit
extern "C" int hostsim_fit_capped(int personality, const double* values, const int* coords, int m, const double* cen,
                                  double delta, double min_w, double max_w, const double* init_w, double weight_sigma,
                                  int maxfev, int cap, double* p_raw, float* ps, int* stats) {
  return run<double>(personality, values, coords, m, cen, delta, min_w, max_w, init_w, weight_sigma, maxfev, p_raw, ps, stats, cap);
}

// reconstruction f0 over arbitrary integer voxels (GaussianFit.get_im, Fitting_v4.py:394-396)
extern "C" void hostsim_get_im(int personality, const double* p_raw, const double* cen, double delta, double min_w,
                               double max_w, const int* coords, int m, double* out) {
  FitParams fp;
  fp.min_w2 = min_w * min_w; fp.max_w2 = max_w * max_w; fp.delta = delta; fp.weight_sigma = 0; fp.personality = personality;
  ModelConsts mc;
  model_consts(fp, cen, p_raw, false, mc);
  VoxConsts<double> vc;
  double origin[3] = {0, 0, 0};
  narrow_consts<double>(mc, origin, false, vc);
  for (int k = 0; k < m; ++k) out[k] = eval_f0<double>(vc, coords[3 * k], coords[3 * k + 1], coords[3 * k + 2]);
}

// ---- experiment: lmder with MINPACK's own Householder qrfac on the full m x 10 Jacobian ------
// (used to separate "normal equations lose digits" from "driver logic differs")
namespace {
void qrfac_full(std::vector<double>& a /*col-major m x n*/, int m, LMState& st, const double* fvec) {
  const int n = NP;
  const double epsmch = DBL_EPSILON;
  double rdiag[NP], wa[NP];
  auto col = [&](int j) { return a.data() + (size_t)j * m; };
  auto enorm = [&](const double* v, int len) { double s = 0; for (int i = 0; i < len; ++i) s += v[i] * v[i]; return sqrt(s); };
  for (int j = 0; j < n; ++j) { st.acn[j] = enorm(col(j), m); rdiag[j] = st.acn[j]; wa[j] = rdiag[j]; st.ipvt[j] = j; }
  for (int j = 0; j < n; ++j) {
    int kmax = j;
    for (int k = j; k < n; ++k) if (rdiag[k] > rdiag[kmax]) kmax = k;
    if (kmax != j) {
      for (int i = 0; i < m; ++i) std::swap(col(j)[i], col(kmax)[i]);
      rdiag[kmax] = rdiag[j]; wa[kmax] = wa[j];
      std::swap(st.ipvt[j], st.ipvt[kmax]);
    }
    double ajnorm = enorm(col(j) + j, m - j);
    if (ajnorm != 0.0) {
      if (col(j)[j] < 0.0) ajnorm = -ajnorm;
      for (int i = j; i < m; ++i) col(j)[i] /= ajnorm;
      col(j)[j] += 1.0;
      for (int k = j + 1; k < n; ++k) {
        double sum = 0; for (int i = j; i < m; ++i) sum += col(j)[i] * col(k)[i];
        double temp = sum / col(j)[j];
        for (int i = j; i < m; ++i) col(k)[i] -= temp * col(j)[i];
        if (rdiag[k] != 0.0) {
          double t = col(k)[j] / rdiag[k];
          rdiag[k] *= sqrt(fmax(0.0, 1.0 - t * t));
          double q = rdiag[k] / wa[k];
          if (0.05 * (q * q) <= epsmch) { rdiag[k] = enorm(col(k) + j + 1, m - j - 1); wa[k] = rdiag[k]; }
        }
      }
    }
    rdiag[j] = -ajnorm;
  }
  // qtf
  std::vector<double> w4(fvec, fvec + m);
  for (int j = 0; j < n; ++j) {
    if (col(j)[j] != 0.0) {
      double sum = 0; for (int i = j; i < m; ++i) sum += col(j)[i] * w4[i];
      double temp = -sum / col(j)[j];
      for (int i = j; i < m; ++i) w4[i] += col(j)[i] * temp;
    }
    col(j)[j] = rdiag[j];
    st.qtf[j] = w4[j];
  }
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) st.R[i][j] = (j >= i) ? col(j)[i] : 0.0;
}
}  // namespace

extern "C" int hostsim_fit_qr(int personality, const double* values, const int* coords, int m, const double* cen,
                              double delta, double min_w, double max_w, const double* init_w, double weight_sigma,
                              int maxfev, double* p_raw, float* ps, int* stats) {
  typedef double T;
  if (m < NP) return 1;
  FitParams fp;
  fp.min_w2 = min_w * min_w; fp.max_w2 = max_w * max_w; fp.delta = delta;
  fp.weight_sigma = weight_sigma; fp.personality = personality;
  LMConfig cfg{1.49012e-8, 1.49012e-8, 0.0, 100.0, maxfev};
  double origin[3];
  std::vector<int> rel(3 * m);
  std::vector<float> data(m);
  for (int i = 0; i < 3; ++i) origin[i] = (double)(long long)cen[i];
  for (int k = 0; k < m; ++k) { for (int i = 0; i < 3; ++i) rel[3 * k + i] = coords[3 * k + i] - (int)origin[i]; data[k] = (float)values[k]; }
  SerialExec ex;
  static SpotShared<T> sh;
  std::vector<double> tmp(m);
  select10_scratch(ex, values, tmp.data(), m, false, sh.small10);
  select10_scratch(ex, values, tmp.data(), m, true, sh.large10);
  initial_guess(fp, sh.small10, sh.large10, init_w, sh.x0);
  ArrayVox<T> vox{m, rel.data(), data.data()};
  LMState& st = sh.st;
  std::vector<double> fvec(m), ftrial(m), J((size_t)m * NP);
  auto evalf = [&](const double* x, std::vector<double>& out) {
    build_consts<T>(fp, cen, origin, x, false, sh.vc);
    for (int k = 0; k < m; ++k) { T a, b, c, d; vox.get(k, a, b, c, d); out[k] = eval_res<T>(sh.vc, a, b, c, d); }
    return pass_residual<T>(ex, sh.vc, vox, (double*)0);
  };
  lm_init(ex, st, sh.x0, evalf(sh.x0, fvec));
  for (;;) {
    build_consts<T>(fp, cen, origin, st.x, true, sh.vc);
    for (int k = 0; k < m; ++k) { T a, b, c, d, r; float Jr[NP]; vox.get(k, a, b, c, d); eval_jac<T>(sh.vc, a, b, c, d, r, Jr); for (int j = 0; j < NP; ++j) J[(size_t)j * m + k] = Jr[j]; }
    // same as lm_outer but with the Householder factorisation
    st.njev += 1;
    qrfac_full(J, m, st, fvec.data());
    lm_post_factor(ex, st);
    if (st.iter == 1) {
      for (int j = 0; j < NP; ++j) { st.diag[j] = st.acn[j]; if (st.acn[j] == 0.0) st.diag[j] = 1.0; }
      for (int j = 0; j < NP; ++j) st.w3[j] = st.diag[j] * st.x[j];
      st.xnorm = enorm_n(st.w3, NP); st.delta = cfg.factor * st.xnorm; if (st.delta == 0.0) st.delta = cfg.factor;
    }
    double gnorm = 0.0;
    if (st.fnorm != 0.0) for (int j = 0; j < NP; ++j) { int l = st.ipvt[j]; if (st.acn[l] != 0.0) { double sum = 0; for (int i = 0; i <= j; ++i) sum += st.R[i][j] * (st.qtf[i] / st.fnorm); gnorm = fmax(gnorm, fabs(sum / st.acn[l])); } }
    st.gnorm = gnorm;
    if (gnorm <= cfg.gtol) { st.info = 4; break; }
    for (int j = 0; j < NP; ++j) st.diag[j] = fmax(st.diag[j], st.acn[j]);
    int action;
    for (;;) {
      lm_propose(ex, st);
      double f1 = evalf(st.xt, ftrial);
      if (getenv("HOSTSIM_DEBUG")) { printf("f par=%.6e delta=%.6e pnorm=%.6e fnorm1=%.17g x:", st.par, st.delta, st.pnorm, f1); for (int j=0;j<NP;++j) printf(" %.5e", st.xt[j]); printf("\n"); }
      action = lm_judge(ex, st, cfg, f1);
      if (action != LM_RETRY) { if (action == LM_ACCEPTED || st.fnorm == f1) fvec = ftrial; }
      if (action != LM_RETRY) break;
    }
    if (action == LM_DONE) break;
  }
  FitResult res;
  finish_fit<T>(ex, fp, cen, origin, vox, sh, &res);
  memcpy(p_raw, res.p_raw, sizeof(double) * NP); memcpy(ps, res.ps, sizeof(float) * NOUT);
  stats[0] = res.nfev; stats[1] = res.njev; stats[2] = res.info;
  return 0;
}
