// Host-callable launchers of the fit stage kernels (fit_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "gauss_model.h"
#include "lm_core.h"

namespace ia3 {

struct LMPause;                       // fit_spot.h

struct FitDev {
  // image / volumes
  const void* im; int im_dtype;       // original stack (u16 / f32 / f64)
  // float64 work volume (im_subtr -> im_add), stored sparsely: only 8x8x8 bricks touched by some
  // seed's window exist.  brick_tab[(bz * nbx + bx) * nby + by] = brick number (or -1); a brick is
  // 512 doubles at vol + 512 * number.  A dense copy would be 1.7 GB per 50x2048x2048 stack.
  double* vol;
  const int* brick_tab;
  int nbx, nby;
  int Z, X, Y;
  // seeds
  long long n;
  const double* centers;              // n x 3
  const int* own_id;                  // v3: lowest index among seeds with identical coordinates
  const int* nbr_start;               // n + 1
  const int* nbr_idx;
  // window
  int K, KW;                          // voxels in the ball; mask words per seed
  const int8_t* offs;                 // K x 3 offsets
  uint32_t* mask;                     // n x KW membership bits
  // ties (v4)
  int* tie_count; int tie_cap;
  int* tie_spot; int* tie_k;
  // results
  float* ps; double* p_raw; uint8_t* success; int* nfev; int* info;
  double* rec;                        // n x K reconstructions (ims_rec)
  // config
  FitParams fp; LMConfig lm; double init_w[3];
  // suspension of long runs (see k_fit): cap = function evaluations per launch (0 = run to the end);
  // pause_ctl[0] = number of suspended spots of the last launch, pause_ctl[1 + i] = spot of slot i;
  // while a spot is suspended info[spot] = -(slot + 1)
  int cap; int pause_slots;
  LMPause* pause_buf; int* pause_ctl;
};

// One entry of the continuation service's table (device-addressable pinned memory): which handle
// (index into the FitDev table), which spot, the fit mode; status is written by the kernel.
struct FitResume { int job; int spot; int mode; int status; };
enum { FIT_DONE = 0, FIT_SUSPENDED = 1 };

int fit_smem_bytes(int K, bool fp32);
int launch_init_window(const FitDev& d, cudaStream_t st);
int launch_voronoi(const FitDev& d, cudaStream_t st);
// mode 0 = firstfit (image data, Voronoi mask), 1 = repeatfit (vol + own rec, full window, write back)
int launch_fit(const FitDev& d, int mode, const int* work, long long n_work, bool fp32, cudaStream_t st);
int launch_subtract(const FitDev& d, const int* work, long long n_work, cudaStream_t st);
// continues suspended spots of several handles in one launch (FP64 model only)
int launch_fit_resume(const FitDev* devs, FitResume* entries, int n, int cap, int smem_bytes, cudaStream_t st);

struct MomentDev {                    // fast_fit_big_image / gfit_fast (Fitting_v4.py:433-556)
  const void* im; int im_dtype;
  int Z, X, Y;
  long long n;
  const double* centers;              // n x 3
  const int* nbr_start;               // seeds within 2r (inclusive, ascending, self included)
  const int* nbr_idx;
  int K; const int8_t* offs;
  int avoid, recenter;
  double bk_f;
  double* out;                        // n x 12
};
int launch_moment_fit(const MomentDev& d, cudaStream_t st);

struct GenericFitDev {
  long long n; const long long* off;
  const double* values; const float* coords; const double* centers;
  double* tmp;                        // scratch, same size as values
  float* ps; double* p_raw; uint8_t* success; int* nfev; int* info; double* rec;
  FitParams fp; LMConfig lm; double init_w[3];
};
int launch_generic_fit(const GenericFitDev& d, cudaStream_t st);

}  // namespace ia3

namespace ia3 {
int launch_apply_ties(uint32_t* mask, int KW, const int* tie_spot, const int* tie_k, const uint8_t* keep, int n, cudaStream_t st);
int launch_window_copy(const FitDev& d, double* snap, double* vol_out, int dir, cudaStream_t st);
int launch_to_f64(const void* im, int dtype, double* out, long long n, cudaStream_t st);
int launch_eval_f0(const FitParams& fp, const double* p_raw, const double* center, const float* coords, long long m,
                   double* out, cudaStream_t st);
}  // namespace ia3
