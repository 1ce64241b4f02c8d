// Host-callable launchers of the alternative Fitting_v4 seeders (aux_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ia3 {
template <typename Tin> int launch_box_norm(const Tin* im, float* out, int Z, int X, int Y, int sz, cudaStream_t st);
template <typename Tin> int launch_fir_axis(const Tin* in, double* out, int Z, int X, int Y, int axis, const double* d_w, int nt, cudaStream_t st);
template <typename Tin> int launch_log_ratio(const Tin* im, const double* blur, double* out, long long n, cudaStream_t st);
template <typename T> int launch_moments(const T* v, long long n, double mean, double* d_acc, cudaStream_t st);
int launch_v2_candidates(const float* v, int Z, int X, int Y, float cutoff, int pix, long long* out_idx, float* out_h, int* count, int cap, cudaStream_t st);
int launch_lr_candidates(const double* v, int Z, int X, int Y, double cutoff, int filt, long long* out_idx, double* out_h, int* count, int cap, cudaStream_t st);
int launch_hist_u16(const uint16_t* im, long long n, unsigned long long* d_out, cudaStream_t st);
}  // namespace ia3
