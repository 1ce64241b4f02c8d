// Kernels of the alternative seeders of External/Fitting_v4.py: get_seed_points_base_v2
// (:95-126, cv2.blur normalisation + 27-neighbour test with wrap-around) and get_seed_points_base (:72-92,
// log-ratio against an FFT Gaussian).  Whole-volume, HBM-bound passes.
//
//   k_box_norm        im_norm = float32(im) - cv2.blur(float32(im), (sz, sz)) per z-slice.  cv2.blur on CV_32F
//                     accumulates the window in double (exact for 24-bit inputs), multiplies by the double
//                     1 / sz^2 and rounds to float32; border = BORDER_REFLECT_101, anchor = sz / 2.  Measured
//                     against cv2 4.13: bit-identical (fixtures from the unmodified reference, tests/test_gpu_extra.py).
//   k_fir_axis        one axis of fft_gaussian_fast (:66-70): the reflect() padding of the reference followed by a
//                     'valid' convolution with the 40-tap window, evaluated directly in FP64 (the reference goes
//                     through a single-precision FFT: parity is by tolerance).
//   k_moments         sum / sum of squared deviations of a volume in FP64 (np.std)
//   k_v2_candidates   voxels with v > cutoff, v > 0 and v >= all 27 neighbours taken modulo the shape
//   k_lr_candidates   voxels that equal the maximum of their in-bounds 3x3x3 neighbourhood (scipy maximum_filter,
//                     reflect) and exceed the cutoff
// Candidates are appended with an atomic counter (order restored on the host by flat index).
#include <algorithm>
#include "ia3_device.h"
#include "aux_kernels.h"

namespace ia3 {

__device__ __forceinline__ int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = (i < 0) ? -i : 2 * (n - 1) - i;
  return i;
}

template <typename Tin>
__device__ __forceinline__ float as_f32(Tin v) { return (float)v; }

// one thread per output pixel; the window is read through the read-only cache (sz^2 loads: 25 for the default 5)
template <typename Tin>
__global__ void __launch_bounds__(256) k_box_norm(const Tin* __restrict__ im, float* __restrict__ out, int Z, int X, int Y, int sz) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n = (long long)Z * X * Y;
  if (t >= n) return;
  const int y = (int)(t % Y);
  const int x = (int)((t / Y) % X);
  const long long zoff = (t / ((long long)X * Y)) * X * Y;
  const int an = sz / 2;
  double s = 0.0;
  for (int i = 0; i < sz; ++i) {
    const int xx = reflect101(x - an + i, X);
    const Tin* row = im + zoff + (long long)xx * Y;
    double rs = 0.0;
    for (int j = 0; j < sz; ++j) rs += (double)as_f32(__ldg(row + reflect101(y - an + j, Y)));
    s += rs;
  }
  const float blur = __double2float_rn(s * (1.0 / ((double)sz * (double)sz)));
  out[t] = __fsub_rn(as_f32(im[t]), blur);
}

// reference reflect() + 'valid' convolution along one axis: out[i] = sum_k w[k] * ext[i + k], ext = (a[h-1..0], a, a[n-1..n-h])[:-1]
template <typename Tin>
__global__ void __launch_bounds__(256) k_fir_axis(const Tin* __restrict__ in, double* __restrict__ out, long long n_lines, int L, long long inner,
                                                  const double* __restrict__ w, int nt) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_lines * L) return;
  // consecutive threads walk the contiguous direction of the volume
  const long long line = (inner == 1) ? t / L : (t / (inner * L)) * inner + t % inner;
  const int i = (inner == 1) ? (int)(t % L) : (int)((t / inner) % L);
  const long long base = (inner == 1) ? line * L : (line / inner) * inner * L + line % inner;
  const int h = nt / 2;
  double acc = 0.0;
  for (int k = 0; k < nt; ++k) {
    int j = i + k - h;                       // index into the unpadded line
    if (j < 0) j = -j - 1;                   // a[h-1 .. 0] in front
    else if (j >= L) j = 2 * L - 1 - j;      // a[n-1 .. ] behind
    j = min(max(j, 0), L - 1);
    acc += w[nt - 1 - k] * (double)in[base + (long long)j * inner];
  }
  out[base + (long long)i * inner] = acc;
}

template <typename Tin>
__global__ void __launch_bounds__(256) k_log_ratio(const Tin* __restrict__ im, const double* __restrict__ blur, double* __restrict__ out, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = log((double)im[i]) - log(blur[i]);
}

template <typename T>
__global__ void __launch_bounds__(256) k_moments(const T* __restrict__ v, long long n, double mean, double* __restrict__ acc /* [0] += sum (v - mean), [1] += sum (v - mean)^2 */) {
  double s = 0.0, q = 0.0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) { const double d = (double)v[i] - mean; s += d; q += d * d; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
  __shared__ double ss[8], qq[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { ss[warp] = s; qq[warp] = q; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 8; ++k) { s += ss[k]; q += qq[k]; }
    atomicAdd(acc, s);
    atomicAdd(acc + 1, q);
  }
}

__global__ void __launch_bounds__(256) k_v2_candidates(const float* __restrict__ v, int Z, int X, int Y, float cutoff, int pix,
                                                       long long* __restrict__ out_idx, float* __restrict__ out_h, int* __restrict__ count, int cap) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n = (long long)Z * X * Y;
  if (t >= n) return;
  const float h = v[t];
  if (!(h > cutoff) || !(h > 0.f)) return;
  const int y = (int)(t % Y), x = (int)((t / Y) % X), z = (int)(t / ((long long)X * Y));
  for (int dz = -pix; dz <= pix; ++dz)
    for (int dx = -pix; dx <= pix; ++dx)
      for (int dy = -pix; dy <= pix; ++dy) {
        const int zz = ((z + dz) % Z + Z) % Z, xx = ((x + dx) % X + X) % X, yy = ((y + dy) % Y + Y) % Y;   // python's modulo
        if (!(h >= v[((long long)zz * X + xx) * Y + yy])) return;
      }
  const int pos = atomicAdd(count, 1);
  if (pos < cap) { out_idx[pos] = t; out_h[pos] = h; }
}

__global__ void __launch_bounds__(256) k_lr_candidates(const double* __restrict__ v, int Z, int X, int Y, double cutoff, int s1, int s2,
                                                       long long* __restrict__ out_idx, double* __restrict__ out_h, int* __restrict__ count, int cap) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n = (long long)Z * X * Y;
  if (t >= n) return;
  const double h = v[t];
  if (!(h > cutoff)) return;
  const int y = (int)(t % Y), x = (int)((t / Y) % X), z = (int)(t / ((long long)X * Y));
  for (int zz = max(z - s1, 0); zz <= min(z + s2, Z - 1); ++zz)
    for (int xx = max(x - s1, 0); xx <= min(x + s2, X - 1); ++xx)
      for (int yy = max(y - s1, 0); yy <= min(y + s2, Y - 1); ++yy)
        if (v[((long long)zz * X + xx) * Y + yy] > h) return;
  const int pos = atomicAdd(count, 1);
  if (pos < cap) { out_idx[pos] = t; out_h[pos] = h; }
}

// 65536-bin histogram of a uint16 volume: the low 8192 values (where the background sits) in shared memory,
// the rest with global atomics; out is 65536 unsigned long long, zeroed by the launcher
__global__ void __launch_bounds__(256) k_hist_u16(const uint16_t* __restrict__ im, long long n, unsigned long long* __restrict__ out) {
  constexpr int LOW = 8192;
  __shared__ unsigned sh[LOW];
  for (int i = threadIdx.x; i < LOW; i += blockDim.x) sh[i] = 0u;
  __syncthreads();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const unsigned v = im[i];
    if (v < LOW) atomicAdd(&sh[v], 1u);
    else atomicAdd(&out[v], 1ull);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < LOW; i += blockDim.x) if (sh[i]) atomicAdd(&out[i], (unsigned long long)sh[i]);
}
int launch_hist_u16(const uint16_t* im, long long n, unsigned long long* d_out, cudaStream_t st) {
  IA3_CUDA(cudaMemsetAsync(d_out, 0, 65536 * sizeof(unsigned long long), st));
  if (n == 0) return 0;
  k_hist_u16<<<148 * 4, 256, 0, st>>>(im, n, d_out);
  IA3_LAUNCH_CHECK();
  return 0;
}

template <typename Tin>
int launch_box_norm(const Tin* im, float* out, int Z, int X, int Y, int sz, cudaStream_t st) {
  const long long n = (long long)Z * X * Y;
  k_box_norm<Tin><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(im, out, Z, X, Y, sz);
  IA3_LAUNCH_CHECK();
  return 0;
}
template int launch_box_norm<uint16_t>(const uint16_t*, float*, int, int, int, int, cudaStream_t);
template int launch_box_norm<float>(const float*, float*, int, int, int, int, cudaStream_t);
template int launch_box_norm<double>(const double*, float*, int, int, int, int, cudaStream_t);

template <typename Tin>
int launch_fir_axis(const Tin* in, double* out, int Z, int X, int Y, int axis, const double* d_w, int nt, cudaStream_t st) {
  const long long n = (long long)Z * X * Y;
  const int L = axis == 0 ? Z : (axis == 1 ? X : Y);
  const long long inner = axis == 0 ? (long long)X * Y : (axis == 1 ? Y : 1);
  k_fir_axis<Tin><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in, out, n / L, L, inner, d_w, nt);
  IA3_LAUNCH_CHECK();
  return 0;
}
template int launch_fir_axis<uint16_t>(const uint16_t*, double*, int, int, int, int, const double*, int, cudaStream_t);
template int launch_fir_axis<float>(const float*, double*, int, int, int, int, const double*, int, cudaStream_t);
template int launch_fir_axis<double>(const double*, double*, int, int, int, int, const double*, int, cudaStream_t);

template <typename Tin>
int launch_log_ratio(const Tin* im, const double* blur, double* out, long long n, cudaStream_t st) {
  k_log_ratio<Tin><<<148 * 8, 256, 0, st>>>(im, blur, out, n);
  IA3_LAUNCH_CHECK();
  return 0;
}
template int launch_log_ratio<uint16_t>(const uint16_t*, const double*, double*, long long, cudaStream_t);
template int launch_log_ratio<float>(const float*, const double*, double*, long long, cudaStream_t);
template int launch_log_ratio<double>(const double*, const double*, double*, long long, cudaStream_t);

template <typename T>
int launch_moments(const T* v, long long n, double mean, double* d_acc, cudaStream_t st) {
  IA3_CUDA(cudaMemsetAsync(d_acc, 0, 2 * sizeof(double), st));
  k_moments<T><<<148 * 8, 256, 0, st>>>(v, n, mean, d_acc);
  IA3_LAUNCH_CHECK();
  return 0;
}
template int launch_moments<float>(const float*, long long, double, double*, cudaStream_t);
template int launch_moments<double>(const double*, long long, double, double*, cudaStream_t);

int launch_v2_candidates(const float* v, int Z, int X, int Y, float cutoff, int pix, long long* out_idx, float* out_h, int* count, int cap, cudaStream_t st) {
  const long long n = (long long)Z * X * Y;
  IA3_CUDA(cudaMemsetAsync(count, 0, sizeof(int), st));
  k_v2_candidates<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(v, Z, X, Y, cutoff, pix, out_idx, out_h, count, cap);
  IA3_LAUNCH_CHECK();
  return 0;
}
int launch_lr_candidates(const double* v, int Z, int X, int Y, double cutoff, int filt, long long* out_idx, double* out_h, int* count, int cap, cudaStream_t st) {
  const long long n = (long long)Z * X * Y;
  const int s1 = filt / 2, s2 = filt - s1 - 1;
  IA3_CUDA(cudaMemsetAsync(count, 0, sizeof(int), st));
  k_lr_candidates<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(v, Z, X, Y, cutoff, s1, s2, out_idx, out_h, count, cap);
  IA3_LAUNCH_CHECK();
  return 0;
}

}  // namespace ia3
