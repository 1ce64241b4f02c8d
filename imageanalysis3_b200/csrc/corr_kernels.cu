// Pre-processing kernels: the compute core of io_tools/load.py correct_fov_image, the step
// that hands the seed / fit stages their stacks.  All whole-volume, HBM-bound passes on uint16 channel stacks.
//
//   k_hot_count      per (x, y) column: in how many planes is the voxel brighter than hot_th x the mean of its four
//                    neighbours (corrections.py:496-499: np.roll, so borders wrap, and the y+1... neighbour is the
//                    y-1 one taken twice, as the reference does)
//   k_hot_fix        replaces the hot columns one after the other, in np.where order, by the float32 mean of their four
//                    neighbours (:505-509); an already replaced neighbour contributes its float32 value, the final
//                    store truncates to uint16 like .astype (one CTA walks the list; a column's position in it is looked
//                    up in an X x Y table, so even thousands of hot columns cost microseconds each)
//   k_zshift         corrections.py:479-487 Z_Shift_Correction: planes scaled to the stack's median; the medians are read off
//                    per-plane device histograms on the host (exact order statistics)
//   k_gauss_nearest  correction_tools/filter.py:14-19 gaussian_high_pass_filter: exact uint16 Gaussian with mode='nearest', then
//   / k_highpass     im - lowpass clipped at 0
//   k_mix            bleed-through mixing sum_j im_j * profile[i, j] in float32, clip, truncate (io_tools/load.py:
//                    347-367) fused with the illumination division (:369-381)
//   k_spline_iir /   scipy.ndimage.spline_filter(np.pad(im, 12, 'edge'), 3, mode='nearest'): the cubic B-spline prefilter
//   k_spline_iir_y   with pole z = sqrt 3 - 2 on the half-sample-symmetric extension of each padded line -- the two-pass
//                    recursion scipy runs, one thread per line, in place on one float64 padded volume; the causal start
//                    is the 28-term sum (|z|^28 < 1e-16).  Along z and x the lines lie side by side in y (every step is a
//                    coalesced row); along y the rows go through transposing shared-memory tiles.  Agrees with scipy's
//                    coefficients to ~1e-15 relative (checked in numpy before it was written)
//   k_warp           map_coordinates(im, coords, order=3, mode='nearest') (:424-459) with coords = grid + chromatic
//                    profile - drift formed exactly as numpy promotes them, 64-tap cubic interpolation on the
//                    prefiltered volume, uint16 output rounded half up and saturated like ndimage's integer store
#include <algorithm>
#include <cmath>
#include "ia3_device.h"
#include "corr_kernels.h"

namespace ia3 {

__global__ void __launch_bounds__(256) k_hot_count(const uint16_t* __restrict__ im, int Z, int X, int Y, float hot_th, int* __restrict__ cnt) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)X * Y) return;
  const int y = (int)(t % Y), x = (int)(t / Y);
  const int xm = (x + X - 1) % X, xp = (x + 1) % X, ym = (y + Y - 1) % Y;
  int c = 0;
  for (int z = 0; z < Z; ++z) {
    const uint16_t* pl = im + (long long)z * X * Y;
    // np.roll(im,1,1)[x] = im[x-1]; roll(im,-1,1)[x] = im[x+1]; roll(im,1,2)[y] = im[y-1], taken twice (:496)
    const float a = (float)pl[(long long)xm * Y + y], b = (float)pl[(long long)xp * Y + y], d = (float)pl[(long long)x * Y + ym];
    const float conv = __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(a, b), d), d), 4.0f);
    if ((float)pl[(long long)x * Y + y] > __fmul_rn(hot_th, conv)) ++c;
  }
  cnt[t] = c;
}

__global__ void __launch_bounds__(256) k_hot_select(const int* __restrict__ cnt, long long n, double thr, int* __restrict__ out, int* __restrict__ count, int cap) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  if ((double)cnt[t] > thr) { const int p = atomicAdd(count, 1); if (p < cap) out[p] = (int)t; }
}

// slot[x * Y + y] = position of the column in the sorted list of hot columns, -1 elsewhere
__global__ void __launch_bounds__(256) k_hot_slots(const int* __restrict__ list, int n, int* __restrict__ slot) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) slot[list[i]] = i;
}

// list: flat x * Y + y indices in ascending (np.where) order; vals: n x Z float32 scratch; slot: see k_hot_slots
__global__ void __launch_bounds__(128) k_hot_fix(uint16_t* __restrict__ im, int Z, int X, int Y, const int* __restrict__ list, int n,
                                                 const int* __restrict__ slot, float* __restrict__ vals) {
  for (int i = 0; i < n; ++i) {
    const int x = list[i] / Y, y = list[i] % Y;
    const bool interior = x > 0 && y > 0 && x < X - 1 && y < Y - 1;
    for (int z = threadIdx.x; z < Z; z += blockDim.x) {
      float v;
      if (interior) {
        float nb[4];
        const int nx[4] = {x + 1, x - 1, x, x}, ny[4] = {y, y, y + 1, y - 1};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int j = slot[nx[k] * Y + ny[k]];
          // a neighbour that was replaced earlier in the sequence contributes its float32 value
          nb[k] = (j >= 0 && j < i) ? vals[(long long)j * Z + z] : (float)im[((long long)z * X + nx[k]) * Y + ny[k]];
        }
        v = __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(nb[0], nb[1]), nb[2]), nb[3]), 4.0f);
      } else {
        v = (float)im[((long long)z * X + x) * Y + y];
      }
      vals[(long long)i * Z + z] = v;
    }
    __syncthreads();
  }
  for (int i = 0; i < n; ++i) {
    const int x = list[i] / Y, y = list[i] % Y;
    for (int z = threadIdx.x; z < Z; z += blockDim.x) im[((long long)z * X + x) * Y + y] = (uint16_t)(int)vals[(long long)i * Z + z];
  }
}

__device__ __forceinline__ float rn_mul(float a, float b) { return __fmul_rn(a, b); }      // no contraction into fma: numpy rounds each step
__device__ __forceinline__ double rn_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float rn_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double rn_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float rn_div(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double rn_div(double a, double b) { return __ddiv_rn(a, b); }

// Tp = the profiles' dtype: uint16 * float32 stays float32 in numpy, uint16 * float64 is float64
template <typename Tp>
__global__ void __launch_bounds__(256) k_mix(const uint16_t* const* __restrict__ ins, int n_in, const Tp* __restrict__ bleed /* n_in x XY or null */,
                                             const Tp* __restrict__ illum /* XY or null */, uint16_t* __restrict__ out, long long XY, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) {
    const long long pix = t % XY;
    uint16_t u;
    if (bleed) {
      Tp acc = rn_mul((Tp)ins[0][t], bleed[pix]);
      for (int j = 1; j < n_in; ++j) acc = rn_add(acc, rn_mul((Tp)ins[j][t], bleed[(long long)j * XY + pix]));
      acc = acc > (Tp)65535 ? (Tp)65535 : acc;
      acc = acc < (Tp)0 ? (Tp)0 : acc;
      u = (uint16_t)(int)acc;
    } else {
      u = ins[0][t];
    }
    // the division runs on float32(im): float32 / float32 profile, or float64 when the profile is float64
    if (illum) u = (uint16_t)(int)rn_div((Tp)(float)u, illum[pix]);
    out[t] = u;
  }
}

// corrections.py:479-487 Z_Shift_Correction: every plane scaled to the whole stack's median, float32 arithmetic in numpy's
// order (im / median_z, then * median_all), truncated to uint16 like .astype
__global__ void __launch_bounds__(256) k_zshift(uint16_t* __restrict__ im, long long XY, long long n, const float* __restrict__ med_z, float med_all) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride)
    im[t] = (uint16_t)(int)__fmul_rn(__fdiv_rn((float)im[t], med_z[t / XY]), med_all);
}

// correction_tools/filter.py:14-19 gaussian_high_pass_filter on a uint16 image: scipy.ndimage.gaussian_filter(im, sigma,
// mode='nearest', truncate) -- three 1-D correlations, double accumulation in scipy's order (centre tap, then the tap pairs
// from the outermost inwards, products and sums rounded separately), each pass stored as uint16 by truncation -- then
// im - lowpass where that is positive, 0 elsewhere.  One thread per output, edge clamped ('nearest').
__global__ void __launch_bounds__(256) k_gauss_nearest(const uint16_t* __restrict__ in, uint16_t* __restrict__ out, int L, long long inner, long long n,
                                                       const double* __restrict__ w, int r) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const long long c = (t / inner) % L;
  const uint16_t* line = in + (t - c * inner);
  double acc = __dmul_rn((double)line[c * inner], w[0]);
  for (int j = r; j >= 1; --j) {
    const long long lo = c - j < 0 ? 0 : c - j, hi = c + j > L - 1 ? L - 1 : c + j;
    acc = __dadd_rn(acc, __dmul_rn(__dadd_rn((double)line[lo * inner], (double)line[hi * inner]), w[j]));
  }
  out[t] = (uint16_t)__double2int_rz(acc);
}

__global__ void __launch_bounds__(256) k_highpass(uint16_t* __restrict__ im, const uint16_t* __restrict__ low, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) {
    const unsigned a = im[t], b = low[t];
    im[t] = (uint16_t)(a > b ? a - b : 0u);
  }
}

constexpr int NPAD = 12;        // scipy.ndimage._interpolation._prepad_for_spline_filter
constexpr int KTAP = 28;

__device__ __forceinline__ int refl(int i, int n) {     // d c b a | a b c d | d c b a
  const int p = 2 * n;
  int m = i % p;
  if (m < 0) m += p;
  return m < n ? m : p - 1 - m;
}

// Axis 0 / axis 1 of the prefilter as the two-pass recursion itself (what scipy runs): one thread per line, lines
// side by side along y so every step of the recursion is a coalesced row of loads / stores.
//   c+[0] = 6 (x[0] + z sum_k z^k x[k])  (half-sample-symmetric start, 28 terms),  c+[i] = 6 x[i] + z c+[i-1]
//   c[n-1] = z / (z - 1) c+[n-1],  c[i] = z (c[i+1] - c+[i])
// Tsrc = uint16: reads the image through the 'edge' padding (first pass); double: the padded volume, in place.
template <typename Tsrc>
__global__ void __launch_bounds__(128) k_spline_iir(const Tsrc* src, double* dst /* may be src */, int Z, int X, int Y, int axis) {
  const int PZ = Z + 2 * NPAD, PX = X + 2 * NPAD, PY = Y + 2 * NPAD;
  const int L = axis == 0 ? PZ : PX, O = axis == 0 ? PX : PZ;          // line length, number of lines per y
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)O * PY) return;
  const int y = (int)(t % PY), o = (int)(t / PY);
  const long long stride = axis == 0 ? (long long)PX * PY : PY;
  const long long base = axis == 0 ? (long long)o * PY + y : (long long)o * PX * PY + y;
  const double zp = -0.26794919243112270647;                            // sqrt(3) - 2
  auto in = [&](int i) -> double {
    if constexpr (sizeof(Tsrc) == 2) {
      const int zz = axis == 0 ? i : o, xx = axis == 0 ? o : i;
      const int oz = min(max(zz - NPAD, 0), Z - 1), ox = min(max(xx - NPAD, 0), X - 1), oy = min(max(y - NPAD, 0), Y - 1);
      return (double)src[((long long)oz * X + ox) * Y + oy];
    } else {
      return src[base + (long long)i * stride];
    }
  };
  double s = 0.0, zk = 1.0;
  for (int k = 0; k < KTAP; ++k) { s += zk * in(refl(k, L)); zk *= zp; }
  double c = 6.0 * (in(0) + zp * s);
  dst[base] = c;
#pragma unroll 8
  for (int i = 1; i < L; ++i) {
    c = 6.0 * in(i) + zp * c;
    dst[base + (long long)i * stride] = c;
  }
  c = zp / (zp - 1.0) * c;
  dst[base + (long long)(L - 1) * stride] = c;
#pragma unroll 8
  for (int i = L - 2; i >= 0; --i) {
    c = zp * (c - dst[base + (long long)i * stride]);
    dst[base + (long long)i * stride] = c;
  }
}

// Axis 2 (the contiguous one): the same recursion, one thread per row, fed through shared memory so that global memory
// is still read and written in coalesced rows: a block owns 128 rows and walks them in chunks of 32 columns (forward for
// the causal pass, backward for the anticausal one); a chunk is loaded row-wise by the warps, each thread runs its row's 32
// recursion steps in the tile (odd pitch: no bank conflicts), and the chunk is written back row-wise.  In place.
constexpr int YR = 128, YC = 32;
__global__ void __launch_bounds__(YR) k_spline_iir_y(double* buf, long long n_rows, int L) {
  __shared__ double tile[YR][YC + 1];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long r0 = (long long)blockIdx.x * YR;
  const long long mine = r0 + tid;
  const bool ok = mine < n_rows;
  const double zp = -0.26794919243112270647;
  double c = 0.0;
  if (ok) {
    const double* r = buf + mine * L;
    double sum = 0.0, zk = 1.0;
    for (int k = 0; k < KTAP; ++k) { sum += zk * r[refl(k, L)]; zk *= zp; }
    c = 6.0 * sum;                                  // c+[0] = 6 x[0] + z (6 sum)
  }
  const int nch = (L + YC - 1) / YC;
  auto load = [&](int y0) {
#pragma unroll 4
    for (int rr = 0; rr < 32; ++rr) {
      const long long row = r0 + warp * 32 + rr;
      if (row < n_rows && y0 + lane < L) tile[warp * 32 + rr][lane] = buf[row * L + y0 + lane];
    }
  };
  auto store = [&](int y0) {
#pragma unroll 4
    for (int rr = 0; rr < 32; ++rr) {
      const long long row = r0 + warp * 32 + rr;
      if (row < n_rows && y0 + lane < L) buf[row * L + y0 + lane] = tile[warp * 32 + rr][lane];
    }
  };
  for (int ch = 0; ch < nch; ++ch) {
    const int y0 = ch * YC, nk = min(YC, L - y0);
    load(y0);
    __syncthreads();
    if (ok) {
#pragma unroll 8
      for (int k = 0; k < nk; ++k) { c = 6.0 * tile[tid][k] + zp * c; tile[tid][k] = c; }
    }
    __syncthreads();
    store(y0);
    __syncthreads();
  }
  c = zp / (zp - 1.0) * c;                           // c[L-1] from c+[L-1]
  for (int ch = nch - 1; ch >= 0; --ch) {
    const int y0 = ch * YC, nk = min(YC, L - y0);
    load(y0);
    __syncthreads();
    if (ok) {
#pragma unroll 8
      for (int k = nk - 1; k >= 0; --k) {
        if (y0 + k != L - 1) c = zp * (c - tile[tid][k]);
        tile[tid][k] = c;
      }
    }
    __syncthreads();
    store(y0);
    __syncthreads();
  }
}

__device__ __forceinline__ void cubic_w(double t, double* w) {       // ndimage's cubic B-spline weights at offset t in [0, 1)
  const double y = t, z = 1.0 - t;
  w[1] = (y * y * (y - 2.0) * 3.0 + 4.0) / 6.0;
  w[2] = (z * z * (z - 2.0) * 3.0 + 4.0) / 6.0;
  w[0] = z * z * z / 6.0;
  w[3] = 1.0 - w[0] - w[1] - w[2];
}

// A thread owns one (x, y) column and walks it along z: three of the four coefficient planes an output needs were touched
// by the previous output of the same block, so they come from L1 / L2 instead of HBM (the padded volume is 1.85 GB; four
// planes of it exceed the L2).  A block covers 8 x 32 columns.
template <typename Tc>
__global__ void __launch_bounds__(256) k_warp(const double* __restrict__ coef, int Z, int X, int Y, const Tc* __restrict__ chroma /* 3 x CZ x X x Y or null */,
                                              int CZ, double d0, double d1, double d2, uint16_t* __restrict__ out) {
  const int PZ = Z + 2 * NPAD, PX = X + 2 * NPAD, PY = Y + 2 * NPAD;
  const int y = blockIdx.x * 32 + (threadIdx.x & 31), x = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= X || y >= Y) return;
  const int dims[3] = {PZ, PX, PY};
  const double drift[3] = {d0, d1, d2};
  const long long plane = (long long)CZ * X * Y;
  for (int z = 0; z < Z; ++z) {
    // coords = int64 grid + profile (-> float64) - drift (float32 or float64, -> float64)     io_tools/load.py:441-449
    double c[3] = {(double)z, (double)x, (double)y};
    if (chroma) {
      const long long off = ((long long)(CZ == 1 ? 0 : z) * X + x) * Y + y;
      c[0] = c[0] + (double)chroma[off]; c[1] = c[1] + (double)chroma[plane + off]; c[2] = c[2] + (double)chroma[2 * plane + off];
    }
    int idx[3][4];
    double w[3][4];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      double cc = (c[a] - drift[a]) + (double)NPAD;
      cc = fmin(fmax(cc, 0.0), (double)(dims[a] - 1));          // mode='nearest' on the padded extent
      const double f = floor(cc);
      const int fl = (int)f;
      cubic_w(cc - f, w[a]);
#pragma unroll
      for (int i = 0; i < 4; ++i) idx[a][i] = min(max(fl - 1 + i, 0), dims[a] - 1);
    }
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      double pi = 0.0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const double* row = coef + ((long long)idx[0][i] * PX + idx[1][j]) * PY;
        double sj = row[idx[2][0]] * w[2][0];
        sj = fma(row[idx[2][1]], w[2][1], sj);
        sj = fma(row[idx[2][2]], w[2][2], sj);
        sj = fma(row[idx[2][3]], w[2][3], sj);
        pi = fma(sj, w[1][j], pi);
      }
      acc = fma(pi, w[0][i], acc);
    }
    double r = acc > 0.0 ? acc + 0.5 : 0.0;                     // ndimage's store to an unsigned integer type
    r = r > 65535.0 ? 65535.0 : r;
    out[((long long)z * X + x) * Y + y] = (uint16_t)(unsigned)r;
  }
}

int launch_highpass(uint16_t* im, uint16_t* bufA, uint16_t* bufB, int Z, int X, int Y, const double* d_w, int r, cudaStream_t st) {
  const long long n = (long long)Z * X * Y;
  if (n == 0) return 0;
  const unsigned g = (unsigned)((n + 255) / 256);
  k_gauss_nearest<<<g, 256, 0, st>>>(im, bufA, Z, (long long)X * Y, n, d_w, r);
  IA3_LAUNCH_CHECK();
  k_gauss_nearest<<<g, 256, 0, st>>>(bufA, bufB, X, (long long)Y, n, d_w, r);
  IA3_LAUNCH_CHECK();
  k_gauss_nearest<<<g, 256, 0, st>>>(bufB, bufA, Y, 1, n, d_w, r);
  IA3_LAUNCH_CHECK();
  k_highpass<<<148 * 8, 256, 0, st>>>(im, bufA, n);
  IA3_LAUNCH_CHECK();
  return 0;
}
int launch_zshift(uint16_t* im, long long XY, long long n, const float* d_med_z, float med_all, cudaStream_t st) {
  if (n == 0) return 0;
  k_zshift<<<148 * 8, 256, 0, st>>>(im, XY, n, d_med_z, med_all);
  IA3_LAUNCH_CHECK();
  return 0;
}
int launch_hot_count(const uint16_t* im, int Z, int X, int Y, float hot_th, int* cnt, cudaStream_t st) {
  const long long n = (long long)X * Y;
  k_hot_count<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(im, Z, X, Y, hot_th, cnt);
  IA3_LAUNCH_CHECK();
  return 0;
}
int launch_hot_select(const int* cnt, long long n, double thr, int* out, int* count, int cap, cudaStream_t st) {
  IA3_CUDA(cudaMemsetAsync(count, 0, sizeof(int), st));
  k_hot_select<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(cnt, n, thr, out, count, cap);
  IA3_LAUNCH_CHECK();
  return 0;
}
int launch_hot_fix(uint16_t* im, int Z, int X, int Y, const int* list, int n, int* slot, float* vals, cudaStream_t st) {
  if (n == 0) return 0;
  IA3_CUDA(cudaMemsetAsync(slot, 0xff, (size_t)X * Y * sizeof(int), st));
  k_hot_slots<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(list, n, slot);
  IA3_LAUNCH_CHECK();
  k_hot_fix<<<1, 128, 0, st>>>(im, Z, X, Y, list, n, slot, vals);
  IA3_LAUNCH_CHECK();
  return 0;
}
template <typename Tp>
int launch_mix(const uint16_t* const* d_ins, int n_in, const Tp* bleed, const Tp* illum, uint16_t* out, long long XY, long long n, cudaStream_t st) {
  if (n == 0) return 0;
  k_mix<Tp><<<148 * 8, 256, 0, st>>>(d_ins, n_in, bleed, illum, out, XY, n);
  IA3_LAUNCH_CHECK();
  return 0;
}
template int launch_mix<float>(const uint16_t* const*, int, const float*, const float*, uint16_t*, long long, long long, cudaStream_t);
template int launch_mix<double>(const uint16_t* const*, int, const double*, const double*, uint16_t*, long long, long long, cudaStream_t);
long long warp_padded_voxels(int Z, int X, int Y) { return (long long)(Z + 2 * NPAD) * (X + 2 * NPAD) * (Y + 2 * NPAD); }
int launch_warp(const uint16_t* im, int Z, int X, int Y, double* buf, const void* chroma, int chroma_f64, int CZ,
                double d0, double d1, double d2, uint16_t* out, cudaStream_t st) {
  const int PZ = Z + 2 * NPAD, PX = X + 2 * NPAD, PY = Y + 2 * NPAD;
  k_spline_iir<uint16_t><<<(unsigned)(((long long)PX * PY + 127) / 128), 128, 0, st>>>(im, buf, Z, X, Y, 0);
  IA3_LAUNCH_CHECK();
  k_spline_iir<double><<<(unsigned)(((long long)PZ * PY + 127) / 128), 128, 0, st>>>(buf, buf, Z, X, Y, 1);
  IA3_LAUNCH_CHECK();
  const long long n_rows = (long long)PZ * PX;
  k_spline_iir_y<<<(unsigned)((n_rows + YR - 1) / YR), YR, 0, st>>>(buf, n_rows, PY);
  IA3_LAUNCH_CHECK();
  const dim3 grid((unsigned)((Y + 31) / 32), (unsigned)((X + 7) / 8));
  if (chroma_f64) k_warp<double><<<grid, 256, 0, st>>>(buf, Z, X, Y, (const double*)chroma, CZ, d0, d1, d2, out);
  else k_warp<float><<<grid, 256, 0, st>>>(buf, Z, X, Y, (const float*)chroma, CZ, d0, d1, d2, out);
  IA3_LAUNCH_CHECK();
  return 0;
}

}  // namespace ia3
