// Seed stage kernels (sm_100a): exact separable Gaussian, 3D rank filters + candidate mask,
// ordered prefix-sum stream compaction.
//
// What "exact" means here (spot_tools/fitting.py:91-125): scipy's
// gaussian_filter on a uint16 stack runs three 1-D correlations (axes 0,1,2); each pass
// accumulates in double as   acc = x[0]*w[0];  for j = r..1: acc += (x[-j] + x[+j]) * w[j]
// (outermost pair first, multiply and add NOT fused) and stores (uint16)acc -- truncation --
// before the next axis reads it; boundary = reflect.  Every kernel below delivers the bits of that
// arithmetic; the filtered volumes, and therefore the seeds, are bit-identical to scipy's.
//
// Kernels, in the order a uint16 stack meets them (one launch per axis per blur):
//   k_gauss_short<R,L>    z pass of the dataset depths (L = 30, 50): the whole line in registers, reflected tap
//                         indices are compile-time constants
//   k_gauss_strided<R>    x pass: a thread slides a register window of 2R + 8 inputs along its line, the lanes
//                         of a warp are consecutive in y (every warp load / store is one 64-byte request)
//   k_gauss_contig<R> /   y pass: same window along the contiguous axis / (7 taps) a warp owns 256 consecutive
//   k_gauss_row<3>        outputs, halo values come from the neighbouring lanes by shuffle
// These do NOT replay scipy's 91 FP64 operations per 61-tap output: window values are kept biased so that a pair
// sum is one integer add that already is the high word of a double, the taps are accumulated with FMA (32 FP64
// instructions per output), and the few outputs whose fast sum lies within a guard band of an integer (where
// truncation could differ) are recomputed in scipy's operation order (exact_from_global).  See DESIGN.md 4.1.
//   k_gauss_axis<R,T>     general tiled kernel, scipy's operation order throughout (__dmul_rn / __dadd_rn are never
//                         contracted to FMA): float32 and float64 stacks, unusual radii, ragged shapes.  A block owns
//                         128 lines x TL axis positions; the (TL + 2r) x 128 input tile sits in shared memory (float for
//                         uint16 / float32 inputs -- exact --, double for float64) with an odd pitch, one thread
//                         walks one line and produces 8 outputs per step from a register window of 8 + 2r doubles.
//   k_flags_u16 / k_flags 3x3x3 (or any size) max / min rank filters, candidate mask, threshold, edge test
//   k_count_bits, k_scan_counts, k_emit   ordered compaction (np.where order, no atomics)
//   k_box_background, k_hist_accum        histogram-mode backgrounds (normalize_local / normalize_background)
// Bound: the 61-tap passes are FP64-pipe / issue bound (measured 0.77-0.85 ms per C2 pass against 0.13 ms of HBM
// time), DRAM traffic per pass = algorithmic; DESIGN.md states both rooflines.
#include <mutex>
#include "ia3_device.h"
#include "seed_kernels.h"

namespace ia3 {

__device__ __forceinline__ int reflect_idx(int i, int n) {
  // scipy NI_EXTEND_REFLECT: (d c b a | a b c d | d c b a), any number of reflections
  if (n <= 1) return 0;
  const int p = 2 * n;
  int m = i % p;
  if (m < 0) m += p;
  return m < n ? m : p - 1 - m;
}

template <typename T> __device__ __forceinline__ T store_cast(double acc);
template <> __device__ __forceinline__ uint16_t store_cast<uint16_t>(double acc) {
  return (uint16_t)__double2int_rz(acc);   // C cast double -> npy_uint16: truncation
}
template <> __device__ __forceinline__ float store_cast<float>(double acc) { return __double2float_rn(acc); }
template <> __device__ __forceinline__ double store_cast<double>(double acc) { return acc; }   // float64 stacks: no rounding between passes
// shared-memory tile element of the general kernel: float holds uint16 / float32 inputs exactly
template <typename Tin> struct TileOf { using type = float; };
template <> struct TileOf<double> { using type = double; };

constexpr size_t kMaxDynSmem = 227 * 1024;   // opt-in maximum per CTA on sm_100a
constexpr int LINES = 128;   // lines per block = threads per block
constexpr int CHMAX = 8;     // outputs per register-window step (4 for the 81-tap legacy filter: register budget)

// R > 0: compile-time radius (fully unrolled register window).  R == 0: runtime radius gw.r,
// taps read straight from shared memory (slow path for unusual sigmas).
template <int R, typename Tin, bool INNER1>
__global__ void __launch_bounds__(LINES, 2)
k_gauss_axis(const Tin* __restrict__ in, Tin* __restrict__ out, int L, long long inner, long long n_lines,
             int TL, GaussW gw) {
  extern __shared__ __align__(16) unsigned char smem_raw_g[];
  using TileT = typename TileOf<Tin>::type;
  TileT* smem = reinterpret_cast<TileT*>(smem_raw_g);
  constexpr int CH = (R > 32) ? 4 : CHMAX;
  const int r = (R > 0) ? R : gw.r;
  const int span = TL + 2 * r;
  const int pitch = span | 1;
  TileT* tile = smem;
  const int tid = threadIdx.x;
  const int a0 = blockIdx.y * TL;
  const long long l0 = (long long)blockIdx.x * LINES;
  const bool interior = (a0 - r >= 0) && (a0 + TL + r <= L);

  long long base = 0;
  bool ok = false;
  if (!INNER1) {
    const long long l = l0 + tid;
    ok = l < n_lines;
    if (ok) base = (l / inner) * ((long long)L * inner) + (l % inner);
    const Tin* src = in + base;
    if (interior) {
      const Tin* s2 = src + (long long)(a0 - r) * inner;
#pragma unroll 4
      for (int a = 0; a < span; ++a) tile[tid * pitch + a] = ok ? (TileT)s2[(long long)a * inner] : (TileT)0;
    } else {
      for (int a = 0; a < span; ++a) {
        const int sidx = reflect_idx(a0 - r + a, L);
        tile[tid * pitch + a] = ok ? (TileT)src[(long long)sidx * inner] : (TileT)0;
      }
    }
  } else {
    const int warp = tid >> 5, lane = tid & 31;
    for (int ll = warp; ll < LINES; ll += LINES / 32) {
      const long long l = l0 + ll;
      if (l >= n_lines) break;
      const Tin* row = in + l * (long long)L;
      if (interior) {
        for (int a = lane; a < span; a += 32) tile[ll * pitch + a] = (TileT)row[a0 - r + a];
      } else {
        for (int a = lane; a < span; a += 32) tile[ll * pitch + a] = (TileT)row[reflect_idx(a0 - r + a, L)];
      }
    }
    ok = (l0 + tid) < n_lines;
  }
  __syncthreads();

  // staging tile for the contiguous-axis pass (outputs must leave coalesced)
  Tin* otile = reinterpret_cast<Tin*>(smem + LINES * pitch);
  const int opitch = (sizeof(Tin) == 2) ? (TL + 2) : (TL + 1);

  const TileT* my = tile + tid * pitch;
  const int nvalid = ok ? min(TL, L - a0) : 0;
  for (int c = 0; c < nvalid; c += CH) {
    double acc[CH];
    if (R > 0) {
      double v[CH + 2 * (R > 0 ? R : 1)];
#pragma unroll
      for (int i = 0; i < CH + 2 * R; ++i) v[i] = (double)my[c + i];
#pragma unroll
      for (int o = 0; o < CH; ++o) {
        double a = __dmul_rn(v[o + R], gw.w[0]);
#pragma unroll
        for (int j = R; j >= 1; --j) a = __dadd_rn(a, __dmul_rn(__dadd_rn(v[o + R - j], v[o + R + j]), gw.w[j]));
        acc[o] = a;
      }
    } else {
#pragma unroll
      for (int o = 0; o < CH; ++o) {
        const TileT* ctr = my + c + o + r;
        double a = __dmul_rn((double)ctr[0], gw.w[0]);
        for (int j = r; j >= 1; --j) a = __dadd_rn(a, __dmul_rn(__dadd_rn((double)ctr[-j], (double)ctr[j]), gw.w[j]));
        acc[o] = a;
      }
    }
    if (!INNER1) {
      if (ok) {
#pragma unroll
        for (int o = 0; o < CH; ++o)
          if (c + o < nvalid) out[base + (long long)(a0 + c + o) * inner] = store_cast<Tin>(acc[o]);
      }
    } else {
#pragma unroll
      for (int o = 0; o < CH; ++o) otile[tid * opitch + c + o] = store_cast<Tin>(acc[o]);
    }
  }
  if (INNER1) {
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31;
    for (int ll = warp; ll < LINES; ll += LINES / 32) {
      const long long l = l0 + ll;
      if (l >= n_lines) break;
      Tin* row = out + l * (long long)L + a0;
      const int nout = min(TL, L - a0);
      for (int a = lane; a < nout; a += 32) row[a] = otile[ll * opitch + a];
    }
  }
}

// ------------------------------------------------------------------------------------------
// uint16 fast path (the production dtype).  Same tiling as k_gauss_axis, but
//   * the tile stays uint16 in shared memory (half the footprint -> longer tiles, less halo);
//   * uint16 -> double by the 2^52 mantissa trick (one FP64 add; F2F/I2F conversions run at a
//     quarter of the FP64 rate and were a third of the old kernel's time);
//   * the taps are accumulated with FMA, 31 FP64 instructions per output instead of 91.  That sum
//     is NOT scipy's sum (different roundings), but both are within 61 * ulp(65535)/2 < 5e-10 of the
//     exact value, so (uint16)acc can only differ when an integer lies within that distance of the
//     FMA sum.  Exactly those outputs (flat regions where the true sum is an integer, ~1e-9 of the
//     rest) are recomputed in scipy's operation order by exact_taps() -- the result is bit-identical
//     to scipy for every voxel, at a third of the FP64 work.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double u16_to_f64(unsigned v) {
  return __hiloint2double(0x43300000, (int)v) - 4503599627370496.0;    // (2^52 + v) - 2^52, exact
}

template <int R>
__device__ __noinline__ unsigned exact_taps(const uint16_t* ctr, const double* w) {
  double a = __dmul_rn((double)ctr[0], w[0]);
  for (int j = R; j >= 1; --j) a = __dadd_rn(a, __dmul_rn(__dadd_rn((double)ctr[-j], (double)ctr[j]), w[j]));
  return (unsigned)__double2int_rz(a);
}

constexpr double kTruncGuard = 2.0e-8;   // > 6x the worst-case |fast sum - scipy sum| for uint16 data (see k_gauss_u16)

template <int R, bool INNER1>
__global__ void __launch_bounds__(LINES, (R > 8) ? 4 : 6)
k_gauss_u16(const uint16_t* __restrict__ in, uint16_t* __restrict__ out, int L, long long inner, long long n_lines,
            int TL, int pitch, GaussW gw) {
  extern __shared__ __align__(16) uint16_t tile16[];
  __shared__ double wsh[R + 1];                    // weights for the (rare) exact path: the kernel
  constexpr int CH = 8;                            // parameter itself must stay in the constant bank
  const int span = TL + 2 * R;
  const int tid = threadIdx.x;
  const int a0 = blockIdx.y * TL;
  const long long l0 = (long long)blockIdx.x * LINES;
  const bool interior = (a0 - R >= 0) && (a0 + TL + R <= L);
  uint16_t* otile = tile16 + LINES * pitch;       // INNER1 only: outputs leave coalesced
  const int opitch = TL + 2;                      // (TL + 2) / 2 odd -> conflict-free 16-bit columns
  // reflected source index of every tile position, once per CTA (the modulo is an integer division);
  // kept behind the tiles
  int* ridx = reinterpret_cast<int*>(tile16 + LINES * pitch + (INNER1 ? LINES * opitch : 0));
  if (tid <= R) wsh[tid] = gw.w[tid];
  if (!interior) {
    for (int a = tid; a < span; a += LINES) ridx[a] = reflect_idx(a0 - R + a, L);
    __syncthreads();
  }

  long long base = 0;
  bool ok = false;
  if (!INNER1) {
    const long long l = l0 + tid;
    ok = l < n_lines;
    if (ok) base = (l / inner) * ((long long)L * inner) + (l % inner);
    const uint16_t* src = in + base;
    uint16_t* mine = tile16 + tid * pitch;
    if (interior) {
      const uint16_t* s2 = src + (long long)(a0 - R) * inner;
      if (ok) {
#pragma unroll 8
        for (int a = 0; a < span; ++a, s2 += inner) mine[a] = *s2;
      }
    } else {
      if (ok) {
#pragma unroll 8
        for (int a = 0; a < span; ++a) mine[a] = src[(long long)ridx[a] * inner];
      }
    }
  } else {
    const int warp = tid >> 5, lane = tid & 31;
    for (int ll = warp; ll < LINES; ll += LINES / 32) {
      const long long l = l0 + ll;
      if (l >= n_lines) break;
      const uint16_t* row = in + l * (long long)L;
      uint16_t* dst = tile16 + ll * pitch;
      if (interior) {
        for (int a = lane; a < span; a += 32) dst[a] = row[a0 - R + a];
      } else {
        for (int a = lane; a < span; a += 32) dst[a] = row[ridx[a]];
      }
    }
    ok = (l0 + tid) < n_lines;
  }
  __syncthreads();

  const uint16_t* my = tile16 + tid * pitch;
  const int nvalid = ok ? min(TL, L - a0) : 0;
  uint16_t* optr = out + base + (long long)a0 * inner;      // !INNER1: walks down the axis
  for (int c = 0; c < nvalid; c += CH) {
    // Integer register window.  A pair sum s = x[-j] + x[+j] (< 2^18) becomes the double 2^20 + s by
    // adding it to the HIGH word of 2^20 (one 3-input integer add, low word 0): no int->fp conversion
    // at all.  The taps are then accumulated as sum_j w_j (2^20 + s_j) and the known constant
    // 2^20 sum_j w_j (gw.w[R + 1], computed by the host) is removed once at the end: 31 DFMA + 1 DADD
    // per output.  The offset costs four bits (accumulator ~6e5 instead of <= 65535): worst-case
    // |a - exact| < 3e-9, inside kTruncGuard.
    int v[CH + 2 * R];
#pragma unroll
    for (int i = 0; i < CH + 2 * R; ++i) v[i] = my[c + i];
    unsigned res[CH];
    double acc[CH];                             // CH independent FMA chains (tap loop outside)
#pragma unroll
    for (int o = 0; o < CH; ++o) acc[o] = __hiloint2double(0x41300000 + v[o + R], 0) * gw.w[0];
#pragma unroll
    for (int j = R; j >= 1; --j) {
#pragma unroll
      for (int o = 0; o < CH; ++o)
        acc[o] = fma(__hiloint2double(0x41300000 + v[o + R - j] + v[o + R + j], 0), gw.w[j], acc[o]);
    }
#pragma unroll
    for (int o = 0; o < CH; ++o) {
      const double a = acc[o] - gw.w[R + 1];
      // trunc(a -+ guard) via round-toward-zero addition of 2^52: the low word is the integer part
      const unsigned rlo = (unsigned)__double2loint(__dadd_rz(fmax(a - kTruncGuard, 0.0), 4503599627370496.0));
      const unsigned rhi = (unsigned)__double2loint(__dadd_rz(a + kTruncGuard, 4503599627370496.0));
      res[o] = rlo;
      if (rlo != rhi) res[o] = exact_taps<R>(my + c + o + R, wsh);
    }
    if (!INNER1) {
#pragma unroll
      for (int o = 0; o < CH; ++o, optr += inner)
        if (c + o < nvalid) *optr = (uint16_t)res[o];
    } else {
#pragma unroll
      for (int o = 0; o < CH; ++o) otile[tid * opitch + c + o] = (uint16_t)res[o];
    }
  }
  if (INNER1) {
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31;
    for (int ll = warp; ll < LINES; ll += LINES / 32) {
      const long long l = l0 + ll;
      if (l >= n_lines) break;
      uint16_t* row = out + l * (long long)L + a0;
      const int nout = min(TL, L - a0);
      for (int a = lane; a < nout; a += 32) row[a] = otile[ll * opitch + a];
    }
  }
}

// ------------------------------------------------------------------------------------------
// Register-window line kernels (uint16): no shared-memory tile at all.  A thread owns one line
// segment and slides a window of 2R (+ alignment) + U input values held in registers along the
// filter axis: per U outputs it loads U new values, computes, stores and shifts the window (moves,
// not loads).  k_gauss_strided: the axis is strided (z and x passes); the lanes of a warp are
// consecutive along the contiguous dimension, so every load/store of the warp is one coalesced
// 64-byte request.  k_gauss_contig: the axis is the contiguous one (y pass); a thread walks its own
// row with 16-byte loads/stores (L1 keeps the 128-byte lines between a thread's consecutive chunks).
// Reflection indices of segments that touch the ends of the axis come from a per-CTA table.
// ------------------------------------------------------------------------------------------
template <int R>
__device__ __noinline__ unsigned exact_from_global(const uint16_t* line, long long stride, int a, int L, const double* w) {
  double acc = __dmul_rn((double)line[(long long)a * stride], w[0]);
  for (int j = R; j >= 1; --j) {
    const double lo = (double)line[(long long)reflect_idx(a - j, L) * stride];
    const double hi = (double)line[(long long)reflect_idx(a + j, L) * stride];
    acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(lo, hi), w[j]));
  }
  return (unsigned)__double2int_rz(acc);
}

// U outputs from the register window; output o is centred on w[C + o].  Window entries are stored
// biased, w = x + kHalfBias (half the high word of 2^20), so a pair sum is ONE integer add and is
// already the high word of the double 2^20 + x- + x+.  gw.w[R + 1] = 2^16 + guard - 2^20 * sum(w)
// turns the accumulator into t = 2^16 + v + guard: the truncated value sits in bits 4..19 of t's high
// word, the next 32 bits are the fraction.  fraction < 2 * guard <=> v is within the guard of an
// integer n -> recompute in scipy's exact order (n = 0 needs no check: the sum cannot be negative).
constexpr int kHalfBias = 0x20980000;
constexpr unsigned kGuardFrac = 172;       // 2 * kTruncGuard in units of 2^-32
template <int R, int U, int C, int NW>
__device__ __forceinline__ void window_outputs(const int (&w)[NW], const GaussW& gw, const double* wsh,
                                               const uint16_t* line, long long stride, int a, int L, unsigned (&res)[U]) {
  double acc[U];
#pragma unroll
  for (int o = 0; o < U; ++o) acc[o] = __hiloint2double(w[C + o] + kHalfBias, 0) * gw.w[0];
#pragma unroll
  for (int j = R; j >= 1; --j) {
#pragma unroll
    for (int o = 0; o < U; ++o)
      acc[o] = fma(__hiloint2double(w[C + o - j] + w[C + o + j], 0), gw.w[j], acc[o]);
  }
  unsigned need = 0;
#pragma unroll
  for (int o = 0; o < U; ++o) {
    const double t = acc[o] + gw.w[R + 1];
    const unsigned hi = (unsigned)__double2hiint(t), lo = (unsigned)__double2loint(t);
    const unsigned frac = __funnelshift_l(lo, hi, 28);
    const unsigned n = (hi >> 4) & 0xffffu;
    res[o] = n;
    if (frac < kGuardFrac && n != 0) need |= 1u << o;
  }
  if (need) {
#pragma unroll
    for (int o = 0; o < U; ++o)
      if ((need >> o & 1u) && a + o < L) res[o] = exact_from_global<R>(line, stride, a + o, L, wsh);
  }
}

template <int R>
__global__ void __launch_bounds__(LINES, (R > 32) ? 3 : 4)
k_gauss_strided(const uint16_t* __restrict__ in, uint16_t* __restrict__ out, int L, long long inner, long long n_lines,
                int seg, GaussW gw) {
  constexpr int U = 8, NW = 2 * R + U;
  extern __shared__ int ridx_s[];                  // seg rounded up to U, + 2R entries
  __shared__ double wsh[R + 1];
  const int tid = threadIdx.x;
  const int a0 = blockIdx.y * seg, a1 = min(L, a0 + seg);
  const int nload = ((a1 - a0 + U - 1) / U) * U + 2 * R;
  const bool edge = (a0 - R < 0) || (a0 - R + nload > L);
  if (tid <= R) wsh[tid] = gw.w[tid];
  if (edge) for (int i = tid; i < nload; i += LINES) ridx_s[i] = reflect_idx(a0 - R + i, L);
  __syncthreads();
  const long long l = (long long)blockIdx.x * LINES + tid;
  if (l >= n_lines) return;
  const long long base = (l / inner) * ((long long)L * inner) + (l % inner);
  const uint16_t* line = in + base;
  uint16_t* op = out + base + (long long)a0 * inner;
  int w[NW];
  unsigned res[U];
  if (!edge) {
    const uint16_t* p = line + (long long)(a0 - R) * inner;
#pragma unroll
    for (int i = 0; i < 2 * R; ++i, p += inner) w[i] = *p | kHalfBias;
    for (int a = a0; a < a1; a += U) {
#pragma unroll
      for (int i = 0; i < U; ++i, p += inner) w[2 * R + i] = *p | kHalfBias;
      window_outputs<R, U, R, NW>(w, gw, wsh, line, inner, a, L, res);
#pragma unroll
      for (int o = 0; o < U; ++o, op += inner) if (a + o < a1) *op = (uint16_t)res[o];
#pragma unroll
      for (int i = 0; i < 2 * R; ++i) w[i] = w[i + U];
    }
  } else {
    const int* ri = ridx_s;
#pragma unroll
    for (int i = 0; i < 2 * R; ++i) w[i] = line[(long long)ri[i] * inner] | kHalfBias;
    ri += 2 * R;
    for (int a = a0; a < a1; a += U, ri += U) {
#pragma unroll
      for (int i = 0; i < U; ++i) w[2 * R + i] = line[(long long)ri[i] * inner] | kHalfBias;
      window_outputs<R, U, R, NW>(w, gw, wsh, line, inner, a, L, res);
#pragma unroll
      for (int o = 0; o < U; ++o, op += inner) if (a + o < a1) *op = (uint16_t)res[o];
#pragma unroll
      for (int i = 0; i < 2 * R; ++i) w[i] = w[i + U];
    }
  }
}

template <int R>
__global__ void __launch_bounds__(LINES, (R > 32) ? 3 : 4)
k_gauss_contig(const uint16_t* __restrict__ in, uint16_t* __restrict__ out, int L, long long n_rows, int seg, GaussW gw) {
  constexpr int U = 8, H = (R + 7) / 8 * 8, NW = 2 * H + U;      // window = [a - H, a + U + H), 16-byte aligned
  extern __shared__ int ridx_s[];
  __shared__ double wsh[R + 1];
  const int tid = threadIdx.x;
  const int a0 = blockIdx.y * seg, a1 = min(L, a0 + seg);          // seg and L are multiples of 8
  const int nload = (a1 - a0) + 2 * H;
  const bool edge = (a0 - H < 0) || (a0 - H + nload > L);
  if (tid <= R) wsh[tid] = gw.w[tid];
  if (edge) for (int i = tid; i < nload; i += LINES) ridx_s[i] = reflect_idx(a0 - H + i, L);
  __syncthreads();
  const long long row = (long long)blockIdx.x * LINES + tid;
  if (row >= n_rows) return;
  const uint16_t* line = in + row * (long long)L;
  uint16_t* op = out + row * (long long)L + a0;
  int w[NW];
  unsigned res[U];
  auto unpack = [&](const uint4& q, int at) {
    w[at + 0] = __byte_perm(q.x, kHalfBias, 0x7610); w[at + 1] = __byte_perm(q.x, kHalfBias, 0x7632);
    w[at + 2] = __byte_perm(q.y, kHalfBias, 0x7610); w[at + 3] = __byte_perm(q.y, kHalfBias, 0x7632);
    w[at + 4] = __byte_perm(q.z, kHalfBias, 0x7610); w[at + 5] = __byte_perm(q.z, kHalfBias, 0x7632);
    w[at + 6] = __byte_perm(q.w, kHalfBias, 0x7610); w[at + 7] = __byte_perm(q.w, kHalfBias, 0x7632);
  };
  const int* ri = ridx_s;
  if (!edge) {
    const uint4* p = reinterpret_cast<const uint4*>(line + (a0 - H));
#pragma unroll
    for (int c = 0; c < 2 * H / 8; ++c) unpack(__ldg(p + c), 8 * c);
  } else {
#pragma unroll
    for (int i = 0; i < 2 * H; ++i) w[i] = line[ri[i]] | kHalfBias;
  }
  for (int a = a0; a < a1; a += U) {
    if (!edge) unpack(__ldg(reinterpret_cast<const uint4*>(line + (a + H))), 2 * H);
    else {
#pragma unroll
      for (int i = 0; i < U; ++i) w[2 * H + i] = line[ri[(a - a0) + 2 * H + i]] | kHalfBias;
    }
    window_outputs<R, U, H, NW>(w, gw, wsh, line, 1, a, L, res);
    uint4 q;
    q.x = res[0] | (res[1] << 16); q.y = res[2] | (res[3] << 16); q.z = res[4] | (res[5] << 16); q.w = res[6] | (res[7] << 16);
    *reinterpret_cast<uint4*>(op) = q;
    op += U;
#pragma unroll
    for (int i = 0; i < 2 * H; ++i) w[i] = w[i + U];
  }
}

// Contiguous axis, small radius (R <= 8): a warp owns 256 consecutive outputs of one row.  Every lane
// loads its 8 values with one 16-byte load (the warp's request is one coalesced 512-byte line) and
// takes the R halo values on each side from its neighbours' registers by shuffle; only the first and
// last lane of a row segment read their halo from memory (reflected at the row ends).
template <int R>
__global__ void __launch_bounds__(LINES)
k_gauss_row(const uint16_t* __restrict__ in, uint16_t* __restrict__ out, int L, long long n_rows, int segs, GaussW gw) {
  static_assert(R <= 8, "halo must come from the adjacent lanes");
  constexpr int U = 8, NW = 2 * R + U;
  __shared__ double wsh[R + 1];
  if (threadIdx.x <= R) wsh[threadIdx.x] = gw.w[threadIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long item = (long long)blockIdx.x * (LINES / 32) + (threadIdx.x >> 5);     // (row, segment)
  if (item >= n_rows * segs) return;
  const long long row = item / segs;
  const int y0 = (int)(item - row * segs) * 256 + lane * U;
  const uint16_t* line = in + row * (long long)L;
  const bool active = y0 < L;                                   // L is a multiple of 8
  uint4 q = make_uint4(0, 0, 0, 0);
  if (active) q = __ldg(reinterpret_cast<const uint4*>(line + y0));
  // neighbours' packed words: the last 8 values of lane - 1, the first 8 of lane + 1
  uint4 ql, qr;
  ql.x = __shfl_up_sync(0xffffffffu, q.x, 1); ql.y = __shfl_up_sync(0xffffffffu, q.y, 1);
  ql.z = __shfl_up_sync(0xffffffffu, q.z, 1); ql.w = __shfl_up_sync(0xffffffffu, q.w, 1);
  qr.x = __shfl_down_sync(0xffffffffu, q.x, 1); qr.y = __shfl_down_sync(0xffffffffu, q.y, 1);
  qr.z = __shfl_down_sync(0xffffffffu, q.z, 1); qr.w = __shfl_down_sync(0xffffffffu, q.w, 1);
  if (!active) return;
  int w[NW];
  auto unpack = [&](const uint4& v, int (&d)[8]) {
    d[0] = __byte_perm(v.x, kHalfBias, 0x7610); d[1] = __byte_perm(v.x, kHalfBias, 0x7632);
    d[2] = __byte_perm(v.y, kHalfBias, 0x7610); d[3] = __byte_perm(v.y, kHalfBias, 0x7632);
    d[4] = __byte_perm(v.z, kHalfBias, 0x7610); d[5] = __byte_perm(v.z, kHalfBias, 0x7632);
    d[6] = __byte_perm(v.w, kHalfBias, 0x7610); d[7] = __byte_perm(v.w, kHalfBias, 0x7632);
  };
  int c[8], l8[8], r8[8];
  unpack(q, c); unpack(ql, l8); unpack(qr, r8);
#pragma unroll
  for (int i = 0; i < U; ++i) w[R + i] = c[i];
  const bool first = (lane == 0), last = (lane == 31) || (y0 + U >= L);
  // R <= 8 <= L: at most one reflection, no modulo needed
#pragma unroll
  for (int i = 0; i < R; ++i) {
    int yl = y0 - R + i, yr = y0 + U + i;
    if (yl < 0) yl = -1 - yl;
    if (yr >= L) yr = 2 * L - 1 - yr;
    w[i] = first ? (int)(line[yl] | kHalfBias) : l8[8 - R + i];
    w[R + U + i] = last ? (int)(line[yr] | kHalfBias) : r8[i];
  }
  unsigned res[U];
  window_outputs<R, U, R, NW>(w, gw, wsh, line, 1, y0, L, res);
  uint4 o;
  o.x = res[0] | (res[1] << 16); o.y = res[2] | (res[3] << 16); o.z = res[4] | (res[5] << 16); o.w = res[6] | (res[7] << 16);
  *reinterpret_cast<uint4*>(out + row * (long long)L + y0) = o;
}

template <int R>
static int launch_row_u16(const uint16_t* in, uint16_t* out, int L, long long n_rows, const GaussW& gw, cudaStream_t st) {
  GaussW gk = gw;
  long double acc = 0.0L;
  for (int j = 0; j <= R; ++j) acc += (long double)gw.w[j];
  gk.w[R + 1] = (double)(65536.0L + (long double)kTruncGuard - acc * 1048576.0L);
  const int segs = (L + 255) / 256;
  const long long items = n_rows * segs;
  k_gauss_row<R><<<(unsigned)((items + LINES / 32 - 1) / (LINES / 32)), LINES, 0, st>>>(in, out, L, n_rows, segs, gk);
  IA3_LAUNCH_CHECK();
  return 0;
}

// Short strided axis whose length is a compile-time constant (the z pass of the depths the
// reference's datasets use): a thread holds its WHOLE line in registers, the reflected tap indices
// are constants, so there is no window shift, no index table and no preload of 2R halo values -- the
// segment kernel above spends half its instructions on those when the axis is shorter than the filter.
__host__ __device__ constexpr int reflect_const(int i, int n) {
  const int p = 2 * n;
  int m = i % p;
  if (m < 0) m += p;
  return m < n ? m : p - 1 - m;
}
template <int R, int L, int A0, int U>
__device__ __forceinline__ void static_outputs(int (&w)[L], const GaussW& gw, const double* wsh, const uint16_t* line,
                                               long long stride, uint16_t* op) {
  constexpr int N = (A0 + U <= L) ? U : (L - A0);
  // the reflection makes index pairs repeat between output groups; without this the compiler keeps
  // those pair sums alive across groups and spills
#pragma unroll
  for (int i = 0; i < L; ++i) asm volatile("" : "+r"(w[i]));
  double acc[N];
#pragma unroll
  for (int o = 0; o < N; ++o) acc[o] = __hiloint2double(w[A0 + o] + kHalfBias, 0) * gw.w[0];
#pragma unroll
  for (int j = R; j >= 1; --j) {
#pragma unroll
    for (int o = 0; o < N; ++o)
      acc[o] = fma(__hiloint2double(w[reflect_const(A0 + o - j, L)] + w[reflect_const(A0 + o + j, L)], 0), gw.w[j], acc[o]);
  }
  unsigned res[N];
  unsigned need = 0;
#pragma unroll
  for (int o = 0; o < N; ++o) {
    const double t = acc[o] + gw.w[R + 1];
    const unsigned hi = (unsigned)__double2hiint(t), lo = (unsigned)__double2loint(t);
    const unsigned frac = __funnelshift_l(lo, hi, 28);
    const unsigned n = (hi >> 4) & 0xffffu;
    res[o] = n;
    if (frac < kGuardFrac && n != 0) need |= 1u << o;
  }
  if (need) {
#pragma unroll
    for (int o = 0; o < N; ++o)
      if (need >> o & 1u) res[o] = exact_from_global<R>(line, stride, A0 + o, L, wsh);
  }
#pragma unroll
  for (int o = 0; o < N; ++o) op[(long long)(A0 + o) * stride] = (uint16_t)res[o];
  if constexpr (A0 + U < L) static_outputs<R, L, A0 + U, U>(w, gw, wsh, line, stride, op);
}

template <int R, int L>
__global__ void __launch_bounds__(LINES, 4)
k_gauss_short(const uint16_t* __restrict__ in, uint16_t* __restrict__ out, long long inner, long long n_lines, GaussW gw) {
  __shared__ double wsh[R + 1];
  if (threadIdx.x <= R) wsh[threadIdx.x] = gw.w[threadIdx.x];
  __syncthreads();
  const long long l = (long long)blockIdx.x * LINES + threadIdx.x;
  if (l >= n_lines) return;
  const long long base = (l / inner) * ((long long)L * inner) + (l % inner);
  const uint16_t* line = in + base;
  int w[L];
#pragma unroll
  for (int i = 0; i < L; ++i) w[i] = line[(long long)i * inner] | kHalfBias;
  static_outputs<R, L, 0, 8>(w, gw, wsh, line, inner, out + base);
}

template <int R, int L>
static int launch_short_u16(const uint16_t* in, uint16_t* out, long long inner, long long n_lines, const GaussW& gw, cudaStream_t st) {
  GaussW gk = gw;
  long double acc = 0.0L;
  for (int j = 0; j <= R; ++j) acc += (long double)gw.w[j];
  gk.w[R + 1] = (double)(65536.0L + (long double)kTruncGuard - acc * 1048576.0L);
  k_gauss_short<R, L><<<(unsigned)((n_lines + LINES - 1) / LINES), LINES, 0, st>>>(in, out, inner, n_lines, gk);
  IA3_LAUNCH_CHECK();
  return 0;
}

template <int R>
static int launch_line_u16(const uint16_t* in, uint16_t* out, int L, long long inner, long long n_lines, bool contig,
                           const GaussW& gw, cudaStream_t st) {
  GaussW gk = gw;                               // w[R + 1] = 2^16 + guard - 2^20 * (w0 + w1 + ... + wR), see window_outputs
  long double acc = 0.0L;
  for (int j = 0; j <= R; ++j) acc += (long double)gw.w[j];
  gk.w[R + 1] = (double)(65536.0L + (long double)kTruncGuard - acc * 1048576.0L);
  const int seg = (L <= 320) ? ((L + 7) / 8) * 8 : 256;
  dim3 grid((unsigned)((n_lines + LINES - 1) / LINES), (unsigned)((L + seg - 1) / seg));
  const size_t smem = (size_t)(seg + 2 * 32 + 2 * R + 16) * sizeof(int);
  if (contig) k_gauss_contig<R><<<grid, LINES, smem, st>>>(in, out, L, n_lines, seg, gk);
  else k_gauss_strided<R><<<grid, LINES, smem, st>>>(in, out, L, inner, n_lines, seg, gk);
  IA3_LAUNCH_CHECK();
  return 0;
}

template <int R, bool INNER1>
static int launch_axis_u16(const uint16_t* in, uint16_t* out, int L, long long inner, long long n_lines, const GaussW& gw,
                           cudaStream_t st) {
  int TL = 128;
  if (L < TL) TL = ((L + 7) / 8) * 8;
  const int span = TL + 2 * R;
  int pitch = span;
  while (pitch % 4 != 2) ++pitch;               // pitch / 2 odd -> conflict-free 16-bit columns
  size_t smem = (size_t)LINES * pitch * sizeof(uint16_t);
  if (INNER1) smem += (size_t)LINES * (TL + 2) * sizeof(uint16_t);
  smem += (size_t)span * sizeof(int);            // reflect index table
  GaussW gk = gw;                               // w[R + 1] = 2^20 * (w0 + w1 + ... + wR): the offset the kernel removes
  {
    long double acc = 0.0L;
    for (int j = 0; j <= R; ++j) acc += (long double)gw.w[j];
    gk.w[R + 1] = (double)(acc * 1048576.0L);
  }
  auto kern = k_gauss_u16<R, INNER1>;
  static std::once_flag once;                  // one flag per template instantiation
  static cudaError_t once_err = cudaSuccess;
  std::call_once(once, [&] { once_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kMaxDynSmem - 1024)); });
  IA3_CUDA(once_err);
  dim3 grid((unsigned)((n_lines + LINES - 1) / LINES), (unsigned)((L + TL - 1) / TL));
  kern<<<grid, LINES, smem, st>>>(in, out, L, inner, n_lines, TL, pitch, gk);
  IA3_LAUNCH_CHECK();
  return 0;
}

template <int R, typename Tin, bool INNER1>
static int launch_axis(const Tin* in, Tin* out, int L, long long inner, long long n_lines, const GaussW& gw,
                       cudaStream_t st) {
  const int r = (R > 0) ? R : gw.r;
  int TL = (sizeof(Tin) == 8 && r > 32) ? 32 : 64;          // float64 tiles are twice as large
  if (L < TL) TL = ((L + CHMAX - 1) / CHMAX) * CHMAX;
  const int span = TL + 2 * r, pitch = span | 1;
  size_t smem = (size_t)LINES * pitch * sizeof(typename TileOf<Tin>::type);
  if (INNER1) smem += (size_t)LINES * (TL + 2) * sizeof(Tin);
  auto kern = k_gauss_axis<R, Tin, INNER1>;
  // always the same (maximum) value: concurrent host threads launch the same instantiation with
  // different tile sizes, and the attribute is per-function state
  if (smem > kMaxDynSmem) { set_error("gaussian tile does not fit in shared memory"); return -1; }
  static std::once_flag once;                  // one flag per template instantiation
  static cudaError_t once_err = cudaSuccess;
  std::call_once(once, [&] { once_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmem); });
  IA3_CUDA(once_err);
  dim3 grid((unsigned)((n_lines + LINES - 1) / LINES), (unsigned)((L + TL - 1) / TL));
  kern<<<grid, LINES, smem, st>>>(in, out, L, inner, n_lines, TL, gw);
  IA3_LAUNCH_CHECK();
  return 0;
}

template <bool INNER1>
static int dispatch_axis_u16(const uint16_t* in, uint16_t* out, int L, long long inner, long long n_lines, const GaussW& gw,
                             cudaStream_t st) {
  // register-window kernels; the contiguous pass needs 16-byte aligned rows
  const bool aligned = (L % 8 == 0) && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) % 16 == 0);
  if (!INNER1) {                                  // stack depths of the reference's datasets: whole line in registers
    if (L == 50 && gw.r == 30) return launch_short_u16<30, 50>(in, out, inner, n_lines, gw, st);
    if (L == 50 && gw.r == 3) return launch_short_u16<3, 50>(in, out, inner, n_lines, gw, st);
    if (L == 30 && gw.r == 30) return launch_short_u16<30, 30>(in, out, inner, n_lines, gw, st);
    if (L == 30 && gw.r == 3) return launch_short_u16<3, 30>(in, out, inner, n_lines, gw, st);
  }
  if (INNER1 && aligned && gw.r == 3) return launch_row_u16<3>(in, out, L, n_lines, gw, st);
  if (!INNER1 || aligned) {
    switch (gw.r) {
      case 3: return launch_line_u16<3>(in, out, L, inner, n_lines, INNER1, gw, st);
      case 30: return launch_line_u16<30>(in, out, L, inner, n_lines, INNER1, gw, st);
      case 40: return launch_line_u16<40>(in, out, L, inner, n_lines, INNER1, gw, st);
      default: break;
    }
  }
  switch (gw.r) {
    case 3: return launch_axis_u16<3, INNER1>(in, out, L, inner, n_lines, gw, st);
    case 30: return launch_axis_u16<30, INNER1>(in, out, L, inner, n_lines, gw, st);
    case 40: return launch_axis_u16<40, INNER1>(in, out, L, inner, n_lines, gw, st);
    default: return launch_axis<0, uint16_t, INNER1>(in, out, L, inner, n_lines, gw, st);
  }
}

template <typename Tin, bool INNER1>
static int dispatch_axis(const Tin* in, Tin* out, int L, long long inner, long long n_lines, const GaussW& gw,
                         cudaStream_t st) {
  if constexpr (sizeof(Tin) == 2) return dispatch_axis_u16<INNER1>(in, out, L, inner, n_lines, gw, st);
  switch (gw.r) {
    case 3: return launch_axis<3, Tin, INNER1>(in, out, L, inner, n_lines, gw, st);
    case 30: return launch_axis<30, Tin, INNER1>(in, out, L, inner, n_lines, gw, st);
    case 40: return launch_axis<40, Tin, INNER1>(in, out, L, inner, n_lines, gw, st);
    default: return launch_axis<0, Tin, INNER1>(in, out, L, inner, n_lines, gw, st);
  }
}

// in -> bufA (z pass) -> bufB (x pass) -> bufA (y pass); result in bufA.
template <typename Tin>
int gaussian_filter_exact(const Tin* in, Tin* bufA, Tin* bufB, int Z, int X, int Y, const GaussW& gw,
                          cudaStream_t st, bool skip_z) {
  if (gw.r > GaussW::MAXR) { set_error("gaussian radius too large (max 95)"); return -1; }
  const long long XY = (long long)X * Y;
  int rc;
  // skip_z: a 2-D image held as a one-plane stack -- scipy filters its two axes only
  if (!skip_z && (rc = dispatch_axis<Tin, false>(in, bufA, Z, XY, XY, gw, st))) return rc;
  if ((rc = dispatch_axis<Tin, false>(skip_z ? in : bufA, bufB, X, Y, (long long)Z * Y, gw, st))) return rc;
  if ((rc = dispatch_axis<Tin, true>(bufB, bufA, Y, 1, (long long)Z * X, gw, st))) return rc;
  return 0;
}
template int gaussian_filter_exact<uint16_t>(const uint16_t*, uint16_t*, uint16_t*, int, int, int, const GaussW&, cudaStream_t, bool);
template int gaussian_filter_exact<float>(const float*, float*, float*, int, int, int, const GaussW&, cudaStream_t, bool);
template int gaussian_filter_exact<double>(const double*, double*, double*, int, int, int, const GaussW&, cudaStream_t, bool);

// ------------------------------------------------------------------------------------------
// Rank filters + candidate mask.  scipy's maximum_filter/minimum_filter(size=s) with reflect
// boundary equal the max/min over the in-bounds part of the window [i - s/2, i + s - s/2 - 1]
// on every axis (a reflected sample is always already inside the window).  One thread owns 8
// consecutive voxels of a row ("chunk"); chunks are numbered in C order so that the later
// prefix sum yields np.where's ordering.
// ------------------------------------------------------------------------------------------
constexpr int FLAG_THREADS = 256;

template <typename Tin> __device__ __forceinline__ void load10(const Tin* row, int y0, int Y, bool vec, Tin* v);
template <>
__device__ __forceinline__ void load10<uint16_t>(const uint16_t* row, int y0, int Y, bool vec, uint16_t* v) {
  if (vec) {
    const uint4 q = *reinterpret_cast<const uint4*>(row + y0);
    v[1] = q.x & 0xffff; v[2] = q.x >> 16; v[3] = q.y & 0xffff; v[4] = q.y >> 16;
    v[5] = q.z & 0xffff; v[6] = q.z >> 16; v[7] = q.w & 0xffff; v[8] = q.w >> 16;
    v[0] = row[max(y0 - 1, 0)];
    v[9] = row[min(y0 + 8, Y - 1)];
  } else {
#pragma unroll
    for (int i = 0; i < 10; ++i) v[i] = row[min(max(y0 - 1 + i, 0), Y - 1)];
  }
}
template <>
__device__ __forceinline__ void load10<float>(const float* row, int y0, int Y, bool vec, float* v) {
  if (vec) {
    const float4 a = *reinterpret_cast<const float4*>(row + y0);
    const float4 b = *reinterpret_cast<const float4*>(row + y0 + 4);
    v[1] = a.x; v[2] = a.y; v[3] = a.z; v[4] = a.w; v[5] = b.x; v[6] = b.y; v[7] = b.z; v[8] = b.w;
    v[0] = row[max(y0 - 1, 0)];
    v[9] = row[min(y0 + 8, Y - 1)];
  } else {
#pragma unroll
    for (int i = 0; i < 10; ++i) v[i] = row[min(max(y0 - 1 + i, 0), Y - 1)];
  }
}

template <>
__device__ __forceinline__ void load10<double>(const double* row, int y0, int Y, bool vec, double* v) {
#pragma unroll
  for (int i = 0; i < 10; ++i) v[i] = row[min(max(y0 - 1 + i, 0), Y - 1)];
}

template <typename Tin> struct DiffT;
template <> struct DiffT<double> {
  // spot_tools/fitting.py:106: float32(max_im) - float32(min_im)
  static __device__ __forceinline__ double diff(double a, double b) { return (double)__fsub_rn(__double2float_rn(a), __double2float_rn(b)); }
  static __device__ __forceinline__ bool nonzero(double a) { return a != 0.0; }
};
template <> struct DiffT<uint16_t> {
  static __device__ __forceinline__ double diff(uint16_t a, uint16_t b) { return (double)((int)a - (int)b); }
  static __device__ __forceinline__ bool nonzero(uint16_t a) { return a != 0; }
};
template <> struct DiffT<float> {
  // numpy: float32 - float32 in float32
  static __device__ __forceinline__ double diff(float a, float b) { return (double)__fsub_rn(a, b); }
  static __device__ __forceinline__ bool nonzero(float a) { return a != 0.f; }
};

// generic (any filt_size) rank value at one voxel
template <typename Tin, bool ISMAX>
__device__ Tin rank_at(const Tin* vol, int z, int x, int y, int Z, int X, int Y, int s1, int s2) {
  Tin best = vol[((long long)z * X + x) * Y + y];
  for (int dz = -s1; dz <= s2; ++dz) {
    const int zz = z + dz; if (zz < 0 || zz >= Z) continue;
    for (int dx = -s1; dx <= s2; ++dx) {
      const int xx = x + dx; if (xx < 0 || xx >= X) continue;
      const Tin* row = vol + ((long long)zz * X + xx) * Y;
      for (int dy = -s1; dy <= s2; ++dy) {
        const int yy = y + dy; if (yy < 0 || yy >= Y) continue;
        const Tin v = row[yy];
        if (ISMAX ? (v > best) : (v < best)) best = v;
      }
    }
  }
  return best;
}

template <typename Tin, int VARIANT>
__global__ void __launch_bounds__(FLAG_THREADS)
k_flags(const Tin* __restrict__ fg, const Tin* __restrict__ bg, SeedDims d, uint8_t* __restrict__ bits,
        int* __restrict__ block_counts) {
  const long long cidx = (long long)blockIdx.x * FLAG_THREADS + threadIdx.x;
  uint32_t mask = 0;
  if (cidx < d.n_chunks) {
    const long long row = cidx / d.cpr;
    const int y0 = (int)(cidx % d.cpr) * 8;
    const int z = (int)(row / d.X), x = (int)(row % d.X);
    const int nv = min(8, d.Y - y0);
    const bool vec = (d.Y % 8 == 0);
    if (d.fs == 3) {
      Tin cmax[10], cmin[10], fc[10], bc[10];
#pragma unroll
      for (int i = 0; i < 10; ++i) { cmax[i] = 0; cmin[i] = 0; }
      bool first = true;
      for (int dz = -1; dz <= 1; ++dz) {
        const int zz = z + dz; if (zz < 0 || zz >= d.Z) continue;
        for (int dx = -1; dx <= 1; ++dx) {
          const int xx = x + dx; if (xx < 0 || xx >= d.X) continue;
          const long long ro = ((long long)zz * d.X + xx) * d.Y;
          Tin vf[10], vb[10];
          load10<Tin>(fg + ro, y0, d.Y, vec, vf);
          load10<Tin>(bg + ro, y0, d.Y, vec, vb);
#pragma unroll
          for (int i = 0; i < 10; ++i) {
            cmax[i] = first ? vf[i] : (vf[i] > cmax[i] ? vf[i] : cmax[i]);
            cmin[i] = first ? vb[i] : (vb[i] < cmin[i] ? vb[i] : cmin[i]);
          }
          if (dz == 0 && dx == 0) {
#pragma unroll
            for (int i = 0; i < 10; ++i) { fc[i] = vf[i]; bc[i] = vb[i]; }
          }
          first = false;
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (i >= nv) break;
        Tin mx = cmax[i + 1]; if (cmax[i] > mx) mx = cmax[i]; if (cmax[i + 2] > mx) mx = cmax[i + 2];
        Tin mn = cmin[i + 1]; if (cmin[i] < mn) mn = cmin[i]; if (cmin[i + 2] < mn) mn = cmin[i + 2];
        const Tin f = fc[i + 1], b = bc[i + 1];
        bool keep = (mx == f) && (mn != b);
        if (VARIANT == 0) {
          keep = keep && (DiffT<Tin>::diff(f, b) >= d.h_min);
        } else {
          keep = keep && DiffT<Tin>::nonzero(mn) && (DiffT<Tin>::diff(f, mn) >= d.h_min);
        }
        if (keep) mask |= (1u << i);
      }
    } else {
      for (int i = 0; i < nv; ++i) {
        const int y = y0 + i;
        const long long li = ((long long)z * d.X + x) * d.Y + y;
        const Tin f = fg[li], b = bg[li];
        const Tin mx = rank_at<Tin, true>(fg, z, x, y, d.Z, d.X, d.Y, d.s1, d.s2);
        if (!(mx == f)) continue;
        const Tin mn = rank_at<Tin, false>(bg, z, x, y, d.Z, d.X, d.Y, d.s1, d.s2);
        bool keep = (mn != b);
        if (VARIANT == 0) keep = keep && (DiffT<Tin>::diff(f, b) >= d.h_min);
        else keep = keep && DiffT<Tin>::nonzero(mn) && (DiffT<Tin>::diff(f, mn) >= d.h_min);
        if (keep) mask |= (1u << i);
      }
    }
    if (VARIANT == 0 && d.edge_on) {
      // remove_edge_points (spot_tools/fitting.py:156-165): d <= c <= size - d, inclusive
      if (z < d.loZ || z > d.hiZ || x < d.lo || x > d.hiX) mask = 0;
      else {
#pragma unroll
        for (int i = 0; i < 8; ++i) { const int y = y0 + i; if (y < d.lo || y > d.hiY) mask &= ~(1u << i); }
      }
    }
    bits[cidx] = (uint8_t)mask;
  }
  // block count (warp popc reduce + shared)
  __shared__ int wsum[FLAG_THREADS / 32];
  int c = __popc(mask);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
#pragma unroll
    for (int i = 0; i < FLAG_THREADS / 32; ++i) t += wsum[i];
    block_counts[blockIdx.x] = t;
  }
}

// ------------------------------------------------------------------------------------------
// uint16 fast path of the candidate mask (filt_size 3, Y % 8 == 0): packed 16-bit SIMD min/max.
// A thread owns one 8-voxel chunk (four u16x2 words) of one (x) row and marches along z, so every
// plane's 3-row (x-1, x, x+1) max / min is formed once and reused for three output planes; the 3-tap
// max / min along y needs one voxel from each neighbouring chunk, taken from the neighbouring lanes
// by shuffle (a warp loads 32 consecutive chunks of one row and produces the inner 30).  Rows, planes
// and chunks outside the stack contribute the identity (0 for max, 0xffff for min), which is what
// scipy's `reflect` amounts to for a rank filter.  ~25 instructions per voxel instead of ~215.
// ------------------------------------------------------------------------------------------
constexpr int FL_ROWS = 8;          // x rows per CTA (one warp each)
constexpr int FL_OUT = 30;          // chunks produced per warp

struct W4 { uint32_t w[4]; };

__device__ __forceinline__ W4 ld_chunk(const uint16_t* vol, long long row_off, int cy, bool ok, uint32_t ident) {
  W4 r;
  if (ok) {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(vol + row_off + (long long)cy * 8));
    r.w[0] = q.x; r.w[1] = q.y; r.w[2] = q.z; r.w[3] = q.w;
  } else {
    r.w[0] = r.w[1] = r.w[2] = r.w[3] = ident;
  }
  return r;
}
__device__ __forceinline__ W4 max3(const W4& a, const W4& b, const W4& c) {
  W4 r;
#pragma unroll
  for (int i = 0; i < 4; ++i) r.w[i] = __vimax3_u16x2(a.w[i], b.w[i], c.w[i]);
  return r;
}
__device__ __forceinline__ W4 min3(const W4& a, const W4& b, const W4& c) {
  W4 r;
#pragma unroll
  for (int i = 0; i < 4; ++i) r.w[i] = __vimin3_u16x2(a.w[i], b.w[i], c.w[i]);
  return r;
}
// 3-tap max / min along y inside the chunk; lh / rh = neighbouring chunks' last / first word
template <bool ISMAX>
__device__ __forceinline__ W4 horiz3(const W4& v, uint32_t left_w3, uint32_t right_w0) {
  W4 r;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t prev = (i == 0) ? left_w3 : v.w[i - 1];
    const uint32_t next = (i == 3) ? right_w0 : v.w[i + 1];
    const uint32_t l = __byte_perm(prev, v.w[i], 0x5432);     // (prev.hi, cur.lo)
    const uint32_t rr = __byte_perm(v.w[i], next, 0x5432);    // (cur.hi, next.lo)
    r.w[i] = ISMAX ? __vimax3_u16x2(l, v.w[i], rr) : __vimin3_u16x2(l, v.w[i], rr);
  }
  return r;
}

template <int VARIANT>
__global__ void __launch_bounds__(FL_ROWS * 32)
k_flags_u16(const uint16_t* __restrict__ fg, const uint16_t* __restrict__ bg, SeedDims d, int zseg,
            uint8_t* __restrict__ bits) {
  const int lane = threadIdx.x & 31;
  const int x = blockIdx.y * FL_ROWS + (threadIdx.x >> 5);
  const int cy = blockIdx.x * FL_OUT + lane - 1;                 // this lane's chunk (may be -1 or >= cpr)
  const int zs = blockIdx.z * zseg, ze = min(d.Z, zs + zseg);
  if (x >= d.X) return;                                          // whole warp
  const bool cok = (cy >= 0) && (cy < d.cpr);
  const bool produce = cok && lane >= 1 && lane <= FL_OUT;
  const long long plane = (long long)d.X * d.Y;
  const bool xm = x > 0, xp = x + 1 < d.X;

  W4 pf0, pf1, pb0, pb1, cf1, cb1;       // 3-row max / min of planes z-1 (0) and z (1); centre row of plane z
  auto load_plane = [&](int zz, W4& pf, W4& pb, W4& cf, W4& cb) {
    if (zz < 0 || zz >= d.Z) {
#pragma unroll
      for (int i = 0; i < 4; ++i) { pf.w[i] = 0u; pb.w[i] = 0xffffffffu; cf.w[i] = 0u; cb.w[i] = 0u; }
      return;
    }
    const long long r0 = (long long)zz * plane + (long long)x * d.Y;
    cf = ld_chunk(fg, r0, cy, cok, 0u);
    cb = ld_chunk(bg, r0, cy, cok, 0xffffffffu);
    const W4 fa = ld_chunk(fg, r0 - d.Y, cy, cok && xm, 0u), fc = ld_chunk(fg, r0 + d.Y, cy, cok && xp, 0u);
    const W4 ba = ld_chunk(bg, r0 - d.Y, cy, cok && xm, 0xffffffffu), bc = ld_chunk(bg, r0 + d.Y, cy, cok && xp, 0xffffffffu);
    pf = max3(fa, cf, fc);
    pb = min3(ba, cb, bc);
  };
  { W4 t0, t1; load_plane(zs - 1, pf0, pb0, t0, t1); }
  load_plane(zs, pf1, pb1, cf1, cb1);
  for (int z = zs; z < ze; ++z) {
    W4 pf2, pb2, cf2, cb2;
    load_plane(z + 1, pf2, pb2, cf2, cb2);
    const W4 vf = max3(pf0, pf1, pf2), vb = min3(pb0, pb1, pb2);
    const uint32_t fl = __shfl_up_sync(0xffffffffu, vf.w[3], 1), fr = __shfl_down_sync(0xffffffffu, vf.w[0], 1);
    const uint32_t bl = __shfl_up_sync(0xffffffffu, vb.w[3], 1), br = __shfl_down_sync(0xffffffffu, vb.w[0], 1);
    if (produce) {
      const W4 mx = horiz3<true>(vf, fl, fr), mn = horiz3<false>(vb, bl, br);
      uint32_t mask = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t eqf = mx.w[i] ^ cf1.w[i];          // half == 0  <=>  local max (plateaus count)
        const uint32_t neb = mn.w[i] ^ cb1.w[i];          // half != 0  <=>  not a local min of the background
        if ((eqf & 0xffffu) == 0 && (neb & 0xffffu) != 0) mask |= 1u << (2 * i);
        if ((eqf >> 16) == 0 && (neb >> 16) != 0) mask |= 2u << (2 * i);
      }
      if (mask) {
        // the (rare) survivors: threshold on exact integers, then the inclusive edge filter
        uint32_t m2 = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (!(mask & (1u << i))) continue;
          const int f = (int)((cf1.w[i >> 1] >> (16 * (i & 1))) & 0xffffu);
          const int b = (int)((cb1.w[i >> 1] >> (16 * (i & 1))) & 0xffffu);
          const int m = (int)((mn.w[i >> 1] >> (16 * (i & 1))) & 0xffffu);
          bool keep;
          if (VARIANT == 0) keep = (double)(f - b) >= d.h_min;
          else keep = (m != 0) && ((double)(f - m) >= d.h_min);
          if (keep) m2 |= 1u << i;
        }
        mask = m2;
        if (VARIANT == 0 && d.edge_on && mask) {
          if (z < d.loZ || z > d.hiZ || x < d.lo || x > d.hiX) mask = 0;
          else {
#pragma unroll
            for (int i = 0; i < 8; ++i) { const int y = cy * 8 + i; if (y < d.lo || y > d.hiY) mask &= ~(1u << i); }
          }
        }
      }
      bits[((long long)z * d.X + x) * d.cpr + cy] = (uint8_t)mask;
    }
    pf0 = pf1; pb0 = pb1; pf1 = pf2; pb1 = pb2; cf1 = cf2; cb1 = cb2;
  }
}

// per-block candidate counts (block = FLAG_THREADS consecutive chunks in C order) from the mask bytes
__global__ void __launch_bounds__(FLAG_THREADS) k_count_bits(const uint8_t* __restrict__ bits, long long n_chunks,
                                                             int* __restrict__ block_counts) {
  const long long cidx = (long long)blockIdx.x * FLAG_THREADS + threadIdx.x;
  int c = (cidx < n_chunks) ? __popc((unsigned)bits[cidx]) : 0;
  __shared__ int wsum[FLAG_THREADS / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
#pragma unroll
    for (int i = 0; i < FLAG_THREADS / 32; ++i) t += wsum[i];
    block_counts[blockIdx.x] = t;
  }
}

// Exclusive scan of the per-block counts by one 1024-thread block (n is ~1e5: one pass of
// contiguous per-thread segments + a block scan of the segment sums).  offsets[n] = total.
__global__ void __launch_bounds__(1024) k_scan_counts(const int* __restrict__ counts, long long* __restrict__ offsets, int n) {
  __shared__ long long part[1024];
  const int t = threadIdx.x;
  const int per = (n + 1023) / 1024;
  const int b = t * per, e = min(n, b + per);
  long long s = 0;
  for (int i = b; i < e; ++i) s += counts[i];
  part[t] = s;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {   // Hillis-Steele inclusive scan
    long long v = (t >= o) ? part[t - o] : 0;
    __syncthreads();
    part[t] += v;
    __syncthreads();
  }
  long long run = part[t] - s;
  for (int i = b; i < e; ++i) { offsets[i] = run; run += counts[i]; }
  if (t == 1023) offsets[n] = part[1023];
}

template <typename Tin, int VARIANT>
__global__ void __launch_bounds__(FLAG_THREADS)
k_emit(const Tin* __restrict__ fg, const Tin* __restrict__ bg, SeedDims d, const uint8_t* __restrict__ bits,
       const long long* __restrict__ offsets, int32_t* __restrict__ out_zxy, float* __restrict__ out_h) {
  const long long cidx = (long long)blockIdx.x * FLAG_THREADS + threadIdx.x;
  const uint32_t mask = (cidx < d.n_chunks) ? bits[cidx] : 0u;
  const int c = __popc(mask);
  // block-exclusive prefix of c
  __shared__ int wsum[FLAG_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  int wbase = 0;
  for (int i = 0; i < warp; ++i) wbase += wsum[i];
  if (mask == 0) return;
  long long pos = offsets[blockIdx.x] + wbase + (incl - c);
  const long long row = cidx / d.cpr;
  const int y0 = (int)(cidx % d.cpr) * 8;
  const int z = (int)(row / d.X), x = (int)(row % d.X);
  for (int i = 0; i < 8; ++i) {
    if (!(mask & (1u << i))) continue;
    const int y = y0 + i;
    const long long li = ((long long)z * d.X + x) * d.Y + y;
    float h;
    if (VARIANT == 0) h = (float)DiffT<Tin>::diff(fg[li], bg[li]);
    else h = (float)DiffT<Tin>::diff(fg[li], rank_at<Tin, false>(bg, z, x, y, d.Z, d.X, d.Y, d.s1, d.s2));
    out_zxy[3 * pos] = z; out_zxy[3 * pos + 1] = x; out_zxy[3 * pos + 2] = y;
    out_h[pos] = h;
    ++pos;
  }
}

template <typename Tin>
int seed_flags(const Tin* fg, const Tin* bg, const SeedDims& d, int variant, uint8_t* bits, int* counts,
               long long* offsets, cudaStream_t st) {
  const int nb = d.n_blocks;
  if constexpr (sizeof(Tin) == 2) {
    if (d.fs == 3 && d.Y % 8 == 0) {
      // z is split so that the grid has a few waves of CTAs even for thin stacks
      const int gx = (d.cpr + FL_OUT - 1) / FL_OUT, gy = (d.X + FL_ROWS - 1) / FL_ROWS;
      int nseg = 1;
      while ((long long)gx * gy * nseg < 148 * 16 && nseg < d.Z) nseg *= 2;
      const int zseg = (d.Z + nseg - 1) / nseg;
      dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)((d.Z + zseg - 1) / zseg));
      const uint16_t* f16 = reinterpret_cast<const uint16_t*>(fg);
      const uint16_t* b16 = reinterpret_cast<const uint16_t*>(bg);
      if (variant == 0) k_flags_u16<0><<<grid, FL_ROWS * 32, 0, st>>>(f16, b16, d, zseg, bits);
      else k_flags_u16<1><<<grid, FL_ROWS * 32, 0, st>>>(f16, b16, d, zseg, bits);
      IA3_LAUNCH_CHECK();
      k_count_bits<<<nb, FLAG_THREADS, 0, st>>>(bits, d.n_chunks, counts);
      IA3_LAUNCH_CHECK();
      k_scan_counts<<<1, 1024, 0, st>>>(counts, offsets, nb);
      IA3_LAUNCH_CHECK();
      return 0;
    }
  }
  if (variant == 0) k_flags<Tin, 0><<<nb, FLAG_THREADS, 0, st>>>(fg, bg, d, bits, counts);
  else k_flags<Tin, 1><<<nb, FLAG_THREADS, 0, st>>>(fg, bg, d, bits, counts);
  IA3_LAUNCH_CHECK();
  k_scan_counts<<<1, 1024, 0, st>>>(counts, offsets, nb);
  IA3_LAUNCH_CHECK();
  return 0;
}
template <typename Tin>
int seed_emit(const Tin* fg, const Tin* bg, const SeedDims& d, int variant, const uint8_t* bits,
              const long long* offsets, int32_t* out_zxy, float* out_h, cudaStream_t st) {
  const int nb = d.n_blocks;
  if (variant == 0) k_emit<Tin, 0><<<nb, FLAG_THREADS, 0, st>>>(fg, bg, d, bits, offsets, out_zxy, out_h);
  else k_emit<Tin, 1><<<nb, FLAG_THREADS, 0, st>>>(fg, bg, d, bits, offsets, out_zxy, out_h);
  IA3_LAUNCH_CHECK();
  return 0;
}
template int seed_flags<uint16_t>(const uint16_t*, const uint16_t*, const SeedDims&, int, uint8_t*, int*, long long*, cudaStream_t);
template int seed_flags<float>(const float*, const float*, const SeedDims&, int, uint8_t*, int*, long long*, cudaStream_t);
template int seed_flags<double>(const double*, const double*, const SeedDims&, int, uint8_t*, int*, long long*, cudaStream_t);
template int seed_emit<uint16_t>(const uint16_t*, const uint16_t*, const SeedDims&, int, const uint8_t*, const long long*, int32_t*, float*, cudaStream_t);
template int seed_emit<float>(const float*, const float*, const SeedDims&, int, const uint8_t*, const long long*, int32_t*, float*, cudaStream_t);
template int seed_emit<double>(const double*, const double*, const SeedDims&, int, const uint8_t*, const long long*, int32_t*, float*, cudaStream_t);

int seed_flag_threads() { return FLAG_THREADS; }

}  // namespace ia3

// ------------------------------------------------------------------------------------------
// Local background of a box of a uint16 stack = mode of its intensity histogram
// (io_tools/load.py:642-686 find_image_background, called per spot by fit_fov_image's
// normalize_local, spot_tools/fitting.py:246-258, on a (2*fit_radius*2+1)^3 crop).
//   counts  np.histogram(crop, bins=arange(first, last, bin)): bin b = [first + b*bin, first + (b+1)*bin),
//           the last bin also holds its right edge, values beyond it are dropped
//   peaks   scipy.signal.find_peaks = strict local maxima, plateaus reported at their midpoint, never the
//           first / last sample; the loop `height = size/50 / 2^k, k = 1..` accepts the first k <= max_iter
//           with a peak of that height, i.e. succeeds iff the highest peak >= size/50/2^max_iter, and then
//           takes the highest peak (lowest index among equals) -> centre of its bin
//   else    np.nanmedian(crop)
// One CTA per box; the histogram lives in shared memory.
// ------------------------------------------------------------------------------------------
namespace ia3 {

constexpr int BG_THREADS = 256;

__device__ __forceinline__ void block_reduce_best(unsigned long long& key, unsigned long long* sh) {
  // max over the block of a packed (height << 32 | ~index) key
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const unsigned long long v = __shfl_xor_sync(0xffffffffu, key, o); if (v > key) key = v; }
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = key;
  __syncthreads();
  if (threadIdx.x < 32) {
    unsigned long long v = (threadIdx.x < BG_THREADS / 32) ? sh[threadIdx.x] : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o); if (w > v) v = w; }
    if (threadIdx.x == 0) sh[0] = v;
  }
  __syncthreads();
  key = sh[0];
  __syncthreads();
}

// k-th smallest (0-based) value of the box by two 256-bin radix passes; every thread returns it
__device__ unsigned box_select(const uint16_t* __restrict__ im, int X, int Y, const int* b, long long nvox, long long k,
                               unsigned* cnt /*256, shared*/) {
  const int ex = b[3] - b[2], ey = b[5] - b[4];
  unsigned prefix = 0;
  for (int pass = 0; pass < 2; ++pass) {
    for (int i = threadIdx.x; i < 256; i += BG_THREADS) cnt[i] = 0;
    __syncthreads();
    for (long long t = threadIdx.x; t < nvox; t += BG_THREADS) {
      const int dy = (int)(t % ey), dx = (int)((t / ey) % ex), dz = (int)(t / ((long long)ey * ex));
      const unsigned v = im[((long long)(b[0] + dz) * X + (b[2] + dx)) * Y + (b[4] + dy)];
      if (pass == 0) atomicAdd(&cnt[v >> 8], 1u);
      else if ((v >> 8) == prefix) atomicAdd(&cnt[v & 255u], 1u);
    }
    __syncthreads();
    unsigned digit = 0;
    long long acc = 0;
    for (int i = 0; i < 256; ++i) {          // every thread scans the same 256 counters
      if (acc + cnt[i] > k) { digit = i; break; }
      acc += cnt[i];
    }
    k -= acc;
    prefix = (pass == 0) ? digit : ((prefix << 8) | digit);
    __syncthreads();
  }
  return prefix;
}

// whole-volume histogram for one large box (normalize_background): many CTAs, shared-memory
// histograms flushed into a global one; k_box_background then starts from that histogram
__global__ void __launch_bounds__(BG_THREADS)
k_hist_accum(const uint16_t* __restrict__ im, long long nvox, int first, int bin, int nbins, unsigned* __restrict__ ghist) {
  extern __shared__ unsigned hist[];
  for (int i = threadIdx.x; i < nbins; i += BG_THREADS) hist[i] = 0;
  __syncthreads();
  const int last_edge = first + nbins * bin;
  const long long stride = (long long)gridDim.x * BG_THREADS;
  for (long long t = (long long)blockIdx.x * BG_THREADS + threadIdx.x; t < nvox; t += stride) {
    const int v = im[t];
    if (v < first || v > last_edge) continue;
    atomicAdd(&hist[(v == last_edge) ? nbins - 1 : (v - first) / bin], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nbins; i += BG_THREADS) if (hist[i]) atomicAdd(&ghist[i], hist[i]);
}

__global__ void __launch_bounds__(BG_THREADS)
k_box_background(const uint16_t* __restrict__ im, int X, int Y, const int* __restrict__ boxes, int first, int bin, int nbins,
                 int max_iter, const unsigned* __restrict__ ghist, double* __restrict__ out) {
  extern __shared__ unsigned hist[];                 // nbins counters
  __shared__ unsigned long long red[BG_THREADS / 32];
  __shared__ unsigned cnt256[256];
  const int* b = boxes + 6 * (long long)blockIdx.x;  // z0 z1 x0 x1 y0 y1 (half open)
  const int ez = b[1] - b[0], ex = b[3] - b[2], ey = b[5] - b[4];
  const long long nvox = (ez > 0 && ex > 0 && ey > 0) ? (long long)ez * ex * ey : 0;
  for (int i = threadIdx.x; i < nbins; i += BG_THREADS) hist[i] = ghist ? ghist[i] : 0u;
  __syncthreads();
  const int last_edge = first + nbins * bin;
  if (!ghist) {
    for (long long t = threadIdx.x; t < nvox; t += BG_THREADS) {
      const int dy = (int)(t % ey), dx = (int)((t / ey) % ex), dz = (int)(t / ((long long)ey * ex));
      const int v = im[((long long)(b[0] + dz) * X + (b[2] + dx)) * Y + (b[4] + dy)];
      if (v < first || v > last_edge) continue;
      const int bi = (v == last_edge) ? nbins - 1 : (v - first) / bin;
      atomicAdd(&hist[bi], 1u);
    }
    __syncthreads();
  }
  // local maxima with plateaus (scipy _local_maxima_1d)
  unsigned long long best = 0ull;
  for (int i = 1 + threadIdx.x; i < nbins - 1; i += BG_THREADS) {
    const unsigned h = hist[i];
    if (h == 0 || !(hist[i - 1] < h)) continue;
    int j = i + 1;
    while (j < nbins - 1 && hist[j] == h) ++j;
    if (hist[j] < h) {
      const unsigned mid = (unsigned)((i + j - 1) / 2);
      const unsigned long long key = ((unsigned long long)h << 32) | (unsigned long long)(0xffffffffu - mid);
      if (key > best) best = key;
    }
  }
  block_reduce_best(best, red);
  const unsigned hmax = (unsigned)(best >> 32);
  double hmin = (double)nvox / 50.0;
  for (int k = 0; k < max_iter; ++k) hmin *= 0.5;   // the last height tried that still counts
  double result;
  if (best != 0ull && (double)hmax >= hmin) {
    const unsigned p = 0xffffffffu - (unsigned)(best & 0xffffffffull);
    result = (double)((long long)first * 2 + (2ll * p + 1) * bin) / 2.0;     // (bins[p] + bins[p+1]) / 2
  } else if (nvox == 0) {
    result = nan("");
  } else {
    const unsigned lo = box_select(im, X, Y, b, nvox, (nvox - 1) / 2, cnt256);
    const unsigned hi = (nvox % 2) ? lo : box_select(im, X, Y, b, nvox, nvox / 2, cnt256);
    result = ((double)lo + (double)hi) / 2.0;                                  // np.nanmedian
  }
  if (threadIdx.x == 0) out[blockIdx.x] = result;
}

int box_background(const uint16_t* im, int X, int Y, const int* d_boxes, long long n, int first, int bin, int nbins, int max_iter,
                   double* d_out, cudaStream_t st) {
  if (n == 0) return 0;
  const size_t smem = (size_t)nbins * sizeof(unsigned);
  if (smem > 200 * 1024) { set_error("background histogram does not fit in shared memory (bin_size too small)"); return -1; }
  static std::once_flag once;
  static cudaError_t once_err = cudaSuccess;
  std::call_once(once, [] { once_err = cudaFuncSetAttribute(k_box_background, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); });
  IA3_CUDA(once_err);
  k_box_background<<<(unsigned)n, BG_THREADS, smem, st>>>(im, X, Y, d_boxes, first, bin, nbins, max_iter, nullptr, d_out);
  IA3_LAUNCH_CHECK();
  return 0;
}

// one box covering the whole (Z, X, Y) volume
int volume_background(const uint16_t* im, int Z, int X, int Y, const int* d_box, int first, int bin, int nbins, int max_iter,
                      unsigned* d_ghist, double* d_out, cudaStream_t st) {
  const size_t smem = (size_t)nbins * sizeof(unsigned);
  if (smem > 200 * 1024) { set_error("background histogram does not fit in shared memory (bin_size too small)"); return -1; }
  static std::once_flag once;
  static cudaError_t once_err = cudaSuccess;
  std::call_once(once, [] {
    once_err = cudaFuncSetAttribute(k_hist_accum, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (once_err == cudaSuccess) once_err = cudaFuncSetAttribute(k_box_background, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  });
  IA3_CUDA(once_err);
  IA3_CUDA(cudaMemsetAsync(d_ghist, 0, smem, st));
  k_hist_accum<<<148 * 4, BG_THREADS, smem, st>>>(im, (long long)Z * X * Y, first, bin, nbins, d_ghist);
  IA3_LAUNCH_CHECK();
  k_box_background<<<1, BG_THREADS, smem, st>>>(im, X, Y, d_box, first, bin, nbins, max_iter, d_ghist, d_out);
  IA3_LAUNCH_CHECK();
  return 0;
}

}  // namespace ia3
