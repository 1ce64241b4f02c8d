// Fit stage kernels (sm_100a): one warp per spot.
//
//   k_init_window  float64 work volume <- image, only for voxels inside some seed's window
//   k_voronoi      firstfit membership of every window voxel (nearest seed), tie detection
//   k_fit          GaussianFit.fit() for one spot per warp: window voxels compacted into shared
//                  memory, initial guess by warp-wide selection, lmder-faithful LM (lm_core.h)
//                  with the per-voxel model / Jacobian strided over the lanes, J^T J, J^T f and
//                  |f|^2 combined by xor-shuffle butterflies (so every lane holds bit-identical
//                  sums), 10x10 algebra in FP64 on lane 0 with its state in shared memory
//   k_subtract     firstfit's "im_subtr[window] -= reconstruction", one dependency level at a time
//
// Sequential semantics of the reference (seed order, in-place im_add updates; Fitting_v4.py:
// 651-675) are preserved by launching one dependency level at a time: two seeds are in the same
// level only if their windows are disjoint, and a seed's level is above that of every
// lower-index seed it overlaps (levels are computed on the host, capi.cu).
#include <algorithm>
#include <mutex>
#include "ia3_device.h"
#include "fit_kernels.h"
#include "fit_spot.h"

namespace ia3 {

constexpr int WARPS = 4;
// One warp (= one spot) per CTA in the fit kernels.  A long fit keeps its CTA resident; with one warp
// per CTA it pins 4.7 K registers and 15 KB of shared memory of its SM instead of four times that, so
// the seed kernels of the other in-flight stacks keep running next to it.
constexpr int FIT_WARPS = 1;
#ifndef IA3_FIT_MINBLOCKS
#define IA3_FIT_MINBLOCKS 8      // lower bound for ptxas; k_fit needs 146 registers (12 spots per SM) since the
#endif                           // normal-equation sums moved to the tensor cores (255 before)
constexpr unsigned FULL = 0xffffffffu;

struct WarpExec {
  static constexpr int W = 32;
  __device__ __forceinline__ int lane() const { return threadIdx.x & 31; }
  __device__ __forceinline__ void sync() const { __syncwarp(); }
  __device__ __forceinline__ double allsum(double v) const {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
  }
  __device__ __forceinline__ int allsum_int(int v) const {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
  }
  __device__ __forceinline__ double bcast(double v, int src) const { return __shfl_sync(FULL, v, src); }
  __device__ __forceinline__ double allmax(double v) const {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
    return v;
  }
  __device__ __forceinline__ void argmin(double& v, int& k) const {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(FULL, v, o);
      const int ok = __shfl_xor_sync(FULL, k, o);
      if (ov < v || (ov == v && ok < k)) { v = ov; k = ok; }
    }
  }
  // lane-private "already selected" bits of select10: slot i = this lane's i-th voxel (K <= 2048)
  struct TakenMask {
    unsigned long long bits;
    __device__ __forceinline__ void clear() { bits = 0ull; }
    __device__ __forceinline__ bool test(int i) const { return (bits >> i) & 1ull; }
    __device__ __forceinline__ void set(int i) { bits |= 1ull << i; }
  };
  // Butterfly reduce-scatter: N per-lane partial sums -> N totals with ~N + 2 log N shuffles instead
  // of 5 N.  At the step with lane offset O a lane keeps the even (bit clear) or odd (bit set) entry of
  // every pair and sends the other one to its partner, halving the live array; after 5 steps entry f of
  // lane l is the total of index 32 f + bitreverse5(l).  Fixed tree order -> deterministic sums.
  template <int N, int O>
  __device__ __forceinline__ void rs_step(double* v) const {
    const bool up = (threadIdx.x & O) != 0;
#pragma unroll
    for (int i = 0; i < (N + 1) / 2; ++i) {
      const double a = v[2 * i];
      const double b = (2 * i + 1 < N) ? v[2 * i + 1] : 0.0;
      const double send = up ? a : b;
      const double keep = up ? b : a;
      v[i] = keep + __shfl_xor_sync(FULL, send, O);
    }
  }
  template <int N>
  __device__ __forceinline__ void reduce_store(double (&v)[N], double* out) const {
    constexpr int N1 = (N + 1) / 2, N2 = (N1 + 1) / 2, N3 = (N2 + 1) / 2, N4 = (N3 + 1) / 2, N5 = (N4 + 1) / 2;
    rs_step<N, 16>(v);
    rs_step<N1, 8>(v);
    rs_step<N2, 4>(v);
    rs_step<N3, 2>(v);
    rs_step<N4, 1>(v);
    const int base = (int)(__brev((unsigned)(threadIdx.x & 31)) >> 27);
#pragma unroll
    for (int f = 0; f < N5; ++f) {
      const int idx = 32 * f + base;
      if (idx < N) out[idx] = v[f];
    }
  }
};

__device__ __forceinline__ long long vol_index(const FitDev& d, int z, int x, int y) {
  const int b = __ldg(d.brick_tab + ((long long)(z >> 3) * d.nbx + (x >> 3)) * d.nby + (y >> 3));
  return (long long)b * 512 + (((z & 7) << 6) | ((x & 7) << 3) | (y & 7));
}

__device__ __forceinline__ double load_im(const void* im, int dtype, long long idx) {
  if (dtype == 0) return (double)reinterpret_cast<const uint16_t*>(im)[idx];
  if (dtype == 1) return (double)reinterpret_cast<const float*>(im)[idx];
  return reinterpret_cast<const double*>(im)[idx];
}

// packed window voxel: offsets (6 bits each, biased by 32) + window index k (14 bits)
__device__ __forceinline__ uint32_t pack_vox(int dz, int dx, int dy, int k) {
  return (uint32_t)(dz + 32) | ((uint32_t)(dx + 32) << 6) | ((uint32_t)(dy + 32) << 12) | ((uint32_t)k << 18);
}

template <typename T>
struct BallVox {
  int m;
  const uint32_t* pk;
  const double* dv;
  __device__ __forceinline__ void get(int k, T& X0, T& X1, T& X2, T& d) const {
    const uint32_t p = pk[k];
    X0 = (T)((int)(p & 63u) - 32);
    X1 = (T)((int)((p >> 6) & 63u) - 32);
    X2 = (T)((int)((p >> 12) & 63u) - 32);
    d = (T)__double2float_rn(dv[k]);              // self.im = float32(im) (Fitting_v4.py:172)
  }
};

// ------------------------------------------------------------------------------------------
__global__ void k_init_window(FitDev d) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long s = t / d.K;
  if (s >= d.n) return;
  const int k = (int)(t % d.K);
  const int z = (int)d.centers[3 * s] + d.offs[3 * k];
  const int x = (int)d.centers[3 * s + 1] + d.offs[3 * k + 1];
  const int y = (int)d.centers[3 * s + 2] + d.offs[3 * k + 2];
  if (z < 0 || z >= d.Z || x < 0 || x >= d.X || y < 0 || y >= d.Y) return;
  const long long idx = ((long long)z * d.X + x) * d.Y + y;
  d.vol[vol_index(d, z, x, y)] = load_im(d.im, d.im_dtype, idx);   // overlapping windows write the same value
}

// squared distance exactly as scipy's sqeuclidean_distance_double for 3 components:
// s = d0*d0; s += d1*d1; s += d2*d2 (no FMA)
__device__ __forceinline__ double sqdist3(double v0, double v1, double v2, const double* c) {
  const double d0 = __dsub_rn(v0, c[0]), d1 = __dsub_rn(v1, c[1]), d2 = __dsub_rn(v2, c[2]);
  double s = __dmul_rn(d0, d0);
  s = __dadd_rn(s, __dmul_rn(d1, d1));
  s = __dadd_rn(s, __dmul_rn(d2, d2));
  return s;
}

__global__ void __launch_bounds__(WARPS * 32) k_voronoi(FitDev d) {
  const int lane = threadIdx.x & 31;
  const long long s = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5);
  if (s >= d.n) return;
  const double c[3] = {d.centers[3 * s], d.centers[3 * s + 1], d.centers[3 * s + 2]};
  const int ic[3] = {(int)c[0], (int)c[1], (int)c[2]};
  const int nb0 = d.nbr_start[s], nb1 = d.nbr_start[s + 1];
  const bool v4 = (d.fp.personality == 4);
  const int own = v4 ? (int)s : d.own_id[s];
  for (int k0 = 0; k0 < d.K; k0 += 32) {
    const int k = k0 + lane;
    bool member = false, tie = false;
    if (k < d.K) {
      const int z = ic[0] + d.offs[3 * k], x = ic[1] + d.offs[3 * k + 1], y = ic[2] + d.offs[3 * k + 2];
      if (z >= 0 && z < d.Z && x >= 0 && x < d.X && y >= 0 && y < d.Y) {
        const double v0 = (double)z, v1 = (double)x, v2 = (double)y;
        if (v4) {
          // cKDTree.query(k=1): nearest by squared distance; a strictly closer seed takes the voxel,
          // an equally close one makes it a tie that the host resolves with the same tree
          const double dm = sqdist3(v0, v1, v2, c);
          bool closer = false;
          for (int e = nb0; e < nb1; ++e) {
            const int j = d.nbr_idx[e];
            const double cj[3] = {d.centers[3 * j], d.centers[3 * j + 1], d.centers[3 * j + 2]};
            const double dj = sqdist3(v0, v1, v2, cj);
            if (dj < dm) { closer = true; break; }
            if (dj == dm) tie = true;
          }
          member = !closer && !tie;
          tie = tie && !closer;
        } else {
          // cdist (sqrt of the squared distance) + argmin: lowest index wins ties
          double best = sqrt(sqdist3(v0, v1, v2, c));
          int bi = (int)s;
          for (int e = nb0; e < nb1; ++e) {
            const int j = d.nbr_idx[e];
            const double cj[3] = {d.centers[3 * j], d.centers[3 * j + 1], d.centers[3 * j + 2]};
            const double dj = sqrt(sqdist3(v0, v1, v2, cj));
            if (dj < best || (dj == best && j < bi)) { best = dj; bi = j; }
          }
          member = (bi == own);
        }
      }
    }
    const unsigned bal = __ballot_sync(FULL, member);
    if (lane == 0) d.mask[s * d.KW + (k0 >> 5)] = bal;
    if (tie) {
      const int pos = atomicAdd(d.tie_count, 1);
      if (pos < d.tie_cap) { d.tie_spot[pos] = (int)s; d.tie_k[pos] = k; }
    }
  }
}

// ------------------------------------------------------------------------------------------
// One spot, one warp.  Returns FIT_SUSPENDED if the run spent `cap` function evaluations without
// finishing: its LM state is then parked in d.pause_buf and a later launch (resume = true) continues
// it bit-identically -- the window is gathered again (nothing that overlaps it changes meanwhile: the
// other spots of its level do not touch it and the next level waits for it).  Why: a junk seed that runs
// MINPACK to maxfev = 1000 keeps its kernel alive for tens of ms; a stack's kernels occupy one of the 32
// hardware queues for as long, and 32 / (sum of a stack's kernel times) capped the pipeline at
// ~150 stacks/s with the GPU mostly idle.  Suspended spots of ALL stacks in flight are continued
// together by one service stream (capi.cu), so a stack's own launches stay short.
template <typename T>
__device__ __forceinline__ int fit_one(const FitDev& d, int mode, long long s, unsigned char* base, int cap, bool resume) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int K = d.K;
  SpotShared<T>& sh = *reinterpret_cast<SpotShared<T>*>(base);
  double* dv = reinterpret_cast<double*>(base + (sizeof(SpotShared<T>) + 15) / 16 * 16);
  uint32_t* pk = reinterpret_cast<uint32_t*>(dv + K);

  WarpExec ex;
  const double c[3] = {d.centers[3 * s], d.centers[3 * s + 1], d.centers[3 * s + 2]};
  const int ic[3] = {(int)c[0], (int)c[1], (int)c[2]};
  const double origin[3] = {(double)ic[0], (double)ic[1], (double)ic[2]};
  const bool had_rec = (mode == 1) && d.success[s];

  // gather the window (in-image voxels, firstfit: Voronoi members only) in window order
  int m = 0;
  for (int k0 = 0; k0 < K; k0 += 32) {
    const int k = k0 + lane;
    bool use = false;
    long long idx = 0;
    int dz = 0, dx = 0, dy = 0;
    if (k < K) {
      dz = d.offs[3 * k]; dx = d.offs[3 * k + 1]; dy = d.offs[3 * k + 2];
      const int z = ic[0] + dz, x = ic[1] + dx, y = ic[2] + dy;
      use = (z >= 0 && z < d.Z && x >= 0 && x < d.X && y >= 0 && y < d.Y);
      idx = ((long long)z * d.X + x) * d.Y + y;
      if (mode == 0) use = use && ((d.mask[s * d.KW + (k0 >> 5)] >> lane) & 1u);
    }
    const unsigned bal = __ballot_sync(FULL, use);
    if (use) {
      const int pos = m + __popc(bal & ((1u << lane) - 1u));
      double v;
      if (mode == 0) v = load_im(d.im, d.im_dtype, idx);
      else { v = d.vol[vol_index(d, ic[0] + dz, ic[1] + dx, ic[2] + dy)]; if (had_rec) v = d.rec[s * K + k] + v; }   // im_ = im_rec + im_  (:662)
      dv[pos] = v;
      pk[pos] = pack_vox(dz, dx, dy, k);
    }
    m += __popc(bal);
  }
  __syncwarp();

  if (m < NP) {
    // len(p_) > len(im): success = False (Fitting_v4.py:382-383); firstfit stores a NaN row
    if (lane == 0) {
      d.success[s] = 0;
      d.nfev[s] = 0; d.info[s] = 0;
      if (mode == 0) {
        for (int i = 0; i < NOUT; ++i) d.ps[s * NOUT + i] = NAN;
        for (int i = 0; i < NP; ++i) d.p_raw[s * NP + i] = NAN;
      }
    }
    return FIT_DONE;
  }

  FitParams fp = d.fp;
  // also when resuming: the v3 width prior lives in fp.init_wt, which initial_guess derives from the window
  select10(ex, dv, m, false, sh.small10);
  select10(ex, dv, m, true, sh.large10);
  if (lane == 0) initial_guess(fp, sh.small10, sh.large10, d.init_w, sh.x0);
  __syncwarp();

  BallVox<T> vox{m, pk, dv};
  {
    constexpr int NW64 = (int)(sizeof(LMPause) / 8);
    static_assert(sizeof(LMPause) % 8 == 0 && sizeof(LMState) % 8 == 0, "suspended state is copied in 8-byte words");
    int slot = -1, start = LM_START_FRESH;
    if (resume) {
      slot = -d.info[s] - 1;
      const unsigned long long* src = reinterpret_cast<const unsigned long long*>(d.pause_buf + slot);
      unsigned long long* st64 = reinterpret_cast<unsigned long long*>(&sh.st);
      unsigned long long* ag64 = reinterpret_cast<unsigned long long*>(sh.Ag);
      constexpr int NS = (int)(sizeof(LMState) / 8);
      for (int i = lane; i < NW64; i += 32) { if (i < NS) st64[i] = src[i]; else ag64[i - NS] = src[i]; }
      __syncwarp();
      start = LM_START_CONTINUE;
    }
    bool suspended = run_lm<T>(ex, fp, d.lm, c, origin, vox, sh, cap, start);
    if (suspended && slot < 0) {
      if (lane == 0) slot = atomicAdd(d.pause_ctl, 1);
      slot = __shfl_sync(FULL, slot, 0);
      if (slot >= d.pause_slots) suspended = run_lm<T>(ex, fp, d.lm, c, origin, vox, sh, 0, LM_START_CONTINUE);   // no room: finish here
      else if (lane == 0) d.pause_ctl[1 + slot] = (int)s;
    }
    if (suspended) {
      unsigned long long* dst = reinterpret_cast<unsigned long long*>(d.pause_buf + slot);
      const unsigned long long* st64 = reinterpret_cast<const unsigned long long*>(&sh.st);
      const unsigned long long* ag64 = reinterpret_cast<const unsigned long long*>(sh.Ag);
      constexpr int NS = (int)(sizeof(LMState) / 8);
      for (int i = lane; i < NW64; i += 32) dst[i] = (i < NS) ? st64[i] : ag64[i - NS];
      if (lane == 0) d.info[s] = -(slot + 1);
      return FIT_SUSPENDED;
    }
  }
  __shared__ FitResult res_s[FIT_WARPS];
  FitResult& res = res_s[warp];
  finish_fit<T>(ex, fp, c, origin, vox, sh, &res);
  if (lane < NOUT) d.ps[s * NOUT + lane] = res.ps[lane];
  if (lane < NP) d.p_raw[s * NP + lane] = res.p_raw[lane];
  if (lane == 0) { d.success[s] = 1; d.nfev[s] = res.nfev; d.info[s] = res.info; }

  if (mode == 1) {
    // im_rec = get_im(); ims_rec[ic] = im_rec; im_add[window] = im_ - im_rec   (:671-675)
    __shared__ VoxConsts<double> vcd_s[FIT_WARPS];
    VoxConsts<double>& vcd = vcd_s[warp];
    if (lane == 0) {
      build_consts<double>(fp, c, origin, sh.st.x, false, vcd);
    }
    __syncwarp();
    for (int pos = lane; pos < m; pos += 32) {
      const uint32_t p = pk[pos];
      const int dz = (int)(p & 63u) - 32, dx = (int)((p >> 6) & 63u) - 32, dy = (int)((p >> 12) & 63u) - 32;
      const int k = (int)(p >> 18);
      const double f0 = eval_f0<double>(vcd, (double)dz, (double)dx, (double)dy);
      d.rec[s * K + k] = f0;
      d.vol[vol_index(d, ic[0] + dz, ic[1] + dx, ic[2] + dy)] = dv[pos] - f0;
    }
  }
  return FIT_DONE;
}

template <typename T>
__global__ void __launch_bounds__(FIT_WARPS * 32, IA3_FIT_MINBLOCKS) k_fit(FitDev d, int mode, const int* __restrict__ work, long long n_work) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5;
  const long long wi = (long long)blockIdx.x * FIT_WARPS + warp;
  if (wi >= n_work) return;
  const long long s = work ? (long long)work[wi] : wi;
  const size_t per_warp = (sizeof(SpotShared<T>) + 15) / 16 * 16 + ((size_t)d.K * (8 + 4) + 15) / 16 * 16;
  fit_one<T>(d, mode, s, smem_raw + per_warp * warp, d.cap, false);
}

// Continuation of suspended spots of several handles (one spot per CTA; FitDev table and entries in
// device-addressable pinned memory).
__global__ void __launch_bounds__(32, IA3_FIT_MINBLOCKS) k_fit_resume(const FitDev* __restrict__ devs, FitResume* entries, int n, int cap) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  if ((int)blockIdx.x >= n) return;
  __shared__ FitDev d_s;
  const FitResume e = entries[blockIdx.x];
  {
    const int* src = reinterpret_cast<const int*>(devs + e.job);
    int* dst = reinterpret_cast<int*>(&d_s);
    for (int i = threadIdx.x; i < (int)(sizeof(FitDev) / 4); i += 32) dst[i] = src[i];
  }
  __syncwarp();
  const int status = fit_one<double>(d_s, e.mode, (long long)e.spot, smem_raw, cap, true);
  if (threadIdx.x == 0) entries[blockIdx.x].status = status;
}

// firstfit: ims_rec[s] = get_im() over the full clipped window; im_subtr[window] -= im_rec (:629-633)
__global__ void __launch_bounds__(WARPS * 32) k_subtract(FitDev d, const int* __restrict__ work, long long n_work) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long wi = (long long)blockIdx.x * WARPS + warp;
  if (wi >= n_work) return;
  const long long s = work[wi];
  if (!d.success[s]) return;
  __shared__ VoxConsts<double> vcd_s[WARPS];
  VoxConsts<double>& vcd = vcd_s[warp];
  const double c[3] = {d.centers[3 * s], d.centers[3 * s + 1], d.centers[3 * s + 2]};
  const int ic[3] = {(int)c[0], (int)c[1], (int)c[2]};
  if (lane == 0) {
    const double origin[3] = {(double)ic[0], (double)ic[1], (double)ic[2]};
    double x[NP];
    for (int i = 0; i < NP; ++i) x[i] = d.p_raw[s * NP + i];
    build_consts<double>(d.fp, c, origin, x, false, vcd);
  }
  __syncwarp();
  for (int k = lane; k < d.K; k += 32) {
    const int dz = d.offs[3 * k], dx = d.offs[3 * k + 1], dy = d.offs[3 * k + 2];
    const int z = ic[0] + dz, x = ic[1] + dx, y = ic[2] + dy;
    if (z < 0 || z >= d.Z || x < 0 || x >= d.X || y < 0 || y >= d.Y) continue;
    const double f0 = eval_f0<double>(vcd, (double)dz, (double)dx, (double)dy);
    d.rec[s * d.K + k] = f0;
    d.vol[vol_index(d, z, x, y)] -= f0;
  }
}

// ------------------------------------------------------------------------------------------
// Standalone GaussianFit on arbitrary voxel lists (values/coords in global memory).
struct GlobalVox {
  int m;
  const double* values;
  const float* coords;
  __device__ __forceinline__ void get(int k, double& X0, double& X1, double& X2, double& d) const {
    X0 = (double)coords[3 * k]; X1 = (double)coords[3 * k + 1]; X2 = (double)coords[3 * k + 2];
    d = (double)(float)values[k];
  }
};

__global__ void __launch_bounds__(FIT_WARPS * 32) k_generic_fit(GenericFitDev d) {
  __shared__ SpotShared<double> sh_s[FIT_WARPS];
  __shared__ FitResult res_s[FIT_WARPS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long b = (long long)blockIdx.x * FIT_WARPS + warp;
  if (b >= d.n) return;
  SpotShared<double>& sh = sh_s[warp];
  const long long o0 = d.off[b];
  const int m = (int)(d.off[b + 1] - o0);
  if (m < NP) {
    if (lane == 0) {
      d.success[b] = 0; d.nfev[b] = 0; d.info[b] = 0;
      for (int i = 0; i < NOUT; ++i) d.ps[b * NOUT + i] = NAN;
      for (int i = 0; i < NP; ++i) d.p_raw[b * NP + i] = NAN;
    }
    return;
  }
  WarpExec ex;
  FitParams fp = d.fp;
  const double c[3] = {d.centers[3 * b], d.centers[3 * b + 1], d.centers[3 * b + 2]};
  const double origin[3] = {0.0, 0.0, 0.0};
  select10_scratch(ex, d.values + o0, d.tmp + o0, m, false, sh.small10);
  select10_scratch(ex, d.values + o0, d.tmp + o0, m, true, sh.large10);
  if (lane == 0) initial_guess(fp, sh.small10, sh.large10, d.init_w, sh.x0);
  __syncwarp();
  GlobalVox vox{m, d.values + o0, d.coords + 3 * o0};
  run_lm<double>(ex, fp, d.lm, c, origin, vox, sh);
  FitResult& res = res_s[warp];
  finish_fit<double>(ex, fp, c, origin, vox, sh, &res);
  if (lane < NOUT) d.ps[b * NOUT + lane] = res.ps[lane];
  if (lane < NP) d.p_raw[b * NP + lane] = res.p_raw[lane];
  if (lane == 0) { d.success[b] = 1; d.nfev[b] = res.nfev; d.info[b] = res.info; }
  if (d.rec) {
    // finish_fit left sh.vc = constants of the final parameters
    for (int k = lane; k < m; k += 32)
      d.rec[o0 + k] = eval_f0<double>(sh.vc, (double)vox.coords[3 * k], (double)vox.coords[3 * k + 1], (double)vox.coords[3 * k + 2]);
  }
}

// ------------------------------------------------------------------------------------------
int fit_smem_bytes(int K, bool fp32) {
  const size_t shs = fp32 ? sizeof(SpotShared<float>) : sizeof(SpotShared<double>);
  const size_t per_warp = (shs + 15) / 16 * 16 + ((size_t)K * (8 + 4) + 15) / 16 * 16;
  return (int)(per_warp * FIT_WARPS);
}

int launch_init_window(const FitDev& d, cudaStream_t st) {
  const long long total = d.n * d.K;
  if (total == 0) return 0;
  k_init_window<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d);
  IA3_LAUNCH_CHECK();
  return 0;
}

int launch_voronoi(const FitDev& d, cudaStream_t st) {
  if (d.n == 0) return 0;
  k_voronoi<<<(unsigned)((d.n + WARPS - 1) / WARPS), WARPS * 32, 0, st>>>(d);
  IA3_LAUNCH_CHECK();
  return 0;
}

int launch_fit(const FitDev& d, int mode, const int* work, long long n_work, bool fp32, cudaStream_t st) {
  if (n_work == 0) return 0;
  const int smem = fit_smem_bytes(d.K, fp32);
  const unsigned grid = (unsigned)((n_work + FIT_WARPS - 1) / FIT_WARPS);
  // same value from every host thread (the attribute is per-function state); k_fit also has ~2 KB static
  constexpr int kMaxDynSmem = 227 * 1024 - 4096;
  if (smem > kMaxDynSmem) { set_error("radius_fit too large for the shared-memory window"); return -1; }
  // once per process and device: changing a function attribute while an instance of the kernel is
  // running (another stack's sweep) would serialise the streams
  static std::once_flag once;
  static cudaError_t once_err = cudaSuccess;
  std::call_once(once, [] {
    once_err = cudaFuncSetAttribute(k_fit<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
    if (once_err == cudaSuccess)
      once_err = cudaFuncSetAttribute(k_fit<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
    if (once_err == cudaSuccess)
      once_err = cudaFuncSetAttribute(k_fit_resume, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem - 1024);
  });
  IA3_CUDA(once_err);
  if (fp32) k_fit<float><<<grid, FIT_WARPS * 32, smem, st>>>(d, mode, work, n_work);
  else k_fit<double><<<grid, FIT_WARPS * 32, smem, st>>>(d, mode, work, n_work);
  IA3_LAUNCH_CHECK();
  return 0;
}

int launch_fit_resume(const FitDev* devs, FitResume* entries, int n, int cap, int smem_bytes, cudaStream_t st) {
  if (n <= 0) return 0;
  if (smem_bytes > 227 * 1024 - 4096 - 1024) { set_error("radius_fit too large for the shared-memory window"); return -1; }
  k_fit_resume<<<(unsigned)n, 32, smem_bytes, st>>>(devs, entries, n, cap);
  IA3_LAUNCH_CHECK();
  return 0;
}

int launch_subtract(const FitDev& d, const int* work, long long n_work, cudaStream_t st) {
  if (n_work == 0) return 0;
  k_subtract<<<(unsigned)((n_work + WARPS - 1) / WARPS), WARPS * 32, 0, st>>>(d, work, n_work);
  IA3_LAUNCH_CHECK();
  return 0;
}

int launch_generic_fit(const GenericFitDev& d, cudaStream_t st) {
  if (d.n == 0) return 0;
  k_generic_fit<<<(unsigned)((d.n + FIT_WARPS - 1) / FIT_WARPS), FIT_WARPS * 32, 0, st>>>(d);
  IA3_LAUNCH_CHECK();
  return 0;
}

}  // namespace ia3

// ------------------------------------------------------------------------------------------
// Moment ("fast") fit of Fitting_v4: fast_fit_big_image + gfit_fast (External/Fitting_v4.py:433-447,
// 494-556), default path (better_fit=False): per seed, the ball voxels that are closer to this seed
// than to any seed within 2r (argmin of cdist, first index wins), optional recentring on the brightest
// voxel, background = the int(n * bk_f)-th smallest value, weights = (values - background) clipped at 0
// IN THE IMAGE'S DTYPE (a uint16 image wraps around, as in the reference), then weighted mean and
// covariance of the voxel coordinates.  One warp per seed.
// ------------------------------------------------------------------------------------------
namespace ia3 {

__global__ void __launch_bounds__(WARPS * 32) k_moment_fit(MomentDev d) {
  extern __shared__ __align__(16) unsigned char msm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long s = (long long)blockIdx.x * WARPS + warp;
  if (s >= d.n) return;
  const int K = d.K;
  const size_t per_warp = ((size_t)K * 12 + 15) / 16 * 16;
  double* vals = reinterpret_cast<double*>(msm + per_warp * warp);
  uint32_t* pk = reinterpret_cast<uint32_t*>(vals + K);
  const double c[3] = {d.centers[3 * s], d.centers[3 * s + 1], d.centers[3 * s + 2]};
  int ic[3] = {(int)c[0], (int)c[1], (int)c[2]};
  const int nb0 = d.nbr_start[s], nb1 = d.nbr_start[s + 1];
  double* out = d.out + 12 * s;

  for (int pass = 0; pass < (d.recenter ? 2 : 1); ++pass) {
    // gather the member voxels that are inside the image, in offset order
    int m = 0;
    for (int k0 = 0; k0 < K; k0 += 32) {
      const int k = k0 + lane;
      bool use = false;
      int dz = 0, dx = 0, dy = 0;
      long long idx = 0;
      if (k < K) {
        dz = d.offs[3 * k]; dx = d.offs[3 * k + 1]; dy = d.offs[3 * k + 2];
        use = true;
        if (d.avoid) {
          // np.argmin(cdist(centers[common] - centre, offsets), 0) == position of this seed in common
          double best = 0.0; int bj = -1;
          for (int e = nb0; e < nb1; ++e) {
            const int j = d.nbr_idx[e];
            const double r[3] = {d.centers[3 * j] - c[0], d.centers[3 * j + 1] - c[1], d.centers[3 * j + 2] - c[2]};
            const double dist = sqrt(sqdist3((double)dz, (double)dx, (double)dy, r));
            if (bj < 0 || dist < best) { best = dist; bj = j; }
          }
          use = (bj == (int)s);
        }
        const int z = ic[0] + dz, x = ic[1] + dx, y = ic[2] + dy;
        use = use && (z >= 0 && z < d.Z && x >= 0 && x < d.X && y >= 0 && y < d.Y);
        idx = ((long long)z * d.X + x) * d.Y + y;
      }
      const unsigned bal = __ballot_sync(FULL, use);
      if (use) {
        const int pos = m + __popc(bal & ((1u << lane) - 1u));
        vals[pos] = load_im(d.im, d.im_dtype, idx);
        pk[pos] = pack_vox(dz, dx, dy, 0);
      }
      m += __popc(bal);
    }
    __syncwarp();
    if (m == 0) {
      if (lane < 12) out[lane] = NAN;
      return;
    }
    if (d.recenter && pass == 0) {
      // zc, xc, yc = coordinates of the brightest member voxel (first maximum), same offsets again
      double bv = -INFINITY; int bp = 0x7fffffff;
      for (int p = lane; p < m; p += 32) if (vals[p] > bv) { bv = vals[p]; bp = p; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(FULL, bv, o);
        const int op = __shfl_xor_sync(FULL, bp, o);
        if (ov > bv || (ov == bv && op < bp)) { bv = ov; bp = op; }
      }
      const uint32_t p = pk[bp];
      ic[0] += (int)(p & 63u) - 32; ic[1] += (int)((p >> 6) & 63u) - 32; ic[2] += (int)((p >> 12) & 63u) - 32;
      __syncwarp();
      continue;
    }
    // background: the kth smallest value, kth = int(m * bk_f)
    const int kth = (int)((double)m * d.bk_f);
    double bk = 0.0;
    for (int p = lane; p < m; p += 32) {
      const double v = vals[p];
      int less = 0, eq_before = 0;
      for (int q = 0; q < m; ++q) { const double u = vals[q]; less += (u < v); eq_before += (u == v && q < p); }
      if (less + eq_before == kth) bk = v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) bk += __shfl_xor_sync(FULL, bk, o);      // exactly one lane holds it
    // weights in the image's dtype
    double hmax = 0.0, wsum = 0.0;
    for (int p = lane; p < m; p += 32) {
      double w = vals[p] - bk;
      if (d.im_dtype == 0) w = (double)(uint16_t)(int)w;        // uint16 - uint16 wraps around
      else if (w < 0) w = 0.0;
      if (d.im_dtype == 1) w = (double)(float)w;                // float32 arithmetic
      vals[p] = w;
      hmax = fmax(hmax, w);
      wsum += w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { hmax = fmax(hmax, __shfl_xor_sync(FULL, hmax, o)); wsum += __shfl_xor_sync(FULL, wsum, o); }
    __syncwarp();
    double mu[3] = {0, 0, 0};
    for (int p = lane; p < m; p += 32) {
      const double w = vals[p] / wsum;
      vals[p] = w;
      const uint32_t q = pk[p];
      mu[0] += (double)(ic[0] + (int)(q & 63u) - 32) * w;
      mu[1] += (double)(ic[1] + (int)((q >> 6) & 63u) - 32) * w;
      mu[2] += (double)(ic[2] + (int)((q >> 12) & 63u) - 32) * w;
    }
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mu[a] += __shfl_xor_sync(FULL, mu[a], o);
    __syncwarp();
    double cv[6] = {0, 0, 0, 0, 0, 0};     // a b c d e f = zz xx yy zx zy xy
    for (int p = lane; p < m; p += 32) {
      const double w = vals[p];
      const uint32_t q = pk[p];
      const double z = (double)(ic[0] + (int)(q & 63u) - 32) - mu[0];
      const double x = (double)(ic[1] + (int)((q >> 6) & 63u) - 32) - mu[1];
      const double y = (double)(ic[2] + (int)((q >> 12) & 63u) - 32) - mu[2];
      cv[0] += z * z * w; cv[1] += x * x * w; cv[2] += y * y * w; cv[3] += z * x * w; cv[4] += z * y * w; cv[5] += x * y * w;
    }
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) cv[a] += __shfl_xor_sync(FULL, cv[a], o);
    if (lane == 0) {
      out[0] = hmax; out[1] = mu[0]; out[2] = mu[1]; out[3] = mu[2]; out[4] = bk;
      for (int a = 0; a < 6; ++a) out[5 + a] = cv[a];
      out[11] = NAN;
    }
    return;
  }
}

int launch_moment_fit(const MomentDev& d, cudaStream_t st) {
  if (d.n == 0) return 0;
  const size_t per_warp = ((size_t)d.K * 12 + 15) / 16 * 16;
  const size_t smem = per_warp * WARPS;
  if (smem > 200 * 1024) { set_error("radius_fit too large for the moment fit"); return -1; }
  static std::once_flag once;
  static cudaError_t once_err = cudaSuccess;
  std::call_once(once, [] { once_err = cudaFuncSetAttribute(k_moment_fit, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); });
  IA3_CUDA(once_err);
  k_moment_fit<<<(unsigned)((d.n + WARPS - 1) / WARPS), WARPS * 32, smem, st>>>(d);
  IA3_LAUNCH_CHECK();
  return 0;
}

}  // namespace ia3

// ---- small helper kernels used by the C ABI ---------------------------------------------------
namespace ia3 {

__global__ void k_apply_ties(uint32_t* mask, int KW, const int* tie_spot, const int* tie_k, const uint8_t* keep, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !keep[i]) return;
  atomicOr(&mask[(long long)tie_spot[i] * KW + (tie_k[i] >> 5)], 1u << (tie_k[i] & 31));
}
int launch_apply_ties(uint32_t* mask, int KW, const int* tie_spot, const int* tie_k, const uint8_t* keep, int n, cudaStream_t st) {
  if (n == 0) return 0;
  k_apply_ties<<<(n + 255) / 256, 256, 0, st>>>(mask, KW, tie_spot, tie_k, keep, n);
  IA3_LAUNCH_CHECK();
  return 0;
}

// gather (dir = 0: snap <- vol) or scatter (dir = 1: vol_out <- snap) the window voxels of all seeds
__global__ void k_window_copy(FitDev d, double* snap, double* vol_out, int dir) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long s = t / d.K;
  if (s >= d.n) return;
  const int k = (int)(t % d.K);
  const int z = (int)d.centers[3 * s] + d.offs[3 * k];
  const int x = (int)d.centers[3 * s + 1] + d.offs[3 * k + 1];
  const int y = (int)d.centers[3 * s + 2] + d.offs[3 * k + 2];
  if (z < 0 || z >= d.Z || x < 0 || x >= d.X || y < 0 || y >= d.Y) return;
  const long long idx = ((long long)z * d.X + x) * d.Y + y;
  if (dir == 0) snap[t] = d.vol[vol_index(d, z, x, y)];
  else vol_out[idx] = snap[t];          // vol_out is a dense (Z, X, Y) volume
}
int launch_window_copy(const FitDev& d, double* snap, double* vol_out, int dir, cudaStream_t st) {
  const long long total = d.n * d.K;
  if (total == 0) return 0;
  k_window_copy<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d, snap, vol_out, dir);
  IA3_LAUNCH_CHECK();
  return 0;
}

__global__ void k_to_f64(const void* im, int dtype, double* out, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = load_im(im, dtype, i);
}
// GaussianFit.get_im() on arbitrary coordinates: f0 = exp(h - xsigmax/2)
__global__ void k_eval_f0(FitParams fp, const double* p_raw, const double* center, const float* coords, long long m, double* out) {
  __shared__ VoxConsts<double> vc;
  if (threadIdx.x == 0) {
    double x[NP], c[3] = {center[0], center[1], center[2]};
    for (int i = 0; i < NP; ++i) x[i] = p_raw[i];
    const double origin[3] = {0.0, 0.0, 0.0};
    build_consts<double>(fp, c, origin, x, false, vc);
  }
  __syncthreads();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < m; k += stride)
    out[k] = eval_f0<double>(vc, (double)coords[3 * k], (double)coords[3 * k + 1], (double)coords[3 * k + 2]);
}
int launch_eval_f0(const FitParams& fp, const double* p_raw, const double* center, const float* coords, long long m,
                   double* out, cudaStream_t st) {
  if (m == 0) return 0;
  const int blocks = (int)std::min<long long>((m + 255) / 256, 1184);
  k_eval_f0<<<blocks, 256, 0, st>>>(fp, p_raw, center, coords, m, out);
  IA3_LAUNCH_CHECK();
  return 0;
}

int launch_to_f64(const void* im, int dtype, double* out, long long n, cudaStream_t st) {
  if (n == 0) return 0;
  k_to_f64<<<148 * 8, 256, 0, st>>>(im, dtype, out, n);
  IA3_LAUNCH_CHECK();
  return 0;
}

}  // namespace ia3
