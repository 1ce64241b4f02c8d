// Fit stage kernels (sm_100a).  See fit_kernels.h for the engine (tasks, rounds, memo).
//
//   k_cells_* / k_nbr_build / k_brick_*   neighbour pools, dependency counts and the sparse work volume's
//                  brick table, all built on the device from the seed list
//   k_init_window  float64 work volume <- image, only for voxels inside some seed's window
//   k_voronoi      firstfit membership of every window voxel (nearest seed), tie detection
//   k_sched        one thread per seed: issues the tasks whose inputs are final (per-seed dataflow),
//                  routes parked runs to the one-warp or to the team kernel, detects completion
//   k_fit_round<W> GaussianFit.fit() for one spot per CTA of W warps (W = 1: bulk, W = TEAM_WARPS:
//                  continuation of long runs): window voxels compacted into shared memory, initial guess
//                  by warp-wide selection, lmder-faithful LM (lm_core.h) -- per-voxel model / Jacobian
//                  spread over all lanes of the CTA, J^T J / J^T f / |f|^2 as FP64 tensor-core tiles per
//                  batch of 32 voxels added in batch order (bit-identical for every W), 10x10 algebra in
//                  FP64 by warp 0 with its state in shared memory -- then the visit's epilogue: results,
//                  reconstruction, im_add update, convergence test of the seed
#include <algorithm>
#include <mutex>
#include "ia3_device.h"
#include "fit_kernels.h"
#include "fit_spot.h"
#include "lm_warp.h"

namespace ia3 {

constexpr int WARPS = 4;
#ifndef IA3_FIT_MINBLOCKS
#define IA3_FIT_MINBLOCKS 8      // lower bound for ptxas; the one-warp fit kernel needs ~150 registers (12 spots per SM)
#endif
constexpr unsigned FULL = 0xffffffffu;
constexpr int BIG_SWEEP = 1 << 20;

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
  // of 5 N (only the scalar fallback of pass_fused uses it on the device).
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
// Neighbour pools and the brick table, built on the device.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int cell_coord(const CellGrid& g, double v, int a) {
  int c = (int)floor((v - g.lo[a]) / g.cs);
  return c < 0 ? 0 : (c >= g.g[a] ? g.g[a] - 1 : c);
}
__device__ __forceinline__ int cell_of(const CellGrid& g, const double* c) {
  return (cell_coord(g, c[0], 0) * g.g[1] + cell_coord(g, c[1], 1)) * g.g[2] + cell_coord(g, c[2], 2);
}

__global__ void k_cells_count(const double* __restrict__ centers, long long n, CellGrid g, int* __restrict__ cell_cnt) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  atomicAdd(&cell_cnt[cell_of(g, centers + 3 * i)], 1);
}

// exclusive scan by ONE block of 1024 threads (out[n] = total); the cell grid has ~10^5 entries
__global__ void __launch_bounds__(1024) k_scan_excl(const int* __restrict__ in, int* __restrict__ out, long long n) {
  constexpr int PER = 8;
  __shared__ int wsum[32];
  __shared__ int carry_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int carry = 0;
  for (long long base = 0; base < n; base += 1024 * PER) {
    const long long i0 = base + (long long)threadIdx.x * PER;
    int v[PER], sum = 0;
#pragma unroll
    for (int e = 0; e < PER; ++e) { v[e] = (i0 + e < n) ? in[i0 + e] : 0; sum += v[e]; }
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int w = wsum[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, w, o); if (lane >= o) w += t; }
      wsum[lane] = w;
    }
    __syncthreads();
    int run = incl - sum + (warp > 0 ? wsum[warp - 1] : 0) + carry;
#pragma unroll
    for (int e = 0; e < PER; ++e) { if (i0 + e < n) out[i0 + e] = run; run += v[e]; }
    if (threadIdx.x == 1023) carry_s = run;
    __syncthreads();
    carry = carry_s;
    __syncthreads();
  }
  if (threadIdx.x == 0) out[n] = carry;
}

__global__ void k_cells_fill(const double* __restrict__ centers, long long n, CellGrid g, const int* __restrict__ cell_start,
                             int* __restrict__ cell_cur, int* __restrict__ order) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = cell_of(g, centers + 3 * i);
  order[cell_start[c] + atomicAdd(&cell_cur[c], 1)] = (int)i;
}

// One thread per seed, two traversals of its 27 cells (count, then fill at a position taken from the
// pool with one atomicAdd).  nbr: seeds that can own a voxel of this seed's window (centres within
// 2 (r + sqrt 3)); dep: seeds whose window can overlap this seed's (integer centre offsets e with
// |e_a| <= 2r - 1 and |e|^2 <= 4 r^2).  The order of the entries is irrelevant to every consumer.
__global__ void k_nbr_build(FitDev d, CellGrid g, const int* __restrict__ cell_start, const int* __restrict__ order) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.n) return;
  const double c[3] = {d.centers[3 * i], d.centers[3 * i + 1], d.centers[3 * i + 2]};
  const int ic[3] = {(int)c[0], (int)c[1], (int)c[2]};
  const int ca = cell_coord(g, c[0], 0), cb = cell_coord(g, c[1], 1), cc = cell_coord(g, c[2], 2);
  const int r = d.radius, lim = 2 * r - 1;
  const double reach = 2.0 * ((double)r + 1.7320508075688772) + 1e-6;
  int cv = 0, co = 0, nl = 0, own = (int)i, bv = 0, bo = 0;
  for (int pass = 0; pass < 2; ++pass) {
    int wv = 0, wo = 0;
    for (int da = -1; da <= 1; ++da) for (int db = -1; db <= 1; ++db) for (int dc = -1; dc <= 1; ++dc) {
      const int a = ca + da, b = cb + db, e3 = cc + dc;
      if (a < 0 || a >= g.g[0] || b < 0 || b >= g.g[1] || e3 < 0 || e3 >= g.g[2]) continue;
      const int id = (a * g.g[1] + b) * g.g[2] + e3;
      for (int e = cell_start[id]; e < cell_start[id + 1]; ++e) {
        const int j = order[e];
        if (j == (int)i) continue;
        const double q0 = d.centers[3 * (long long)j], q1 = d.centers[3 * (long long)j + 1], q2 = d.centers[3 * (long long)j + 2];
        const double d0 = q0 - c[0], d1 = q1 - c[1], d2 = q2 - c[2];
        const bool isv = d0 * d0 + d1 * d1 + d2 * d2 <= reach * reach;
        const int e0 = abs((int)q0 - ic[0]), e1 = abs((int)q1 - ic[1]), e2 = abs((int)q2 - ic[2]);
        const bool iso = e0 <= lim && e1 <= lim && e2 <= lim && e0 * e0 + e1 * e1 + e2 * e2 <= 4 * r * r;
        if (pass == 0) {
          cv += isv; co += iso; nl += (iso && j < (int)i);
          if (d0 == 0 && d1 == 0 && d2 == 0 && j < own) own = j;
        } else {
          if (isv) d.nbr_idx[bv + wv++] = j;
          if (iso) d.dep_idx[bo + wo++] = j;
        }
      }
    }
    if (pass == 0) {
      bv = cv ? atomicAdd(&d.ctl->pool_v, cv) : 0;
      bo = co ? atomicAdd(&d.ctl->pool_o, co) : 0;
      const bool ovf = (bv + cv > d.pool_cap_v) || (bo + co > d.pool_cap_o);
      d.own_id[i] = own; d.n_lower[i] = nl;
      d.nbr_start[i] = bv; d.dep_start[i] = bo;
      d.nbr_cnt[i] = ovf ? 0 : cv; d.dep_cnt[i] = ovf ? 0 : co;
      if (ovf) { d.ctl->overflow = 1; return; }     // the host enlarges the pools (pool_v / pool_o keep counting) and rebuilds
    }
  }
}

// brick table: -1 = no window touches the brick; marked bricks get a number in [0, n_bricks)
__global__ void k_brick_mark(FitDev d) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= d.n) return;
  const int dims[3] = {d.Z, d.X, d.Y};
  int lo[3], hi[3];
  for (int a = 0; a < 3; ++a) {
    const int ic = (int)d.centers[3 * s + a];
    lo[a] = max(ic - d.radius, 0);
    hi[a] = min(ic + d.radius - 1, dims[a] - 1);
    if (lo[a] > hi[a]) return;
  }
  for (int bz = lo[0] >> 3; bz <= hi[0] >> 3; ++bz)
    for (int bx = lo[1] >> 3; bx <= hi[1] >> 3; ++bx)
      for (int by = lo[2] >> 3; by <= hi[2] >> 3; ++by) d.brick_tab[((long long)bz * d.nbx + bx) * d.nby + by] = -2;
}
__global__ void k_brick_assign(FitDev d, long long tab_n) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= tab_n) return;
  if (d.brick_tab[t] == -2) d.brick_tab[t] = atomicAdd(&d.ctl->n_bricks, 1);
}

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
  const int nb0 = d.nbr_start[s], nb1 = nb0 + d.nbr_cnt[s];
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
            const double cj[3] = {d.centers[3 * (long long)j], d.centers[3 * (long long)j + 1], d.centers[3 * (long long)j + 2]};
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
            const double cj[3] = {d.centers[3 * (long long)j], d.centers[3 * (long long)j + 1], d.centers[3 * (long long)j + 2]};
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
      const int pos = atomicAdd(&d.ctl->tie_count, 1);
      if (pos < d.tie_cap) { d.tie_spot[pos] = (int)s; d.tie_k[pos] = k; }
    }
  }
}

// all_ok0: every seed has >= 10 member voxels (firstfit succeeds everywhere); all_ok1: every seed has >= 10
// window voxels inside the image (every seed holds a fit after the first repeat sweep).  These decide whether
// np.array(centers_fit) is float32 or float64 in the reference's distance test.
__global__ void __launch_bounds__(WARPS * 32) k_member_stats(FitDev d) {
  const int lane = threadIdx.x & 31;
  const long long s = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5);
  if (s >= d.n) return;
  const int ic[3] = {(int)d.centers[3 * s], (int)d.centers[3 * s + 1], (int)d.centers[3 * s + 2]};
  int m0 = 0, mf = 0;
  for (int k0 = 0; k0 < d.K; k0 += 32) {
    const int k = k0 + lane;
    bool in = false;
    if (k < d.K) {
      const int z = ic[0] + d.offs[3 * k], x = ic[1] + d.offs[3 * k + 1], y = ic[2] + d.offs[3 * k + 2];
      in = (z >= 0 && z < d.Z && x >= 0 && x < d.X && y >= 0 && y < d.Y);
    }
    mf += __popc(__ballot_sync(FULL, in));
    m0 += __popc(d.mask[s * d.KW + (k0 >> 5)]);
  }
  if (lane == 0) {
    if (m0 < NP) d.ctl->all_ok0 = 0;
    if (mf < NP) d.ctl->all_ok1 = 0;
  }
}

// ------------------------------------------------------------------------------------------
// Engine: scheduler
// ------------------------------------------------------------------------------------------
__global__ void k_engine_init(FitDev d) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) {
    EngineCtl* c = d.ctl;
    for (int p = 0; p < 2; ++p) { c->n_bulk[p] = c->cur_bulk[p] = c->n_team[p] = c->cur_team[p] = c->n_park[p] = c->alive[p] = c->blocks_done[p] = 0; }
    c->cap_team_now = d.cap_team_short; c->done = 0; c->all_ok0 = 1; c->all_ok1 = 1;
    c->st_lm_runs = c->st_memo_hits = c->st_spec_runs = c->st_spec_hits = c->st_parked = c->st_team_tasks = c->st_tasks = c->st_rounds = 0;
    c->st_evals = 0ull;
    for (int q = 0; q < 8; ++q) c->prof[q] = 0ull;
  }
  if (i >= d.n) return;
  d.stage[i] = -2; d.fin[i] = 0; d.busy[i] = 0; d.conv[i] = 0; d.specst[i] = 0; d.succ_prev[i] = 0;
  d.memo_valid[i] = 0; d.memo_committed[i] = 0; d.success[i] = 0; d.succ_first[i] = 0;
  d.dists[i] = INFINITY;
  d.cen_prev[3 * i] = d.cen_prev[3 * i + 1] = d.cen_prev[3 * i + 2] = NAN;
}

__device__ __forceinline__ unsigned* list_bulk(const FitDev& d, int p) { return d.lists + (size_t)(0 + p) * d.list_cap; }
__device__ __forceinline__ unsigned* list_team(const FitDev& d, int p) { return d.lists + (size_t)(2 + p) * d.list_cap; }
__device__ __forceinline__ unsigned* list_park(const FitDev& d, int p) { return d.lists + (size_t)(4 + p) * d.list_cap; }

__global__ void __launch_bounds__(256) k_sched(FitDev d, int round, int phases, int sweep_cap) {
  EngineCtl* c = d.ctl;
  const int pr = round & 1, nx = pr ^ 1;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t == 0) {      // nothing touches the other parity's counters during this round's k_sched
    c->n_bulk[nx] = 0; c->cur_bulk[nx] = 0; c->n_team[nx] = 0; c->cur_team[nx] = 0; c->alive[nx] = 0; c->blocks_done[nx] = 0;
    c->n_park[nx] = 0;     // filled by this round's fit kernels
  }
  unsigned* bulk = list_bulk(d, pr);
  unsigned* team = list_team(d, pr);
  int alive = 0;
  if (t < d.n) {
    const int i = (int)t;
    const int st = d.stage[i];
    if (d.busy[i]) alive = 1;
    else if (st == -2) {
      if (phases & 1) {
        alive = 1;
        d.busy[i] = 1;
        bulk[atomicAdd(&c->n_bulk[pr], 1)] = make_task(i, TASK_F, false);
        if ((phases & 4) && d.dep_cnt[i] == 0 && d.specst[i] == 0) {
          d.specst[i] = 1;
          bulk[atomicAdd(&c->n_bulk[pr], 1)] = make_task(i, TASK_S, false);
        }
      }
    } else if (st == -1) {
      alive = 1;
      bool ready = true;
      const int e0 = d.dep_start[i], e1 = e0 + d.dep_cnt[i];
      for (int e = e0; e < e1 && ready; ++e) { const int j = d.dep_idx[e]; if (j < i && d.stage[j] < 0) ready = false; }
      if (ready) { d.busy[i] = 1; bulk[atomicAdd(&c->n_bulk[pr], 1)] = make_task(i, TASK_B, false); }
    } else if ((phases & 2) && !d.fin[i]) {
      const int k = st + 1;
      if (k > d.max_sweeps) d.fin[i] = 1;
      else if (k <= sweep_cap) {
        alive = 1;
        bool ready = d.specst[i] != 1;
        const int e0 = d.dep_start[i], e1 = e0 + d.dep_cnt[i];
        for (int e = e0; e < e1 && ready; ++e) {
          const int j = d.dep_idx[e];
          const int eff = d.fin[j] ? BIG_SWEEP : d.stage[j];
          if (eff < (j < i ? k : k - 1)) ready = false;
        }
        if (ready) { d.busy[i] = 1; bulk[atomicAdd(&c->n_bulk[pr], 1)] = make_task(i, TASK_R, false); }
      }
    }
    if (d.specst[i] == 1) alive = 1;
  }
  if (t < c->n_park[pr]) {
    const unsigned task = list_park(d, pr)[t];
    const long long seed = task & TASK_SEED_MASK;
    const long long slot = (((task >> 28) & 3u) == TASK_S) ? d.n + seed : seed;
    alive = 1;
    if (d.live[slot].nfev >= d.team_after) team[atomicAdd(&c->n_team[pr], 1)] = task;
    else bulk[atomicAdd(&c->n_bulk[pr], 1)] = task;
  }
  if (alive) { c->alive[pr] = 1; __threadfence(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const int tk = atomicAdd(&c->blocks_done[pr], 1);
    if (tk == (int)gridDim.x - 1) {
      __threadfence();
      const int al = *(volatile int*)&c->alive[pr];
      const int nb = *(volatile int*)&c->n_bulk[pr];
      c->cap_team_now = (nb == 0) ? d.cap_team_long : d.cap_team_short;
      // A short work list leaves the GPU to this stack's critical path: the team kernel and the one-warp kernel
      // of a round run one after the other, and a team evaluates twice as fast.  With few tasks in all, the one-warp
      // ones join the team list (results do not depend on who runs a task) and the one-warp kernel finds nothing.
      const int nt = *(volatile int*)&c->n_team[pr];
      if (nb > 0 && nb + nt <= d.merge_small) {
        for (int q = 0; q < nb; ++q) team[nt + q] = *(volatile unsigned*)&bulk[q];
        c->n_team[pr] = nt + nb;
        c->n_bulk[pr] = 0;
        __threadfence();
      }
      if (!al) { c->done = 1; *(volatile int*)d.h_done = 1; __threadfence_system(); }
      else {
        const int slot = c->st_rounds & 511;
        unsigned long long ns;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns));
        c->trace_work[slot] = ((unsigned)min(*(volatile int*)&c->n_bulk[pr], 65535) << 16) | (unsigned)min(*(volatile int*)&c->n_team[pr], 65535);
        c->trace_ns[slot] = ns;
        c->st_rounds += 1;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Engine: one task on a CTA of TW warps
// ------------------------------------------------------------------------------------------
struct CtaScratch {
  int go, action, w, hit;
  int old_succ, pad0, pad1, pad2;
  float old_c[4];
  FitResult res;
  int cnt[64];                 // voxels per 32-chunk of the window (K <= 2048)
};

template <int TW> __device__ __forceinline__ void cta_sync() { if constexpr (TW == 1) __syncwarp(); else __syncthreads(); }
template <int TW> __device__ __forceinline__ bool cta_all(bool p) {
  if constexpr (TW == 1) return __all_sync(FULL, p);
  else return __syncthreads_and(p ? 1 : 0) != 0;
}
template <int TW> __device__ __forceinline__ bool cta_any(bool p) {
  if constexpr (TW == 1) return __any_sync(FULL, p);
  else return __syncthreads_or(p ? 1 : 0) != 0;
}

__host__ __device__ constexpr size_t al16(size_t b) { return (b + 15) / 16 * 16; }
constexpr int GRAM_DOUBLES = (NP + 1) * GRAM_PITCH;
struct EngineSmem {
  size_t off_cs, off_gram, off_part, off_dv, off_pk, total;
  __host__ __device__ EngineSmem(int K, int tw) {
    off_cs = al16(sizeof(SpotShared<double>));
    off_gram = off_cs + al16(sizeof(CtaScratch));
    off_part = off_gram + (size_t)(tw - 1) * GRAM_DOUBLES * 8;
    off_dv = off_part + (tw > 1 ? (size_t)tw * 6 * 32 * 8 : 0);
    off_pk = off_dv + (size_t)K * 8;
    total = al16(off_pk + (size_t)K * 4);
  }
};

// Residual norm and normal equations at sh.vc over the window, by all TW warps of the CTA.  Batches of
// 32 voxels go round-robin over the warps; every batch tile starts from zero (mma_batch_tile) and the
// tiles are added in batch order by warp 0 -- the same additions, in the same order, as one warp alone
// performs (pass_fused_mma).  The return value is valid on warp 0.
template <int TW, typename Vox>
__device__ __forceinline__ double pass_cta(WarpExec& ex, SpotShared<double>& sh, const Vox& vox, double* gram_w, double* part) {
  if constexpr (TW == 1) {
    return pass_fused_mma<double>(ex, sh.vc, vox, sh.Ag, sh.gram);
  } else {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double agiant = 1.304e19 / (double)vox.m;
    const int nb = (vox.m + 31) >> 5;
    double acc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    int nbad = 0;
    for (int g0 = 0; g0 < nb; g0 += TW) {
      const int b = g0 + warp;
      if (b < nb) {
        double t[6];
        mma_batch_tile<double>(sh.vc, vox, b * 32, lane, gram_w, t, nbad, agiant);
#pragma unroll
        for (int e = 0; e < 6; ++e) part[(warp * 6 + e) * 32 + lane] = t[e];
      }
      __syncthreads();
      if (warp == 0) {
        const int cnt = min(TW, nb - g0);
        for (int w = 0; w < cnt; ++w) {
#pragma unroll
          for (int e = 0; e < 6; ++e) acc[e] += part[(w * 6 + e) * 32 + lane];
        }
      }
      __syncthreads();
    }
    const bool bad = __syncthreads_or(nbad) != 0;
    double fn = 0.0;
    if (warp == 0) {
      const double s2 = mma_scatter(acc, lane, sh.Ag);
      fn = bad ? pass_residual<double>(ex, sh.vc, vox, (double*)0) : sqrt(s2);
    }
    return fn;
  }
}

// lmder, driven by warp 0; the other warps of the CTA only join the voxel passes.  Same control flow
// as run_lm (fit_spot.h).  Returns true if the run was suspended after `cap` evaluations.
#ifndef IA3_LM_WARP
#define IA3_LM_WARP 1            // 1: register / shuffle algebra of lm_warp.h; 0: the generic shared-memory algebra of lm_core.h
#endif
#if IA3_LM_WARP
#define LM_OUTER(ex, st, cfg, A, g) lw::outer(st, cfg, A, g)
#define LM_PROPOSE(ex, st) lw::propose(st)
#define LM_JUDGE(ex, st, cfg, fn) lw::judge(st, cfg, fn)
#define LM_CONSTS(ex, fp, cen, origin, x, sh) lw::build_consts<double>(fp, cen, origin, x, (sh).etab, (sh).scal, (sh).vc)
#else
#define LM_CONSTS(ex, fp, cen, origin, x, sh) build_consts_par<double>(ex, fp, cen, origin, x, sh)
#define LM_OUTER(ex, st, cfg, A, g) lm_outer(ex, st, cfg, A, g)
#define LM_PROPOSE(ex, st) lm_propose(ex, st)
#define LM_JUDGE(ex, st, cfg, fn) lm_judge(ex, st, cfg, fn)
#endif
#ifdef IA3_FIT_PROF
#define PROF_T(v) const long long v = clock64()
#define PROF_ADD(i, a, b) do { if (TW > 1 && threadIdx.x == 0) prof_acc[i] += (unsigned long long)((b) - (a)); } while (0)
#else
#define PROF_T(v) do {} while (0)
#define PROF_ADD(i, a, b) do {} while (0)
#endif
template <int TW, typename Vox>
__device__ __forceinline__ bool run_lm_cta(WarpExec& ex, const FitParams& fp, const LMConfig& cfg, const double* cen,
                                           const double* origin, const Vox& vox, SpotShared<double>& sh, CtaScratch& cs,
                                           double* gram_w, double* part, int cap, int start, unsigned long long* prof_out) {
  const int warp = threadIdx.x >> 5;
  LMState& st = sh.st;
#ifdef IA3_FIT_PROF
  unsigned long long prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
  if (start == LM_START_FRESH) {
    if (warp == 0) LM_CONSTS(ex, fp, cen, origin, sh.x0, sh);
    cta_sync<TW>();
    const double fn0 = pass_cta<TW>(ex, sh, vox, gram_w, part);
    if (warp == 0) lm_init(ex, st, sh.x0, fn0);
  }
  cta_sync<TW>();
  const int nfev_entry = st.nfev;
  bool suspended = false;
  for (;;) {
    // sh.Ag holds J^T J, J^T f at st.x (x0, or the trial point that was just accepted)
    PROF_T(t0);
    if (warp == 0) {
      int go = 1;
      __syncwarp();
      if (cap > 0 && st.nfev - nfev_entry >= cap) go = 2;
      else if (!LM_OUTER(ex, st, cfg, sh.Ag, sh.Ag + NTRI)) go = 0;
      if (ex.lane() == 0) cs.go = go;
    }
    cta_sync<TW>();
    PROF_T(t1);
    PROF_ADD(0, t0, t1);
    const int go = cs.go;
    if (go == 2) { suspended = true; break; }
    if (go == 0) break;
    int action;
    for (;;) {
      PROF_T(t2);
      if (warp == 0) LM_PROPOSE(ex, st);
      PROF_T(t3);
      if (warp == 0) LM_CONSTS(ex, fp, cen, origin, st.xt, sh);
      cta_sync<TW>();
      PROF_T(t4);
      const double fn1 = pass_cta<TW>(ex, sh, vox, gram_w, part);     // lm_outer has consumed the old sums
      PROF_T(t5);
      if (warp == 0) {
        const int a = LM_JUDGE(ex, st, cfg, fn1);
        if (ex.lane() == 0) cs.action = a;
      }
      cta_sync<TW>();
      PROF_T(t6);
      PROF_ADD(1, t2, t3); PROF_ADD(2, t3, t4); PROF_ADD(3, t4, t5); PROF_ADD(4, t5, t6); PROF_ADD(5, 0, 1);
      action = cs.action;
      if (action != LM_RETRY) break;
    }
    if (action == LM_DONE) break;
  }
  cta_sync<TW>();
#ifdef IA3_FIT_PROF
  if (TW > 1 && threadIdx.x == 0) for (int q = 0; q < 6; ++q) atomicAdd(&prof_out[q], prof_acc[q]);
#endif
  return suspended;
}

// np.sum((old - new) ** 2, axis=-1) of the reference's convergence test (Fitting_v4.py:678), in the
// dtype np.array(centers_fit) has there: float32 when every seed holds a fit, float64 otherwise;
// numpy adds the three squares left to right.
__device__ __forceinline__ double centre_dist2(const float* o, const float* nw, bool f32) {
  if (f32) {
    const float d0 = __fsub_rn(o[0], nw[0]), d1 = __fsub_rn(o[1], nw[1]), d2 = __fsub_rn(o[2], nw[2]);
    return (double)__fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2));
  }
  const double d0 = __dsub_rn((double)o[0], (double)nw[0]), d1 = __dsub_rn((double)o[1], (double)nw[1]), d2 = __dsub_rn((double)o[2], (double)nw[2]);
  return __dadd_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)), __dmul_rn(d2, d2));
}

template <int TW>
__device__ void process_task(const FitDev& d, unsigned task, unsigned char* base, int cap, int nx) {
  constexpr int NT = TW * 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int kind = (int)((task >> 28) & 3u);
  const bool resume = (task & TASK_RESUME) != 0;
  const long long s = (long long)(task & TASK_SEED_MASK);
  const int K = d.K;
  const EngineSmem lay(K, TW);
  SpotShared<double>& sh = *reinterpret_cast<SpotShared<double>*>(base);
  CtaScratch& cs = *reinterpret_cast<CtaScratch*>(base + lay.off_cs);
  double* gram_w = (warp == 0) ? sh.gram : reinterpret_cast<double*>(base + lay.off_gram) + (size_t)(warp - 1) * GRAM_DOUBLES;
  double* part = reinterpret_cast<double*>(base + lay.off_part);
  double* dv = reinterpret_cast<double*>(base + lay.off_dv);
  uint32_t* pk = reinterpret_cast<uint32_t*>(base + lay.off_pk);
  EngineCtl* ctl = d.ctl;

  const double c[3] = {d.centers[3 * s], d.centers[3 * s + 1], d.centers[3 * s + 2]};
  const int ic[3] = {(int)c[0], (int)c[1], (int)c[2]};
  const double origin[3] = {(double)ic[0], (double)ic[1], (double)ic[2]};
  FitParams fp = d.fp;
  fp.delta = (kind == TASK_R || kind == TASK_S) ? d.delta_repeat : d.delta_first;
  WarpExec ex;

  // ---- B: im_subtr[window] -= firstfit's reconstruction (Fitting_v4.py:629-633), full clipped window ----
  auto subtract_full = [&]() {               // sh.vc = per-voxel constants of the firstfit parameters
    for (int k = tid; k < K; k += NT) {
      const int dz = d.offs[3 * k], dx = d.offs[3 * k + 1], dy = d.offs[3 * k + 2];
      const int z = ic[0] + dz, x = ic[1] + dx, y = ic[2] + dy;
      if (z < 0 || z >= d.Z || x < 0 || x >= d.X || y < 0 || y >= d.Y) continue;
      const double f0 = eval_f0<double>(sh.vc, (double)dz, (double)dx, (double)dy);
      d.rec[s * K + k] = f0;
      d.vol[vol_index(d, z, x, y)] -= f0;
    }
  };
  if (kind == TASK_B) {
    if (d.succ_first[s]) {
      if (tid == 0) {
        double x[NP];
        for (int i = 0; i < NP; ++i) x[i] = d.praw_first[s * NP + i];
        build_consts<double>(fp, c, origin, x, false, sh.vc);
      }
      cta_sync<TW>();
      subtract_full();
    }
    cta_sync<TW>();
    if (tid == 0) { d.stage[s] = 0; d.busy[s] = 0; }
    return;
  }

  // ---- gather the window (in-image voxels; firstfit: Voronoi members only) in window order ----
  const bool had_rec = (kind == TASK_R) && d.success[s];
  if (tid == 0) {
    cs.old_succ = d.success[s];
    for (int i = 0; i < 3; ++i) cs.old_c[i] = d.ps[s * NOUT + 1 + i];
  }
  const int nch = (K + 31) >> 5;
  for (int pass = 0; pass < 2; ++pass) {
    for (int ch = warp; ch < nch; ch += TW) {
      const int k = ch * 32 + lane;
      bool use = false;
      int dz = 0, dx = 0, dy = 0;
      if (k < K) {
        dz = d.offs[3 * k]; dx = d.offs[3 * k + 1]; dy = d.offs[3 * k + 2];
        const int z = ic[0] + dz, x = ic[1] + dx, y = ic[2] + dy;
        use = (z >= 0 && z < d.Z && x >= 0 && x < d.X && y >= 0 && y < d.Y);
        if (kind == TASK_F) use = use && ((d.mask[s * d.KW + ch] >> lane) & 1u);
      }
      const unsigned bal = __ballot_sync(FULL, use);
      if (pass == 0) { if (lane == 0) cs.cnt[ch] = __popc(bal); continue; }
      if (use) {
        int pos = __popc(bal & ((1u << lane) - 1u));
        for (int q = 0; q < ch; ++q) pos += cs.cnt[q];
        const int z = ic[0] + dz, x = ic[1] + dx, y = ic[2] + dy;
        double v;
        if (kind == TASK_R) { v = d.vol[vol_index(d, z, x, y)]; if (had_rec) v = d.rec[s * K + k] + v; }   // im_ = im_rec + im_  (:662)
        else v = load_im(d.im, d.im_dtype, ((long long)z * d.X + x) * d.Y + y);
        dv[pos] = v;
        pk[pos] = pack_vox(dz, dx, dy, k);
      }
    }
    cta_sync<TW>();
  }
  int m = 0;
  for (int q = 0; q < nch; ++q) m += cs.cnt[q];

  if (m < NP) {
    // len(p_) > len(im): success = False (Fitting_v4.py:382-383)
    if (tid == 0) {
      if (kind == TASK_S) { d.specst[s] = 2; return; }
      d.success[s] = 0; d.nfev[s] = 0; d.info[s] = 0;
      if (kind == TASK_F) {      // firstfit stores a NaN row; nothing is subtracted
        for (int i = 0; i < NOUT; ++i) d.ps[s * NOUT + i] = NAN;
        for (int i = 0; i < NP; ++i) { d.p_raw[s * NP + i] = NAN; d.praw_first[s * NP + i] = NAN; }
        d.succ_first[s] = 0;
        d.stage[s] = 0;
      } else {                   // a failed repeat visit keeps the old row; keep = False -> dists = 0 -> converged
        const int k = d.stage[s] + 1;
        d.succ_prev[s] = (uint8_t)cs.old_succ;
        for (int i = 0; i < 3; ++i) d.cen_prev[3 * s + i] = cs.old_c[i];
        d.dists[s] = 0.0;
        d.conv[s] = (0.0 < d.th2) ? 1 : 0;
        d.stage[s] = k;
        if (0.0 < d.th2 || k >= d.max_sweeps) d.fin[s] = 1;
      }
      d.busy[s] = 0;
    }
    return;
  }

  // ---- initial guess (GaussianFit.__init__); also when resuming: the v3 width prior lives in fp.init_wt ----
  if (warp == 0) {
    select10(ex, dv, m, false, sh.small10);
    select10(ex, dv, m, true, sh.large10);
    if (lane == 0) initial_guess(fp, sh.small10, sh.large10, d.init_w, sh.x0);
    __syncwarp();
  }
  cta_sync<TW>();
  BallVox<double> vox{m, pk, dv};
  const long long slot = (kind == TASK_S) ? d.n + s : s;

  int start = LM_START_FRESH;
  bool hit = false;
  if (resume) {
    if (warp == 0) {
      const LMLive* g = d.live + slot;
      for (int i = lane; i < NP; i += 32) { sh.st.x[i] = g->x[i]; sh.st.diag[i] = g->diag[i]; }
      for (int i = lane; i < NTRI + NP; i += 32) sh.Ag[i] = g->Ag[i];
      if (lane == 0) {
        sh.st.fnorm = g->fnorm; sh.st.xnorm = g->xnorm; sh.st.delta = g->delta; sh.st.par = g->par;
        sh.st.iter = g->iter; sh.st.nfev = g->nfev; sh.st.njev = g->njev; sh.st.info = g->info;
      }
    }
    start = LM_START_CONTINUE;
  } else if (kind != TASK_F && d.memo_on) {
    // memo: same float32 window values and same x0 as the seed's previous run in repeat mode?
    float* kd = d.key_d32 + s * K;
    double* kx = d.key_x0 + s * NP;
    if (kind == TASK_R) {
      bool same = d.memo_valid[s] == 2;
      if (same) {
        for (int pos = tid; pos < m; pos += NT) same = same && (__float_as_uint(kd[pos]) == __float_as_uint(__double2float_rn(dv[pos])));
        if (tid < NP) same = same && (__double_as_longlong(kx[tid]) == __double_as_longlong(sh.x0[tid]));
      }
      hit = cta_all<TW>(same);
    }
    if (!hit) {
      for (int pos = tid; pos < m; pos += NT) kd[pos] = __double2float_rn(dv[pos]);
      if (tid < NP) kx[tid] = sh.x0[tid];
      if (tid == 0) { d.memo_valid[s] = 1; d.memo_committed[s] = 0; }
    }
  }

  bool suspended = false;
  if (!hit) suspended = run_lm_cta<TW>(ex, fp, d.lm, c, origin, vox, sh, cs, gram_w, part, cap, start, ctl->prof);
  if (suspended) {
    if (warp == 0) {
      LMLive* g = d.live + slot;
      for (int i = lane; i < NP; i += 32) { g->x[i] = sh.st.x[i]; g->diag[i] = sh.st.diag[i]; }
      for (int i = lane; i < NTRI + NP; i += 32) g->Ag[i] = sh.Ag[i];
      if (lane == 0) {
        g->fnorm = sh.st.fnorm; g->xnorm = sh.st.xnorm; g->delta = sh.st.delta; g->par = sh.st.par;
        g->iter = sh.st.iter; g->nfev = sh.st.nfev; g->njev = sh.st.njev; g->info = sh.st.info;
        list_park(d, nx)[atomicAdd(&ctl->n_park[nx], 1)] = make_task((int)s, kind, true);
        atomicAdd(&ctl->st_parked, 1);
      }
    }
    return;
  }

  if (tid == 0 && !hit) {
    atomicAdd(&ctl->st_lm_runs, 1);
    atomicAdd(&ctl->st_evals, (unsigned long long)sh.st.nfev);
    if (kind == TASK_S) atomicAdd(&ctl->st_spec_runs, 1);
  }
  if (kind != TASK_F && d.memo_on && !hit) {
    if (tid < NP) d.memo_praw[s * NP + tid] = sh.st.x[tid];
    if (tid == 0) {
      d.memo_meta[4 * s] = sh.st.nfev; d.memo_meta[4 * s + 1] = sh.st.njev; d.memo_meta[4 * s + 2] = sh.st.info;
      d.memo_valid[s] = 2; d.memo_committed[s] = (kind == TASK_R) ? 1 : 0;
    }
  }
  if (kind == TASK_S) {
    if (tid == 0) d.specst[s] = 2;
    return;
  }

  // a hit on a result that ps / rec already hold needs no epilogue arithmetic: same parameters, same
  // reconstruction, centre unmoved
  const bool cheap = hit && d.memo_committed[s];
  if (hit && !cheap) {           // adopt the speculative run's result
    if (tid < NP) sh.st.x[tid] = d.memo_praw[s * NP + tid];
    if (tid == 0) {
      sh.st.nfev = d.memo_meta[4 * s]; sh.st.njev = d.memo_meta[4 * s + 1]; sh.st.info = d.memo_meta[4 * s + 2];
      atomicAdd(&ctl->st_spec_hits, 1);
    }
  }
  if (tid == 0 && hit) atomicAdd(&ctl->st_memo_hits, 1);
  cta_sync<TW>();
  if (!cheap) {
    if (warp == 0) finish_fit<double>(ex, fp, c, origin, vox, sh, &cs.res);      // leaves sh.vc = constants of the final parameters
    cta_sync<TW>();
    if (tid < NOUT) d.ps[s * NOUT + tid] = cs.res.ps[tid];
    if (tid < NP) d.p_raw[s * NP + tid] = cs.res.p_raw[tid];
    if (tid == 0) { d.success[s] = 1; d.nfev[s] = cs.res.nfev; d.info[s] = cs.res.info; }
  }

  if (kind == TASK_F) {
    if (tid < NP) d.praw_first[s * NP + tid] = cs.res.p_raw[tid];
    const bool sub_now = d.n_lower[s] == 0;       // no lower-index window overlaps: subtract right away
    if (sub_now) subtract_full();
    cta_sync<TW>();
    if (tid == 0) { d.succ_first[s] = 1; d.stage[s] = sub_now ? 0 : -1; d.busy[s] = 0; }
    return;
  }

  // ---- R epilogue: im_rec = get_im(); ims_rec[ic] = im_rec; im_add[window] = im_ - im_rec (:671-675) ----
  for (int pos = tid; pos < m; pos += NT) {
    const uint32_t p = pk[pos];
    const int dz = (int)(p & 63u) - 32, dx = (int)((p >> 6) & 63u) - 32, dy = (int)((p >> 12) & 63u) - 32;
    const int k = (int)(p >> 18);
    double f0;
    if (cheap) f0 = d.rec[s * K + k];
    else { f0 = eval_f0<double>(sh.vc, (double)dz, (double)dx, (double)dy); d.rec[s * K + k] = f0; }
    d.vol[vol_index(d, ic[0] + dz, ic[1] + dx, ic[2] + dy)] = dv[pos] - f0;
  }
  cta_sync<TW>();
  if (tid == 0) {
    const int k = d.stage[s] + 1;
    float nc[3];
    for (int i = 0; i < 3; ++i) nc[i] = cheap ? cs.old_c[i] : cs.res.ps[1 + i];
    const bool keep = cs.old_succ != 0;          // success & success_old; this visit succeeded
    const bool f32 = (k == 1) ? (ctl->all_ok0 != 0) : (ctl->all_ok1 != 0);
    const double dist = keep ? centre_dist2(cs.old_c, nc, f32) : 0.0;
    const bool cv = dist < d.th2;
    d.succ_prev[s] = (uint8_t)cs.old_succ;
    for (int i = 0; i < 3; ++i) d.cen_prev[3 * s + i] = cs.old_c[i];
    d.dists[s] = dist;
    d.conv[s] = cv ? 1 : 0;
    if (hit) d.memo_committed[s] = 1;            // ps / rec now hold the memoised run's result
    d.stage[s] = k;
    if (cv || k >= d.max_sweeps) d.fin[s] = 1;
    d.busy[s] = 0;
  }
}

template <int TW>
__global__ void __launch_bounds__(TW * 32, TW == 1 ? IA3_FIT_MINBLOCKS : 1) k_fit_round(FitDev d, int round) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int w_s;
  EngineCtl* c = d.ctl;
  const int pr = round & 1;
  const int n_work = (TW == 1) ? c->n_bulk[pr] : c->n_team[pr];
  if (n_work == 0) return;
  int* cursor = (TW == 1) ? &c->cur_bulk[pr] : &c->cur_team[pr];
  const unsigned* list = (TW == 1) ? list_bulk(d, pr) : list_team(d, pr);
  const int cap = (TW == 1) ? d.cap_bulk : c->cap_team_now;
  for (;;) {
    cta_sync<TW>();
    if (threadIdx.x == 0) w_s = atomicAdd(cursor, 1);
    cta_sync<TW>();
    const int w = w_s;
    if (w >= n_work) break;
    if (threadIdx.x == 0) { atomicAdd(&c->st_tasks, 1); if (TW > 1) atomicAdd(&c->st_team_tasks, 1); }
    process_task<TW>(d, list[w], smem_raw, cap, pr ^ 1);
  }
}

// firstfit's subtraction replayed into a dense float64 volume (im_subtr on demand): one pass over the
// seeds whose lower-index overlapping seeds were done BEFORE this launch
__global__ void __launch_bounds__(WARPS * 32) k_subtract_dense(FitDev d, double* dense, const uint8_t* done, uint8_t* done_next, int* pending) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long s = (long long)blockIdx.x * WARPS + warp;
  if (s >= d.n) return;
  if (done[s]) { if (lane == 0) done_next[s] = 1; return; }
  bool ready = true;
  const int e0 = d.dep_start[s], e1 = e0 + d.dep_cnt[s];
  for (int e = e0; e < e1; ++e) { const int j = d.dep_idx[e]; if (j < s && !done[j]) ready = false; }
  if (!ready) { if (lane == 0) { done_next[s] = 0; atomicAdd(pending, 1); } return; }
  if (d.succ_first[s]) {
    __shared__ VoxConsts<double> vcd_s[WARPS];
    VoxConsts<double>& vcd = vcd_s[warp];
    const double c[3] = {d.centers[3 * s], d.centers[3 * s + 1], d.centers[3 * s + 2]};
    const int ic[3] = {(int)c[0], (int)c[1], (int)c[2]};
    if (lane == 0) {
      const double origin[3] = {(double)ic[0], (double)ic[1], (double)ic[2]};
      double x[NP];
      for (int i = 0; i < NP; ++i) x[i] = d.praw_first[s * NP + i];
      FitParams fp = d.fp;
      fp.delta = d.delta_first;
      build_consts<double>(fp, c, origin, x, false, vcd);
    }
    __syncwarp();
    for (int k = lane; k < d.K; k += 32) {
      const int dz = d.offs[3 * k], dx = d.offs[3 * k + 1], dy = d.offs[3 * k + 2];
      const int z = ic[0] + dz, x = ic[1] + dx, y = ic[2] + dy;
      if (z < 0 || z >= d.Z || x < 0 || x >= d.X || y < 0 || y >= d.Y) continue;
      dense[((long long)z * d.X + x) * d.Y + y] -= eval_f0<double>(vcd, (double)dz, (double)dx, (double)dy);
    }
  }
  if (lane == 0) done_next[s] = 1;
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

__global__ void __launch_bounds__(32) k_generic_fit(GenericFitDev d) {
  __shared__ SpotShared<double> sh_s[1];
  __shared__ FitResult res_s[1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long b = (long long)blockIdx.x + warp;
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
int engine_smem_bytes(int K, int team_warps) { return (int)EngineSmem(K, team_warps).total; }

int launch_init_window(const FitDev& d, cudaStream_t st) {
  const long long total = d.n * d.K;
  if (total == 0) return 0;
  k_init_window<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d);
  IA3_LAUNCH_CHECK();
  return 0;
}

int launch_build_neighbours(const FitDev& d, const CellGrid& g, int* cell_cnt, int* cell_start, int* cell_cur, int* order, cudaStream_t st) {
  if (d.n == 0) return 0;
  const unsigned nb = (unsigned)((d.n + 255) / 256);
  IA3_CUDA(cudaMemsetAsync(cell_cnt, 0, sizeof(int) * (size_t)g.ncell, st));
  IA3_CUDA(cudaMemsetAsync(cell_cur, 0, sizeof(int) * (size_t)g.ncell, st));
  k_cells_count<<<nb, 256, 0, st>>>(d.centers, d.n, g, cell_cnt);
  IA3_LAUNCH_CHECK();
  k_scan_excl<<<1, 1024, 0, st>>>(cell_cnt, cell_start, g.ncell);
  IA3_LAUNCH_CHECK();
  k_cells_fill<<<nb, 256, 0, st>>>(d.centers, d.n, g, cell_start, cell_cur, order);
  IA3_LAUNCH_CHECK();
  k_nbr_build<<<nb, 256, 0, st>>>(d, g, cell_start, order);
  IA3_LAUNCH_CHECK();
  return 0;
}

int launch_build_bricks(const FitDev& d, cudaStream_t st) {
  const long long tab_n = (long long)d.nbz * d.nbx * d.nby;
  IA3_CUDA(cudaMemsetAsync(d.brick_tab, 0xFF, sizeof(int) * (size_t)tab_n, st));
  if (d.n == 0) return 0;
  k_brick_mark<<<(unsigned)((d.n + 255) / 256), 256, 0, st>>>(d);
  IA3_LAUNCH_CHECK();
  k_brick_assign<<<(unsigned)((tab_n + 255) / 256), 256, 0, st>>>(d, tab_n);
  IA3_LAUNCH_CHECK();
  return 0;
}

int launch_voronoi(const FitDev& d, cudaStream_t st) {
  if (d.n == 0) return 0;
  k_voronoi<<<(unsigned)((d.n + WARPS - 1) / WARPS), WARPS * 32, 0, st>>>(d);
  IA3_LAUNCH_CHECK();
  return 0;
}

int launch_member_stats(const FitDev& d, cudaStream_t st) {
  if (d.n == 0) return 0;
  k_member_stats<<<(unsigned)((d.n + WARPS - 1) / WARPS), WARPS * 32, 0, st>>>(d);
  IA3_LAUNCH_CHECK();
  return 0;
}

int lw_prof_read(unsigned long long* out16) {
#ifdef IA3_FIT_PROF
  IA3_CUDA(cudaMemcpyFromSymbol(out16, lw::g_lw_prof, sizeof(unsigned long long) * 16));
  return 0;
#else
  for (int i = 0; i < 16; ++i) out16[i] = 0;
  return 0;
#endif
}

int launch_engine_reset(const FitDev& d, cudaStream_t st) {
  k_engine_init<<<(unsigned)((std::max<long long>(d.n, 1) + 255) / 256), 256, 0, st>>>(d);
  IA3_LAUNCH_CHECK();
  return 0;
}

int launch_sched(const FitDev& d, int round, int phases, int sweep_cap, cudaStream_t st) {
  const long long threads = std::max<long long>(std::max<long long>(d.n, d.list_cap), 1);
  k_sched<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(d, round, phases, sweep_cap);
  IA3_LAUNCH_CHECK();
  return 0;
}

// grid_hint: upper bound of the CTAs worth launching (the kernels claim tasks from a list until it is
// empty, so any grid is correct).  A CTA that finds the list empty exits at once, but launching and
// retiring it still occupies the GPU's work distributor, which every stack in flight shares: the first
// rounds of a run (one task per seed) get a grid for the whole GPU, later rounds (a few % of the seeds)
// a small one.
int launch_fit_round(const FitDev& d, int round, bool team, int grid_hint, cudaStream_t st) {
  if (d.n == 0) return 0;
  constexpr int kMaxDynSmem = 227 * 1024 - 2048;
  const int smem1 = engine_smem_bytes(d.K, 1), smemT = engine_smem_bytes(d.K, TEAM_WARPS);
  if (smem1 > kMaxDynSmem || smemT > kMaxDynSmem) { set_error("radius_fit too large for the shared-memory window"); return -1; }
  // once per process and device: changing a function attribute while an instance of the kernel is
  // running (another stack's round) would serialise the streams
  static std::once_flag once;
  static cudaError_t once_err = cudaSuccess;
  std::call_once(once, [] {
    once_err = cudaFuncSetAttribute(k_fit_round<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
    if (once_err == cudaSuccess)
      once_err = cudaFuncSetAttribute(k_fit_round<TEAM_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
  });
  IA3_CUDA(once_err);
  const long long want = std::max<long long>(1, std::min<long long>(2 * d.n, grid_hint));
  if (team) k_fit_round<TEAM_WARPS><<<(unsigned)want, TEAM_WARPS * 32, smemT, st>>>(d, round);
  else k_fit_round<1><<<(unsigned)want, 32, smem1, st>>>(d, round);
  IA3_LAUNCH_CHECK();
  return 0;
}

int launch_subtract_dense(const FitDev& d, double* dense, uint8_t* done, uint8_t* done_next, int* pending, cudaStream_t st) {
  if (d.n == 0) return 0;
  k_subtract_dense<<<(unsigned)((d.n + WARPS - 1) / WARPS), WARPS * 32, 0, st>>>(d, dense, done, done_next, pending);
  IA3_LAUNCH_CHECK();
  return 0;
}

int launch_generic_fit(const GenericFitDev& d, cudaStream_t st) {
  if (d.n == 0) return 0;
  k_generic_fit<<<(unsigned)d.n, 32, 0, st>>>(d);
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

// out[i] = vol[idx[i]] for a list of voxel indices (the stack's dtype: 2 or 4 bytes per voxel)
namespace ia3 {
template <typename T>
__global__ void k_gather_vox(const T* __restrict__ vol, const long long* __restrict__ idx, long long n, T* __restrict__ out) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = vol[idx[i]];
}
int launch_gather_u16(const void* vol, int dtype, const long long* idx, long long n, void* out, cudaStream_t st) {
  if (n == 0) return 0;
  const int blocks = (int)std::min<long long>((n + 255) / 256, 1184);
  if (dtype == 0) k_gather_vox<uint16_t><<<blocks, 256, 0, st>>>((const uint16_t*)vol, idx, n, (uint16_t*)out);
  else if (dtype == 1) k_gather_vox<float><<<blocks, 256, 0, st>>>((const float*)vol, idx, n, (float*)out);
  else k_gather_vox<double><<<blocks, 256, 0, st>>>((const double*)vol, idx, n, (double*)out);
  IA3_LAUNCH_CHECK();
  return 0;
}
}  // namespace ia3
