// Host-callable launchers of the fit stage kernels (fit_kernels.cu) and the device-side state of the
// fit engine.
//
// The engine runs firstfit and ALL sweeps of repeatfit (External/Fitting_v4.py:590-683) without a host
// round trip per level or per sweep.  Work is cut into tasks, one per (seed, step):
//   F  firstfit of a seed        (image data, Voronoi members, delta = min_delta_center)
//   B  im_subtr[window] -= rec   (in seed order among overlapping windows)
//   R  one repeatfit visit       (im_add + own reconstruction, full window, delta = max_delta_center)
//   S  speculative R of sweep 1  (see "memo" below)
// and a task is issued as soon as the tasks the reference would have executed before it AND whose
// output it reads are done (per-seed dataflow): R_k of seed i needs R_k of every lower-index seed whose
// window overlaps i's, and R_{k-1} of every higher-index one -- exactly the in-order Gauss-Seidel sweep
// of the reference, without its global order.  The convergence rule (distance between
// consecutive centres < max_dist_th, frozen once converged, at most n_max_iter + 1 sweeps) is per seed,
// so it is evaluated on the device when a seed's visit ends.
// Execution is in ROUNDS of three stream-ordered launches: k_sched builds the round's work lists from
// the per-seed state; k_fit_round<1> runs every listed task with one warp, for at most cap_bulk function
// evaluations; k_fit_round<TEAM_WARPS> continues long runs (>= team_after evaluations so far) with a
// team of warps per spot.  A run that is not finished when its cap is reached parks the live part of
// its lmder state (LMLive, 728 B) and is continued next round.  Every kernel terminates; nothing spins.
// memo: GaussianFit is a deterministic function of (float32 window values, x0, delta).  A seed that is
// revisited with exactly the inputs of its previous visit (every isolated seed in sweep 2; every seed
// whose neighbours have frozen) gets that visit's result without running lmder again.  The same
// comparison validates the speculative task S: for a seed whose window overlaps no other window the
// inputs of its first repeat visit are, almost surely, the image values themselves, so that visit is
// started together with firstfit and adopted when the real inputs turn out to be identical.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "gauss_model.h"
#include "lm_core.h"

namespace ia3 {

struct LMLive;                        // fit_spot.h

constexpr int TEAM_WARPS = 8;         // warps per spot in the continuation kernel

enum { TASK_F = 0, TASK_R = 1, TASK_S = 2, TASK_B = 3 };
constexpr unsigned TASK_SEED_MASK = 0x0fffffffu;
constexpr unsigned TASK_RESUME = 1u << 30;
__host__ __device__ inline unsigned make_task(int seed, int kind, bool resume) {
  return (unsigned)seed | ((unsigned)kind << 28) | (resume ? TASK_RESUME : 0u);
}

struct EngineCtl {
  // per round parity: work lists filled by k_sched, consumed through the cursors by the fit kernels;
  // park list of parity p^1 filled by the fit kernels of round parity p
  int n_bulk[2], cur_bulk[2], n_team[2], cur_team[2], n_park[2], alive[2], blocks_done[2];
  int cap_team_now;                   // evaluations per team round (long when nothing else is pending)
  int done;
  int all_ok0, all_ok1;               // every seed holds a fit after firstfit / after sweep 1 (dtype of the distance test)
  int pool_v, pool_o, overflow;       // neighbour pools (k_nbr_build)
  int n_bricks;
  int tie_count;
  // statistics (ia3_fit_engine_stats)
  int st_lm_runs, st_memo_hits, st_spec_runs, st_spec_hits, st_parked, st_team_tasks, st_tasks, st_rounds;
  unsigned long long st_evals;
  // per-round trace (ring of 512): (n_bulk << 16 | n_team), and the device clock in ns when the round's k_sched ended
  unsigned trace_work[512];
  unsigned long long trace_ns[512];
  // cycle counters of the team kernel's phases (IA3_FIT_PROF builds): outer, lmpar, consts, pass, judge, evals
  unsigned long long prof[8];
};

struct FitDev {
  // image / volumes
  const void* im; int im_dtype;       // original stack (u16 / f32 / f64)
  // float64 work volume (im_subtr -> im_add), stored sparsely: only 8x8x8 bricks touched by some
  // seed's window exist.  brick_tab[(bz * nbx + bx) * nby + by] = brick number (or < 0); a brick is
  // 512 doubles at vol + 512 * number.  A dense copy would be 1.7 GB per 50x2048x2048 stack.
  double* vol;
  int* brick_tab;
  int nbz, nbx, nby;
  int Z, X, Y;
  // seeds
  long long n;
  const double* centers;              // n x 3
  int* own_id;                        // v3: lowest index among seeds with identical coordinates
  int* nbr_start; int* nbr_cnt; int* nbr_idx;     // seeds that can own a voxel of this seed's window (pool)
  int* dep_start; int* dep_cnt; int* dep_idx;     // seeds whose window overlaps this seed's window (pool)
  int* n_lower;                       // how many of those have a lower index
  int pool_cap_v, pool_cap_o;
  // window
  int K, KW, radius;                  // voxels in the ball; mask words per seed
  const int8_t* offs;                 // K x 3 offsets
  uint32_t* mask;                     // n x KW membership bits
  // ties (v4)
  int tie_cap;
  int* tie_spot; int* tie_k;
  // results
  float* ps; double* p_raw; uint8_t* success; int* nfev; int* info;
  double* rec;                        // n x K reconstructions (ims_rec)
  double* praw_first; uint8_t* succ_first;        // firstfit's raw parameters (im_subtr is rebuilt from them on demand)
  // config
  FitParams fp; LMConfig lm; double init_w[3];
  double delta_first, delta_repeat, th2;
  int max_sweeps;
  // engine
  EngineCtl* ctl; int* h_done;
  int* stage;                         // -2 firstfit pending, -1 firstfit done / subtraction pending, k >= 0: k repeat visits done
  uint8_t* fin;                       // no more visits (converged, or max_sweeps reached)
  uint8_t* busy;                      // a task of this seed is listed, running or parked
  uint8_t* conv;                      // dists < max_dist_th^2 after the last visit
  uint8_t* specst;                    // speculative task: 0 none, 1 in flight, 2 finished
  uint8_t* succ_prev; float* cen_prev;            // success / centre before the seed's last visit
  double* dists;
  float* key_d32; double* key_x0;     // inputs of the seed's last lmder run in repeat mode (n x K, n x NP)
  double* memo_praw; int* memo_meta;  // its result: raw parameters; nfev, njev, info, unused
  uint8_t* memo_valid;                // 0 none, 1 running, 2 result available
  uint8_t* memo_committed;            // that result is what ps / rec of the seed hold
  LMLive* live;                       // 2 n parked-run slots: seed (F, R), n + seed (S)
  unsigned* lists; int list_cap;      // bulk[2], team[2], park[2], list_cap entries each
  int cap_bulk, cap_team_short, cap_team_long, team_after;
  int merge_small;            // a round whose one-warp + team work lists together hold <= this many tasks runs them all as team tasks
  int memo_on;
};

struct CellGrid { double lo[3]; double cs; int g[3]; long long ncell; };

int engine_smem_bytes(int K, int team_warps);
int launch_init_window(const FitDev& d, cudaStream_t st);
// neighbour pools, dependency counts, brick table: all on the device (no host pass over the seeds)
int launch_build_neighbours(const FitDev& d, const CellGrid& g, int* cell_cnt, int* cell_start, int* cell_cur, int* order, cudaStream_t st);
int launch_build_bricks(const FitDev& d, cudaStream_t st);
int launch_voronoi(const FitDev& d, cudaStream_t st);
int launch_member_stats(const FitDev& d, cudaStream_t st);
// one engine round.  phases: bit 0 firstfit (F, B), bit 1 repeatfit (R), bit 2 speculation (S);
// sweep_cap: no R visit beyond this sweep number (host-driven sweep-by-sweep mode)
int launch_sched(const FitDev& d, int round, int phases, int sweep_cap, cudaStream_t st);
int launch_fit_round(const FitDev& d, int round, bool team, int grid_hint, cudaStream_t st);
int launch_engine_reset(const FitDev& d, cudaStream_t st);
int lw_prof_read(unsigned long long* out16);    // IA3_FIT_PROF builds: cycle counters of lm_warp.h since process start

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

int launch_apply_ties(uint32_t* mask, int KW, const int* tie_spot, const int* tie_k, const uint8_t* keep, int n, cudaStream_t st);
int launch_window_copy(const FitDev& d, double* snap, double* vol_out, int dir, cudaStream_t st);
// one pass of "dense[window] -= reconstruction of firstfit" over the seeds whose lower-index overlapping
// seeds are done (done: n bytes, pending: device counter of seeds still to do)
int launch_subtract_dense(const FitDev& d, double* dense, uint8_t* done, uint8_t* done_next, int* pending, cudaStream_t st);
int launch_to_f64(const void* im, int dtype, double* out, long long n, cudaStream_t st);
int launch_eval_f0(const FitParams& fp, const double* p_raw, const double* center, const float* coords, long long m,
                   double* out, cudaStream_t st);
int launch_gather_u16(const void* vol, int dtype, const long long* idx, long long n, void* out, cudaStream_t st);

}  // namespace ia3
