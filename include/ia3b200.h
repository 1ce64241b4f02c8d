/* ia3b200.h -- C ABI of libia3b200.so: B200 (sm_100a) spot finding for 3D FISH stacks.
 *
 * The reference (zhengpuas47/ImageAnalysis3) has no FFI for this path: the boundary is a set of
 * Python call signatures.  Each entry point below is what the Python mirror in
 * imageanalysis3_b200/{spot_tools/fitting.py, External/Fitting_v4.py, External/Fitting_v3.py,
 * visual_tools.py} binds through ctypes, and cites the reference code it replaces.
 *
 * Conventions: plain pointers and sizes only; every pointer is HOST memory unless the name
 * starts with d_; stacks are C-contiguous (Z, X, Y); return value 0 = OK, negative = error
 * (ia3_last_error() gives the text).  Calls on one handle are not thread-safe; different
 * handles may be used from different threads.  CUDA is initialised lazily inside the calling
 * process (fork-safe as long as the parent has not called into the library).
 */
#ifndef IA3B200_H
#define IA3B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IA3_DTYPE_U16 0
#define IA3_DTYPE_F32 1
#define IA3_DTYPE_F64 2   /* fit stage only */

typedef struct ia3_stack ia3_stack;   /* one image stack resident in HBM + its work buffers */
typedef struct ia3_fit ia3_fit;       /* one iter_fit_seed_points object on a stack */

/* ---- process / device ---------------------------------------------------------------- */
int ia3_init(int device);                 /* select device (default: IA3_DEVICE env or 0) */
const char* ia3_last_error(void);
int ia3_version(void);
int ia3_device_sm_count(void);
/* number of kernels launched by this library in this process (bench.py's gpu_launches) */
int64_t ia3_launch_count(void);
/* wall-clock time spent inside the entry points so far (text table; only filled when the process was
 * started with IA3_STATS=1) */
int ia3_debug_stats(char* buf, int cap);
/* CUDA-event stopwatch on the library's stream (all kernels and copies of this process are issued
 * on it): start records an event after a device sync, stop returns the elapsed device time. */
int ia3_timer_start(void);
int ia3_timer_stop(float* ms);

/* ---- stacks ---------------------------------------------------------------------------- */
/* Upload a host stack -- one H2D pass per stack; seed and fit stages then share the resident copy.
 * Pinned memory goes to the copy engine directly; pageable memory (a plain numpy array, what the
 * reference's callers hold) is staged through pinned chunks by a few worker threads.  Replaces the implicit "im" argument of get_seeds / fit_fov_image /
 * iter_fit_seed_points (spot_tools/fitting.py:20,169; External/Fitting_v4.py:560). */
int ia3_stack_create(const void* im, int dtype, int Z, int X, int Y, ia3_stack** out);
/* Wrap a stack that is already in device memory (no copy, not owned). */
int ia3_stack_wrap_device(const void* d_im, int dtype, int Z, int X, int Y, ia3_stack** out);
int ia3_stack_destroy(ia3_stack* s);
/* Give device memory back early (to the library's allocation cache, i.e. to the other stacks in
 * flight): what & 1 = the seed stage's work volumes and candidate list (after ia3_seed_fetch);
 * what & 2 = the library's own copy of the image (after ia3_fit_first_run: repeat sweeps only touch
 * the sparse work volume).  A later call that needs what was released fails with an error. */
int ia3_stack_trim(ia3_stack* s, int what);

/* counts[v] = number of voxels with value v of a uint16 stack (65536 entries).  The percentile thresholds of the
 * seeders (scipy.stats.scoreatpercentile over the whole image: spot_tools/fitting.py:75-76, visual_tools.py:
 * 1808-1811) are order statistics, which the host reads off this histogram exactly. */
int ia3_stack_histogram(ia3_stack* s, uint64_t* counts);

/* ---- pre-processing (the caller upstream of the seed stage) ------------------------------- */
/* The compute core of io_tools/load.py:185-520 correct_fov_image on resident uint16 channel stacks; the file reading
 * (DaxReader, io_tools/load.py:290-300) and the profile loading stay on the host side of this boundary. */
/* An owned, uninitialised stack: the destination of ia3_corr_mix / ia3_corr_warp. */
int ia3_stack_alloc(int dtype, int Z, int X, int Y, ia3_stack** out);
/* Per-dataset constants (correction profiles) kept in device memory across calls: ia3_corr_mix / ia3_corr_warp take either
 * host pointers (copied on every call) or pointers returned here.  Pageable sources are staged like stacks. */
int ia3_device_upload(const void* host, size_t bytes, void** dev);
int ia3_device_free(void* dev);
/* Copy a resident stack's image to the host (Z*X*Y elements of its dtype). */
int ia3_stack_fetch(ia3_stack* s, void* out);
/* corrections.py:490-510 Remove_Hot_Pixels(im.astype(float32), dtype=uint16, hot_pix_th, hot_th), in place: columns
 * brighter than hot_th x the mean of their np.roll neighbours in more than hot_pix_th x Z planes are replaced, in
 * np.where order, by the float32 mean of their four neighbours.  *n_hot = number of such columns. */
int ia3_corr_hot_pixels(ia3_stack* s, double hot_th, double hot_pix_th, int64_t* n_hot);
/* corrections.py:479-487 Z_Shift_Correction(im.astype(float32), dtype=uint16), in place: every plane divided by its median and multiplied
 * by the stack's median (float32, numpy's order), truncated to uint16.  io_tools/load.py:336-345. */
int ia3_corr_zshift(ia3_stack* s);
/* correction_tools/filter.py:14-19 gaussian_high_pass_filter, in place: im - gaussian_filter(im, sigma, mode='nearest', truncate) where
 * positive, else 0 (io_tools/load.py:487-497).  w_half: r + 1 taps, centre to edge, of scipy's normalised kernel (radius r = int(truncate * sigma + 0.5)). */
int ia3_corr_highpass(ia3_stack* s, const double* w_half, int r);
/* out = illumination(bleed-through(ins)) (io_tools/load.py:347-381): with ``bleed`` (host, n_in x X x Y: row i of the
 * (n, n, X, Y) profile) out = clip(sum_j ins[j] * bleed[j]) truncated to uint16, else out = ins[0]; with ``illum``
 * (host, X x Y) that result is divided by it and truncated again.  profile_f64: the profiles are float64 (numpy then
 * computes in float64) instead of float32.  bleed / illum: host or ia3_device_upload pointers.  Bleed-through mixing cannot
 * run in place. */
int ia3_corr_mix(ia3_stack* const* ins, int n_in, const void* bleed, const void* illum, int profile_f64, ia3_stack* out);
/* out = scipy.ndimage.map_coordinates(in, grid + chroma - drift, order 3, mode 'nearest') rounded to uint16
 * (io_tools/load.py:424-459).  drift: 3 doubles (z, x, y) -- the reference's float32 drift widened, or align_image's float64 result -- or null; chroma: float32 (or float64 with chroma_f64) array
 * (3, chroma_z, X, Y) with chroma_z = 1 or Z (what correction_tools/chromatic.py:282-289 saves), host or ia3_device_upload
 * pointer, or null.  Floating-point path: the spline coefficients agree with scipy's recursion to ~1e-11
 * of a count, so the rounded uint16 output is equal except where a value falls within that of a half-integer. */
int ia3_corr_warp(ia3_stack* in, const double* drift, const void* chroma, int chroma_f64, int chroma_z, ia3_stack* out);

/* ---- seed stage ------------------------------------------------------------------------ */
typedef struct {
  /* half kernels: w[j] = weight at distance j from the centre tap, j = 0..r (scipy
   * _gaussian_kernel1d, truncate=4 -> r = int(4*sigma+0.5)); r < 0 = "no blur" (gfilt_size
   * falsy, spot_tools/fitting.py:93-94,100-101) */
  const double* w_fg; int r_fg;
  const double* w_bg; int r_bg;
  int filt_size;          /* maximum_filter / minimum_filter size (3) */
  int variant;            /* 0: spot_tools.fitting.get_seeds (:91-125)
                             1: visual_tools.get_seed_points_base (visual_tools.py:348-367) */
  double edge;            /* variant 0: min_edge_distance (keep d <= c <= size-d); <=0 = off */
  double h_min;           /* keep candidates with h >= h_min (host folds >, >=, float32/float64
                             comparison semantics into this number) */
  int two_d;              /* the stack is a 2-D image held as (1, X, Y): filters and edge test on x, y only */
} ia3_seed_cfg;

typedef struct {
  float ms_gauss_fg, ms_gauss_bg, ms_rank, ms_compact, ms_total;
} ia3_seed_timing;

/* Runs the whole device part of the seed stage: two separable Gaussian blurs with scipy's exact
 * integer semantics (per-axis uint16 truncation, reflect boundary, FP64 symmetric-pair order),
 * 3D max / min rank filters, mask, threshold, edge filter, and an ordered (C order: z, x, y)
 * prefix-sum stream compaction.  Candidates stay on the device until fetched.
 * Replaces spot_tools/fitting.py:91-125 and visual_tools.py:350-367. */
int ia3_seed_run(ia3_stack* s, const ia3_seed_cfg* cfg, int64_t* n_candidates, ia3_seed_timing* t);
/* Copy candidates to the host: zxy is n x 3 int32 (C order of the stack), h is n float32
 * (variant 0: max_im - min_im; variant 1: max_im - min_filter(min_im)). */
int ia3_seed_fetch(ia3_stack* s, int32_t* zxy, float* h, int64_t cap);
/* Debug/parity taps: copy an intermediate volume back (which: 0 = foreground blur,
 * 1 = background blur), same dtype as the stack. */
int ia3_seed_fetch_volume(ia3_stack* s, int which, void* out);
/* Values of such a volume at n voxels given as flat C-order indices (External/Fitting_v3.py:276-283
 * reads its two blurs at the candidate voxels only); out: n values of the stack's dtype. */
int ia3_seed_gather_volume(ia3_stack* s, int which, const int64_t* flat_idx, int64_t n, void* out);

/* ---- alternative seeders of External/Fitting_v4.py (no caller inside the reference) ------------- */
/* get_seed_points_base_v2 (Fitting_v4.py:95-126): im_norm = float32(im) - cv2.blur(float32(im), (gfilt_size,
 * gfilt_size)) per z-slice (bit-identical to cv2 4.13), std = np.std(im_norm) (FP64 on the device: agrees with
 * numpy's float32 pairwise value to ~1e-7 relative), candidates = voxels with im_norm > float32(std) * th_seed that
 * are >= all (2 (filt_size / 2) + 1)^3 neighbours taken modulo the shape.  flat_idx / h: C-order voxel indices and
 * im_norm values, UNORDERED (the caller sorts); *n_out = number found (-2 is returned if it exceeds cap). */
int ia3_seed_v2(ia3_stack* s, int gfilt_size, int filt_size, double th_seed, double* std_out, int64_t* flat_idx, float* h,
                int64_t cap, int64_t* n_out);
/* fft_gaussian_fast (Fitting_v4.py:66-70): the reference's reflect() padding and 'valid' convolution with the
 * normalised int(gaus * exp)-tap Gaussian windows, evaluated directly in FP64 (the reference multiplies
 * single-precision FFTs: agreement ~1e-6 relative); out = Z*X*Y float64 on the host. */
int ia3_fft_gaussian(ia3_stack* s, const double* gaus3, int exp, double* out);
/* get_seed_points_base (Fitting_v4.py:72-92): im_diff = log(im) - log(fft_gaussian_fast(im, [gfilt_size] * 3)),
 * candidates = voxels equal to the maximum of their filt_size^3 neighbourhood with im_diff > th_seed * std(im_diff). */
int ia3_seed_logratio(ia3_stack* s, double gfilt_size, int filt_size, double th_seed, double* std_out, int64_t* flat_idx, double* h,
                      int64_t cap, int64_t* n_out);

/* ---- local background (fit_fov_image's normalize_local / normalize_background) ------------ */
/* find_image_background (io_tools/load.py:642-686) for n boxes of a uint16 stack: mode of the
 * histogram with bins arange(first, last, bin_size) found with scipy.signal.find_peaks' rule and the
 * reference's height loop, np.nanmedian of the box if that loop fails.  boxes is n x 6 int32
 * (z0, z1, x0, x1, y0, y1; half open, as the slices of generate_neighboring_crop,
 * io_tools/crop.py:59-88); out receives n float64. */
int ia3_box_background(ia3_stack* s, const int32_t* boxes, int64_t n, int first, int last, int bin_size, int max_iter,
                       double* out);

/* ---- fit stage ------------------------------------------------------------------------- */
typedef struct {
  int personality;      /* 4: External/Fitting_v4.py GaussianFit; 3: External/Fitting_v3.py */
  int radius;           /* radius_fit (5): window = offsets -r..r-1 with d^2 <= r^2 */
  double min_w, max_w;  /* sigma bounds (0.5, 4) */
  double init_w[3];     /* v4: scalar init_w replicated; v3: per-axis init_w (_sigma_zxy) */
  double weight_sigma;  /* v3 width prior (0 = off) */
  int maxfev;           /* 1000 (v4, Fitting_v4.py:388) / 1100 (v3, leastsq default) */
  int eval_fp32;        /* reserved, must be 0 (the model is evaluated in FP64, the reference's precision) */
} ia3_fit_cfg;

/* iter_fit_seed_points.__init__ (Fitting_v4.py:560-588 / Fitting_v3.py:313-335):
 * centers is n x 3 float64 (z, x, y). */
int ia3_fit_create(ia3_stack* s, const double* centers_zxy, int64_t n, const ia3_fit_cfg* cfg, ia3_fit** out);
int ia3_fit_destroy(ia3_fit* f);

/* firstfit, step 1: Voronoi membership of every window voxel (v4: cKDTree nearest seed,
 * Fitting_v4.py:422-424,612; v3: brute-force argmin, lowest index wins, Fitting_v3.py:40-47).
 * v4 only: voxels whose nearest seed is not unique are reported so that the caller can resolve
 * them with the very same cKDTree (its tie-break is an implementation detail of scipy). */
int ia3_fit_first_prepare(ia3_fit* f, int64_t* n_ties);
int ia3_fit_first_ties(ia3_fit* f, int32_t* spot, int32_t* zxy, int64_t cap);
int ia3_fit_first_resolve(ia3_fit* f, const uint8_t* keep, int64_t n);
/* firstfit / repeatfit on the device, without a host round trip per sweep (Fitting_v4.py:590-683,
 * Fitting_v3.py:337-421).  phases: 1 = firstfit, 2 = repeatfit (firstfit must have run), 3 = both in one
 * go.  The per-seed rule of the reference's loop is evaluated where a seed's visit ends: a seed is
 * visited in sweep k+1 iff sum((centre_k - centre_{k-1})^2) >= max_dist_th2 (in the dtype
 * np.array(centers_fit) has in the reference), and at most n_max_iter + 1 sweeps are made; a visit
 * starts as soon as the visits the reference orders before it, and whose output it reads, are done.
 * Outputs (each pointer may be NULL): final rows of ps / p_raw / success / nfev / info; converged and
 * dists as after the last sweep; n_visits = sweeps in which the seed was visited (n_iter is their
 * maximum); success_old / centers_old = the seed's success flag / centre before its last visit. */
typedef struct {
  float* ps; double* p_raw; uint8_t* success; int32_t* nfev; int32_t* info;   /* n x 11, n x 10, n, n, n */
  uint8_t* converged; double* dists; int32_t* n_visits;                       /* n each */
  uint8_t* success_old; float* centers_old;                                   /* n, n x 3 */
} ia3_fit_out;
int ia3_fit_run(ia3_fit* f, int phases, double min_delta_center, double max_delta_center, double max_dist_th2,
                int n_max_iter, const ia3_fit_out* out);
/* firstfit alone (= ia3_fit_run(phases = 1)): all fits + ordered subtraction of the reconstructions
 * (Fitting_v4.py:606-639).  Outputs (each may be NULL): ps n x 11 float32 (NaN rows on
 * failure), p_raw n x 10 float64, success n, nfev n, info n. */
int ia3_fit_first_run(ia3_fit* f, double delta_center, float* ps, double* p_raw, uint8_t* success,
                      int32_t* nfev, int32_t* info);
/* ONE sweep of repeatfit over the seeds with active[i] != 0, in seed order, for callers that run the
 * reference's loop themselves (Fitting_v4.py:651-675).  Only rows of active seeds change. */
int ia3_fit_repeat_sweep(ia3_fit* f, double delta_center, const uint8_t* active, float* ps, double* p_raw,
                         uint8_t* success, int32_t* nfev, int32_t* info);
/* counters of the fit engine since ia3_fit_create: rounds, tasks, lmder runs, function evaluations, memo
 * hits, speculative runs, speculative runs adopted, suspensions, team continuations, bricks */
int ia3_fit_engine_stats(ia3_fit* f, int64_t* out, int cap);
/* lazily materialised attributes: which = 0 -> im_subtr (rebuilt from firstfit's parameters), 1 -> im_add
 * (current work volume); float64 volume of the stack's shape. */
int ia3_fit_get_volume(ia3_fit* f, int which, double* out);
/* ims_rec[i]: reconstruction over the clipped window of seed i (float64, <= K values);
 * also returns the window voxel coordinates (count x 3 int32) if zxy != NULL. */
int ia3_fit_get_rec(ia3_fit* f, int64_t i, double* rec, int32_t* zxy, int32_t* count);
int ia3_fit_num_levels(ia3_fit* f);
float ia3_fit_last_ms(ia3_fit* f);        /* device time of the last ia3_fit_run / first_run / repeat_sweep */

/* ---- standalone GaussianFit ------------------------------------------------------------ */
/* A batch of independent GaussianFit(im, X, center, ...).fit() calls (Fitting_v4.py:165-396,
 * Fitting_v3.py:50-257).  Problem b uses values[off[b]:off[b+1]] (float64, as handed to
 * GaussianFit) and coords[3*off[b] : 3*off[b+1]] (k x 3, float32-representable).
 * rec (optional, same layout as values) receives get_im() on the same coordinates. */
int ia3_gaussfit_batch(const ia3_fit_cfg* cfg, double delta_center, int64_t n_problems, const int64_t* off,
                       const double* values, const float* coords, const double* centers,
                       float* ps, double* p_raw, uint8_t* success, int32_t* nfev, int32_t* info, double* rec);

/* Fitting_v4.fast_fit_big_image with better_fit=False (External/Fitting_v4.py:494-556) = gfit_fast
 * (:433-447) for every seed: weighted-moment estimate [h, z, x, y, background, cov_zz, cov_xx, cov_yy,
 * cov_zx, cov_zy, cov_xy, nan] over the seed's ball (offsets -r..r-1, d^2 <= r^2), restricted to the
 * voxels nearer to it than to any other seed within 2r when avoid_neighbors != 0, recentred on the
 * brightest voxel when recenter != 0.  centers is n x 3 float64, out n x 12 float64 (NaN rows for
 * empty windows).  Weights are formed in the image's dtype like the reference does (uint16 wraps). */
typedef struct {
  int radius;            /* radius_fit (4) */
  int avoid_neighbors;   /* avoid_neigbors (sic) */
  int recenter;
  double bk_f;           /* background quantile (0.1) */
} ia3_moment_cfg;
int ia3_moment_fit(ia3_stack* s, const double* centers_zxy, int64_t n, const ia3_moment_cfg* cfg, double* out);

/* GaussianFit.get_im() (Fitting_v4.py:394-396): Gaussian part exp(h - q/2) of the model with raw
 * parameters p_raw (10) on m coordinates (m x 3 float32). */
int ia3_gauss_eval(const ia3_fit_cfg* cfg, double delta_center, const double* p_raw, const double* center,
                   const float* coords, int64_t m, double* out);

#ifdef __cplusplus
}
#endif
#endif
