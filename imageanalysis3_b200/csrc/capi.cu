// C ABI of libia3b200.so (declared in include/ia3b200.h): handles, host-side orchestration
// (buffer management, neighbour lists, dependency levels) and kernel launches.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <thread>
#include <time.h>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "../../include/ia3b200.h"
#include "fit_kernels.h"
#include "fit_spot.h"
#include "ia3_device.h"
#include "seed_kernels.h"

namespace ia3 {

static thread_local std::string g_err;
std::atomic<int64_t> g_launches{0};
static int g_device = -1;
static cudaStream_t g_stream = nullptr;          // standalone GaussianFit batches and the stopwatch
static cudaEvent_t g_t0 = nullptr, g_t1 = nullptr;
static std::mutex g_mu;

// Every stack owns a stream (taken from a small pool): stacks handled by different host threads
// overlap on the device -- the copy of one stack with the seed kernels of the next and with the
// long tail of a third one's fit sweeps (a few junk seeds run MINPACK to maxfev while the rest of
// the GPU would otherwise idle).
static std::vector<cudaStream_t> g_stream_pool;
static int acquire_stream(cudaStream_t* out) {
  {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_stream_pool.empty()) { *out = g_stream_pool.back(); g_stream_pool.pop_back(); return 0; }
  }
  IA3_CUDA(cudaStreamCreateWithFlags(out, cudaStreamNonBlocking));
  return 0;
}
static cudaStream_t g_upload_stream = nullptr;
static std::mutex g_upload_mu;
static int upload_stream(cudaStream_t* out) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!g_upload_stream) IA3_CUDA(cudaStreamCreateWithFlags(&g_upload_stream, cudaStreamNonBlocking));
  *out = g_upload_stream;
  return 0;
}
static int global_stream(cudaStream_t* out) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!g_stream) IA3_CUDA(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking));
  *out = g_stream;
  return 0;
}
// Events and pinned staging buffers are pooled like streams and device blocks: in steady state a stack
// in flight makes no call that creates or destroys a driver object (those calls wait for the device
// while other stacks' kernels run, and every other host thread then queues behind them).
static std::vector<cudaEvent_t> g_event_pool;
static int acquire_event(cudaEvent_t* out) {
  {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_event_pool.empty()) { *out = g_event_pool.back(); g_event_pool.pop_back(); return 0; }
  }
  IA3_CUDA(cudaEventCreate(out));
  return 0;
}
static void release_event(cudaEvent_t e) {
  if (!e) return;
  std::lock_guard<std::mutex> lk(g_mu);
  g_event_pool.push_back(e);
}
static std::multimap<size_t, void*> g_host_free;
static std::unordered_map<void*, size_t> g_host_sizes;
static int host_alloc(void** p, size_t bytes) {
  size_t cls = 4096;
  while (cls < bytes) cls <<= 1;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_host_free.find(cls);
    if (it != g_host_free.end()) { *p = it->second; g_host_free.erase(it); return 0; }
  }
  IA3_CUDA(cudaMallocHost(p, cls));
  std::lock_guard<std::mutex> lk(g_mu);
  g_host_sizes[*p] = cls;
  return 0;
}
static void host_free(void* p) {
  if (!p) return;
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_host_sizes.find(p);
  if (it != g_host_sizes.end()) g_host_free.emplace(it->second, p);
}
static int reserve_pinned(void** buf, size_t* cap, size_t bytes) {
  if (bytes <= *cap) return 0;
  host_free(*buf);
  *buf = nullptr; *cap = 0;
  if (host_alloc(buf, bytes)) return -1;
  size_t cls = 4096;
  while (cls < bytes) cls <<= 1;
  *cap = cls;
  return 0;
}

// ---- small transfers without the copy engines ------------------------------------------------
// With many stacks in flight the H2D copy engine is busy with other stacks' 400 MB image uploads,
// and a cudaMemcpyAsync of a few KB (work lists, seed tables, results) queues behind them in the
// engine's FIFO: every such copy then costs one or more image uploads of latency.  Pinned host
// memory is device-addressable under UVA, so small transfers are done by a copy KERNEL on the
// stack's own stream instead (loads / stores over PCIe); only the image itself (and the optional
// whole-volume fetches) use the copy engines.  The host side of every transfer is a pinned arena.
constexpr size_t kSmallCopyMax = (size_t)16 << 20;
template <typename W>
__global__ void k_copy(W* __restrict__ dst, const W* __restrict__ src, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
static int small_copy(void* dst, const void* src, size_t bytes, cudaStream_t st) {
  if (bytes == 0) return 0;
  if (bytes > kSmallCopyMax) { IA3_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, st)); return 0; }
  const uintptr_t al = reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src) | bytes;
  auto blocks = [](size_t n) { return (unsigned)std::min<size_t>((n + 255) / 256, 296); };
  if ((al & 15) == 0) { const size_t n = bytes / 16; k_copy<uint4><<<blocks(n), 256, 0, st>>>((uint4*)dst, (const uint4*)src, n); }
  else if ((al & 3) == 0) { const size_t n = bytes / 4; k_copy<uint32_t><<<blocks(n), 256, 0, st>>>((uint32_t*)dst, (const uint32_t*)src, n); }
  else k_copy<uint8_t><<<blocks(bytes), 256, 0, st>>>((uint8_t*)dst, (const uint8_t*)src, bytes);
  IA3_LAUNCH_CHECK();
  return 0;
}

static void release_stream(cudaStream_t st) {
  if (!st) return;
  std::lock_guard<std::mutex> lk(g_mu);
  g_stream_pool.push_back(st);
}

void set_error(const std::string& msg) { g_err = msg; }

// ---- wall-clock accounting of the entry points (IA3_STATS=1; tools/trace_pipeline.py prints it) ----
struct StatSlot { const char* name; std::atomic<long long> ns{0}; std::atomic<long long> cpu_ns{0}; std::atomic<long long> calls{0}; };
static inline long long thread_cpu_ns() {
  timespec ts;
  clock_gettime(CLOCK_THREAD_CPUTIME_ID, &ts);
  return (long long)ts.tv_sec * 1000000000ll + ts.tv_nsec;
}
static StatSlot g_stats[24];
static std::atomic<int> g_nstats{0};
static const bool g_stats_on = getenv("IA3_STATS") != nullptr;
static StatSlot* stat_slot(const char* name) {
  const int n = g_nstats.load();
  for (int i = 0; i < n; ++i) if (g_stats[i].name == name) return &g_stats[i];
  std::lock_guard<std::mutex> lk(g_mu);
  const int m = g_nstats.load();
  for (int i = 0; i < m; ++i) if (g_stats[i].name == name) return &g_stats[i];
  if (m >= 24) return &g_stats[23];
  g_stats[m].name = name;
  g_nstats.store(m + 1);
  return &g_stats[m];
}
struct StatScope {
  StatSlot* s = nullptr;
  std::chrono::steady_clock::time_point t0;
  long long c0 = 0;
  explicit StatScope(const char* name) { if (g_stats_on) { s = stat_slot(name); t0 = std::chrono::steady_clock::now(); c0 = thread_cpu_ns(); } }
  ~StatScope() {
    if (s) {
      s->ns += std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
      s->cpu_ns += thread_cpu_ns() - c0;
      s->calls += 1;
    }
  }
};
#define IA3_STAT(name) ia3::StatScope _stat_scope(name)

// ---- caching allocator --------------------------------------------------------------------
static std::multimap<size_t, void*> g_free;
static std::unordered_map<void*, size_t> g_sizes;
static size_t g_cached_bytes = 0;
static const size_t kCacheLimit = (size_t)150 << 30;

// Request sizes are rounded up to size classes (powers of two up to 1 MiB, then eight classes per
// octave: at most 12.5 % over-allocation), and a freed block is reused only for its own class.  After
// a warm-up every allocation of a stack in flight is a cache hit: cudaMalloc while other stacks'
// kernels are running was measured at ~150 ms per call (it waits for the device), which throttled the
// whole pipeline.
static size_t size_class(size_t bytes) {
  if (bytes < 256) bytes = 256;
  size_t p2 = 256;
  while (p2 < bytes) p2 <<= 1;
  if (p2 <= ((size_t)1 << 20)) return p2;
  const size_t step = p2 >> 4;                   // p2/2 < bytes <= p2: eight steps of p2/16 above p2/2
  return (bytes + step - 1) / step * step;
}

int dev_alloc(void** p, size_t bytes) {
  IA3_STAT("dev_alloc");
  bytes = size_class(bytes);
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_free.lower_bound(bytes);         // best fit, at most twice the class (or 1 MiB for small ones)
    if (it != g_free.end() && it->first <= std::max<size_t>(2 * bytes, (size_t)1 << 20)) {
      *p = it->second;
      g_cached_bytes -= it->first;
      g_free.erase(it);
      return 0;
    }
  }
  StatScope _m("cudaMalloc");
  cudaError_t e = cudaMalloc(p, bytes);
  if (e != cudaSuccess) {
    dev_cache_clear();
    e = cudaMalloc(p, bytes);
  }
  if (e != cudaSuccess) {
    set_error(std::string("cudaMalloc failed: ") + cudaGetErrorString(e));
    return -1;
  }
  std::lock_guard<std::mutex> lk(g_mu);
  g_sizes[*p] = bytes;
  return 0;
}

void dev_free(void* p) {
  if (!p) return;
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_sizes.find(p);
  if (it == g_sizes.end()) return;
  if (g_cached_bytes + it->second > kCacheLimit) {
    cudaFree(p);
    g_sizes.erase(it);
    return;
  }
  g_free.emplace(it->second, p);
  g_cached_bytes += it->second;
}

void dev_cache_clear() {
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto& kv : g_free) { cudaFree(kv.second); g_sizes.erase(kv.second); }
  g_free.clear();
  g_cached_bytes = 0;
}

// Host threads that wait for a stack's stream should sleep, not spin: dozens of stacks are in flight
// per GPU (one host thread each), and spinning threads starve each other and the other ranks of the
// box.  The scheduling flag belongs to the device's primary context, so it has to be set for THAT device
// (cudaSetDeviceFlags would act on the calling thread's current device, i.e. device 0) and before the
// context exists; if the application created the context first the call fails harmlessly.
static void init_device_flags(int dev) {
  static std::once_flag once;
  std::call_once(once, [dev] {
    if (cudaInitDevice(dev, cudaDeviceScheduleBlockingSync, cudaInitDeviceFlagsAreValid) != cudaSuccess) cudaGetLastError();
  });
}

static int ensure_device() {
  if (g_device >= 0) { IA3_CUDA(cudaSetDevice(g_device)); return 0; }
  int dev = 0;
  if (const char* e = getenv("IA3_DEVICE")) dev = atoi(e);
  int n = 0;
  IA3_CUDA(cudaGetDeviceCount(&n));
  if (n <= 0) { set_error("no CUDA device visible: libia3b200 has no CPU fallback"); return -1; }
  if (dev >= n) dev = dev % n;
  init_device_flags(dev);
  IA3_CUDA(cudaSetDevice(dev));
  g_device = dev;
  return 0;
}

static size_t dtype_size(int dt) { return dt == IA3_DTYPE_U16 ? 2 : (dt == IA3_DTYPE_F32 ? 4 : 8); }

template <typename T>
static int upload(T** d, const std::vector<T>& h, cudaStream_t st) {
  if (dev_alloc((void**)d, h.size() * sizeof(T))) return -1;
  if (!h.empty()) IA3_CUDA(cudaMemcpyAsync(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, st));
  return 0;
}

}  // namespace ia3

using namespace ia3;

// A cudaMemcpyAsync into pageable host memory blocks the calling thread until the stream reaches the
// copy, and the driver serialises other host threads' calls behind it meanwhile.  With several stacks
// in flight that turned one stack's 80 ms straggler kernel into a stall of every other stack.  So:
// wait for the stream first (cudaStreamSynchronize waits without blocking other threads), copy after.
#define IA3_DRAIN(st) IA3_CUDA(cudaStreamSynchronize(st))

struct ia3_stack {
  int dtype = 0, Z = 0, X = 0, Y = 0;
  size_t nvox = 0;
  void* d_im = nullptr;
  bool owns = false;
  cudaStream_t stream = nullptr;
  // seed stage buffers
  void* fg = nullptr; void* bg = nullptr; void* scratch = nullptr;
  const void* fg_final = nullptr; const void* bg_final = nullptr;
  uint8_t* bits = nullptr; int* counts = nullptr; long long* offsets = nullptr;
  int32_t* cand_zxy = nullptr; float* cand_h = nullptr;
  int64_t n_cand = 0;
  void* h_mail = nullptr;            // pinned, device-addressable: scalars the seed stage hands back
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

struct ia3_fit {
  ia3_stack* s = nullptr;
  ia3_fit_cfg cfg;
  FitDev d;
  int64_t n = 0;
  std::vector<double> centers;
  std::vector<int> level;            // per seed, 0-based
  int n_levels = 0;
  std::vector<int8_t> offs;
  bool prepared = false, first_done = false;
  int64_t n_ties = 0;
  // device arrays owned here
  double* d_centers = nullptr; int* d_own = nullptr; int* d_nbr_start = nullptr; int* d_nbr_idx = nullptr;
  int8_t* d_offs = nullptr; uint32_t* d_mask = nullptr;
  int* d_tie_count = nullptr; int* d_tie_spot = nullptr; int* d_tie_k = nullptr; int tie_cap = 0;
  float* d_ps = nullptr; double* d_praw = nullptr; uint8_t* d_succ = nullptr; int* d_nfev = nullptr; int* d_info = nullptr;
  double* d_rec = nullptr; double* d_snap = nullptr; double* d_vol = nullptr; int* d_brick_tab = nullptr;
  int64_t n_bricks = 0;
  int* d_work = nullptr; size_t work_cap = 0;
  void* h_stage = nullptr; size_t stage_cap = 0;      // pinned staging, device -> host (results, ties)
  void* h_up = nullptr; size_t up_cap = 0;            // pinned staging, host -> device (inputs, work lists)
  uint8_t* d_keep = nullptr; size_t keep_cap = 0;
  LMPause* d_pause = nullptr; int* d_pause_ctl = nullptr; int* h_pause_ctl = nullptr;   // suspended long runs (k_fit)
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  float last_ms = 0.f;
};

// ---- continuation service ---------------------------------------------------------------------
// k_fit suspends a run after g_fit_cap function evaluations (fit_kernels.cu: fit_one).  The suspended
// spots of every stack in flight are continued here, together, by one thread on one stream: a round =
// one k_fit_resume launch over all pending spots, g_fit_cap more evaluations each.  A stack's own
// launches therefore stay short, and the few junk seeds that run MINPACK to maxfev hold one hardware
// queue in total instead of one per stack.
static int fit_cap_from_env() {
  const char* e = getenv("IA3_FIT_CAP");
  const int v = e ? atoi(e) : 100;
  return v < 0 ? 0 : v;
}
static const int g_fit_cap = fit_cap_from_env();
struct SvcJob {
  FitDev d;
  int mode = 0;
  std::vector<int> spots;
  bool done = false, failed = false;
  std::condition_variable cv;
};
// never destroyed: the service thread is detached and may be waiting on them when the process exits
// (destroying a condition variable with a waiter blocks the exit)
static std::mutex& g_svc_mu = *new std::mutex;
static std::condition_variable& g_svc_cv = *new std::condition_variable;
static std::vector<SvcJob*>& g_svc_in = *new std::vector<SvcJob*>;
static std::once_flag g_svc_once;

static void svc_main(int device) {
  constexpr int MAXJ = 512, MAXE = 16384;
  FitDev* devs = nullptr;
  FitResume* ent = nullptr;
  cudaStream_t st = nullptr;
  // highest priority: a round's few CTAs should not wait for free slots behind thousands of ordinary
  // fits of later-launched kernels -- every stack with a suspended spot is waiting for this stream
  int prio_lo = 0, prio_hi = 0;
  cudaSetDevice(device);
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  bool ok = cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, prio_hi) == cudaSuccess &&
            cudaMallocHost((void**)&devs, sizeof(FitDev) * MAXJ) == cudaSuccess && cudaMallocHost((void**)&ent, sizeof(FitResume) * MAXE) == cudaSuccess;
  std::vector<SvcJob*> active, batch;
  for (;;) {
    {
      std::unique_lock<std::mutex> lk(g_svc_mu);
      if (active.empty()) g_svc_cv.wait(lk, [] { return !g_svc_in.empty(); });
      active.insert(active.end(), g_svc_in.begin(), g_svc_in.end());
      g_svc_in.clear();
    }
    batch.clear();
    int ne = 0, smem = 0;
    for (SvcJob* j : active) {
      if ((int)batch.size() == MAXJ || ne + (int)j->spots.size() > MAXE) break;
      devs[batch.size()] = j->d;
      for (int sp : j->spots) ent[ne++] = FitResume{(int)batch.size(), sp, j->mode, -1};
      smem = std::max(smem, fit_smem_bytes(j->d.K, false));
      batch.push_back(j);
    }
    bool fail = !ok;
    if (!fail) {
      IA3_STAT("  service round");
      fail = launch_fit_resume(devs, ent, ne, g_fit_cap, smem, st) != 0 || cudaStreamSynchronize(st) != cudaSuccess;
    }
    if (g_stats_on) { static StatSlot* spots_slot = stat_slot("  service spots (calls = spots)"); spots_slot->calls += ne; }
    if (fail) {
      std::lock_guard<std::mutex> lk(g_svc_mu);
      for (SvcJob* j : active) { j->failed = true; j->done = true; j->cv.notify_all(); }
      active.clear();
      continue;
    }
    for (SvcJob* j : batch) j->spots.clear();
    for (int e = 0; e < ne; ++e)
      if (ent[e].status != FIT_DONE) batch[ent[e].job]->spots.push_back(ent[e].spot);
    std::vector<SvcJob*> still;
    {
      std::lock_guard<std::mutex> lk(g_svc_mu);
      for (SvcJob* j : active) {
        if (j->spots.empty()) { j->done = true; j->cv.notify_all(); }
        else still.push_back(j);
      }
    }
    active.swap(still);
  }
}

// blocks until the suspended spots (h_pause_ctl[1..np]) of this handle have finished
static int service_run(ia3_fit* f, int mode, int np) {
  std::call_once(g_svc_once, [] { std::thread(svc_main, g_device).detach(); });
  SvcJob job;
  job.d = f->d;
  job.mode = mode;
  job.spots.assign(f->h_pause_ctl + 1, f->h_pause_ctl + 1 + np);
  std::unique_lock<std::mutex> lk(g_svc_mu);
  g_svc_in.push_back(&job);
  g_svc_cv.notify_one();
  job.cv.wait(lk, [&] { return job.done; });
  if (job.failed) { set_error("continuation of suspended fits failed"); return -1; }
  return 0;
}

// one k_fit launch on the handle's stream, then (cap > 0) hand the spots it suspended to the service
static int run_fit_launch(ia3_fit* f, int mode, const int* work, long long n_work) {
  if (n_work <= 0) return 0;
  cudaStream_t st = f->s->stream;
  const bool capped = f->d.cap > 0;
  if (capped) IA3_CUDA(cudaMemsetAsync(f->d_pause_ctl, 0, sizeof(int), st));
  if (launch_fit(f->d, mode, work, n_work, f->cfg.eval_fp32 != 0, st)) return -1;
  if (capped && small_copy(f->h_pause_ctl, f->d_pause_ctl, sizeof(int) * (size_t)(1 + f->d.pause_slots), st)) return -1;
  // A sweep kernel lasts as long as its slowest fit; nothing is queued behind it (an item waiting for it
  // would also hold back the other stacks that share the hardware queue).
  IA3_DRAIN(st);
  if (!capped) return 0;
  const int np = std::min(f->h_pause_ctl[0], f->d.pause_slots);
  { IA3_STAT("  suspended spots -> service"); if (np > 0 && service_run(f, mode, np)) return -1; }
  return 0;
}

extern "C" {

int ia3_init(int device) {
  if (device >= 0) {
    int n = 0;
    IA3_CUDA(cudaGetDeviceCount(&n));
    if (n <= 0) { set_error("no CUDA device visible: libia3b200 has no CPU fallback"); return -1; }
    g_device = device % n;
    init_device_flags(g_device);
  }
  return ensure_device();
}
const char* ia3_last_error(void) { return g_err.c_str(); }
int ia3_version(void) { return 100; }
int ia3_device_sm_count(void) {
  if (ensure_device()) return -1;
  int v = 0;
  cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, g_device);
  return v;
}
int64_t ia3_launch_count(void) { return g_launches.load(); }
int ia3_debug_stats(char* buf, int cap) {
  if (!buf || cap <= 0) {                        // reset
    for (int i = 0; i < g_nstats.load(); ++i) { g_stats[i].ns = 0; g_stats[i].cpu_ns = 0; g_stats[i].calls = 0; }
    return 0;
  }
  int off = 0;
  const int n = g_nstats.load();
  for (int i = 0; i < n && off < cap - 96; ++i)
    off += snprintf(buf + off, cap - off, "%-24s calls %8lld  wall %10.2f ms  cpu %10.2f ms\n", g_stats[i].name, g_stats[i].calls.load(),
                    g_stats[i].ns.load() * 1e-6, g_stats[i].cpu_ns.load() * 1e-6);
  return off;
}

int ia3_timer_start(void) {
  if (ensure_device()) return -1;
  cudaStream_t gs;
  if (global_stream(&gs)) return -1;
  if (!g_t0) { IA3_CUDA(cudaEventCreate(&g_t0)); IA3_CUDA(cudaEventCreate(&g_t1)); }
  IA3_CUDA(cudaDeviceSynchronize());
  IA3_CUDA(cudaEventRecord(g_t0, g_stream));
  return 0;
}
int ia3_timer_stop(float* ms) {
  if (ensure_device()) return -1;
  if (!g_t0) { set_error("timer not started"); return -1; }
  IA3_CUDA(cudaDeviceSynchronize());               // work of every stream of this process is done
  IA3_CUDA(cudaEventRecord(g_t1, g_stream));
  IA3_CUDA(cudaEventSynchronize(g_t1));
  IA3_CUDA(cudaEventElapsedTime(ms, g_t0, g_t1));
  return 0;
}

// ---- stacks ---------------------------------------------------------------------------------
static int stack_common(ia3_stack* s, int dtype, int Z, int X, int Y) {
  if (dtype < 0 || dtype > 2 || Z <= 0 || X <= 0 || Y <= 0) { set_error("bad stack dtype/shape"); return -1; }
  s->dtype = dtype; s->Z = Z; s->X = X; s->Y = Y;
  s->nvox = (size_t)Z * X * Y;
  if (acquire_stream(&s->stream)) return -1;
  for (auto& e : s->ev) if (acquire_event(&e)) return -1;
  if (host_alloc(&s->h_mail, 4096)) return -1;
  return 0;
}

int ia3_stack_create(const void* im, int dtype, int Z, int X, int Y, ia3_stack** out) {
  IA3_STAT("ia3_stack_create");
  if (ensure_device()) return -1;
  if (!im || !out) { set_error("null argument"); return -1; }
  ia3_stack* s = new ia3_stack();
  if (stack_common(s, dtype, Z, X, Y)) { delete s; return -1; }
  if (dev_alloc(&s->d_im, s->nvox * dtype_size(dtype))) { delete s; return -1; }
  s->owns = true;
  // The image goes up on ONE upload stream shared by all stacks (the copy engine serialises the
  // uploads anyway): with more stacks in flight than hardware queues (32), a stack's own stream
  // shares its queue with another stack's, and a 400 MB copy waiting its turn behind other uploads
  // would hold back that other stack's kernels for as long.
  {
    cudaStream_t us;
    if (upload_stream(&us)) { ia3_stack_destroy(s); return -1; }
    cudaError_t e;
    {
      std::lock_guard<std::mutex> lk(g_upload_mu);
      e = cudaMemcpyAsync(s->d_im, im, s->nvox * dtype_size(dtype), cudaMemcpyHostToDevice, us);
      if (e == cudaSuccess) e = cudaEventRecord(s->ev[5], us);
    }
    if (e == cudaSuccess) e = cudaStreamWaitEvent(s->stream, s->ev[5], 0);
    if (e == cudaSuccess) e = cudaEventSynchronize(s->ev[5]);
    if (e != cudaSuccess) { set_error(std::string("image upload failed: ") + cudaGetErrorString(e)); ia3_stack_destroy(s); return -1; }
  }
  *out = s;
  return 0;
}

int ia3_stack_wrap_device(const void* d_im, int dtype, int Z, int X, int Y, ia3_stack** out) {
  IA3_STAT("ia3_stack_wrap_device");
  if (ensure_device()) return -1;
  if (!d_im || !out) { set_error("null argument"); return -1; }
  ia3_stack* s = new ia3_stack();
  if (stack_common(s, dtype, Z, X, Y)) { delete s; return -1; }
  s->d_im = const_cast<void*>(d_im);
  s->owns = false;
  *out = s;
  return 0;
}

int ia3_stack_destroy(ia3_stack* s) {
  IA3_STAT("ia3_stack_destroy");
  if (!s) return 0;
  if (g_device >= 0) cudaSetDevice(g_device);
  if (s->stream) cudaStreamSynchronize(s->stream);
  if (s->owns) dev_free(s->d_im);
  dev_free(s->fg); dev_free(s->bg); dev_free(s->scratch);
  dev_free(s->bits); dev_free(s->counts); dev_free(s->offsets);
  dev_free(s->cand_zxy); dev_free(s->cand_h);
  for (auto& e : s->ev) release_event(e);
  release_stream(s->stream);
  host_free(s->h_mail);
  delete s;
  return 0;
}

int ia3_stack_trim(ia3_stack* s, int what) {
  IA3_STAT("ia3_stack_trim");
  if (!s) return 0;
  if (ensure_device()) return -1;
  IA3_CUDA(cudaStreamSynchronize(s->stream));
  if (what & 1) {
    dev_free(s->fg); dev_free(s->bg); dev_free(s->scratch);
    dev_free(s->bits); dev_free(s->counts); dev_free(s->offsets);
    dev_free(s->cand_zxy); dev_free(s->cand_h);
    s->fg = s->bg = s->scratch = nullptr;
    s->bits = nullptr; s->counts = nullptr; s->offsets = nullptr;
    s->cand_zxy = nullptr; s->cand_h = nullptr;
    s->fg_final = s->bg_final = nullptr;
    s->n_cand = 0;
  }
  if ((what & 2) && s->owns) {
    dev_free(s->d_im);
    s->d_im = nullptr;
    s->owns = false;
  }
  return 0;
}

// ---- seed stage -----------------------------------------------------------------------------
}  // extern "C"

template <typename Tin>
static int seed_run_t(ia3_stack* s, const ia3_seed_cfg* cfg, int64_t* n_candidates, ia3_seed_timing* t) {
  cudaStream_t st = s->stream;
  const size_t bytes = s->nvox * sizeof(Tin);
  const Tin* im = reinterpret_cast<const Tin*>(s->d_im);
  if (!s->scratch && dev_alloc(&s->scratch, bytes)) return -1;
  IA3_CUDA(cudaEventRecord(s->ev[0], st));
  auto blur = [&](const double* w, int r, void** buf, const void** fin) -> int {
    if (r < 0) { *fin = im; return 0; }
    if (r > GaussW::MAXR) { set_error("gaussian radius too large (max 95)"); return -1; }
    if (!*buf && dev_alloc(buf, bytes)) return -1;
    GaussW gw;
    memset(&gw, 0, sizeof(gw));
    gw.r = r;
    for (int j = 0; j <= r; ++j) gw.w[j] = w[j];
    if (gaussian_filter_exact<Tin>(im, reinterpret_cast<Tin*>(*buf), reinterpret_cast<Tin*>(s->scratch), s->Z, s->X, s->Y, gw, st)) return -1;
    *fin = *buf;
    return 0;
  };
  if (blur(cfg->w_fg, cfg->r_fg, &s->fg, &s->fg_final)) return -1;
  IA3_CUDA(cudaEventRecord(s->ev[1], st));
  if (blur(cfg->w_bg, cfg->r_bg, &s->bg, &s->bg_final)) return -1;
  IA3_CUDA(cudaEventRecord(s->ev[2], st));

  SeedDims d;
  d.Z = s->Z; d.X = s->X; d.Y = s->Y;
  d.cpr = (s->Y + 7) / 8;
  d.n_chunks = (long long)s->Z * s->X * d.cpr;
  const int T = seed_flag_threads();
  d.n_blocks = (int)((d.n_chunks + T - 1) / T);
  d.fs = cfg->filt_size < 1 ? 1 : cfg->filt_size;
  d.s1 = d.fs / 2; d.s2 = d.fs - d.s1 - 1;
  d.edge_on = (cfg->variant == 0 && cfg->edge > 0) ? 1 : 0;
  d.lo = (int)std::ceil(cfg->edge);
  d.hiZ = (int)std::floor((double)s->Z - cfg->edge);
  d.hiX = (int)std::floor((double)s->X - cfg->edge);
  d.hiY = (int)std::floor((double)s->Y - cfg->edge);
  d.h_min = cfg->h_min;
  if (!s->bits && dev_alloc((void**)&s->bits, (size_t)d.n_chunks)) return -1;
  if (!s->counts && dev_alloc((void**)&s->counts, sizeof(int) * (size_t)d.n_blocks)) return -1;
  if (!s->offsets && dev_alloc((void**)&s->offsets, sizeof(long long) * ((size_t)d.n_blocks + 1))) return -1;
  const Tin* fg = reinterpret_cast<const Tin*>(s->fg_final);
  const Tin* bg = reinterpret_cast<const Tin*>(s->bg_final);
  if (seed_flags<Tin>(fg, bg, d, cfg->variant, s->bits, s->counts, s->offsets, st)) return -1;
  IA3_CUDA(cudaEventRecord(s->ev[3], st));
  if (small_copy(s->h_mail, s->offsets + d.n_blocks, sizeof(long long), st)) return -1;
  IA3_CUDA(cudaStreamSynchronize(st));
  const long long total = *static_cast<const long long*>(s->h_mail);
  dev_free(s->cand_zxy); dev_free(s->cand_h);
  s->cand_zxy = nullptr; s->cand_h = nullptr;
  if (dev_alloc((void**)&s->cand_zxy, sizeof(int32_t) * 3 * (size_t)std::max<long long>(total, 1))) return -1;
  if (dev_alloc((void**)&s->cand_h, sizeof(float) * (size_t)std::max<long long>(total, 1))) return -1;
  if (total > 0 && seed_emit<Tin>(fg, bg, d, cfg->variant, s->bits, s->offsets, s->cand_zxy, s->cand_h, st)) return -1;
  IA3_CUDA(cudaEventRecord(s->ev[4], st));
  IA3_CUDA(cudaStreamSynchronize(st));
  s->n_cand = total;
  if (n_candidates) *n_candidates = total;
  if (t) {
    cudaEventElapsedTime(&t->ms_gauss_fg, s->ev[0], s->ev[1]);
    cudaEventElapsedTime(&t->ms_gauss_bg, s->ev[1], s->ev[2]);
    cudaEventElapsedTime(&t->ms_rank, s->ev[2], s->ev[3]);
    cudaEventElapsedTime(&t->ms_compact, s->ev[3], s->ev[4]);
    cudaEventElapsedTime(&t->ms_total, s->ev[0], s->ev[4]);
  }
  return 0;
}

extern "C" {

int ia3_seed_run(ia3_stack* s, const ia3_seed_cfg* cfg, int64_t* n_candidates, ia3_seed_timing* t) {
  IA3_STAT("ia3_seed_run");
  if (ensure_device()) return -1;
  if (!s || !cfg) { set_error("null argument"); return -1; }
  if (!s->d_im) { set_error("the stack's image was released (ia3_stack_trim)"); return -1; }
  if (s->dtype == IA3_DTYPE_U16) return seed_run_t<uint16_t>(s, cfg, n_candidates, t);
  if (s->dtype == IA3_DTYPE_F32) return seed_run_t<float>(s, cfg, n_candidates, t);
  set_error("seed stage supports uint16 and float32 stacks");
  return -1;
}

int ia3_seed_fetch(ia3_stack* s, int32_t* zxy, float* h, int64_t cap) {
  IA3_STAT("ia3_seed_fetch");
  if (ensure_device()) return -1;
  if (!s) { set_error("null argument"); return -1; }
  const int64_t n = std::min<int64_t>(cap, s->n_cand);
  if (n > 0) {
    const size_t bz = sizeof(int32_t) * 3 * (size_t)n, bh = sizeof(float) * (size_t)n, oh = (bz + 255) / 256 * 256;
    void* stage = nullptr;
    if (host_alloc(&stage, oh + bh)) return -1;
    char* hp = static_cast<char*>(stage);
    int rc = 0;
    if (zxy) rc |= small_copy(hp, s->cand_zxy, bz, s->stream);
    if (h) rc |= small_copy(hp + oh, s->cand_h, bh, s->stream);
    if (cudaStreamSynchronize(s->stream) != cudaSuccess) { set_error("seed_fetch: stream failed"); rc = -1; }
    if (!rc) {
      if (zxy) memcpy(zxy, hp, bz);
      if (h) memcpy(h, hp + oh, bh);
    }
    host_free(stage);
    return rc;
  }
  return 0;
}

int ia3_seed_fetch_volume(ia3_stack* s, int which, void* out) {
  if (ensure_device()) return -1;
  const void* src = which == 0 ? s->fg_final : s->bg_final;
  if (!src) { set_error("seed stage has not run"); return -1; }
  IA3_CUDA(cudaMemcpyAsync(out, src, s->nvox * dtype_size(s->dtype), cudaMemcpyDeviceToHost, s->stream));
  IA3_CUDA(cudaStreamSynchronize(s->stream));
  return 0;
}

int ia3_box_background(ia3_stack* s, const int32_t* boxes, int64_t n, int first, int last, int bin_size, int max_iter,
                       double* out) {
  IA3_STAT("ia3_box_background");
  if (ensure_device()) return -1;
  if (!s || (n > 0 && (!boxes || !out))) { set_error("null argument"); return -1; }
  if (!s->d_im) { set_error("the stack's image was released (ia3_stack_trim)"); return -1; }
  if (s->dtype != IA3_DTYPE_U16) { set_error("ia3_box_background supports uint16 stacks"); return -1; }
  if (bin_size < 1 || last <= first) { set_error("bad histogram range"); return -1; }
  if (n <= 0) return 0;
  // np.arange(first, last, bin_size) has ceil((last - first) / bin_size) edges -> one bin fewer
  const int nedges = (last - first + bin_size - 1) / bin_size;
  const int nbins = nedges - 1;
  if (nbins < 1) { set_error("bad histogram range"); return -1; }
  for (int64_t i = 0; i < n; ++i) {
    const int32_t* b = boxes + 6 * i;
    if (b[0] < 0 || b[1] > s->Z || b[2] < 0 || b[3] > s->X || b[4] < 0 || b[5] > s->Y) { set_error("box outside the stack"); return -1; }
  }
  cudaStream_t st = s->stream;
  void* h = nullptr;
  int* d_boxes = nullptr;
  double* d_out = nullptr;
  const size_t bb = (size_t)n * 6 * 4, ob = (size_t)n * 8;
  if (host_alloc(&h, bb + ob + 256) || dev_alloc((void**)&d_boxes, bb) || dev_alloc((void**)&d_out, ob)) return -1;
  memcpy(h, boxes, bb);
  char* ho = static_cast<char*>(h) + (bb + 255) / 256 * 256;
  if (small_copy(d_boxes, h, bb, st)) return -1;
  const bool whole = (n == 1 && boxes[0] == 0 && boxes[1] == s->Z && boxes[2] == 0 && boxes[3] == s->X && boxes[4] == 0 && boxes[5] == s->Y);
  unsigned* d_ghist = nullptr;
  if (whole) {
    if (dev_alloc((void**)&d_ghist, (size_t)nbins * 4)) return -1;
    if (volume_background(reinterpret_cast<const uint16_t*>(s->d_im), s->Z, s->X, s->Y, d_boxes, first, bin_size, nbins, max_iter,
                          d_ghist, d_out, st)) return -1;
  } else if (box_background(reinterpret_cast<const uint16_t*>(s->d_im), s->X, s->Y, d_boxes, n, first, bin_size, nbins, max_iter, d_out, st)) return -1;
  if (small_copy(ho, d_out, ob, st)) return -1;
  IA3_CUDA(cudaStreamSynchronize(st));
  memcpy(out, ho, ob);
  host_free(h); dev_free(d_boxes); dev_free(d_out); dev_free(d_ghist);
  return 0;
}

// ---- fit stage ------------------------------------------------------------------------------
static inline long long cell_key(long long a, long long b, long long c) {
  return ((a + (1LL << 20)) << 42) | ((b + (1LL << 20)) << 21) | (c + (1LL << 20));
}

int ia3_fit_create(ia3_stack* s, const double* centers_zxy, int64_t n, const ia3_fit_cfg* cfg, ia3_fit** out) {
  IA3_STAT("ia3_fit_create");
  if (ensure_device()) return -1;
  if (!s || !cfg || !out || (n > 0 && !centers_zxy)) { set_error("null argument"); return -1; }
  if (!s->d_im) { set_error("the stack's image was released (ia3_stack_trim)"); return -1; }
  if (cfg->radius < 1 || cfg->radius > 7) { set_error("radius_fit must be in 1..7 (window of at most 2048 voxels)"); return -1; }
  if (cfg->personality != 3 && cfg->personality != 4) { set_error("personality must be 3 or 4"); return -1; }
  for (int64_t i = 0; i < 3 * n; ++i)
    if (!(std::fabs(centers_zxy[i]) < 1e6)) { set_error("seed coordinates must be finite and |c| < 1e6"); return -1; }
  ia3_fit* f = new ia3_fit();
  f->s = s; f->cfg = *cfg; f->n = n;
  f->centers.assign(centers_zxy, centers_zxy + 3 * n);
  cudaStream_t st = s->stream;
  const int r = cfg->radius;
  // window offsets: np.indices([2r]*3) - r, kept where d^2 <= r^2, C order (Fitting_v4.py:580-583)
  for (int a = -r; a < r; ++a) for (int b = -r; b < r; ++b) for (int c = -r; c < r; ++c)
    if (a * a + b * b + c * c <= r * r) { f->offs.push_back((int8_t)a); f->offs.push_back((int8_t)b); f->offs.push_back((int8_t)c); }
  const int K = (int)(f->offs.size() / 3);
  const int KW = (K + 31) / 32;

  // neighbour lists (seeds that can own a voxel of this seed's window) and dependency levels
  const double reach = 2.0 * ((double)r + 1.7320508075688772) + 1e-6;
  const double cs = std::ceil(reach);
  // uniform grid of cells of edge >= reach, as a counting sort (cell -> [start, end) in `order`)
  double lo3[3] = {0, 0, 0}, hi3[3] = {0, 0, 0};
  for (int64_t i = 0; i < n; ++i)
    for (int a = 0; a < 3; ++a) {
      const double v = f->centers[3 * i + a];
      if (i == 0 || v < lo3[a]) lo3[a] = v;
      if (i == 0 || v > hi3[a]) hi3[a] = v;
    }
  long long gdim[3];
  for (int a = 0; a < 3; ++a) gdim[a] = (long long)std::floor((hi3[a] - lo3[a]) / cs) + 1;
  double cse = cs;
  while ((double)gdim[0] * (double)gdim[1] * (double)gdim[2] > 4.0e7) {     // absurdly sparse seeds: coarser cells
    cse *= 2.0;
    for (int a = 0; a < 3; ++a) gdim[a] = (long long)std::floor((hi3[a] - lo3[a]) / cse) + 1;
  }
  auto cellc = [&](double v, int a) { return (long long)std::floor((v - lo3[a]) / cse); };
  const long long ncell = (n > 0) ? gdim[0] * gdim[1] * gdim[2] : 0;
  std::vector<int> cell_start((size_t)ncell + 1, 0), order((size_t)n), cell_of((size_t)n);
  for (int64_t i = 0; i < n; ++i) {
    const double* c = &f->centers[3 * i];
    const long long id = (cellc(c[0], 0) * gdim[1] + cellc(c[1], 1)) * gdim[2] + cellc(c[2], 2);
    cell_of[i] = (int)id;
    cell_start[id + 1] += 1;
  }
  for (long long k = 0; k < ncell; ++k) cell_start[k + 1] += cell_start[k];
  {
    std::vector<int> fill(cell_start.begin(), cell_start.end() - (ncell > 0 ? 1 : 0));
    for (int64_t i = 0; i < n; ++i) order[fill[cell_of[i]]++] = (int)i;
  }
  std::vector<int> nbr_start(n + 1, 0), nbr_idx, own(n);
  f->level.assign(n, 0);
  int n_levels = 0;
  const int lim = 2 * r - 1;
  for (int64_t i = 0; i < n; ++i) {
    const double* c = &f->centers[3 * i];
    const long long a = cellc(c[0], 0), b = cellc(c[1], 1), cc = cellc(c[2], 2);
    const int ic[3] = {(int)c[0], (int)c[1], (int)c[2]};
    int lvl = 0, ownid = (int)i;
    const size_t begin = nbr_idx.size();
    for (long long da = -1; da <= 1; ++da) for (long long db = -1; db <= 1; ++db) for (long long dc = -1; dc <= 1; ++dc) {
      const long long ca = a + da, cb = b + db, ccc = cc + dc;
      if (ca < 0 || ca >= gdim[0] || cb < 0 || cb >= gdim[1] || ccc < 0 || ccc >= gdim[2]) continue;
      const long long id = (ca * gdim[1] + cb) * gdim[2] + ccc;
      for (int e = cell_start[id]; e < cell_start[id + 1]; ++e) {
        const int j = order[e];
        if (j == (int)i) continue;
        const double* q = &f->centers[3 * j];
        const double d0 = q[0] - c[0], d1 = q[1] - c[1], d2 = q[2] - c[2];
        if (d0 * d0 + d1 * d1 + d2 * d2 <= reach * reach) nbr_idx.push_back(j);
        if (d0 == 0 && d1 == 0 && d2 == 0 && j < ownid) ownid = j;
        if (j < (int)i) {
          const int e0 = std::abs((int)q[0] - ic[0]), e1 = std::abs((int)q[1] - ic[1]), e2 = std::abs((int)q[2] - ic[2]);
          if (e0 <= lim && e1 <= lim && e2 <= lim && e0 * e0 + e1 * e1 + e2 * e2 <= 4 * r * r)
            lvl = std::max(lvl, f->level[j] + 1);
        }
      }
    }
    std::sort(nbr_idx.begin() + begin, nbr_idx.end());
    nbr_start[i + 1] = (int)nbr_idx.size();
    own[i] = ownid;
    f->level[i] = lvl;
    n_levels = std::max(n_levels, lvl + 1);
  }
  f->n_levels = n_levels;

  // inputs go through one pinned arena (no pageable copies), the brick table included
  const int nbz = (s->Z + 7) / 8, nbx = (s->X + 7) / 8, nby = (s->Y + 7) / 8;
  const size_t tab_n = (size_t)nbz * nbx * nby;
  auto al = [](size_t b) { return (b + 255) / 256 * 256; };
  const size_t up_total = al(f->centers.size() * 8) + al(own.size() * 4) + al(nbr_start.size() * 4) + al(nbr_idx.size() * 4) +
                          al(f->offs.size()) + al(tab_n * 4) + al((size_t)std::max<int64_t>(n, 1) * 4);
  if (reserve_pinned(&f->h_up, &f->up_cap, up_total)) { ia3_fit_destroy(f); return -1; }
  size_t up_off = 0;
  auto put = [&](void** d, const void* src, size_t bytes) -> int {
    if (dev_alloc(d, bytes)) return -1;
    if (bytes) {
      char* h = static_cast<char*>(f->h_up) + up_off;
      memcpy(h, src, bytes);
      if (small_copy(*d, h, bytes, st)) return -1;
      up_off += (bytes + 255) / 256 * 256;
    }
    return 0;
  };
  if (put((void**)&f->d_centers, f->centers.data(), f->centers.size() * 8) || put((void**)&f->d_own, own.data(), own.size() * 4) ||
      put((void**)&f->d_nbr_start, nbr_start.data(), nbr_start.size() * 4) || put((void**)&f->d_nbr_idx, nbr_idx.data(), nbr_idx.size() * 4) ||
      put((void**)&f->d_offs, f->offs.data(), f->offs.size())) { ia3_fit_destroy(f); return -1; }
  const size_t nn = (size_t)std::max<int64_t>(n, 1);
  if (dev_alloc((void**)&f->d_mask, nn * KW * 4) || dev_alloc((void**)&f->d_ps, nn * NOUT * 4) ||
      dev_alloc((void**)&f->d_praw, nn * NP * 8) || dev_alloc((void**)&f->d_succ, nn) ||
      dev_alloc((void**)&f->d_nfev, nn * 4) || dev_alloc((void**)&f->d_info, nn * 4) ||
      dev_alloc((void**)&f->d_rec, nn * K * 8) || dev_alloc((void**)&f->d_snap, nn * K * 8) ||
      dev_alloc((void**)&f->d_tie_count, 256)) { ia3_fit_destroy(f); return -1; }
  // sparse float64 work volume: the 8x8x8 bricks touched by some seed's (clipped) window
  {
    std::vector<int> tab(tab_n, -1);
    int next = 0;
    for (int64_t i = 0; i < n; ++i) {
      int lo[3], hi[3];
      const int dims[3] = {s->Z, s->X, s->Y};
      bool empty = false;
      for (int a = 0; a < 3; ++a) {
        const int ic = (int)f->centers[3 * i + a];
        lo[a] = std::max(ic - r, 0);
        hi[a] = std::min(ic + r - 1, dims[a] - 1);
        if (lo[a] > hi[a]) empty = true;
      }
      if (empty) continue;
      for (int bz = lo[0] >> 3; bz <= hi[0] >> 3; ++bz)
        for (int bx = lo[1] >> 3; bx <= hi[1] >> 3; ++bx)
          for (int by = lo[2] >> 3; by <= hi[2] >> 3; ++by) {
            int& t = tab[((size_t)bz * nbx + bx) * nby + by];
            if (t < 0) t = next++;
          }
    }
    f->n_bricks = next;
    if (put((void**)&f->d_brick_tab, tab.data(), tab_n * 4) || dev_alloc((void**)&f->d_vol, (size_t)std::max(next, 1) * 512 * 8)) { ia3_fit_destroy(f); return -1; }
  }
  IA3_CUDA(cudaMemsetAsync(f->d_succ, 0, nn, st));
  IA3_CUDA(cudaMemsetAsync(f->d_rec, 0, nn * K * 8, st));
  if (acquire_event(&f->e0) || acquire_event(&f->e1)) { ia3_fit_destroy(f); return -1; }

  FitDev& d = f->d;
  memset(&d, 0, sizeof(d));
  d.im = s->d_im; d.im_dtype = s->dtype; d.vol = f->d_vol; d.brick_tab = f->d_brick_tab; d.nbx = nbx; d.nby = nby;
  d.Z = s->Z; d.X = s->X; d.Y = s->Y;
  d.n = n; d.centers = f->d_centers; d.own_id = f->d_own; d.nbr_start = f->d_nbr_start; d.nbr_idx = f->d_nbr_idx;
  d.K = K; d.KW = KW; d.offs = f->d_offs; d.mask = f->d_mask;
  d.tie_count = f->d_tie_count;
  d.ps = f->d_ps; d.p_raw = f->d_praw; d.success = f->d_succ; d.nfev = f->d_nfev; d.info = f->d_info;
  d.rec = f->d_rec;
  d.fp.min_w2 = cfg->min_w * cfg->min_w; d.fp.max_w2 = cfg->max_w * cfg->max_w;
  d.fp.delta = 1.0; d.fp.weight_sigma = cfg->weight_sigma; d.fp.personality = cfg->personality;
  d.fp.init_wt[0] = d.fp.init_wt[1] = d.fp.init_wt[2] = 0.0;
  d.lm.ftol = 1.49012e-8; d.lm.xtol = 1.49012e-8; d.lm.gtol = 0.0; d.lm.factor = 100.0;
  d.lm.maxfev = cfg->maxfev > 0 ? cfg->maxfev : (cfg->personality == 4 ? 1000 : 1100);
  for (int i = 0; i < 3; ++i) d.init_w[i] = cfg->init_w[i];
  d.cap = (cfg->eval_fp32 != 0) ? 0 : g_fit_cap;
  if (d.cap > 0) {
    d.pause_slots = (int)std::min<int64_t>(std::max<int64_t>(n, 1), 1024);
    void* hp = nullptr;
    if (dev_alloc((void**)&f->d_pause, sizeof(LMPause) * (size_t)d.pause_slots) ||
        dev_alloc((void**)&f->d_pause_ctl, sizeof(int) * (size_t)(1 + d.pause_slots)) ||
        host_alloc(&hp, sizeof(int) * (size_t)(1 + d.pause_slots))) { ia3_fit_destroy(f); return -1; }
    f->h_pause_ctl = static_cast<int*>(hp);
    d.pause_buf = f->d_pause; d.pause_ctl = f->d_pause_ctl;
  }
  IA3_CUDA(cudaStreamSynchronize(st));
  *out = f;
  return 0;
}

int ia3_fit_destroy(ia3_fit* f) {
  IA3_STAT("ia3_fit_destroy");
  if (!f) return 0;
  if (g_device >= 0) cudaSetDevice(g_device);
  if (f->s && f->s->stream) cudaStreamSynchronize(f->s->stream);
  void* ptrs[] = {f->d_centers, f->d_own, f->d_nbr_start, f->d_nbr_idx, f->d_offs, f->d_mask, f->d_tie_count,
                  f->d_tie_spot, f->d_tie_k, f->d_ps, f->d_praw, f->d_succ, f->d_nfev, f->d_info, f->d_rec,
                  f->d_snap, f->d_vol, f->d_work, f->d_brick_tab};
  for (void* p : ptrs) dev_free(p);
  release_event(f->e0);
  release_event(f->e1);
  host_free(f->h_stage);
  host_free(f->h_up);
  host_free(f->h_pause_ctl);
  dev_free(f->d_keep); dev_free(f->d_pause); dev_free(f->d_pause_ctl);
  delete f;
  return 0;
}

int ia3_fit_first_prepare(ia3_fit* f, int64_t* n_ties) {
  IA3_STAT("ia3_fit_first_prepare");
  if (ensure_device()) return -1;
  if (!f) { set_error("null argument"); return -1; }
  cudaStream_t st = f->s->stream;
  int cap = (int)std::min<int64_t>(std::max<int64_t>(f->n * 8, 4096), (int64_t)1 << 24);
  for (int attempt = 0; attempt < 2; ++attempt) {
    if (cap > f->tie_cap) {
      dev_free(f->d_tie_spot); dev_free(f->d_tie_k);
      f->d_tie_spot = f->d_tie_k = nullptr;
      if (dev_alloc((void**)&f->d_tie_spot, sizeof(int) * (size_t)cap) || dev_alloc((void**)&f->d_tie_k, sizeof(int) * (size_t)cap)) return -1;
      f->tie_cap = cap;
    }
    f->d.tie_cap = f->tie_cap; f->d.tie_spot = f->d_tie_spot; f->d.tie_k = f->d_tie_k;
    IA3_CUDA(cudaMemsetAsync(f->d_tie_count, 0, sizeof(int), st));
    if (launch_voronoi(f->d, st)) return -1;
    if (reserve_pinned(&f->h_stage, &f->stage_cap, 256)) return -1;
    if (small_copy(f->h_stage, f->d_tie_count, sizeof(int), st)) return -1;
    IA3_CUDA(cudaStreamSynchronize(st));
    const int cnt = *static_cast<int*>(f->h_stage);
    f->n_ties = cnt;
    if (cnt <= f->tie_cap) break;
    cap = cnt;
  }
  f->prepared = true;
  if (n_ties) *n_ties = f->n_ties;
  return 0;
}

int ia3_fit_first_ties(ia3_fit* f, int32_t* spot, int32_t* zxy, int64_t cap) {
  IA3_STAT("ia3_fit_first_ties");
  if (ensure_device()) return -1;
  if (!f || !f->prepared) { set_error("first_prepare has not run"); return -1; }
  const int64_t n = std::min<int64_t>(cap, f->n_ties);
  if (n <= 0) return 0;
  if (reserve_pinned(&f->h_stage, &f->stage_cap, 2 * sizeof(int) * (size_t)n)) return -1;
  int* sp = static_cast<int*>(f->h_stage);
  int* kk = sp + n;
  if (small_copy(sp, f->d_tie_spot, sizeof(int) * (size_t)n, f->s->stream) ||
      small_copy(kk, f->d_tie_k, sizeof(int) * (size_t)n, f->s->stream)) return -1;
  IA3_CUDA(cudaStreamSynchronize(f->s->stream));
  for (int64_t i = 0; i < n; ++i) {
    const int sidx = sp[i], k = kk[i];
    spot[i] = sidx;
    for (int a = 0; a < 3; ++a) zxy[3 * i + a] = (int)f->centers[3 * (size_t)sidx + a] + f->offs[3 * (size_t)k + a];
  }
  return 0;
}

int ia3_fit_first_resolve(ia3_fit* f, const uint8_t* keep, int64_t n) {
  IA3_STAT("ia3_fit_first_resolve");
  if (ensure_device()) return -1;
  if (!f || !f->prepared) { set_error("first_prepare has not run"); return -1; }
  n = std::min<int64_t>(n, f->n_ties);
  if (n <= 0) return 0;
  if ((size_t)n > f->keep_cap) {
    dev_free(f->d_keep);
    f->d_keep = nullptr;
    if (dev_alloc((void**)&f->d_keep, (size_t)n)) return -1;
    f->keep_cap = (size_t)n;
  }
  cudaStream_t st = f->s->stream;
  if (reserve_pinned(&f->h_up, &f->up_cap, (size_t)n)) return -1;
  memcpy(f->h_up, keep, (size_t)n);
  if (small_copy(f->d_keep, f->h_up, (size_t)n, st)) return -1;
  if (launch_apply_ties(f->d_mask, f->d.KW, f->d_tie_spot, f->d_tie_k, f->d_keep, (int)n, st)) return -1;
  return 0;        // stream-ordered before first_run; the arena is next touched after first_run's synchronisation
}

static int fetch_results(ia3_fit* f, float* ps, double* p_raw, uint8_t* success, int32_t* nfev, int32_t* info) {
  cudaStream_t st = f->s->stream;
  const size_t n = (size_t)f->n;
  if (n == 0) return 0;
  // device -> pinned staging (asynchronous copies queued behind the kernels), one wait, then plain
  // memcpy into the caller's arrays: no pageable-memory copy ever enters the driver
  const size_t sz[5] = {n * NOUT * 4, n * NP * 8, n, n * 4, n * 4};
  const void* src[5] = {f->d_ps, f->d_praw, f->d_succ, f->d_nfev, f->d_info};
  void* dst[5] = {ps, p_raw, success, nfev, info};
  size_t off[5], total = 0;
  for (int i = 0; i < 5; ++i) { off[i] = total; total += (sz[i] + 255) / 256 * 256; }
  if (reserve_pinned(&f->h_stage, &f->stage_cap, total)) return -1;
  char* h = static_cast<char*>(f->h_stage);
  for (int i = 0; i < 5; ++i)
    if (dst[i] && small_copy(h + off[i], src[i], sz[i], st)) return -1;
  { IA3_STAT("  fetch: wait for stream"); IA3_DRAIN(st); }
  for (int i = 0; i < 5; ++i) if (dst[i]) memcpy(dst[i], h + off[i], sz[i]);
  return 0;
}

// builds per-level work lists of the selected seeds; returns level boundaries
static int build_work(ia3_fit* f, const uint8_t* active, std::vector<int>& bounds) {
  IA3_STAT("  build_work");
  std::vector<std::vector<int>> per(f->n_levels);
  for (int64_t i = 0; i < f->n; ++i)
    if (!active || active[i]) per[f->level[i]].push_back((int)i);
  std::vector<int> flat;
  bounds.assign(1, 0);
  for (auto& v : per) { flat.insert(flat.end(), v.begin(), v.end()); bounds.push_back((int)flat.size()); }
  if (flat.size() > f->work_cap) {
    dev_free(f->d_work);
    f->d_work = nullptr;
    if (dev_alloc((void**)&f->d_work, sizeof(int) * std::max<size_t>(flat.size(), (size_t)f->n))) return -1;
    f->work_cap = std::max<size_t>(flat.size(), (size_t)f->n);
  }
  if (!flat.empty()) {
    // the previous call on this handle ended with a stream synchronisation, so the arena is free
    if (reserve_pinned(&f->h_up, &f->up_cap, sizeof(int) * flat.size())) return -1;
    memcpy(f->h_up, flat.data(), sizeof(int) * flat.size());
    if (small_copy(f->d_work, f->h_up, sizeof(int) * flat.size(), f->s->stream)) return -1;
  }
  return 0;
}

int ia3_fit_first_run(ia3_fit* f, double delta_center, float* ps, double* p_raw, uint8_t* success, int32_t* nfev,
                      int32_t* info) {
  IA3_STAT("ia3_fit_first_run");
  if (ensure_device()) return -1;
  if (!f) { set_error("null argument"); return -1; }
  if (!f->s->d_im) { set_error("the stack's image was released (ia3_stack_trim)"); return -1; }
  if (!f->prepared && ia3_fit_first_prepare(f, nullptr)) return -1;
  cudaStream_t st = f->s->stream;
  f->d.fp.delta = delta_center;
  std::vector<int> bounds;
  if (build_work(f, nullptr, bounds)) return -1;
  IA3_CUDA(cudaEventRecord(f->e0, st));
  if (launch_init_window(f->d, st)) return -1;
  if (run_fit_launch(f, 0, nullptr, f->n)) return -1;
  for (int l = 0; l < f->n_levels; ++l)
    if (launch_subtract(f->d, f->d_work + bounds[l], bounds[l + 1] - bounds[l], st)) return -1;
  if (launch_window_copy(f->d, f->d_snap, nullptr, 0, st)) return -1;
  IA3_CUDA(cudaEventRecord(f->e1, st));
  if (fetch_results(f, ps, p_raw, success, nfev, info)) return -1;
  cudaEventElapsedTime(&f->last_ms, f->e0, f->e1);
  f->first_done = true;
  return 0;
}

int ia3_fit_repeat_sweep(ia3_fit* f, double delta_center, const uint8_t* active, float* ps, double* p_raw,
                         uint8_t* success, int32_t* nfev, int32_t* info) {
  IA3_STAT("ia3_fit_repeat_sweep");
  if (ensure_device()) return -1;
  if (!f || !f->first_done) { set_error("firstfit has not run"); return -1; }
  cudaStream_t st = f->s->stream;
  f->d.fp.delta = delta_center;
  std::vector<int> bounds;
  if (build_work(f, active, bounds)) return -1;
  IA3_CUDA(cudaEventRecord(f->e0, st));
  for (int l = 0; l < f->n_levels; ++l)
    if (run_fit_launch(f, 1, f->d_work + bounds[l], bounds[l + 1] - bounds[l])) return -1;
  IA3_CUDA(cudaEventRecord(f->e1, st));
  if (fetch_results(f, ps, p_raw, success, nfev, info)) return -1;
  cudaEventElapsedTime(&f->last_ms, f->e0, f->e1);
  return 0;
}

int ia3_fit_get_volume(ia3_fit* f, int which, double* out) {
  if (ensure_device()) return -1;
  if (!f || !f->first_done) { set_error("firstfit has not run"); return -1; }
  if (!f->s->d_im) { set_error("the stack's image was released (ia3_stack_trim)"); return -1; }
  cudaStream_t st = f->s->stream;
  double* tmp = nullptr;
  if (dev_alloc((void**)&tmp, f->s->nvox * 8)) return -1;
  if (launch_to_f64(f->s->d_im, f->s->dtype, tmp, (long long)f->s->nvox, st)) return -1;
  if (which == 0) {
    if (launch_window_copy(f->d, f->d_snap, tmp, 1, st)) return -1;
  } else {
    // im_add: current work volume on the window voxels
    double* snap2 = nullptr;
    if (dev_alloc((void**)&snap2, (size_t)std::max<int64_t>(f->n, 1) * f->d.K * 8)) return -1;
    if (launch_window_copy(f->d, snap2, nullptr, 0, st)) return -1;
    if (launch_window_copy(f->d, snap2, tmp, 1, st)) return -1;
    IA3_CUDA(cudaStreamSynchronize(st));
    dev_free(snap2);
  }
  IA3_DRAIN(st);
  IA3_CUDA(cudaMemcpyAsync(out, tmp, f->s->nvox * 8, cudaMemcpyDeviceToHost, st));
  IA3_CUDA(cudaStreamSynchronize(st));
  dev_free(tmp);
  return 0;
}

int ia3_fit_get_rec(ia3_fit* f, int64_t i, double* rec, int32_t* zxy, int32_t* count) {
  if (ensure_device()) return -1;
  if (!f || i < 0 || i >= f->n) { set_error("bad seed index"); return -1; }
  const int K = f->d.K;
  std::vector<double> full(K);
  IA3_CUDA(cudaMemcpyAsync(full.data(), f->d_rec + (size_t)i * K, sizeof(double) * K, cudaMemcpyDeviceToHost, f->s->stream));
  IA3_CUDA(cudaStreamSynchronize(f->s->stream));
  int m = 0;
  for (int k = 0; k < K; ++k) {
    int v[3];
    for (int a = 0; a < 3; ++a) v[a] = (int)f->centers[3 * (size_t)i + a] + f->offs[3 * (size_t)k + a];
    if (v[0] < 0 || v[0] >= f->s->Z || v[1] < 0 || v[1] >= f->s->X || v[2] < 0 || v[2] >= f->s->Y) continue;
    if (rec) rec[m] = full[k];
    if (zxy) { zxy[3 * m] = v[0]; zxy[3 * m + 1] = v[1]; zxy[3 * m + 2] = v[2]; }
    ++m;
  }
  if (count) *count = m;
  return 0;
}

int ia3_fit_num_levels(ia3_fit* f) { return f ? f->n_levels : 0; }
float ia3_fit_last_ms(ia3_fit* f) { return f ? f->last_ms : 0.f; }

// ---- standalone GaussianFit -----------------------------------------------------------------
int ia3_gaussfit_batch(const ia3_fit_cfg* cfg, double delta_center, int64_t n_problems, const int64_t* off,
                       const double* values, const float* coords, const double* centers, float* ps, double* p_raw,
                       uint8_t* success, int32_t* nfev, int32_t* info, double* rec) {
  if (ensure_device()) return -1;
  if (!cfg || n_problems < 0 || (n_problems > 0 && (!off || !values || !coords || !centers))) { set_error("null argument"); return -1; }
  if (n_problems == 0) return 0;
  const int64_t total = off[n_problems];
  cudaStream_t st;
  if (global_stream(&st)) return -1;
  GenericFitDev d;
  memset(&d, 0, sizeof(d));
  long long* d_off; double* d_val; float* d_co; double* d_cen; double* d_tmp; double* d_rec = nullptr;
  float* d_ps; double* d_praw; uint8_t* d_s; int* d_nf; int* d_in;
  const size_t np_ = (size_t)n_problems, tt = (size_t)std::max<int64_t>(total, 1);
  if (dev_alloc((void**)&d_off, (np_ + 1) * 8) || dev_alloc((void**)&d_val, tt * 8) || dev_alloc((void**)&d_co, tt * 12) ||
      dev_alloc((void**)&d_cen, np_ * 24) || dev_alloc((void**)&d_tmp, tt * 8) || dev_alloc((void**)&d_ps, np_ * NOUT * 4) ||
      dev_alloc((void**)&d_praw, np_ * NP * 8) || dev_alloc((void**)&d_s, np_) || dev_alloc((void**)&d_nf, np_ * 4) ||
      dev_alloc((void**)&d_in, np_ * 4) || (rec && dev_alloc((void**)&d_rec, tt * 8))) return -1;
  IA3_CUDA(cudaMemcpyAsync(d_off, off, (np_ + 1) * 8, cudaMemcpyHostToDevice, st));
  IA3_CUDA(cudaMemcpyAsync(d_val, values, (size_t)total * 8, cudaMemcpyHostToDevice, st));
  IA3_CUDA(cudaMemcpyAsync(d_co, coords, (size_t)total * 12, cudaMemcpyHostToDevice, st));
  IA3_CUDA(cudaMemcpyAsync(d_cen, centers, np_ * 24, cudaMemcpyHostToDevice, st));
  d.n = n_problems; d.off = d_off; d.values = d_val; d.coords = d_co; d.centers = d_cen; d.tmp = d_tmp;
  d.ps = d_ps; d.p_raw = d_praw; d.success = d_s; d.nfev = d_nf; d.info = d_in; d.rec = d_rec;
  d.fp.min_w2 = cfg->min_w * cfg->min_w; d.fp.max_w2 = cfg->max_w * cfg->max_w; d.fp.delta = delta_center;
  d.fp.weight_sigma = cfg->weight_sigma; d.fp.personality = cfg->personality;
  d.lm.ftol = 1.49012e-8; d.lm.xtol = 1.49012e-8; d.lm.gtol = 0.0; d.lm.factor = 100.0;
  d.lm.maxfev = cfg->maxfev > 0 ? cfg->maxfev : (cfg->personality == 4 ? 1000 : 1100);
  for (int i = 0; i < 3; ++i) d.init_w[i] = cfg->init_w[i];
  if (launch_generic_fit(d, st)) return -1;
  IA3_DRAIN(st);
  if (ps) IA3_CUDA(cudaMemcpyAsync(ps, d_ps, np_ * NOUT * 4, cudaMemcpyDeviceToHost, st));
  if (p_raw) IA3_CUDA(cudaMemcpyAsync(p_raw, d_praw, np_ * NP * 8, cudaMemcpyDeviceToHost, st));
  if (success) IA3_CUDA(cudaMemcpyAsync(success, d_s, np_, cudaMemcpyDeviceToHost, st));
  if (nfev) IA3_CUDA(cudaMemcpyAsync(nfev, d_nf, np_ * 4, cudaMemcpyDeviceToHost, st));
  if (info) IA3_CUDA(cudaMemcpyAsync(info, d_in, np_ * 4, cudaMemcpyDeviceToHost, st));
  if (rec) IA3_CUDA(cudaMemcpyAsync(rec, d_rec, (size_t)total * 8, cudaMemcpyDeviceToHost, st));
  IA3_CUDA(cudaStreamSynchronize(st));
  void* ptrs[] = {d_off, d_val, d_co, d_cen, d_tmp, d_rec, d_ps, d_praw, d_s, d_nf, d_in};
  for (void* p : ptrs) dev_free(p);
  return 0;
}

int ia3_moment_fit(ia3_stack* s, const double* centers_zxy, int64_t n, const ia3_moment_cfg* cfg, double* out) {
  IA3_STAT("ia3_moment_fit");
  if (ensure_device()) return -1;
  if (!s || !cfg || (n > 0 && (!centers_zxy || !out))) { set_error("null argument"); return -1; }
  if (!s->d_im) { set_error("the stack's image was released (ia3_stack_trim)"); return -1; }
  if (cfg->radius < 1 || cfg->radius > 7) { set_error("radius_fit must be in 1..7"); return -1; }
  if (n <= 0) return 0;
  for (int64_t i = 0; i < 3 * n; ++i)
    if (!(std::fabs(centers_zxy[i]) < 1e6)) { set_error("seed coordinates must be finite and |c| < 1e6"); return -1; }
  const int r = cfg->radius;
  std::vector<int8_t> offs;
  for (int a = -r; a < r; ++a) for (int b = -r; b < r; ++b) for (int c = -r; c < r; ++c)
    if (a * a + b * b + c * c <= r * r) { offs.push_back((int8_t)a); offs.push_back((int8_t)b); offs.push_back((int8_t)c); }
  const int K = (int)(offs.size() / 3);
  // cKDTree.query_ball_tree(tree, 2r): seeds within 2r (inclusive), ascending, self included
  std::vector<int> nbr_start(n + 1, 0), nbr_idx;
  if (cfg->avoid_neighbors) {
    const double reach = 2.0 * r, cs = reach;
    std::unordered_map<long long, std::vector<int>> cells;
    cells.reserve((size_t)n * 2 + 16);
    auto cellc = [&](double v) { return (long long)std::floor(v / cs); };
    for (int64_t i = 0; i < n; ++i) {
      const double* c = centers_zxy + 3 * i;
      cells[cell_key(cellc(c[0]), cellc(c[1]), cellc(c[2]))].push_back((int)i);
    }
    for (int64_t i = 0; i < n; ++i) {
      const double* c = centers_zxy + 3 * i;
      const long long a = cellc(c[0]), b = cellc(c[1]), cc = cellc(c[2]);
      const size_t begin = nbr_idx.size();
      for (long long da = -1; da <= 1; ++da) for (long long db = -1; db <= 1; ++db) for (long long dc = -1; dc <= 1; ++dc) {
        auto it = cells.find(cell_key(a + da, b + db, cc + dc));
        if (it == cells.end()) continue;
        for (int j : it->second) {
          const double* q = centers_zxy + 3 * (size_t)j;
          const double d0 = q[0] - c[0], d1 = q[1] - c[1], d2 = q[2] - c[2];
          if (std::sqrt(d0 * d0 + d1 * d1 + d2 * d2) <= reach) nbr_idx.push_back(j);
        }
      }
      std::sort(nbr_idx.begin() + begin, nbr_idx.end());
      nbr_start[i + 1] = (int)nbr_idx.size();
    }
  }
  cudaStream_t st = s->stream;
  auto al = [](size_t b) { return (b + 255) / 256 * 256; };
  const size_t b_cen = (size_t)n * 24, b_ns = (size_t)(n + 1) * 4, b_ni = std::max<size_t>(nbr_idx.size(), 1) * 4, b_off = offs.size(), b_out = (size_t)n * 96;
  void* h = nullptr;
  if (host_alloc(&h, al(b_cen) + al(b_ns) + al(b_ni) + al(b_off) + al(b_out))) return -1;
  double* d_cen = nullptr; int* d_ns = nullptr; int* d_ni = nullptr; int8_t* d_off = nullptr; double* d_out = nullptr;
  if (dev_alloc((void**)&d_cen, b_cen) || dev_alloc((void**)&d_ns, b_ns) || dev_alloc((void**)&d_ni, b_ni) ||
      dev_alloc((void**)&d_off, b_off) || dev_alloc((void**)&d_out, b_out)) return -1;
  char* hp = static_cast<char*>(h);
  size_t off = 0;
  auto put = [&](void* d, const void* src, size_t bytes) -> int {
    memcpy(hp + off, src, bytes);
    if (small_copy(d, hp + off, bytes, st)) return -1;
    off += (bytes + 255) / 256 * 256;
    return 0;
  };
  if (put(d_cen, centers_zxy, b_cen) || put(d_ns, nbr_start.data(), b_ns) || (nbr_idx.size() && put(d_ni, nbr_idx.data(), nbr_idx.size() * 4)) ||
      put(d_off, offs.data(), b_off)) return -1;
  MomentDev d;
  memset(&d, 0, sizeof(d));
  d.im = s->d_im; d.im_dtype = s->dtype; d.Z = s->Z; d.X = s->X; d.Y = s->Y; d.n = n; d.centers = d_cen;
  d.nbr_start = d_ns; d.nbr_idx = d_ni; d.K = K; d.offs = d_off; d.avoid = cfg->avoid_neighbors ? 1 : 0;
  d.recenter = cfg->recenter ? 1 : 0; d.bk_f = cfg->bk_f; d.out = d_out;
  if (launch_moment_fit(d, st)) return -1;
  char* ho = hp + off;
  if (small_copy(ho, d_out, b_out, st)) return -1;
  IA3_CUDA(cudaStreamSynchronize(st));
  memcpy(out, ho, b_out);
  host_free(h);
  dev_free(d_cen); dev_free(d_ns); dev_free(d_ni); dev_free(d_off); dev_free(d_out);
  return 0;
}

int ia3_gauss_eval(const ia3_fit_cfg* cfg, double delta_center, const double* p_raw, const double* center,
                   const float* coords, int64_t m, double* out) {
  if (ensure_device()) return -1;
  if (!cfg || !p_raw || !center || (m > 0 && (!coords || !out))) { set_error("null argument"); return -1; }
  if (m == 0) return 0;
  FitParams fp;
  memset(&fp, 0, sizeof(fp));
  fp.min_w2 = cfg->min_w * cfg->min_w; fp.max_w2 = cfg->max_w * cfg->max_w; fp.delta = delta_center;
  fp.weight_sigma = 0.0; fp.personality = cfg->personality;
  cudaStream_t st;
  if (global_stream(&st)) return -1;
  double* d_p; double* d_c; float* d_co; double* d_out;
  if (dev_alloc((void**)&d_p, 80) || dev_alloc((void**)&d_c, 24) || dev_alloc((void**)&d_co, (size_t)m * 12) ||
      dev_alloc((void**)&d_out, (size_t)m * 8)) return -1;
  IA3_CUDA(cudaMemcpyAsync(d_p, p_raw, 80, cudaMemcpyHostToDevice, st));
  IA3_CUDA(cudaMemcpyAsync(d_c, center, 24, cudaMemcpyHostToDevice, st));
  IA3_CUDA(cudaMemcpyAsync(d_co, coords, (size_t)m * 12, cudaMemcpyHostToDevice, st));
  if (launch_eval_f0(fp, d_p, d_c, d_co, m, d_out, st)) return -1;
  IA3_DRAIN(st);
  IA3_CUDA(cudaMemcpyAsync(out, d_out, (size_t)m * 8, cudaMemcpyDeviceToHost, st));
  IA3_CUDA(cudaStreamSynchronize(st));
  dev_free(d_p); dev_free(d_c); dev_free(d_co); dev_free(d_out);
  return 0;
}

}  // extern "C"
