// C ABI of libia3b200.so (declared in include/ia3b200.h): handles, host-side orchestration (pooled
// streams / events / device and pinned blocks, image upload, the round loop of the fit engine) and
// kernel launches.  Neighbour lists, dependency relations and the brick table are built on the device.
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
#include "aux_kernels.h"
#include "corr_kernels.h"

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
// Fit rounds are many short kernels on a dependent chain; the seed stage of the other stacks in flight is a few
// long kernels that fill every SM.  Fit streams get the highest priority, so a round's CTAs are placed as soon
// as an SM slot frees instead of waiting for a whole seed kernel to drain (a round would otherwise pick up
// ~1 ms of queueing, fifty times per stack).
static std::vector<cudaStream_t> g_hi_stream_pool;
static int acquire_stream_hi(cudaStream_t* out) {
  {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_hi_stream_pool.empty()) { *out = g_hi_stream_pool.back(); g_hi_stream_pool.pop_back(); return 0; }
  }
  int lo = 0, hi = 0;
  IA3_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  IA3_CUDA(cudaStreamCreateWithPriority(out, cudaStreamNonBlocking, hi));
  return 0;
}
static void release_stream_hi(cudaStream_t st) {
  if (!st) return;
  std::lock_guard<std::mutex> lk(g_mu);
  g_hi_stream_pool.push_back(st);
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
static std::vector<cudaEvent_t> g_bs_event_pool;  // blocking-sync events: a waiting host thread sleeps
static int acquire_event_bs(cudaEvent_t* out) {
  {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_bs_event_pool.empty()) { *out = g_bs_event_pool.back(); g_bs_event_pool.pop_back(); return 0; }
  }
  IA3_CUDA(cudaEventCreateWithFlags(out, cudaEventBlockingSync | cudaEventDisableTiming));
  return 0;
}
static void release_event_bs(cudaEvent_t e) {
  if (!e) return;
  std::lock_guard<std::mutex> lk(g_mu);
  g_bs_event_pool.push_back(e);
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

struct ia3_fit;

// device and pinned blocks that a call must hand back on every path out of it
struct Scoped {
  std::vector<void*> dev, host;
  ~Scoped() { for (void* p : dev) dev_free(p); for (void* p : host) host_free(p); }
  template <typename T> int dalloc(T** p, size_t bytes) { if (dev_alloc((void**)p, bytes)) return -1; dev.push_back(*p); return 0; }
  int halloc(void** p, size_t bytes) { if (host_alloc(p, bytes)) return -1; host.push_back(*p); return 0; }
};

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

}  // extern "C"

// ---- image upload ---------------------------------------------------------------------------
// A cudaMemcpyAsync from PAGEABLE memory is staged by the driver through its own small bounce buffer,
// synchronously, at a fraction of the link rate, and it stalls the other host threads' CUDA calls while
// it runs.  The reference's callers hand over plain numpy arrays, so pageable input is the normal
// case: it goes through pinned staging chunks filled by a few worker threads (one thread's memcpy is
// ~10 GB/s, the link takes 55) -- the host memcpy of one chunk overlaps the DMA of the others, every
// chunk is one cudaMemcpyAsync from pinned memory on the shared upload stream.  Pinned input (allocated
// or registered with CUDA by the caller) is handed to the copy engine directly.
constexpr size_t kStageChunk = (size_t)4 << 20;     // one DMA
constexpr size_t kStageJob = (size_t)16 << 20;      // one queue entry (contiguous part of an image)
struct StageJob { const char* src; char* dst; size_t bytes; std::atomic<int>* pending; std::atomic<int>* err; bool down = false; };
static std::mutex& g_sq_mu = *new std::mutex;
static std::condition_variable& g_sq_cv = *new std::condition_variable;
static std::condition_variable& g_sq_done = *new std::condition_variable;
static std::vector<StageJob>& g_sq = *new std::vector<StageJob>;
static size_t g_sq_head = 0;
static std::once_flag g_sq_once;

static void stage_worker(int device, cudaStream_t us, cudaStream_t ds) {
  void* buf[2] = {nullptr, nullptr};
  cudaEvent_t ev[2] = {nullptr, nullptr};
  bool used[2] = {false, false};
  bool ok = cudaSetDevice(device) == cudaSuccess;
  for (int i = 0; i < 2 && ok; ++i)
    ok = cudaMallocHost(&buf[i], kStageChunk) == cudaSuccess && cudaEventCreateWithFlags(&ev[i], cudaEventBlockingSync | cudaEventDisableTiming) == cudaSuccess;
  int k = 0;
  for (;;) {
    StageJob j;
    {
      std::unique_lock<std::mutex> lk(g_sq_mu);
      g_sq_cv.wait(lk, [] { return g_sq_head < g_sq.size(); });
      j = g_sq[g_sq_head++];
      if (g_sq_head == g_sq.size()) { g_sq.clear(); g_sq_head = 0; }
    }
    cudaError_t e = ok ? cudaSuccess : cudaErrorMemoryAllocation;
    if (j.down) {
      // device -> pageable host: the DMA of one chunk runs while the previous one is copied out of its pinned buffer
      for (int i = 0; i < 2 && e == cudaSuccess; ++i) if (used[i]) { e = cudaEventSynchronize(ev[i]); used[i] = false; }
      size_t prev_off = 0, prev_nb = 0; int prev_k = -1;
      for (size_t off = 0; off < j.bytes && e == cudaSuccess; off += kStageChunk) {
        const size_t nb = std::min(kStageChunk, j.bytes - off);
        e = cudaMemcpyAsync(buf[k], j.src + off, nb, cudaMemcpyDeviceToHost, ds);
        if (e == cudaSuccess) e = cudaEventRecord(ev[k], ds);
        if (e == cudaSuccess && prev_k >= 0) { e = cudaEventSynchronize(ev[prev_k]); if (e == cudaSuccess) memcpy(j.dst + prev_off, buf[prev_k], prev_nb); }
        prev_k = k; prev_off = off; prev_nb = nb;
        k ^= 1;
      }
      if (e == cudaSuccess && prev_k >= 0) { e = cudaEventSynchronize(ev[prev_k]); if (e == cudaSuccess) memcpy(j.dst + prev_off, buf[prev_k], prev_nb); }
      if (e != cudaSuccess) j.err->store(1);
      if (j.pending->fetch_sub(1) == 1) { std::lock_guard<std::mutex> lk(g_sq_mu); g_sq_done.notify_all(); }
      continue;
    }
    for (size_t off = 0; off < j.bytes && e == cudaSuccess; off += kStageChunk) {
      const size_t nb = std::min(kStageChunk, j.bytes - off);
      if (used[k]) e = cudaEventSynchronize(ev[k]);            // the DMA that last read this chunk is done
      if (e != cudaSuccess) break;
      memcpy(buf[k], j.src + off, nb);
      e = cudaMemcpyAsync(j.dst + off, buf[k], nb, cudaMemcpyHostToDevice, us);
      if (e == cudaSuccess) e = cudaEventRecord(ev[k], us);
      used[k] = true;
      k ^= 1;
    }
    if (e != cudaSuccess) j.err->store(1);
    if (j.pending->fetch_sub(1) == 1) { std::lock_guard<std::mutex> lk(g_sq_mu); g_sq_done.notify_all(); }
  }
}

static bool is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged || a.type == cudaMemoryTypeDevice;
}

static cudaStream_t g_down_stream = nullptr;
static int start_stage_workers() {
  cudaStream_t us;
  if (upload_stream(&us)) return -1;
  std::call_once(g_sq_once, [us] {
    int nthr = 6;
    if (const char* ev = getenv("IA3_STAGE_THREADS")) nthr = atoi(ev);
    nthr = std::max(1, std::min(nthr, 32));
    if (cudaStreamCreateWithFlags(&g_down_stream, cudaStreamNonBlocking) != cudaSuccess) g_down_stream = us;
    for (int i = 0; i < nthr; ++i) std::thread(stage_worker, g_device, us, g_down_stream).detach();
  });
  return 0;
}

// device -> host copy of a whole image; pageable destinations go through the staging workers
static int download_image(const void* d_src, void* dst, size_t bytes, cudaStream_t st) {
  IA3_CUDA(cudaStreamSynchronize(st));                              // the stack's kernels are done
  if (is_pinned(dst)) {
    IA3_CUDA(cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToHost, st));
    IA3_CUDA(cudaStreamSynchronize(st));
    return 0;
  }
  if (start_stage_workers()) return -1;
  std::atomic<int> pending{0}, err{0};
  const int njobs = (int)((bytes + kStageJob - 1) / kStageJob);
  pending.store(njobs);
  {
    std::lock_guard<std::mutex> lk(g_sq_mu);
    for (int j = 0; j < njobs; ++j) {
      const size_t off = (size_t)j * kStageJob;
      g_sq.push_back(StageJob{static_cast<const char*>(d_src) + off, static_cast<char*>(dst) + off, std::min(kStageJob, bytes - off), &pending, &err, true});
    }
  }
  g_sq_cv.notify_all();
  {
    std::unique_lock<std::mutex> lk(g_sq_mu);
    g_sq_done.wait(lk, [&] { return pending.load() == 0; });
  }
  if (err.load()) { set_error("image download failed in the staging workers"); return -1; }
  return 0;
}

// host -> device copy of `bytes` on the shared upload stream; `ev` (blocking sync) is recorded behind the last DMA and
// waited for.  Pinned sources go to the copy engine directly, pageable ones through the staging workers.
static int upload_bytes(void* d_dst, const void* src, size_t bytes, cudaEvent_t ev, cudaStream_t then_stream) {
  cudaStream_t us;
  if (upload_stream(&us)) return -1;
  cudaError_t e = cudaSuccess;
  if (is_pinned(src)) {
    std::lock_guard<std::mutex> lk(g_upload_mu);
    e = cudaMemcpyAsync(d_dst, src, bytes, cudaMemcpyHostToDevice, us);
    if (e == cudaSuccess) e = cudaEventRecord(ev, us);
  } else {
    IA3_STAT("  upload: pinned staging");
    if (start_stage_workers()) return -1;
    std::atomic<int> pending{0}, err{0};
    const int njobs = (int)((bytes + kStageJob - 1) / kStageJob);
    pending.store(njobs);
    {
      std::lock_guard<std::mutex> lk(g_sq_mu);
      for (int j = 0; j < njobs; ++j) {
        const size_t off = (size_t)j * kStageJob;
        g_sq.push_back(StageJob{static_cast<const char*>(src) + off, static_cast<char*>(d_dst) + off, std::min(kStageJob, bytes - off), &pending, &err});
      }
    }
    g_sq_cv.notify_all();
    {
      std::unique_lock<std::mutex> lk(g_sq_mu);
      g_sq_done.wait(lk, [&] { return pending.load() == 0; });
    }
    if (err.load()) { set_error("upload failed in the staging workers"); return -1; }
    std::lock_guard<std::mutex> lk(g_upload_mu);
    e = cudaEventRecord(ev, us);                                // behind every chunk's DMA
  }
  if (e == cudaSuccess && then_stream) e = cudaStreamWaitEvent(then_stream, ev, 0);
  if (e == cudaSuccess) e = cudaEventSynchronize(ev);
  if (e != cudaSuccess) { set_error(std::string("upload failed: ") + cudaGetErrorString(e)); return -1; }
  return 0;
}

static int upload_image(ia3_stack* s, const void* im, size_t bytes) { return upload_bytes(s->d_im, im, bytes, s->ev[5], s->stream); }

static bool is_device_ptr(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice;
}

extern "C" {

// ---- stacks ---------------------------------------------------------------------------------
static int stack_common(ia3_stack* s, int dtype, int Z, int X, int Y) {
  if (dtype < 0 || dtype > 2 || Z <= 0 || X <= 0 || Y <= 0) { set_error("bad stack dtype/shape"); return -1; }
  s->dtype = dtype; s->Z = Z; s->X = X; s->Y = Y;
  s->nvox = (size_t)Z * X * Y;
  if (acquire_stream(&s->stream)) return -1;
  for (int i = 0; i < 5; ++i) if (acquire_event(&s->ev[i])) return -1;        // timing events of the seed stage
  if (acquire_event_bs(&s->ev[5])) return -1;                                   // end of the upload: waited on without spinning
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
  if (upload_image(s, im, s->nvox * dtype_size(dtype))) { ia3_stack_destroy(s); return -1; }
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
  for (int i = 0; i < 5; ++i) release_event(s->ev[i]);
  release_event_bs(s->ev[5]);
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
    if (gaussian_filter_exact<Tin>(im, reinterpret_cast<Tin*>(*buf), reinterpret_cast<Tin*>(s->scratch), s->Z, s->X, s->Y, gw, st, cfg->two_d != 0)) return -1;
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
  d.loZ = cfg->two_d ? -(1 << 30) : d.lo;
  d.hiZ = cfg->two_d ? (1 << 30) : (int)std::floor((double)s->Z - cfg->edge);
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
  if (cfg->two_d && s->Z != 1) { set_error("two_d: the image must be held as a one-plane stack"); return -1; }
  if (s->dtype == IA3_DTYPE_U16) return seed_run_t<uint16_t>(s, cfg, n_candidates, t);
  if (s->dtype == IA3_DTYPE_F32) return seed_run_t<float>(s, cfg, n_candidates, t);
  if (s->dtype == IA3_DTYPE_F64) return seed_run_t<double>(s, cfg, n_candidates, t);
  set_error("bad stack dtype");
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
  Scoped sc;
  void* h = nullptr;
  int* d_boxes = nullptr;
  double* d_out = nullptr;
  const size_t bb = (size_t)n * 6 * 4, ob = (size_t)n * 8;
  if (sc.halloc(&h, bb + ob + 256) || sc.dalloc(&d_boxes, bb) || sc.dalloc(&d_out, ob)) return -1;
  memcpy(h, boxes, bb);
  char* ho = static_cast<char*>(h) + (bb + 255) / 256 * 256;
  if (small_copy(d_boxes, h, bb, st)) return -1;
  const bool whole = (n == 1 && boxes[0] == 0 && boxes[1] == s->Z && boxes[2] == 0 && boxes[3] == s->X && boxes[4] == 0 && boxes[5] == s->Y);
  if (whole) {
    unsigned* d_ghist = nullptr;
    if (sc.dalloc(&d_ghist, (size_t)nbins * 4)) return -1;
    if (volume_background(reinterpret_cast<const uint16_t*>(s->d_im), s->Z, s->X, s->Y, d_boxes, first, bin_size, nbins, max_iter,
                          d_ghist, d_out, st)) return -1;
  } else if (box_background(reinterpret_cast<const uint16_t*>(s->d_im), s->X, s->Y, d_boxes, n, first, bin_size, nbins, max_iter, d_out, st)) return -1;
  if (small_copy(ho, d_out, ob, st)) return -1;
  IA3_CUDA(cudaStreamSynchronize(st));
  memcpy(out, ho, ob);
  return 0;
}

// Values of an intermediate seed-stage volume at n voxels (flat C-order indices): which = 0 foreground
// blur, 1 background blur; out has the stack's dtype.  (Fitting_v3.get_seed_points_base reads its two
// blurs at the candidate voxels only, External/Fitting_v3.py:276-283.)
int ia3_seed_gather_volume(ia3_stack* s, int which, const int64_t* flat_idx, int64_t n, void* out) {
  if (ensure_device()) return -1;
  if (!s || (n > 0 && (!flat_idx || !out))) { set_error("null argument"); return -1; }
  const void* src = which == 0 ? s->fg_final : s->bg_final;
  if (!src) { set_error("seed stage has not run"); return -1; }
  if (n <= 0) return 0;
  for (int64_t i = 0; i < n; ++i)
    if (flat_idx[i] < 0 || (size_t)flat_idx[i] >= s->nvox) { set_error("voxel index outside the stack"); return -1; }
  Scoped sc;
  const size_t bi = (size_t)n * 8, bo = (size_t)n * dtype_size(s->dtype), obo = (bi + 255) / 256 * 256;
  void* h = nullptr; long long* d_idx = nullptr; void* d_out = nullptr;
  if (sc.halloc(&h, obo + bo) || sc.dalloc(&d_idx, bi) || sc.dalloc(&d_out, bo)) return -1;
  memcpy(h, flat_idx, bi);
  cudaStream_t st = s->stream;
  if (small_copy(d_idx, h, bi, st)) return -1;
  if (launch_gather_u16(src, s->dtype, d_idx, n, d_out, st)) return -1;
  if (small_copy(static_cast<char*>(h) + obo, d_out, bo, st)) return -1;
  IA3_CUDA(cudaStreamSynchronize(st));
  memcpy(out, static_cast<char*>(h) + obo, bo);
  return 0;
}

// ---- fit stage ------------------------------------------------------------------------------
}  // extern "C"

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
// function evaluations per round for one-warp runs; >= team_after evaluations so far: a team of warps
// continues the run, team_cap (long: nothing else pending) evaluations per round.  0 = never suspend.
static const int g_cap_bulk = std::max(0, env_int("IA3_FIT_CAP", 12));
static const int g_team_after = std::max(1, env_int("IA3_FIT_TEAM_AFTER", 48));
static const int g_team_cap = std::max(1, env_int("IA3_FIT_TEAM_CAP", 32));
static const int g_team_cap_long = std::max(1, env_int("IA3_FIT_TEAM_CAP_LONG", 160));
static const int g_memo_on = env_int("IA3_FIT_MEMO", 1) != 0;
static const int g_spec_on = env_int("IA3_FIT_SPEC", 1) != 0;
static const int g_chunk = std::max(1, env_int("IA3_FIT_CHUNK", 4));
static const int g_merge_small = std::max(0, env_int("IA3_FIT_MERGE", 74));   // <= this many tasks in a round: all of them run as team tasks
// CTAs per launch: the first two rounds of a run hold one task per seed, later ones a few % of that
static const int g_grid_bulk = std::max(1, env_int("IA3_FIT_GRID", 148 * 16));
static const int g_grid_bulk_late = std::max(1, env_int("IA3_FIT_GRID_LATE", 148 * 3));
static const int g_grid_team = std::max(1, env_int("IA3_FIT_GRID_TEAM", 148));
static const int g_grid_team_late = std::max(1, env_int("IA3_FIT_GRID_TEAM_LATE", 74));

struct ia3_fit {
  ia3_stack* s = nullptr;
  cudaStream_t st = nullptr;                      // high-priority stream of the fit rounds (the stack's own stream carries upload + seed stage)
  ia3_fit_cfg cfg;
  FitDev d;
  int64_t n = 0;
  std::vector<double> centers;
  std::vector<int8_t> offs;
  bool prepared = false, started = false, first_done = false;
  int64_t n_ties = 0;
  int ties_prefetched = 0;
  cudaEvent_t ev_chunk[2] = {nullptr, nullptr}, e0 = nullptr, e1 = nullptr;
  void* arena = nullptr;                          // all per-seed device arrays
  int* d_nbr_idx = nullptr; int* d_dep_idx = nullptr;
  int* d_tie_spot = nullptr; int* d_tie_k = nullptr;
  double* d_vol = nullptr;
  int* d_cells = nullptr;                         // cell_cnt, cell_start, cell_cur
  uint8_t* d_keep = nullptr; size_t keep_cap = 0;
  CellGrid grid;
  void* h_pin = nullptr;                          // pinned: done flag, copy of EngineCtl, prefetched ties
  void* h_stage = nullptr; size_t stage_cap = 0;  // pinned staging, device -> host (results)
  void* h_up = nullptr; size_t up_cap = 0;        // pinned staging, host -> device (centres, offsets)
  void* h_keep = nullptr; size_t hkeep_cap = 0;   // pinned staging of the tie decisions (its own buffer: the copy is asynchronous)
  int round = 0, gsweep = 0;
  float last_ms = 0.f;
  int n_levels = -1;
  int* h_done() const { return static_cast<int*>(h_pin); }
  EngineCtl* h_ctl() const { return reinterpret_cast<EngineCtl*>(static_cast<char*>(h_pin) + 256); }
  int* h_ties() const { return reinterpret_cast<int*>(static_cast<char*>(h_pin) + 1024); }
};
constexpr int kTiePrefetch = 16384;               // ties copied with the count (one wait instead of two)
constexpr size_t kPinBytes = 1024 + 2 * sizeof(int) * kTiePrefetch;

static int build_neighbours(ia3_fit* f) {
  cudaStream_t st = f->st;
  const size_t nc = (size_t)f->grid.ncell;
  int* cnt = f->d_cells;
  int* start = cnt + nc;
  int* cur = start + nc + 1;
  int* order = cur + nc;
  IA3_CUDA(cudaMemsetAsync(&f->d.ctl->pool_v, 0, 3 * sizeof(int), st));      // pool_v, pool_o, overflow
  return launch_build_neighbours(f->d, f->grid, cnt, start, cur, order, st);
}

extern "C" {

int ia3_fit_create(ia3_stack* s, const double* centers_zxy, int64_t n, const ia3_fit_cfg* cfg, ia3_fit** out) {
  IA3_STAT("ia3_fit_create");
  if (ensure_device()) return -1;
  if (!s || !cfg || !out || (n > 0 && !centers_zxy)) { set_error("null argument"); return -1; }
  if (!s->d_im) { set_error("the stack's image was released (ia3_stack_trim)"); return -1; }
  if (cfg->radius < 1 || cfg->radius > 7) { set_error("radius_fit must be in 1..7 (window of at most 2048 voxels)"); return -1; }
  if (cfg->personality != 3 && cfg->personality != 4) { set_error("personality must be 3 or 4"); return -1; }
  if (cfg->eval_fp32) { set_error("eval_fp32 was removed: the fit evaluates in FP64 like the reference"); return -1; }
  if (n >= (int64_t)TASK_SEED_MASK / 2) { set_error("too many seeds"); return -1; }
  // window centres are int(c): keep them representable
  double lo3[3] = {0, 0, 0}, hi3[3] = {0, 0, 0};
  for (int64_t i = 0; i < n; ++i)
    for (int a = 0; a < 3; ++a) {
      const double v = centers_zxy[3 * i + a];
      if (!(std::fabs(v) < 1e6)) { set_error("seed coordinates must be finite and |c| < 1e6"); return -1; }
      if (i == 0 || v < lo3[a]) lo3[a] = v;
      if (i == 0 || v > hi3[a]) hi3[a] = v;
    }
  IA3_CUDA(cudaStreamSynchronize(s->stream));      // everything queued on the stack (upload, seed stage) is done before the fit reads the image
  cudaStream_t st = nullptr;
  if (acquire_stream_hi(&st)) return -1;
  ia3_fit* f = new ia3_fit();
  f->s = s; f->cfg = *cfg; f->n = n; f->st = st;
  f->centers.assign(centers_zxy, centers_zxy + 3 * n);
  const int r = cfg->radius;
  // window offsets: np.indices([2r]*3) - r, kept where d^2 <= r^2, C order (Fitting_v4.py:580-583)
  for (int a = -r; a < r; ++a) for (int b = -r; b < r; ++b) for (int c = -r; c < r; ++c)
    if (a * a + b * b + c * c <= r * r) { f->offs.push_back((int8_t)a); f->offs.push_back((int8_t)b); f->offs.push_back((int8_t)c); }
  const int K = (int)(f->offs.size() / 3);
  const int KW = (K + 31) / 32;
  const size_t nn = (size_t)std::max<int64_t>(n, 1);

  // uniform grid of cells of edge >= the neighbour reach (coarser if the seeds are absurdly sparse)
  CellGrid& g = f->grid;
  g.cs = std::ceil(2.0 * ((double)r + 1.7320508075688772) + 1e-6);
  for (;;) {
    double prod = 1.0;
    for (int a = 0; a < 3; ++a) { g.lo[a] = lo3[a]; g.g[a] = (int)std::floor((hi3[a] - lo3[a]) / g.cs) + 1; prod *= (double)g.g[a]; }
    if (prod <= 4.0e6) break;
    g.cs *= 2.0;
  }
  g.ncell = (long long)g.g[0] * g.g[1] * g.g[2];

#define FAIL() do { ia3_fit_destroy(f); return -1; } while (0)
  if (acquire_event_bs(&f->ev_chunk[0]) || acquire_event_bs(&f->ev_chunk[1]) ||
      acquire_event(&f->e0) || acquire_event(&f->e1)) FAIL();
  if (host_alloc(&f->h_pin, kPinBytes)) FAIL();
  memset(f->h_pin, 0, 1024);

  FitDev& d = f->d;
  memset(&d, 0, sizeof(d));
  const int nbz = (s->Z + 7) / 8, nbx = (s->X + 7) / 8, nby = (s->Y + 7) / 8;
  const size_t tab_n = (size_t)nbz * nbx * nby;
  d.list_cap = (int)(2 * nn + 64);
  // one arena for every per-seed array
  size_t off = 0;
  auto carve = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
  const size_t o_ctl = carve(sizeof(EngineCtl)), o_cen = carve(nn * 24), o_offs = carve(f->offs.size()), o_own = carve(nn * 4),
               o_ns = carve(nn * 4), o_nc = carve(nn * 4), o_ds = carve(nn * 4), o_dc = carve(nn * 4), o_nl = carve(nn * 4),
               o_mask = carve(nn * KW * 4), o_ps = carve(nn * NOUT * 4), o_praw = carve(nn * NP * 8), o_succ = carve(nn),
               o_nfev = carve(nn * 4), o_info = carve(nn * 4), o_rec = carve(nn * K * 8), o_pf = carve(nn * NP * 8), o_sf = carve(nn),
               o_stage = carve(nn * 4), o_fin = carve(nn), o_busy = carve(nn), o_conv = carve(nn), o_spec = carve(nn),
               o_sprev = carve(nn), o_cprev = carve(nn * 12), o_dists = carve(nn * 8), o_kd = carve(nn * K * 4), o_kx = carve(nn * NP * 8),
               o_mp = carve(nn * NP * 8), o_mm = carve(nn * 16), o_mv = carve(nn), o_mc = carve(nn), o_live = carve(2 * nn * sizeof(LMLive)),
               o_lists = carve((size_t)6 * d.list_cap * 4), o_tab = carve(tab_n * 4);
  if (dev_alloc(&f->arena, off)) FAIL();
  char* A = static_cast<char*>(f->arena);
  d.ctl = reinterpret_cast<EngineCtl*>(A + o_ctl);
  d.centers = reinterpret_cast<double*>(A + o_cen); d.offs = reinterpret_cast<int8_t*>(A + o_offs);
  d.own_id = reinterpret_cast<int*>(A + o_own); d.nbr_start = reinterpret_cast<int*>(A + o_ns); d.nbr_cnt = reinterpret_cast<int*>(A + o_nc);
  d.dep_start = reinterpret_cast<int*>(A + o_ds); d.dep_cnt = reinterpret_cast<int*>(A + o_dc); d.n_lower = reinterpret_cast<int*>(A + o_nl);
  d.mask = reinterpret_cast<uint32_t*>(A + o_mask); d.ps = reinterpret_cast<float*>(A + o_ps); d.p_raw = reinterpret_cast<double*>(A + o_praw);
  d.success = reinterpret_cast<uint8_t*>(A + o_succ); d.nfev = reinterpret_cast<int*>(A + o_nfev); d.info = reinterpret_cast<int*>(A + o_info);
  d.rec = reinterpret_cast<double*>(A + o_rec); d.praw_first = reinterpret_cast<double*>(A + o_pf); d.succ_first = reinterpret_cast<uint8_t*>(A + o_sf);
  d.stage = reinterpret_cast<int*>(A + o_stage); d.fin = reinterpret_cast<uint8_t*>(A + o_fin); d.busy = reinterpret_cast<uint8_t*>(A + o_busy);
  d.conv = reinterpret_cast<uint8_t*>(A + o_conv); d.specst = reinterpret_cast<uint8_t*>(A + o_spec); d.succ_prev = reinterpret_cast<uint8_t*>(A + o_sprev);
  d.cen_prev = reinterpret_cast<float*>(A + o_cprev); d.dists = reinterpret_cast<double*>(A + o_dists);
  d.key_d32 = reinterpret_cast<float*>(A + o_kd); d.key_x0 = reinterpret_cast<double*>(A + o_kx); d.memo_praw = reinterpret_cast<double*>(A + o_mp);
  d.memo_meta = reinterpret_cast<int*>(A + o_mm); d.memo_valid = reinterpret_cast<uint8_t*>(A + o_mv); d.memo_committed = reinterpret_cast<uint8_t*>(A + o_mc);
  d.live = reinterpret_cast<LMLive*>(A + o_live); d.lists = reinterpret_cast<unsigned*>(A + o_lists); d.brick_tab = reinterpret_cast<int*>(A + o_tab);
  d.h_done = f->h_done();

  d.pool_cap_v = (int)std::min<size_t>(std::max<size_t>(32 * nn, 4096), (size_t)1 << 28);
  d.pool_cap_o = d.pool_cap_v;
  d.tie_cap = (int)std::min<int64_t>(std::max<int64_t>(n * 8, 4096), (int64_t)1 << 24);
  if (dev_alloc((void**)&f->d_nbr_idx, sizeof(int) * (size_t)d.pool_cap_v) || dev_alloc((void**)&f->d_dep_idx, sizeof(int) * (size_t)d.pool_cap_o) ||
      dev_alloc((void**)&f->d_tie_spot, sizeof(int) * (size_t)d.tie_cap) || dev_alloc((void**)&f->d_tie_k, sizeof(int) * (size_t)d.tie_cap) ||
      dev_alloc((void**)&f->d_cells, sizeof(int) * ((size_t)3 * g.ncell + 1 + nn))) FAIL();
  d.nbr_idx = f->d_nbr_idx; d.dep_idx = f->d_dep_idx; d.tie_spot = f->d_tie_spot; d.tie_k = f->d_tie_k;

  d.im = s->d_im; d.im_dtype = s->dtype; d.nbz = nbz; d.nbx = nbx; d.nby = nby;
  d.Z = s->Z; d.X = s->X; d.Y = s->Y;
  d.n = n; d.K = K; d.KW = KW; d.radius = r;
  d.fp.min_w2 = cfg->min_w * cfg->min_w; d.fp.max_w2 = cfg->max_w * cfg->max_w;
  d.fp.delta = 1.0; d.fp.weight_sigma = cfg->weight_sigma; d.fp.personality = cfg->personality;
  d.fp.init_wt[0] = d.fp.init_wt[1] = d.fp.init_wt[2] = 0.0;
  d.lm.ftol = 1.49012e-8; d.lm.xtol = 1.49012e-8; d.lm.gtol = 0.0; d.lm.factor = 100.0;
  d.lm.maxfev = cfg->maxfev > 0 ? cfg->maxfev : (cfg->personality == 4 ? 1000 : 1100);
  for (int i = 0; i < 3; ++i) d.init_w[i] = cfg->init_w[i];
  d.delta_first = 1.0; d.delta_repeat = 2.5; d.th2 = 0.01; d.max_sweeps = 11;
  d.cap_bulk = g_cap_bulk; d.team_after = g_team_after; d.cap_team_short = g_team_cap; d.cap_team_long = g_team_cap_long;
  d.merge_small = g_merge_small;
  d.memo_on = g_memo_on;

  // inputs go through one pinned arena (no pageable copies); it is next written after a synchronisation
  const size_t b_cen = (size_t)n * 24, b_off = f->offs.size();
  if (reserve_pinned(&f->h_up, &f->up_cap, (b_cen + 255) / 256 * 256 + b_off)) FAIL();
  char* hu = static_cast<char*>(f->h_up);
  if (b_cen) memcpy(hu, f->centers.data(), b_cen);
  memcpy(hu + (b_cen + 255) / 256 * 256, f->offs.data(), b_off);
  if (small_copy(const_cast<double*>(d.centers), hu, b_cen, st) || small_copy(const_cast<int8_t*>(d.offs), hu + (b_cen + 255) / 256 * 256, b_off, st)) FAIL();
  if (cudaMemsetAsync(d.ctl, 0, sizeof(EngineCtl), st) != cudaSuccess) { set_error("cudaMemsetAsync failed"); FAIL(); }
  if (build_neighbours(f) || launch_build_bricks(d, st)) FAIL();
#undef FAIL
  *out = f;
  return 0;
}

int ia3_fit_destroy(ia3_fit* f) {
  IA3_STAT("ia3_fit_destroy");
  if (!f) return 0;
  if (g_device >= 0) cudaSetDevice(g_device);
  if (f->st) cudaStreamSynchronize(f->st);
  void* ptrs[] = {f->arena, f->d_nbr_idx, f->d_dep_idx, f->d_tie_spot, f->d_tie_k, f->d_vol, f->d_cells, f->d_keep};
  for (void* p : ptrs) dev_free(p);
  release_event(f->e0); release_event(f->e1);
  release_event_bs(f->ev_chunk[0]); release_event_bs(f->ev_chunk[1]);
  release_stream_hi(f->st);
  host_free(f->h_pin); host_free(f->h_stage); host_free(f->h_up); host_free(f->h_keep);
  delete f;
  return 0;
}

int ia3_fit_first_prepare(ia3_fit* f, int64_t* n_ties) {
  IA3_STAT("ia3_fit_first_prepare");
  if (ensure_device()) return -1;
  if (!f) { set_error("null argument"); return -1; }
  if (f->prepared) { if (n_ties) *n_ties = f->n_ties; return 0; }
  cudaStream_t st = f->st;
  FitDev& d = f->d;
  for (int attempt = 0; attempt < 4; ++attempt) {
    IA3_CUDA(cudaMemsetAsync(&d.ctl->tie_count, 0, sizeof(int), st));
    if (launch_voronoi(d, st)) return -1;
    const int npre = std::min(d.tie_cap, kTiePrefetch);
    if (small_copy(f->h_ctl(), d.ctl, sizeof(EngineCtl), st) || small_copy(f->h_ties(), d.tie_spot, sizeof(int) * (size_t)npre, st) ||
        small_copy(f->h_ties() + kTiePrefetch, d.tie_k, sizeof(int) * (size_t)npre, st)) return -1;
    IA3_CUDA(cudaStreamSynchronize(st));
    const EngineCtl& c = *f->h_ctl();
    if (c.overflow) {               // a neighbour pool was too small (dense clusters): the counts are the sizes needed
      dev_free(f->d_nbr_idx); dev_free(f->d_dep_idx);
      f->d_nbr_idx = f->d_dep_idx = nullptr;
      d.pool_cap_v = std::max(c.pool_v, 1); d.pool_cap_o = std::max(c.pool_o, 1);
      if (dev_alloc((void**)&f->d_nbr_idx, sizeof(int) * (size_t)d.pool_cap_v) || dev_alloc((void**)&f->d_dep_idx, sizeof(int) * (size_t)d.pool_cap_o)) return -1;
      d.nbr_idx = f->d_nbr_idx; d.dep_idx = f->d_dep_idx;
      if (build_neighbours(f)) return -1;
      continue;
    }
    f->n_ties = c.tie_count;
    if (c.tie_count > d.tie_cap) {
      dev_free(f->d_tie_spot); dev_free(f->d_tie_k);
      f->d_tie_spot = f->d_tie_k = nullptr;
      d.tie_cap = c.tie_count;
      if (dev_alloc((void**)&f->d_tie_spot, sizeof(int) * (size_t)d.tie_cap) || dev_alloc((void**)&f->d_tie_k, sizeof(int) * (size_t)d.tie_cap)) return -1;
      d.tie_spot = f->d_tie_spot; d.tie_k = f->d_tie_k;
      continue;
    }
    f->ties_prefetched = std::min<int>(c.tie_count, npre);
    // sparse float64 work volume: the 8x8x8 bricks touched by some seed's (clipped) window
    if (!f->d_vol && dev_alloc((void**)&f->d_vol, (size_t)std::max(c.n_bricks, 1) * 512 * 8)) return -1;
    d.vol = f->d_vol;
    f->prepared = true;
    if (n_ties) *n_ties = f->n_ties;
    return 0;
  }
  set_error("first_prepare: neighbour / tie buffers did not settle");
  return -1;
}

int ia3_fit_first_ties(ia3_fit* f, int32_t* spot, int32_t* zxy, int64_t cap) {
  IA3_STAT("ia3_fit_first_ties");
  if (ensure_device()) return -1;
  if (!f || !f->prepared) { set_error("first_prepare has not run"); return -1; }
  const int64_t n = std::min<int64_t>(cap, f->n_ties);
  if (n <= 0) return 0;
  const int* sp = f->h_ties();
  const int* kk = f->h_ties() + kTiePrefetch;
  if (n > f->ties_prefetched) {
    if (reserve_pinned(&f->h_stage, &f->stage_cap, 2 * sizeof(int) * (size_t)n)) return -1;
    int* a = static_cast<int*>(f->h_stage);
    if (small_copy(a, f->d_tie_spot, sizeof(int) * (size_t)n, f->st) ||
        small_copy(a + n, f->d_tie_k, sizeof(int) * (size_t)n, f->st)) return -1;
    IA3_CUDA(cudaStreamSynchronize(f->st));
    sp = a; kk = a + n;
  }
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
  if (f->started) { set_error("first_resolve after the fit has started"); return -1; }
  n = std::min<int64_t>(n, f->n_ties);
  if (n <= 0) return 0;
  if ((size_t)n > f->keep_cap) {
    dev_free(f->d_keep);
    f->d_keep = nullptr;
    if (dev_alloc((void**)&f->d_keep, (size_t)n)) return -1;
    f->keep_cap = (size_t)n;
  }
  cudaStream_t st = f->st;
  // the copy is asynchronous: the flags get a pinned buffer of their own, which nothing else writes
  if (reserve_pinned(&f->h_keep, &f->hkeep_cap, (size_t)n)) return -1;
  memcpy(f->h_keep, keep, (size_t)n);
  if (small_copy(f->d_keep, f->h_keep, (size_t)n, st)) return -1;
  if (launch_apply_ties(f->d.mask, f->d.KW, f->d_tie_spot, f->d_tie_k, f->d_keep, (int)n, st)) return -1;
  return 0;
}

}  // extern "C"

// rounds until the scheduler reports that nothing is left to do.  Rounds are enqueued a chunk ahead of
// the host's look at the done flag, so the device never waits for the host; the rounds still queued
// when the flag is seen find empty work lists.
static int engine_run(ia3_fit* f, int phases, int sweep_cap) {
  cudaStream_t st = f->st;
  const FitDev& d = f->d;
  if (f->n == 0) return 0;
  *(volatile int*)f->h_done() = 0;
  IA3_CUDA(cudaMemsetAsync(&d.ctl->done, 0, sizeof(int), st));
  int rounds_done = 0;
  for (int chunk = 0;; ++chunk) {
    if (chunk > 100000) { set_error("fit engine did not finish"); return -1; }
    for (int r = 0; r < g_chunk; ++r) {
      // one stream per stack: with dozens of stacks in flight every extra stream shares a hardware queue
      // with another stack's, and a cross-stream wait parked in such a queue holds that stack back too
      const bool early = rounds_done < 2;
      if (launch_sched(d, f->round, phases, sweep_cap, st) ||
          launch_fit_round(d, f->round, true, early ? g_grid_team : g_grid_team_late, st) ||
          launch_fit_round(d, f->round, false, early ? g_grid_bulk : g_grid_bulk_late, st)) return -1;
      ++f->round;
      ++rounds_done;
    }
    IA3_CUDA(cudaEventRecord(f->ev_chunk[chunk & 1], st));
    if (chunk >= 1) {
      IA3_STAT("  engine: wait for chunk");
      IA3_CUDA(cudaEventSynchronize(f->ev_chunk[(chunk - 1) & 1]));
      if (*(volatile int*)f->h_done()) break;
    }
  }
  return 0;
}

static int engine_start(ia3_fit* f) {
  if (f->started) return 0;
  cudaStream_t st = f->st;
  if (!f->prepared && ia3_fit_first_prepare(f, nullptr)) return -1;
  if (launch_engine_reset(f->d, st) || launch_member_stats(f->d, st) || launch_init_window(f->d, st)) return -1;
  f->started = true;
  return 0;
}

static int fetch_results(ia3_fit* f, const ia3_fit_out* o) {
  cudaStream_t st = f->st;
  const size_t n = (size_t)f->n;
  if (n == 0 || !o) return 0;
  // device -> pinned staging (copy kernels queued behind the rounds), one wait, then plain memcpy
  // into the caller's arrays: no pageable-memory copy ever enters the driver
  const FitDev& d = f->d;
  constexpr int NA = 10;
  const size_t sz[NA] = {n * NOUT * 4, n * NP * 8, n, n * 4, n * 4, n, n * 8, n * 4, n, n * 12};
  const void* src[NA] = {d.ps, d.p_raw, d.success, d.nfev, d.info, d.conv, d.dists, d.stage, d.succ_prev, d.cen_prev};
  void* dst[NA] = {o->ps, o->p_raw, o->success, o->nfev, o->info, o->converged, o->dists, o->n_visits, o->success_old, o->centers_old};
  size_t off[NA], total = 0;
  for (int i = 0; i < NA; ++i) { off[i] = total; total += (sz[i] + 255) / 256 * 256; }
  if (reserve_pinned(&f->h_stage, &f->stage_cap, total)) return -1;
  char* h = static_cast<char*>(f->h_stage);
  for (int i = 0; i < NA; ++i)
    if (dst[i] && small_copy(h + off[i], src[i], sz[i], st)) return -1;
  { IA3_STAT("  fetch: wait for stream"); IA3_DRAIN(st); }
  for (int i = 0; i < NA; ++i) if (dst[i]) memcpy(dst[i], h + off[i], sz[i]);
  return 0;
}

extern "C" {

int ia3_fit_run(ia3_fit* f, int phases, double min_delta_center, double max_delta_center, double max_dist_th2,
                int n_max_iter, const ia3_fit_out* out) {
  IA3_STAT("ia3_fit_run");
  if (ensure_device()) return -1;
  if (!f) { set_error("null argument"); return -1; }
  if (!(phases & 3)) { set_error("phases: 1 = firstfit, 2 = repeatfit, 3 = both"); return -1; }
  if ((phases & 1) && f->first_done) phases &= ~1;
  if ((phases & 2) && !(phases & 1) && !f->first_done) { set_error("firstfit has not run"); return -1; }
  if ((phases & 1) && !f->s->d_im) { set_error("the stack's image was released (ia3_stack_trim)"); return -1; }
  cudaStream_t st = f->st;
  FitDev& d = f->d;
  if (phases & 1) d.delta_first = min_delta_center;
  if (phases & 2) {
    if (f->started && d.delta_repeat != max_delta_center) IA3_CUDA(cudaMemsetAsync(d.memo_valid, 0, (size_t)std::max<int64_t>(f->n, 1), st));
    d.delta_repeat = max_delta_center;
    d.th2 = max_dist_th2;
    d.max_sweeps = std::max(1, n_max_iter + 1);       // n_iter > n_max_iter stops the reference's loop
  }
  if (engine_start(f)) return -1;
  int ph = phases & 3;
  if (phases == 3 && g_spec_on && g_memo_on) ph |= 4;
  IA3_CUDA(cudaEventRecord(f->e0, st));
  if (ph && engine_run(f, ph, 1 << 20)) return -1;
  IA3_CUDA(cudaEventRecord(f->e1, st));
  if (fetch_results(f, out)) return -1;
  cudaEventElapsedTime(&f->last_ms, f->e0, f->e1);
  if (phases & 1) f->first_done = true;
  return 0;
}

int ia3_fit_first_run(ia3_fit* f, double delta_center, float* ps, double* p_raw, uint8_t* success, int32_t* nfev,
                      int32_t* info) {
  ia3_fit_out o;
  memset(&o, 0, sizeof(o));
  o.ps = ps; o.p_raw = p_raw; o.success = success; o.nfev = nfev; o.info = info;
  return ia3_fit_run(f, 1, delta_center, 0.0, 0.0, 0, &o);
}

int ia3_fit_repeat_sweep(ia3_fit* f, double delta_center, const uint8_t* active, float* ps, double* p_raw,
                         uint8_t* success, int32_t* nfev, int32_t* info) {
  IA3_STAT("ia3_fit_repeat_sweep");
  if (ensure_device()) return -1;
  if (!f || !f->first_done) { set_error("firstfit has not run"); return -1; }
  cudaStream_t st = f->st;
  FitDev& d = f->d;
  const size_t n = (size_t)f->n;
  if (d.delta_repeat != delta_center && f->gsweep > 0) IA3_CUDA(cudaMemsetAsync(d.memo_valid, 0, std::max<size_t>(n, 1), st));
  d.delta_repeat = delta_center;
  d.th2 = -1.0;                        // the caller decides who is visited again
  d.max_sweeps = 1 << 20;
  if (n) {
    // fin = !active (all seeds if active == NULL); the previous call ended with a synchronisation
    if (reserve_pinned(&f->h_up, &f->up_cap, n)) return -1;
    uint8_t* h = static_cast<uint8_t*>(f->h_up);
    for (size_t i = 0; i < n; ++i) h[i] = (active && !active[i]) ? 1 : 0;
    if (small_copy(d.fin, h, n, st)) return -1;
  }
  f->gsweep += 1;
  IA3_CUDA(cudaEventRecord(f->e0, st));
  if (engine_run(f, 2, f->gsweep)) return -1;
  IA3_CUDA(cudaEventRecord(f->e1, st));
  ia3_fit_out o;
  memset(&o, 0, sizeof(o));
  o.ps = ps; o.p_raw = p_raw; o.success = success; o.nfev = nfev; o.info = info;
  if (fetch_results(f, &o)) return -1;
  cudaEventElapsedTime(&f->last_ms, f->e0, f->e1);
  return 0;
}

int ia3_fit_engine_stats(ia3_fit* f, int64_t* out, int cap) {
  if (ensure_device()) return -1;
  if (!f || !out) { set_error("null argument"); return -1; }
  cudaStream_t st = f->st;
  Scoped sc;
  void* h = nullptr;
  if (sc.halloc(&h, sizeof(EngineCtl))) return -1;
  if (small_copy(h, f->d.ctl, sizeof(EngineCtl), st)) return -1;
  IA3_CUDA(cudaStreamSynchronize(st));
  const EngineCtl& c = *static_cast<const EngineCtl*>(h);
  std::vector<int64_t> v = {c.st_rounds, c.st_tasks, c.st_lm_runs, (int64_t)c.st_evals, c.st_memo_hits, c.st_spec_runs, c.st_spec_hits,
                            c.st_parked, c.st_team_tasks, c.n_bricks};
  for (int q = 0; q < 6; ++q) v.push_back((int64_t)c.prof[q]);                       // 10..15
  if (getenv("IA3_LW_PROF")) {
    unsigned long long lp[16];
    if (lw_prof_read(lp) == 0) {
      fprintf(stderr, "lm_warp cycles (process total): factor steps %llu, factor B0+store %llu, outer rest %llu, GN %llu, parl %llu, "
                      "damped solve %llu (its Cholesky %llu) x %llu; factor calls %llu, propose calls %llu, team evals %llu\n", lp[0], lp[1], lp[2], lp[3], lp[4], lp[5], lp[6], lp[7],
              lp[8], lp[9], (unsigned long long)c.prof[5]);
    }
  }
  const int nr = std::min(c.st_rounds, 512);
  for (int r = 0; r < nr; ++r) { v.push_back((int64_t)c.trace_work[r]); v.push_back((int64_t)c.trace_ns[r]); }   // 16 + 2 r
  for (int i = 0; i < cap; ++i) out[i] = i < (int)v.size() ? v[i] : -1;
  return 0;
}

int ia3_fit_get_volume(ia3_fit* f, int which, double* out) {
  if (ensure_device()) return -1;
  if (!f || !f->first_done) { set_error("firstfit has not run"); return -1; }
  if (!f->s->d_im) { set_error("the stack's image was released (ia3_stack_trim)"); return -1; }
  cudaStream_t st = f->st;
  Scoped sc;
  double* tmp = nullptr;
  if (sc.dalloc(&tmp, f->s->nvox * 8)) return -1;
  if (launch_to_f64(f->s->d_im, f->s->dtype, tmp, (long long)f->s->nvox, st)) return -1;
  const size_t nn = (size_t)std::max<int64_t>(f->n, 1);
  if (which == 0) {
    // im_subtr: the image minus firstfit's reconstructions, subtracted in seed order where windows
    // overlap (replayed from firstfit's parameters; passes = dependency depth)
    uint8_t* done = nullptr; int* pending = nullptr; void* hp = nullptr;
    if (sc.dalloc(&done, 2 * nn) || sc.dalloc(&pending, 256) || sc.halloc(&hp, 256)) return -1;
    IA3_CUDA(cudaMemsetAsync(done, 0, 2 * nn, st));
    for (int pass = 0;; ++pass) {
      if (pass > f->n + 1) { set_error("im_subtr: dependency replay did not finish"); return -1; }
      IA3_CUDA(cudaMemsetAsync(pending, 0, sizeof(int), st));
      uint8_t* a = done + (size_t)(pass & 1) * nn;
      uint8_t* b = done + (size_t)((pass & 1) ^ 1) * nn;
      if (launch_subtract_dense(f->d, tmp, a, b, pending, st)) return -1;
      if (small_copy(hp, pending, sizeof(int), st)) return -1;
      IA3_CUDA(cudaStreamSynchronize(st));
      if (*static_cast<int*>(hp) == 0) break;
    }
  } else {
    // im_add: current work volume on the window voxels
    double* snap = nullptr;
    if (sc.dalloc(&snap, nn * f->d.K * 8)) return -1;
    if (launch_window_copy(f->d, snap, nullptr, 0, st)) return -1;
    if (launch_window_copy(f->d, snap, tmp, 1, st)) return -1;
  }
  IA3_DRAIN(st);
  IA3_CUDA(cudaMemcpyAsync(out, tmp, f->s->nvox * 8, cudaMemcpyDeviceToHost, st));
  IA3_CUDA(cudaStreamSynchronize(st));
  return 0;
}

int ia3_fit_get_rec(ia3_fit* f, int64_t i, double* rec, int32_t* zxy, int32_t* count) {
  if (ensure_device()) return -1;
  if (!f || i < 0 || i >= f->n) { set_error("bad seed index"); return -1; }
  const int K = f->d.K;
  std::vector<double> full(K);
  IA3_CUDA(cudaStreamSynchronize(f->st));
  IA3_CUDA(cudaMemcpyAsync(full.data(), f->d.rec + (size_t)i * K, sizeof(double) * K, cudaMemcpyDeviceToHost, f->st));
  IA3_CUDA(cudaStreamSynchronize(f->st));
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

// depth of the dependency order (seed j is one level above every lower-index seed whose window
// overlaps its own): how many seeds in a row the reference's in-order sweep forces apart
int ia3_fit_num_levels(ia3_fit* f) {
  if (!f) return 0;
  if (f->n_levels >= 0) return f->n_levels;
  if (f->n == 0) { f->n_levels = 0; return 0; }
  if (ensure_device() || (!f->prepared && ia3_fit_first_prepare(f, nullptr))) return -1;
  const size_t n = (size_t)f->n;
  const size_t pool = (size_t)std::max(f->h_ctl()->pool_o, 1);
  std::vector<int> ds(n), dc(n), di(pool);
  cudaStream_t st = f->st;
  if (cudaStreamSynchronize(st) != cudaSuccess ||
      cudaMemcpy(ds.data(), f->d.dep_start, n * 4, cudaMemcpyDeviceToHost) != cudaSuccess ||
      cudaMemcpy(dc.data(), f->d.dep_cnt, n * 4, cudaMemcpyDeviceToHost) != cudaSuccess ||
      cudaMemcpy(di.data(), f->d.dep_idx, pool * 4, cudaMemcpyDeviceToHost) != cudaSuccess) { set_error("num_levels: copy failed"); return -1; }
  std::vector<int> level(n, 0);
  int nl = 0;
  for (size_t i = 0; i < n; ++i) {
    int lvl = 0;
    for (int e = ds[i]; e < ds[i] + dc[i]; ++e) if (di[e] < (int)i) lvl = std::max(lvl, level[di[e]] + 1);
    level[i] = lvl;
    nl = std::max(nl, lvl + 1);
  }
  f->n_levels = nl;
  return nl;
}
float ia3_fit_last_ms(ia3_fit* f) { return f ? f->last_ms : 0.f; }

// ---- alternative seeders of Fitting_v4 (a12) ----------------------------------------------------
}  // extern "C"

// np.std(volume) in FP64: two passes (mean, then squared deviations)
template <typename T>
static int volume_std(const T* v, long long n, double* d_acc, void* h_pin, cudaStream_t st, double* mean_out, double* std_out) {
  double* h = static_cast<double*>(h_pin);
  if (launch_moments<T>(v, n, 0.0, d_acc, st) || small_copy(h, d_acc, 16, st)) return -1;
  IA3_CUDA(cudaStreamSynchronize(st));
  const double mean = h[0] / (double)n;
  if (launch_moments<T>(v, n, mean, d_acc, st) || small_copy(h, d_acc, 16, st)) return -1;
  IA3_CUDA(cudaStreamSynchronize(st));
  *mean_out = mean;
  *std_out = std::sqrt(h[1] / (double)n);
  return 0;
}

template <typename Tin>
static int fft_gaussian_dev(ia3_stack* s, const double* gaus, int expo, double* d_a, double* d_b, Scoped& sc) {
  // gker (Fitting_v4.py:11-16): outer product of scipy.signal.gaussian(int(g * exp), g) windows over its sum
  cudaStream_t st = s->stream;
  const Tin* im = reinterpret_cast<const Tin*>(s->d_im);
  const int dims[3] = {s->Z, s->X, s->Y};
  for (int a = 0; a < 3; ++a) {
    const int nt = (int)(gaus[a] * expo);
    if (nt < 2 || nt > 4096 || nt / 2 > dims[a]) { set_error("fft_gaussian: window does not fit the stack"); return -1; }
    std::vector<double> w(nt);
    double sum = 0.0;
    for (int k = 0; k < nt; ++k) { const double x = k - (nt - 1) / 2.0; w[k] = std::exp(-0.5 * (x / gaus[a]) * (x / gaus[a])); sum += w[k]; }
    for (double& v : w) v /= sum;
    void* h = nullptr; double* d_w = nullptr;
    if (sc.halloc(&h, sizeof(double) * nt) || sc.dalloc(&d_w, sizeof(double) * nt)) return -1;
    memcpy(h, w.data(), sizeof(double) * nt);
    if (small_copy(d_w, h, sizeof(double) * nt, st)) return -1;
    int rc;
    if (a == 0) rc = launch_fir_axis<Tin>(im, d_a, s->Z, s->X, s->Y, 0, d_w, nt, st);
    else if (a == 1) rc = launch_fir_axis<double>(d_a, d_b, s->Z, s->X, s->Y, 1, d_w, nt, st);
    else rc = launch_fir_axis<double>(d_b, d_a, s->Z, s->X, s->Y, 2, d_w, nt, st);
    if (rc) return -1;
  }
  return 0;          // result in d_a
}

extern "C" {

int ia3_stack_histogram(ia3_stack* s, uint64_t* counts) {
  if (ensure_device()) return -1;
  if (!s || !counts) { set_error("null argument"); return -1; }
  if (!s->d_im) { set_error("the stack's image was released (ia3_stack_trim)"); return -1; }
  if (s->dtype != IA3_DTYPE_U16) { set_error("ia3_stack_histogram supports uint16 stacks"); return -1; }
  Scoped sc;
  unsigned long long* d_h = nullptr; void* h = nullptr;
  const size_t bytes = 65536 * sizeof(unsigned long long);
  if (sc.dalloc(&d_h, bytes) || sc.halloc(&h, bytes)) return -1;
  cudaStream_t st = s->stream;
  if (launch_hist_u16((const uint16_t*)s->d_im, (long long)s->nvox, d_h, st) || small_copy(h, d_h, bytes, st)) return -1;
  IA3_CUDA(cudaStreamSynchronize(st));
  memcpy(counts, h, bytes);
  return 0;
}

int ia3_seed_v2(ia3_stack* s, int gfilt_size, int filt_size, double th_seed, double* std_out, int64_t* flat_idx, float* h_out,
                int64_t cap, int64_t* n_out) {
  if (ensure_device()) return -1;
  if (!s || !std_out || !n_out || (cap > 0 && (!flat_idx || !h_out))) { set_error("null argument"); return -1; }
  if (!s->d_im) { set_error("the stack's image was released (ia3_stack_trim)"); return -1; }
  if (gfilt_size < 0 || gfilt_size > 255 || filt_size < 1) { set_error("bad filter size"); return -1; }
  cudaStream_t st = s->stream;
  Scoped sc;
  float* d_norm = nullptr; double* d_acc = nullptr; long long* d_idx = nullptr; float* d_h = nullptr; int* d_cnt = nullptr;
  void* h = nullptr;
  const size_t capc = (size_t)std::max<int64_t>(std::min<int64_t>(cap, (int64_t)1 << 24), 1);
  if (sc.dalloc(&d_norm, s->nvox * 4) || sc.dalloc(&d_acc, 256) || sc.dalloc(&d_idx, capc * 8) || sc.dalloc(&d_h, capc * 4) || sc.dalloc(&d_cnt, 256) ||
      sc.halloc(&h, 256 + capc * 12)) return -1;
  const int sz = gfilt_size == 0 ? 1 : gfilt_size;         // a 1x1 box blur leaves im - im = 0: handled below
  int rc;
  if (s->dtype == IA3_DTYPE_U16) rc = launch_box_norm<uint16_t>((const uint16_t*)s->d_im, d_norm, s->Z, s->X, s->Y, sz, st);
  else if (s->dtype == IA3_DTYPE_F32) rc = launch_box_norm<float>((const float*)s->d_im, d_norm, s->Z, s->X, s->Y, sz, st);
  else rc = launch_box_norm<double>((const double*)s->d_im, d_norm, s->Z, s->X, s->Y, sz, st);
  if (rc) return -1;
  if (gfilt_size == 0) { set_error("gfilt_size = 0 (no normalisation) is not supported on the device"); return -1; }
  double mean = 0.0, sd = 0.0;
  if (volume_std<float>(d_norm, (long long)s->nvox, d_acc, h, st, &mean, &sd)) return -1;
  const float std32 = (float)sd;
  const float cutoff = std32 * (float)th_seed;              // np.float32 * python float -> float32
  if (launch_v2_candidates(d_norm, s->Z, s->X, s->Y, cutoff, filt_size / 2, d_idx, d_h, d_cnt, (int)capc, st)) return -1;
  if (small_copy(h, d_cnt, sizeof(int), st)) return -1;
  IA3_CUDA(cudaStreamSynchronize(st));
  const int n = *static_cast<int*>(h);
  *std_out = (double)std32;
  *n_out = n;
  if (n > cap) { set_error("ia3_seed_v2: more candidates than the caller's capacity"); return -2; }
  if (n > 0) {
    char* hp = static_cast<char*>(h) + 256;
    if (small_copy(hp, d_idx, (size_t)n * 8, st) || small_copy(hp + capc * 8, d_h, (size_t)n * 4, st)) return -1;
    IA3_CUDA(cudaStreamSynchronize(st));
    memcpy(flat_idx, hp, (size_t)n * 8);
    memcpy(h_out, hp + capc * 8, (size_t)n * 4);
  }
  return 0;
}

int ia3_fft_gaussian(ia3_stack* s, const double* gaus, int expo, double* out) {
  if (ensure_device()) return -1;
  if (!s || !gaus || !out) { set_error("null argument"); return -1; }
  if (!s->d_im) { set_error("the stack's image was released (ia3_stack_trim)"); return -1; }
  Scoped sc;
  double* d_a = nullptr; double* d_b = nullptr;
  if (sc.dalloc(&d_a, s->nvox * 8) || sc.dalloc(&d_b, s->nvox * 8)) return -1;
  int rc;
  if (s->dtype == IA3_DTYPE_U16) rc = fft_gaussian_dev<uint16_t>(s, gaus, expo, d_a, d_b, sc);
  else if (s->dtype == IA3_DTYPE_F32) rc = fft_gaussian_dev<float>(s, gaus, expo, d_a, d_b, sc);
  else rc = fft_gaussian_dev<double>(s, gaus, expo, d_a, d_b, sc);
  if (rc) return -1;
  IA3_CUDA(cudaStreamSynchronize(s->stream));
  IA3_CUDA(cudaMemcpyAsync(out, d_a, s->nvox * 8, cudaMemcpyDeviceToHost, s->stream));
  IA3_CUDA(cudaStreamSynchronize(s->stream));
  return 0;
}

int ia3_seed_logratio(ia3_stack* s, double gfilt_size, int filt_size, double th_seed, double* std_out, int64_t* flat_idx, double* h_out,
                      int64_t cap, int64_t* n_out) {
  if (ensure_device()) return -1;
  if (!s || !std_out || !n_out || (cap > 0 && (!flat_idx || !h_out))) { set_error("null argument"); return -1; }
  if (!s->d_im) { set_error("the stack's image was released (ia3_stack_trim)"); return -1; }
  if (filt_size < 1) { set_error("bad filter size"); return -1; }
  cudaStream_t st = s->stream;
  Scoped sc;
  double* d_a = nullptr; double* d_b = nullptr; double* d_acc = nullptr; long long* d_idx = nullptr; double* d_h = nullptr; int* d_cnt = nullptr;
  void* h = nullptr;
  const size_t capc = (size_t)std::max<int64_t>(std::min<int64_t>(cap, (int64_t)1 << 24), 1);
  if (sc.dalloc(&d_a, s->nvox * 8) || sc.dalloc(&d_b, s->nvox * 8) || sc.dalloc(&d_acc, 256) || sc.dalloc(&d_idx, capc * 8) || sc.dalloc(&d_h, capc * 8) ||
      sc.dalloc(&d_cnt, 256) || sc.halloc(&h, 256 + capc * 16)) return -1;
  const double gaus[3] = {gfilt_size, gfilt_size, gfilt_size};
  int rc;
  const long long n = (long long)s->nvox;
  if (s->dtype == IA3_DTYPE_U16) rc = fft_gaussian_dev<uint16_t>(s, gaus, 8, d_a, d_b, sc) || launch_log_ratio<uint16_t>((const uint16_t*)s->d_im, d_a, d_b, n, st);
  else if (s->dtype == IA3_DTYPE_F32) rc = fft_gaussian_dev<float>(s, gaus, 8, d_a, d_b, sc) || launch_log_ratio<float>((const float*)s->d_im, d_a, d_b, n, st);
  else rc = fft_gaussian_dev<double>(s, gaus, 8, d_a, d_b, sc) || launch_log_ratio<double>((const double*)s->d_im, d_a, d_b, n, st);
  if (rc) return -1;
  double mean = 0.0, sd = 0.0;
  if (volume_std<double>(d_b, n, d_acc, h, st, &mean, &sd)) return -1;
  if (launch_lr_candidates(d_b, s->Z, s->X, s->Y, th_seed * sd, filt_size, d_idx, d_h, d_cnt, (int)capc, st)) return -1;
  if (small_copy(h, d_cnt, sizeof(int), st)) return -1;
  IA3_CUDA(cudaStreamSynchronize(st));
  const int nc = *static_cast<int*>(h);
  *std_out = sd;
  *n_out = nc;
  if (nc > cap) { set_error("ia3_seed_logratio: more candidates than the caller's capacity"); return -2; }
  if (nc > 0) {
    char* hp = static_cast<char*>(h) + 256;
    if (small_copy(hp, d_idx, (size_t)nc * 8, st) || small_copy(hp + capc * 8, d_h, (size_t)nc * 8, st)) return -1;
    IA3_CUDA(cudaStreamSynchronize(st));
    memcpy(flat_idx, hp, (size_t)nc * 8);
    memcpy(h_out, hp + capc * 8, (size_t)nc * 8);
  }
  return 0;
}

// ---- standalone GaussianFit -----------------------------------------------------------------
}  // extern "C"

// a pooled stream for the duration of a call (concurrent callers do not queue behind each other)
struct ScopedStream {
  cudaStream_t st = nullptr;
  ~ScopedStream() { if (st) release_stream(st); }
  int get() { return acquire_stream(&st); }
};

extern "C" {

int ia3_gaussfit_batch(const ia3_fit_cfg* cfg, double delta_center, int64_t n_problems, const int64_t* off,
                       const double* values, const float* coords, const double* centers, float* ps, double* p_raw,
                       uint8_t* success, int32_t* nfev, int32_t* info, double* rec) {
  if (ensure_device()) return -1;
  if (!cfg || n_problems < 0 || (n_problems > 0 && (!off || !values || !coords || !centers))) { set_error("null argument"); return -1; }
  if (n_problems == 0) return 0;
  const int64_t total = off[n_problems];
  ScopedStream ss;
  if (ss.get()) return -1;
  cudaStream_t st = ss.st;
  Scoped sc;
  GenericFitDev d;
  memset(&d, 0, sizeof(d));
  long long* d_off; double* d_val; float* d_co; double* d_cen; double* d_tmp; double* d_rec = nullptr;
  float* d_ps; double* d_praw; uint8_t* d_s; int* d_nf; int* d_in;
  const size_t np_ = (size_t)n_problems, tt = (size_t)std::max<int64_t>(total, 1);
  if (sc.dalloc(&d_off, (np_ + 1) * 8) || sc.dalloc(&d_val, tt * 8) || sc.dalloc(&d_co, tt * 12) ||
      sc.dalloc(&d_cen, np_ * 24) || sc.dalloc(&d_tmp, tt * 8) || sc.dalloc(&d_ps, np_ * NOUT * 4) ||
      sc.dalloc(&d_praw, np_ * NP * 8) || sc.dalloc(&d_s, np_) || sc.dalloc(&d_nf, np_ * 4) ||
      sc.dalloc(&d_in, np_ * 4) || (rec && sc.dalloc(&d_rec, tt * 8))) return -1;
  // inputs and outputs through one pinned arena (no pageable copy enters the driver)
  auto al = [](size_t b) { return (b + 255) / 256 * 256; };
  const size_t in_sz[4] = {(np_ + 1) * 8, (size_t)total * 8, (size_t)total * 12, np_ * 24};
  const void* in_src[4] = {off, values, coords, centers};
  void* in_dst[4] = {d_off, d_val, d_co, d_cen};
  const size_t out_sz[6] = {np_ * NOUT * 4, np_ * NP * 8, np_, np_ * 4, np_ * 4, (size_t)total * 8};
  const void* out_src[6] = {d_ps, d_praw, d_s, d_nf, d_in, d_rec};
  void* out_dst[6] = {ps, p_raw, success, nfev, info, rec};
  size_t need = 0;
  for (size_t b : in_sz) need += al(b);
  for (size_t b : out_sz) need += al(b);
  void* h = nullptr;
  if (sc.halloc(&h, need)) return -1;
  char* hp = static_cast<char*>(h);
  size_t o = 0;
  for (int i = 0; i < 4; ++i) {
    memcpy(hp + o, in_src[i], in_sz[i]);
    if (small_copy(in_dst[i], hp + o, in_sz[i], st)) return -1;
    o += al(in_sz[i]);
  }
  d.n = n_problems; d.off = d_off; d.values = d_val; d.coords = d_co; d.centers = d_cen; d.tmp = d_tmp;
  d.ps = d_ps; d.p_raw = d_praw; d.success = d_s; d.nfev = d_nf; d.info = d_in; d.rec = d_rec;
  d.fp.min_w2 = cfg->min_w * cfg->min_w; d.fp.max_w2 = cfg->max_w * cfg->max_w; d.fp.delta = delta_center;
  d.fp.weight_sigma = cfg->weight_sigma; d.fp.personality = cfg->personality;
  d.lm.ftol = 1.49012e-8; d.lm.xtol = 1.49012e-8; d.lm.gtol = 0.0; d.lm.factor = 100.0;
  d.lm.maxfev = cfg->maxfev > 0 ? cfg->maxfev : (cfg->personality == 4 ? 1000 : 1100);
  for (int i = 0; i < 3; ++i) d.init_w[i] = cfg->init_w[i];
  if (launch_generic_fit(d, st)) return -1;
  size_t oo[6];
  for (int i = 0; i < 6; ++i) {
    oo[i] = o;
    if (out_dst[i] && small_copy(hp + o, out_src[i], out_sz[i], st)) return -1;
    o += al(out_sz[i]);
  }
  IA3_CUDA(cudaStreamSynchronize(st));
  for (int i = 0; i < 6; ++i) if (out_dst[i]) memcpy(out_dst[i], hp + oo[i], out_sz[i]);
  return 0;
}

static inline long long cell_key(long long a, long long b, long long c) {
  return ((a + (1LL << 20)) << 42) | ((b + (1LL << 20)) << 21) | (c + (1LL << 20));
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
  Scoped sc;
  void* h = nullptr;
  if (sc.halloc(&h, al(b_cen) + al(b_ns) + al(b_ni) + al(b_off) + al(b_out))) return -1;
  double* d_cen = nullptr; int* d_ns = nullptr; int* d_ni = nullptr; int8_t* d_off = nullptr; double* d_out = nullptr;
  if (sc.dalloc(&d_cen, b_cen) || sc.dalloc(&d_ns, b_ns) || sc.dalloc(&d_ni, b_ni) ||
      sc.dalloc(&d_off, b_off) || sc.dalloc(&d_out, b_out)) return -1;
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
  ScopedStream ss;
  if (ss.get()) return -1;
  cudaStream_t st = ss.st;
  Scoped sc;
  double* d_p; double* d_c; float* d_co; double* d_out;
  void* h = nullptr;
  const size_t b_co = (size_t)m * 12, b_out = (size_t)m * 8, o_co = 512, o_out = o_co + (b_co + 255) / 256 * 256;
  if (sc.dalloc(&d_p, 80) || sc.dalloc(&d_c, 24) || sc.dalloc(&d_co, b_co) || sc.dalloc(&d_out, b_out) || sc.halloc(&h, o_out + b_out)) return -1;
  char* hp = static_cast<char*>(h);
  memcpy(hp, p_raw, 80); memcpy(hp + 256, center, 24); memcpy(hp + o_co, coords, b_co);
  if (small_copy(d_p, hp, 80, st) || small_copy(d_c, hp + 256, 24, st) || small_copy(d_co, hp + o_co, b_co, st)) return -1;
  if (launch_eval_f0(fp, d_p, d_c, d_co, m, d_out, st)) return -1;
  if (small_copy(hp + o_out, d_out, b_out, st)) return -1;
  IA3_CUDA(cudaStreamSynchronize(st));
  memcpy(out, hp + o_out, b_out);
  return 0;
}

}  // extern "C"

// ---- pre-processing: correct_fov_image's compute core (corr_kernels.cu) -----------------------
extern "C" {

int ia3_stack_alloc(int dtype, int Z, int X, int Y, ia3_stack** out) {
  if (ensure_device()) return -1;
  if (!out) { set_error("null argument"); return -1; }
  ia3_stack* s = new ia3_stack();
  if (stack_common(s, dtype, Z, X, Y)) { ia3_stack_destroy(s); return -1; }
  if (dev_alloc(&s->d_im, s->nvox * dtype_size(dtype))) { ia3_stack_destroy(s); return -1; }
  s->owns = true;
  *out = s;
  return 0;
}

int ia3_device_upload(const void* host, size_t bytes, void** dev) {
  if (ensure_device()) return -1;
  if (!host || !dev || bytes == 0) { set_error("null argument"); return -1; }
  void* d = nullptr;
  if (dev_alloc(&d, bytes)) return -1;
  cudaEvent_t ev = nullptr;
  if (acquire_event_bs(&ev)) { dev_free(d); return -1; }
  const int rc = upload_bytes(d, host, bytes, ev, nullptr);
  release_event_bs(ev);
  if (rc) { dev_free(d); return -1; }
  *dev = d;
  return 0;
}

int ia3_device_free(void* dev) {
  if (!dev) return 0;
  if (g_device >= 0) cudaSetDevice(g_device);
  cudaDeviceSynchronize();                      // nothing in flight may still read it
  dev_free(dev);
  return 0;
}

int ia3_stack_fetch(ia3_stack* s, void* out) {
  if (ensure_device()) return -1;
  if (!s || !out) { set_error("null argument"); return -1; }
  if (!s->d_im) { set_error("the stack's image was released (ia3_stack_trim)"); return -1; }
  return download_image(s->d_im, out, s->nvox * dtype_size(s->dtype), s->stream);
}

static int corr_check(const ia3_stack* s, const char* who) {
  if (!s) { set_error("null argument"); return -1; }
  if (!s->d_im) { set_error("the stack's image was released (ia3_stack_trim)"); return -1; }
  if (s->dtype != IA3_DTYPE_U16) { set_error(std::string(who) + " works on uint16 stacks (what the microscope files hold)"); return -1; }
  return 0;
}

int ia3_corr_hot_pixels(ia3_stack* s, double hot_th, double hot_pix_th, int64_t* n_hot) {
  IA3_STAT("ia3_corr_hot_pixels");
  if (ensure_device()) return -1;
  if (corr_check(s, "ia3_corr_hot_pixels")) return -1;
  cudaStream_t st = s->stream;
  Scoped sc;
  const long long nxy = (long long)s->X * s->Y;
  const int cap = 1 << 16;
  int* d_cnt = nullptr; int* d_list = nullptr; int* d_n = nullptr; float* d_vals = nullptr; void* h = nullptr;
  if (sc.dalloc(&d_cnt, (size_t)nxy * 4) || sc.dalloc(&d_list, (size_t)cap * 4) || sc.dalloc(&d_n, 256) || sc.halloc(&h, 256 + (size_t)cap * 4)) return -1;
  if (launch_hot_count((const uint16_t*)s->d_im, s->Z, s->X, s->Y, (float)hot_th, d_cnt, st)) return -1;
  if (launch_hot_select(d_cnt, nxy, hot_pix_th * (double)s->Z, d_list, d_n, cap, st)) return -1;
  if (small_copy(h, d_n, sizeof(int), st)) return -1;
  IA3_CUDA(cudaStreamSynchronize(st));
  const int n = *static_cast<int*>(h);
  if (n_hot) *n_hot = n;
  if (n == 0) return 0;
  if (n > cap) { set_error("ia3_corr_hot_pixels: more than 65536 hot columns -- check hot_th"); return -1; }
  int* hl = reinterpret_cast<int*>(static_cast<char*>(h) + 256);
  if (small_copy(hl, d_list, (size_t)n * 4, st)) return -1;
  IA3_CUDA(cudaStreamSynchronize(st));
  std::sort(hl, hl + n);                                     // np.where order: the fix is sequential and in place
  if (small_copy(d_list, hl, (size_t)n * 4, st)) return -1;
  if (sc.dalloc(&d_vals, (size_t)n * s->Z * 4)) return -1;
  if (launch_hot_fix((uint16_t*)s->d_im, s->Z, s->X, s->Y, d_list, n, d_cnt /* the counts are no longer needed */, d_vals, st)) return -1;
  IA3_CUDA(cudaStreamSynchronize(st));
  return 0;
}

// median of N values given as a 65536-bin histogram, as np.median returns it for a float32 array of integers: the middle
// element, or the float32 mean of the two middle elements
static float median_from_counts(const unsigned long long* c, unsigned long long N) {
  const unsigned long long k_lo = (N - 1) / 2, k_hi = N / 2;
  unsigned long long run = 0;
  int v_lo = -1, v_hi = -1;
  for (int v = 0; v < 65536 && v_hi < 0; ++v) {
    run += c[v];
    if (v_lo < 0 && run > k_lo) v_lo = v;
    if (run > k_hi) v_hi = v;
  }
  return ((float)v_lo + (float)v_hi) / 2.0f;
}

int ia3_corr_zshift(ia3_stack* s) {
  IA3_STAT("ia3_corr_zshift");
  if (ensure_device()) return -1;
  if (corr_check(s, "ia3_corr_zshift")) return -1;
  cudaStream_t st = s->stream;
  Scoped sc;
  const size_t hb = 65536 * sizeof(unsigned long long), nxy = (size_t)s->X * s->Y;
  unsigned long long* d_h = nullptr; float* d_med = nullptr; void* h = nullptr;
  if (sc.dalloc(&d_h, hb * s->Z) || sc.dalloc(&d_med, sizeof(float) * s->Z + 256) || sc.halloc(&h, hb * s->Z + sizeof(float) * s->Z + 256)) return -1;
  for (int z = 0; z < s->Z; ++z)
    if (launch_hist_u16((const uint16_t*)s->d_im + (size_t)z * nxy, (long long)nxy, d_h + (size_t)z * 65536, st)) return -1;
  if (small_copy(h, d_h, hb * s->Z, st)) return -1;
  IA3_CUDA(cudaStreamSynchronize(st));
  const unsigned long long* hc = static_cast<const unsigned long long*>(h);
  float* med = reinterpret_cast<float*>(static_cast<char*>(h) + hb * s->Z);
  std::vector<unsigned long long> all(65536, 0);
  for (int z = 0; z < s->Z; ++z) {
    med[z] = median_from_counts(hc + (size_t)z * 65536, nxy);
    for (int v = 0; v < 65536; ++v) all[v] += hc[(size_t)z * 65536 + v];
  }
  const float med_all = median_from_counts(all.data(), (unsigned long long)s->nvox);
  if (small_copy(d_med, med, sizeof(float) * s->Z, st)) return -1;
  if (launch_zshift((uint16_t*)s->d_im, (long long)nxy, (long long)s->nvox, d_med, med_all, st)) return -1;
  IA3_CUDA(cudaStreamSynchronize(st));
  return 0;
}

int ia3_corr_highpass(ia3_stack* s, const double* w_half, int r) {
  IA3_STAT("ia3_corr_highpass");
  if (ensure_device()) return -1;
  if (corr_check(s, "ia3_corr_highpass")) return -1;
  if (!w_half || r < 0 || r > 255) { set_error("bad Gaussian kernel"); return -1; }
  cudaStream_t st = s->stream;
  Scoped sc;
  uint16_t* a = nullptr; uint16_t* b = nullptr; double* d_w = nullptr; void* h = nullptr;
  if (sc.dalloc(&a, s->nvox * 2) || sc.dalloc(&b, s->nvox * 2) || sc.dalloc(&d_w, 4096) || sc.halloc(&h, 4096)) return -1;
  memcpy(h, w_half, sizeof(double) * (size_t)(r + 1));
  if (small_copy(d_w, h, sizeof(double) * (size_t)(r + 1), st)) return -1;
  if (launch_highpass((uint16_t*)s->d_im, a, b, s->Z, s->X, s->Y, d_w, r, st)) return -1;
  IA3_CUDA(cudaStreamSynchronize(st));
  return 0;
}

int ia3_corr_mix(ia3_stack* const* ins, int n_in, const void* bleed, const void* illum, int profile_f64, ia3_stack* out) {
  IA3_STAT("ia3_corr_mix");
  if (ensure_device()) return -1;
  if (!ins || n_in < 1 || n_in > 16 || !out) { set_error("bad argument"); return -1; }
  if (corr_check(out, "ia3_corr_mix")) return -1;
  for (int j = 0; j < n_in; ++j) {
    if (corr_check(ins[j], "ia3_corr_mix")) return -1;
    if (ins[j]->Z != out->Z || ins[j]->X != out->X || ins[j]->Y != out->Y) { set_error("ia3_corr_mix: stacks differ in shape"); return -1; }
    if (bleed && ins[j] == out) { set_error("ia3_corr_mix: bleed-through mixing cannot run in place"); return -1; }
  }
  if (!bleed && n_in != 1) { set_error("ia3_corr_mix: several inputs need a bleed-through profile"); return -1; }
  cudaStream_t st = out->stream;
  Scoped sc;
  const size_t es = profile_f64 ? 8 : 4, nxy = (size_t)out->X * out->Y;
  const size_t b_bleed = bleed ? (size_t)n_in * nxy * es : 0, b_illum = illum ? nxy * es : 0;
  char* d_bleed = nullptr; char* d_illum = nullptr; const uint16_t** d_ptrs = nullptr; void* h = nullptr;
  if (sc.dalloc(&d_ptrs, 256) || sc.halloc(&h, 256)) return -1;
  // the profiles are per-dataset constants: a caller that corrects many fields of view uploads them once
  // (ia3_device_upload) and passes the device pointers; host arrays are copied here, every call
  if (bleed) {
    if (is_device_ptr(bleed)) d_bleed = const_cast<char*>(static_cast<const char*>(bleed));
    else { if (sc.dalloc(&d_bleed, b_bleed)) return -1; IA3_CUDA(cudaMemcpyAsync(d_bleed, bleed, b_bleed, cudaMemcpyHostToDevice, st)); }
  }
  if (illum) {
    if (is_device_ptr(illum)) d_illum = const_cast<char*>(static_cast<const char*>(illum));
    else { if (sc.dalloc(&d_illum, b_illum)) return -1; IA3_CUDA(cudaMemcpyAsync(d_illum, illum, b_illum, cudaMemcpyHostToDevice, st)); }
  }
  const uint16_t** hp = static_cast<const uint16_t**>(h);
  for (int j = 0; j < n_in; ++j) hp[j] = (const uint16_t*)ins[j]->d_im;
  if (small_copy(d_ptrs, hp, (size_t)n_in * sizeof(void*), st)) return -1;
  int rc;
  if (profile_f64) rc = launch_mix<double>(d_ptrs, n_in, (const double*)d_bleed, (const double*)d_illum, (uint16_t*)out->d_im, (long long)nxy, (long long)out->nvox, st);
  else rc = launch_mix<float>(d_ptrs, n_in, (const float*)d_bleed, (const float*)d_illum, (uint16_t*)out->d_im, (long long)nxy, (long long)out->nvox, st);
  if (rc) return -1;
  IA3_CUDA(cudaStreamSynchronize(st));
  return 0;
}

int ia3_corr_warp(ia3_stack* in, const double* drift, const void* chroma, int chroma_f64, int chroma_z, ia3_stack* out) {
  IA3_STAT("ia3_corr_warp");
  if (ensure_device()) return -1;
  if (corr_check(in, "ia3_corr_warp") || corr_check(out, "ia3_corr_warp")) return -1;
  if (in == out) { set_error("ia3_corr_warp cannot run in place"); return -1; }
  if (in->Z != out->Z || in->X != out->X || in->Y != out->Y) { set_error("ia3_corr_warp: stacks differ in shape"); return -1; }
  if (chroma && chroma_z != 1 && chroma_z != in->Z) { set_error("ia3_corr_warp: the chromatic profile has 1 or Z planes per axis"); return -1; }
  cudaStream_t st = out->stream;
  Scoped sc;
  const long long np_ = warp_padded_voxels(in->Z, in->X, in->Y);
  double* buf = nullptr; char* d_ch = nullptr;
  const size_t b_ch = chroma ? (size_t)3 * chroma_z * in->X * in->Y * (chroma_f64 ? 8 : 4) : 0;
  if (sc.dalloc(&buf, (size_t)np_ * 8)) return -1;
  if (chroma) {
    // a profile that already lives in device memory (ia3_device_upload) is used where it is
    if (is_device_ptr(chroma)) d_ch = const_cast<char*>(static_cast<const char*>(chroma));
    else {
      if (sc.dalloc(&d_ch, b_ch)) return -1;
      IA3_CUDA(cudaMemcpyAsync(d_ch, chroma, b_ch, cudaMemcpyHostToDevice, st));
    }
  }
  const double d0 = drift ? drift[0] : 0.0, d1 = drift ? drift[1] : 0.0, d2 = drift ? drift[2] : 0.0;
  if (launch_warp((const uint16_t*)in->d_im, in->Z, in->X, in->Y, buf, d_ch, chroma_f64, chroma_z, d0, d1, d2, (uint16_t*)out->d_im, st)) return -1;
  IA3_CUDA(cudaStreamSynchronize(st));
  return 0;
}

}  // extern "C"
