// Device-side plumbing shared by the .cu translation units: error handling, launch counting,
// a small caching allocator for the large per-stack work volumes.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include <string>
#include "ia3_common.h"

namespace ia3 {

void set_error(const std::string& msg);
extern std::atomic<int64_t> g_launches;

#define IA3_CUDA(expr)                                                                          \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      char _b[512];                                                                             \
      snprintf(_b, sizeof(_b), "%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,             \
               cudaGetErrorString(_e));                                                         \
      ia3::set_error(_b);                                                                       \
      return -1;                                                                                \
    }                                                                                           \
  } while (0)

#define IA3_LAUNCH_CHECK()                                                                      \
  do {                                                                                          \
    ++ia3::g_launches;                                                                          \
    IA3_CUDA(cudaGetLastError());                                                               \
  } while (0)

// cached device allocations (cudaMalloc of GB-sized buffers costs milliseconds)
int dev_alloc(void** p, size_t bytes);
void dev_free(void* p);
void dev_cache_clear();

}  // namespace ia3
