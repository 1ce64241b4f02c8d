// Common macros shared by the sm_100a kernels and the host-side simulation harness
// (tests/hostsim) that compiles the same numerical core with g++ for CPU-only debugging.
#pragma once
#include <stdint.h>
#include <math.h>
#include <float.h>

#if defined(__CUDACC__)
#define IA3_HD __host__ __device__ __forceinline__
#define IA3_HDN __host__ __device__ __noinline__ inline  /* big scalar routines: one copy each, the fit kernel must fit the I-cache */
#define IA3_D __device__ __forceinline__
#else
#define IA3_HD inline
#define IA3_HDN inline
#define IA3_D inline
#endif

namespace ia3 {

constexpr int NP = 10;          // raw parameters of the 3D Gaussian model: bk,h,xp,yp,zp,w1,w2,w3,pp,tp
constexpr int NOUT = 11;        // natural parameters + eps
constexpr int NTRI = NP * (NP + 1) / 2;  // 55 entries of the symmetric J^T J

// index of (i,j), i<=j, in the packed upper triangle (row-major)
IA3_HD int tri(int i, int j) { return i * NP - (i * (i - 1)) / 2 + (j - i); }

}  // namespace ia3
