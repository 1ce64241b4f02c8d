// Host-callable launchers of the seed stage kernels (seed_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ia3 {

struct GaussW {
  static constexpr int MAXR = 95;
  double w[MAXR + 1];   // w[j]: weight at distance j from the centre tap
  int r;
};

struct SeedDims {
  int Z, X, Y;
  int cpr;              // 8-voxel chunks per row
  long long n_chunks;   // Z*X*cpr
  int n_blocks;
  int fs, s1, s2;       // rank filter size; window [i-s1, i+s2]
  int edge_on, lo, loZ, hiZ, hiX, hiY;
  double h_min;
};

template <typename Tin>
int gaussian_filter_exact(const Tin* in, Tin* bufA, Tin* bufB, int Z, int X, int Y, const GaussW& gw, cudaStream_t st, bool skip_z = false);

template <typename Tin>
int seed_flags(const Tin* fg, const Tin* bg, const SeedDims& d, int variant, uint8_t* bits, int* counts,
               long long* offsets, cudaStream_t st);
template <typename Tin>
int seed_emit(const Tin* fg, const Tin* bg, const SeedDims& d, int variant, const uint8_t* bits,
              const long long* offsets, int32_t* out_zxy, float* out_h, cudaStream_t st);
int seed_flag_threads();

}  // namespace ia3

namespace ia3 {
// mode-of-histogram background of n boxes (z0 z1 x0 x1 y0 y1, half open) of a uint16 stack
int box_background(const uint16_t* im, int X, int Y, const int* d_boxes, long long n, int first, int bin, int nbins, int max_iter,
                   double* d_out, cudaStream_t st);
int volume_background(const uint16_t* im, int Z, int X, int Y, const int* d_box, int first, int bin, int nbins, int max_iter,
                      unsigned* d_ghist, double* d_out, cudaStream_t st);
}  // namespace ia3
