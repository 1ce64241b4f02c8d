// Host-callable launchers of the pre-processing kernels (corr_kernels.cu): correct_fov_image's compute core.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ia3 {
int launch_highpass(uint16_t* im, uint16_t* bufA, uint16_t* bufB, int Z, int X, int Y, const double* d_w, int r, cudaStream_t st);
int launch_zshift(uint16_t* im, long long XY, long long n, const float* d_med_z, float med_all, cudaStream_t st);
int launch_hot_count(const uint16_t* im, int Z, int X, int Y, float hot_th, int* cnt, cudaStream_t st);
int launch_hot_select(const int* cnt, long long n, double thr, int* out, int* count, int cap, cudaStream_t st);
int launch_hot_fix(uint16_t* im, int Z, int X, int Y, const int* list, int n, int* slot /* X*Y ints, scratch */, float* vals, cudaStream_t st);
// Tp = float or double: the dtype the profile arrays were saved in decides numpy's arithmetic type
template <typename Tp>
int launch_mix(const uint16_t* const* d_ins, int n_in, const Tp* bleed, const Tp* illum, uint16_t* out, long long XY, long long n, cudaStream_t st);
long long warp_padded_voxels(int Z, int X, int Y);
int launch_warp(const uint16_t* im, int Z, int X, int Y, double* buf, const void* chroma, int chroma_f64, int CZ,
                double d0, double d1, double d2, uint16_t* out, cudaStream_t st);
}  // namespace ia3
