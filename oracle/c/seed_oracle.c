/* ORACLE (test infrastructure, CPU): C restatement of the scipy.ndimage routines the
 * reference's seed stage runs on (spot_tools/fitting.py:92,95,99,102):
 *
 *   gaussian_filter(u16|f32 volume, sigma)  == three passes of NI_Correlate1D (scipy 1.18.1,
 *       src/ni_filters.c; third-party, not vendored in the reference) with the symmetric-kernel
 *       branch:  acc = x[0]*w[0]; for j = -r..-1: acc += (x[j] + x[-j]) * w[j]; double
 *       accumulation, line copied to a double buffer and extended with NI_EXTEND_REFLECT, result
 *       cast back to the array dtype ((npy_uint16)acc truncates, (float)acc rounds) after EACH
 *       axis, axes in order 0, 1, 2.
 *   maximum_filter / minimum_filter(size)   == separable NI_MinOrMaxFilter1D, origin 0, reflect.
 *
 * Compile with -ffp-contract=off (scipy's x86-64 wheels contain no fused multiply-add).
 * Checked bit-for-bit against scipy itself in tests/test_oracle_pinned.py.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static int reflect_index(long i, long n) {
  if (n <= 1) return 0;
  long p = 2 * n;
  long m = i % p;
  if (m < 0) m += p;
  return (int)(m < n ? m : p - 1 - m);
}

/* one axis pass over a volume viewed as [outer][L][inner] */
#define DEFINE_PASS(NAME, T, STORE)                                                              \
  static void NAME(const T* in, T* out, long outer, long L, long inner, const double* w, int r,  \
                   int threads) {                                                                \
    const double* fw = w + r; /* fw[0] = centre tap */                                           \
    long nlines = outer * inner;                                                                 \
    (void)threads;                                                                               \
    _Pragma("omp parallel num_threads(threads)")                                                 \
    {                                                                                            \
      double* buf = (double*)malloc(sizeof(double) * (size_t)(L + 2 * r));                       \
      _Pragma("omp for schedule(static)")                                                        \
      for (long l = 0; l < nlines; ++l) {                                                        \
        long o = l / inner, i = l % inner;                                                       \
        const T* src = in + o * L * inner + i;                                                   \
        T* dst = out + o * L * inner + i;                                                        \
        for (long a = -r; a < L + r; ++a) buf[a + r] = (double)src[(long)reflect_index(a, L) * inner]; \
        for (long ll = 0; ll < L; ++ll) {                                                        \
          const double* x = buf + r + ll;                                                        \
          double acc = x[0] * fw[0];                                                             \
          for (int j = -r; j < 0; ++j) acc += (x[j] + x[-j]) * fw[j];                            \
          dst[ll * inner] = STORE(acc);                                                          \
        }                                                                                        \
      }                                                                                          \
      free(buf);                                                                                 \
    }                                                                                            \
  }

#define STORE_U16(a) ((uint16_t)(a))
#define STORE_F32(a) ((float)(a))
DEFINE_PASS(pass_u16, uint16_t, STORE_U16)
DEFINE_PASS(pass_f32, float, STORE_F32)

static int nthreads(int t) {
#ifdef _OPENMP
  return t > 0 ? t : omp_get_max_threads();
#else
  (void)t;
  return 1;
#endif
}

void gauss3d_u16(const uint16_t* in, uint16_t* out, int Z, int X, int Y, const double* w, int r, int threads) {
  size_t n = (size_t)Z * X * Y;
  uint16_t* tmp = (uint16_t*)malloc(n * sizeof(uint16_t));
  int t = nthreads(threads);
  pass_u16(in, out, 1, Z, (long)X * Y, w, r, t);
  pass_u16(out, tmp, Z, X, Y, w, r, t);
  pass_u16(tmp, out, (long)Z * X, Y, 1, w, r, t);
  free(tmp);
}

void gauss3d_f32(const float* in, float* out, int Z, int X, int Y, const double* w, int r, int threads) {
  size_t n = (size_t)Z * X * Y;
  float* tmp = (float*)malloc(n * sizeof(float));
  int t = nthreads(threads);
  pass_f32(in, out, 1, Z, (long)X * Y, w, r, t);
  pass_f32(out, tmp, Z, X, Y, w, r, t);
  pass_f32(tmp, out, (long)Z * X, Y, 1, w, r, t);
  free(tmp);
}

/* 1-D running min/max of window [i - size/2, i + size - size/2 - 1] with reflect boundary */
#define DEFINE_RANK(NAME, T)                                                                     \
  static void NAME(const T* in, T* out, long outer, long L, long inner, int size, int is_max,    \
                   int threads) {                                                                \
    int s1 = size / 2, s2 = size - s1 - 1;                                                       \
    long nlines = outer * inner;                                                                 \
    (void)threads;                                                                               \
    _Pragma("omp parallel for schedule(static) num_threads(threads)")                            \
    for (long l = 0; l < nlines; ++l) {                                                          \
      long o = l / inner, i = l % inner;                                                         \
      const T* src = in + o * L * inner + i;                                                     \
      T* dst = out + o * L * inner + i;                                                          \
      for (long ll = 0; ll < L; ++ll) {                                                          \
        T best = src[(long)reflect_index(ll - s1, L) * inner];                                   \
        for (long a = ll - s1 + 1; a <= ll + s2; ++a) {                                          \
          T v = src[(long)reflect_index(a, L) * inner];                                          \
          if (is_max ? (v > best) : (v < best)) best = v;                                        \
        }                                                                                        \
        dst[ll * inner] = best;                                                                  \
      }                                                                                          \
    }                                                                                            \
  }
DEFINE_RANK(rank_u16, uint16_t)
DEFINE_RANK(rank_f32, float)

void rank3d_u16(const uint16_t* in, uint16_t* out, int Z, int X, int Y, int size, int is_max, int threads) {
  size_t n = (size_t)Z * X * Y;
  uint16_t* tmp = (uint16_t*)malloc(n * sizeof(uint16_t));
  int t = nthreads(threads);
  rank_u16(in, out, 1, Z, (long)X * Y, size, is_max, t);
  rank_u16(out, tmp, Z, X, Y, size, is_max, t);
  rank_u16(tmp, out, (long)Z * X, Y, 1, size, is_max, t);
  free(tmp);
}

void rank3d_f32(const float* in, float* out, int Z, int X, int Y, int size, int is_max, int threads) {
  size_t n = (size_t)Z * X * Y;
  float* tmp = (float*)malloc(n * sizeof(float));
  int t = nthreads(threads);
  rank_f32(in, out, 1, Z, (long)X * Y, size, is_max, t);
  rank_f32(out, tmp, Z, X, Y, size, is_max, t);
  rank_f32(tmp, out, (long)Z * X, Y, 1, size, is_max, t);
  free(tmp);
}
