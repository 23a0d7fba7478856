// Shared device helpers for the jaxpm_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/jaxpm_b200.h"

namespace jpm {

// ---- host-side error plumbing ------------------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define JPM_CHECK_ARG(cond, msg)                  \
  do {                                            \
    if (!(cond)) {                                \
      jpm::set_error("invalid argument: %s", msg); \
      return JPM_ERR_INVALID;                     \
    }                                             \
  } while (0)

#define JPM_CUDA(call)                                                                   \
  do {                                                                                   \
    cudaError_t e__ = (call);                                                            \
    if (e__ != cudaSuccess) {                                                            \
      jpm::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, \
                     __LINE__);                                                          \
      return JPM_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

#define JPM_LAUNCH_CHECK()                    \
  do {                                        \
    jpm::count_launch();                      \
    JPM_CUDA(cudaPeekAtLastError());          \
  } while (0)

constexpr int kNumSMs = 148;  // B200

// ---- CIC index / weight rules --------------------------------------------------
// One dimension of the 8-corner stencil: two cell indices and two weights.
struct Cic1 {
  int i0, i1;    // wrapped cell indices; -1 == dropped by the reference (relative rule only)
  float w0, w1;  // 1-|x-corner|
  float s0, s1;  // d w / d x = -sign(x-corner), sign(0) = 0
};

__device__ __forceinline__ int pymod(int a, int n) {
  if (a >= 0 && a < n) return a;
  int r = a % n;
  return r < 0 ? r + n : r;
}

__device__ __forceinline__ float sgn(float t) { return (t > 0.f) ? 1.f : ((t < 0.f) ? -1.f : 0.f); }

// Absolute rule, jaxpm/painting.py:22-37: floor, +{0,1}, 1-|x-c|, int32 cast, python mod.
template <bool GRAD>
__device__ __forceinline__ Cic1 cic_abs(float p, int n) {
  Cic1 c;
  const float f = floorf(p);
  const float d0 = p - f;
  const float d1 = p - (f + 1.0f);
  c.w0 = 1.0f - fabsf(d0);
  c.w1 = 1.0f - fabsf(d1);
  if (GRAD) {
    c.s0 = -sgn(d0);
    c.s1 = -sgn(d1);
  }
  const int i = (int)f;
  c.i0 = pymod(i, n);
  c.i1 = pymod((int)(f + 1.0f), n);
  return c;
}

// Relative rule, jaxpm/painting_utils.py:48-65 with cell_size = 1, offset = 0:
//   pp = base + d ; x = pp + o ; r = x mod L (python-sign float mod) ; idx = floor(r) ;
//   nd = pp - idx ; nd -= rint(nd / L) * L ; w = 1-|nd|.   idx == n is dropped (mode='drop').
__device__ __forceinline__ void cic_rel_corner(float pp, float o, float L, int n, int& idx, float& w,
                                               float& s) {
  const float x = pp + o;
  float r = x;
  if (!(x >= 0.0f && x < L)) {
    r = fmodf(x, L);
    if (r != 0.0f && r < 0.0f) r = r + L;
  }
  const float fi = floorf(r);
  float nd = pp - fi;
  if (!(fabsf(nd) < 0.25f * L)) nd = __fsub_rn(nd, __fmul_rn(rintf(__fdiv_rn(nd, L)), L));
  w = 1.0f - fabsf(nd);
  s = -sgn(nd);
  const int i = (int)fi;
  idx = (i >= 0 && i < n) ? i : -1;
}

template <bool GRAD>
__device__ __forceinline__ Cic1 cic_rel(int base, float d, int n) {
  Cic1 c;
  const float pp = (float)base + d;
  const float L = (float)n;
  cic_rel_corner(pp, 0.0f, L, n, c.i0, c.w0, c.s0);
  cic_rel_corner(pp, 1.0f, L, n, c.i1, c.w1, c.s1);
  return c;
}

template <bool REL, bool GRAD>
__device__ __forceinline__ Cic1 cic_1d(int base, float v, int n) {
  if (REL) return cic_rel<GRAD>(base, v, n);
  return cic_abs<GRAD>(v, n);
}

// streaming loads/stores of the [np][3] particle stream (touched once per kernel)
__device__ __forceinline__ float ld_stream(const float* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(float* p, float v) { __stcs(p, v); }

inline int div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace jpm
