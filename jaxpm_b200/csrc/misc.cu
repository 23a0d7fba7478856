// Error plumbing, device info, small elementwise helpers and halo pack/unpack kernels.
//   reference for pack/unpack: jaxpm/distributed.py:68-113 (slice_pad / slice_unpad_impl)
#include <cstring>

#include "common.cuh"

namespace jpm {

static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

__global__ void __launch_bounds__(256)
axpby_kernel(float* out, float a, const float* x, float b, const float* y, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = y ? (a * x[i] + b * y[i]) : a * x[i];
}

__global__ void __launch_bounds__(256)
grid_plus_disp_kernel(float* __restrict__ out, const float* __restrict__ disp, int nx, int ny, int nz,
                      int ox, int oy) {
  const long long n = (long long)nx * ny * nz;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
    const int k = (int)(p % nz);
    const long long t = p / nz;
    const int j = (int)(t % ny), i = (int)(t / ny);
    out[3 * p + 0] = (float)(i + ox) + disp[3 * p + 0];
    out[3 * p + 1] = (float)(j + oy) + disp[3 * p + 1];
    out[3 * p + 2] = (float)k + disp[3 * p + 2];
  }
}

// box <-> packed copies; rows along z are contiguous, so threads run along z.
template <int DIR>  // 0: pack (mesh->buf), 1: unpack copy, 2: unpack add
__global__ void __launch_bounds__(256)
box_kernel(float* __restrict__ mesh, float* __restrict__ buf, int ny, int nz, int x0, int y0, int bx,
           int by) {
  const long long rows = (long long)bx * by;
  for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
    const int i = (int)(r / by) + x0, j = (int)(r % by) + y0;
    float* m = mesh + ((long long)i * ny + j) * nz;
    float* b = buf + r * nz;
    for (int k = threadIdx.x; k < nz; k += blockDim.x) {
      if (DIR == 0) b[k] = m[k];
      if (DIR == 1) m[k] = b[k];
      if (DIR == 2) m[k] += b[k];
    }
  }
}

// ---- counter-based Gaussian field generator ------------------------------------------------------------------
// N(0,1) white noise for linear_field (jaxpm/pm.py:129-144 draws it with normal_field, distributed.py:193-223).
// Philox4x32-10 (Salmon et al. 2011) keyed by the 64-bit seed, counter = (index of the 4-cell group, stream id):
// cell c of the GLOBAL mesh always gets the same number, whatever the launch shape or the decomposition - a sharded
// run therefore draws exactly the single-device field (the reference does not: one key per device,
// distributed.py:204-215).  NOT JAX's threefry stream: a seed does not reproduce a JAX run (parity runs share the
// IC array).  Box-Muller on (x + 0.5) 2^-32: no 0 / 1 arguments, |z| <= 6.66.
__device__ __forceinline__ void philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0,
                                              unsigned k1, unsigned out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const unsigned n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__global__ void __launch_bounds__(256)
normal_field_kernel(float* __restrict__ out, int lx, int ly, int nz, int ox, int oy, int gny, unsigned seed_lo,
                    unsigned seed_hi, unsigned stream_id) {
  // local block [lx][ly][nz] of the global [*][gny][nz] mesh starting at (ox, oy, 0); 4 consecutive z cells per thread
  const long long n4 = (long long)lx * ly * (nz / 4);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n4; t += stride) {
    const int k4 = (int)(t % (nz / 4));
    const long long r = t / (nz / 4);
    const int j = (int)(r % ly), i = (int)(r / ly);
    const unsigned long long g4 = ((unsigned long long)(i + ox) * gny + (j + oy)) * (nz / 4) + k4;   // global group index
    unsigned x[4];
    philox4x32_10((unsigned)g4, (unsigned)(g4 >> 32), stream_id, 0u, seed_lo, seed_hi, x);
    float4 z;
    {
      const float u1 = ((float)(x[0] >> 8) + 0.5f) * (1.0f / 16777216.0f), u2 = ((float)(x[1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
      const float rad = sqrtf(-2.0f * logf(u1));
      float sn, cs;
      sincospif(2.0f * u2, &sn, &cs);
      z.x = rad * cs; z.y = rad * sn;
    }
    {
      const float u1 = ((float)(x[2] >> 8) + 0.5f) * (1.0f / 16777216.0f), u2 = ((float)(x[3] >> 8) + 0.5f) * (1.0f / 16777216.0f);
      const float rad = sqrtf(-2.0f * logf(u1));
      float sn, cs;
      sincospif(2.0f * u2, &sn, &cs);
      z.z = rad * cs; z.w = rad * sn;
    }
    reinterpret_cast<float4*>(out)[t] = z;
  }
}

// same numbers for row lengths that are not a multiple of 4: one cell per thread (group = global flat index / 4)
__global__ void __launch_bounds__(256)
normal_field_cell_kernel(float* __restrict__ out, int lx, int ly, int nz, int ox, int oy, int gny, unsigned seed_lo,
                         unsigned seed_hi, unsigned stream_id) {
  const long long n = (long long)lx * ly * nz;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) {
    const int k = (int)(t % nz);
    const long long r = t / nz;
    const int j = (int)(r % ly), i = (int)(r / ly);
    const unsigned long long flat = ((unsigned long long)(i + ox) * gny + (j + oy)) * nz + k;
    const unsigned long long g4 = flat >> 2;
    const int e = (int)(flat & 3);
    unsigned x[4];
    philox4x32_10((unsigned)g4, (unsigned)(g4 >> 32), stream_id, 0u, seed_lo, seed_hi, x);
    const unsigned a = x[e & 2], b = x[(e & 2) + 1];
    const float u1 = ((float)(a >> 8) + 0.5f) * (1.0f / 16777216.0f), u2 = ((float)(b >> 8) + 0.5f) * (1.0f / 16777216.0f);
    const float rad = sqrtf(-2.0f * logf(u1));
    float sn, cs;
    sincospif(2.0f * u2, &sn, &cs);
    out[t] = rad * ((e & 1) ? sn : cs);
  }
}

static int ew_grid(long long n) {
  long long blocks = (n + 255) / 256;
  const long long cap = (long long)kNumSMs * 16;
  return (int)(blocks > cap ? cap : blocks);
}

}  // namespace jpm

using namespace jpm;

extern "C" int32_t jpm_abi_version(void) { return JPM_ABI_VERSION; }

extern "C" int32_t jpm_normal_field_f32(void* stream, float* out, int32_t lx, int32_t ly, int32_t nz, int32_t ox,
                                        int32_t oy, int32_t global_ny, uint64_t seed, uint32_t stream_id) {
  JPM_CHECK_ARG(out && lx > 0 && ly > 0 && nz > 0, "bad arguments");
  JPM_CHECK_ARG(ox >= 0 && oy >= 0 && oy + ly <= global_ny, "block outside the global mesh");
  const unsigned slo = (unsigned)(seed & 0xffffffffull), shi = (unsigned)(seed >> 32);
  if (nz % 4 == 0) {
    const long long n4 = (long long)lx * ly * (nz / 4);
    normal_field_kernel<<<ew_grid(n4), 256, 0, (cudaStream_t)stream>>>(out, lx, ly, nz, ox, oy, global_ny, slo, shi,
                                                                      stream_id);
  } else {
    normal_field_cell_kernel<<<ew_grid((long long)lx * ly * nz), 256, 0, (cudaStream_t)stream>>>(
        out, lx, ly, nz, ox, oy, global_ny, slo, shi, stream_id);
  }
  JPM_LAUNCH_CHECK();
  return JPM_OK;
}
extern "C" const char* jpm_last_error_string(void) { return g_err; }
extern "C" int64_t jpm_kernel_launch_count(void) { return g_launches.load(); }

extern "C" int32_t jpm_device_info(char* name, int32_t name_len, int32_t* sm_count, int32_t* cc_major,
                                   int32_t* cc_minor) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    set_error("no CUDA device: %s", cudaGetErrorString(e));
    return JPM_ERR_NOGPU;
  }
  cudaDeviceProp prop;
  JPM_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (name && name_len > 0) {
    strncpy(name, prop.name, name_len - 1);
    name[name_len - 1] = 0;
  }
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  if (prop.major != 10) {
    set_error("device %s is sm_%d%d; this library is built for sm_100a only", prop.name, prop.major,
              prop.minor);
    return JPM_ERR_NOGPU;
  }
  return JPM_OK;
}

extern "C" int32_t jpm_axpby_f32(void* stream, float* out, float a, const float* x, float b,
                                 const float* y, int64_t n) {
  JPM_CHECK_ARG(out && x && n >= 0, "null pointer");
  if (n == 0) return JPM_OK;
  axpby_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(out, a, x, b, y, n);
  JPM_LAUNCH_CHECK();
  return JPM_OK;
}

// 4th-order central-difference divergence of three periodic meshes, out = sum_d D_d g_d with
// D f(x) = [8 (f(x+1) - f(x-1)) - (f(x+2) - f(x-2))] / 12 - the operator the reference's gradient kernel
// i (8 sin w - sin 2w) / 6 (kernels.py:62-66) is the symbol of.  The adjoint of pm_forces needs
// sum_d L_d^T G_d = -Phi (sum_d D_d G_d) (D_d antisymmetric, Phi = IFFT 1/k^2 FFT symmetric): forming the divergence
// here leaves ONE transform pair (the potential chain) instead of three forward transforms + a k-space pass.
__global__ void __launch_bounds__(256)
fd_div3_kernel(float* __restrict__ out, const float* __restrict__ g0, const float* __restrict__ g1,
               const float* __restrict__ g2, int nx, int ny, int nz) {
  const long long n = (long long)nx * ny * nz;
  const long long stride = (long long)gridDim.x * blockDim.x;
  constexpr float c8 = 2.0f / 3.0f, c1 = 1.0f / 12.0f;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) {
    const int k = (int)(t % nz);
    const long long r = t / nz;
    const int j = (int)(r % ny), i = (int)(r / ny);
    auto wrap = [](int a, int m) { return a < 0 ? a + m : (a >= m ? a - m : a); };
    const long long sx = (long long)ny * nz;
    const long long row = ((long long)i * ny + j) * nz, col = (long long)j * nz + k, pl = (long long)i * sx;
    const float dx = c8 * (g0[wrap(i + 1, nx) * sx + col] - g0[wrap(i - 1, nx) * sx + col]) -
                     c1 * (g0[wrap(i + 2, nx) * sx + col] - g0[wrap(i - 2, nx) * sx + col]);
    const float dy = c8 * (g1[pl + (long long)wrap(j + 1, ny) * nz + k] - g1[pl + (long long)wrap(j - 1, ny) * nz + k]) -
                     c1 * (g1[pl + (long long)wrap(j + 2, ny) * nz + k] - g1[pl + (long long)wrap(j - 2, ny) * nz + k]);
    const float dz = c8 * (g2[row + wrap(k + 1, nz)] - g2[row + wrap(k - 1, nz)]) -
                     c1 * (g2[row + wrap(k + 2, nz)] - g2[row + wrap(k - 2, nz)]);
    out[t] = (dx + dy) + dz;
  }
}

extern "C" int32_t jpm_fd_divergence3_f32(void* stream, float* out, const float* g3, int32_t nx, int32_t ny,
                                          int32_t nz) {
  JPM_CHECK_ARG(out && g3 && nx >= 4 && ny >= 4 && nz >= 4, "bad arguments (every axis needs >= 4 cells)");
  const long long n = (long long)nx * ny * nz;
  fd_div3_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(out, g3, g3 + n, g3 + 2 * n, nx, ny, nz);
  JPM_LAUNCH_CHECK();
  return JPM_OK;
}

extern "C" int32_t jpm_grid_plus_disp_f32(void* stream, float* out, const float* disp, int32_t nx,
                                          int32_t ny, int32_t nz, int32_t ox, int32_t oy) {
  JPM_CHECK_ARG(out && disp && nx > 0 && ny > 0 && nz > 0, "bad arguments");
  grid_plus_disp_kernel<<<ew_grid((long long)nx * ny * nz), 256, 0, (cudaStream_t)stream>>>(
      out, disp, nx, ny, nz, ox, oy);
  JPM_LAUNCH_CHECK();
  return JPM_OK;
}

extern "C" int32_t jpm_pack_box_f32(void* stream, float* packed, const float* mesh, int32_t ny,
                                    int32_t nz, int32_t x0, int32_t x1, int32_t y0, int32_t y1) {
  JPM_CHECK_ARG(packed && mesh && x1 >= x0 && y1 >= y0 && x0 >= 0 && y0 >= 0 && y1 <= ny, "bad box");
  const long long rows = (long long)(x1 - x0) * (y1 - y0);
  if (rows == 0) return JPM_OK;
  const int grid = (int)(rows < kNumSMs * 16 ? rows : kNumSMs * 16);
  box_kernel<0><<<grid, 256, 0, (cudaStream_t)stream>>>(const_cast<float*>(mesh), packed, ny, nz, x0,
                                                        y0, x1 - x0, y1 - y0);
  JPM_LAUNCH_CHECK();
  return JPM_OK;
}

extern "C" int32_t jpm_unpack_box_f32(void* stream, float* mesh, const float* packed, int32_t ny,
                                      int32_t nz, int32_t x0, int32_t x1, int32_t y0, int32_t y1,
                                      int32_t accumulate) {
  JPM_CHECK_ARG(packed && mesh && x1 >= x0 && y1 >= y0 && x0 >= 0 && y0 >= 0 && y1 <= ny, "bad box");
  const long long rows = (long long)(x1 - x0) * (y1 - y0);
  if (rows == 0) return JPM_OK;
  const int grid = (int)(rows < kNumSMs * 16 ? rows : kNumSMs * 16);
  if (accumulate)
    box_kernel<2><<<grid, 256, 0, (cudaStream_t)stream>>>(mesh, const_cast<float*>(packed), ny, nz, x0,
                                                          y0, x1 - x0, y1 - y0);
  else
    box_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(mesh, const_cast<float*>(packed), ny, nz, x0,
                                                          y0, x1 - x0, y1 - y0);
  JPM_LAUNCH_CHECK();
  return JPM_OK;
}
