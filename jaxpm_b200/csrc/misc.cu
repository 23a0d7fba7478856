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

static int ew_grid(long long n) {
  long long blocks = (n + 255) / 256;
  const long long cap = (long long)kNumSMs * 16;
  return (int)(blocks > cap ? cap : blocks);
}

}  // namespace jpm

using namespace jpm;

extern "C" int32_t jpm_abi_version(void) { return JPM_ABI_VERSION; }
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
