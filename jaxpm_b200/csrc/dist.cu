// Building blocks of the multi-GPU path (one process per GPU): 1-D batched FFT plans for the
// pencil/slab decomposition, pack/unpack kernels around the all-to-all transposes, and the fused
// k-space pass on an axis-permuted local spectrum block.
//   reference: jaxpm/distributed.py:37-42 (fft3d/ifft3d -> jaxdecomp.pfft3d/pifft3d, [ext]),
//              jaxpm/kernels.py:10-23 (fftk on the transposed layout), jaxpm/pm.py:49-56
#include <cufft.h>

#include "common.cuh"

struct jpm_fft1d {
  cufftHandle h = 0;
  void* work = nullptr;
  int kind = 0;  // 0 R2C, 1 C2R, 2 C2C
  int n = 0;
  long long batch = 0;
};

namespace jpm {

#define JPM_CUFFT(call)                                                                          \
  do {                                                                                           \
    cufftResult r__ = (call);                                                                    \
    if (r__ != CUFFT_SUCCESS) {                                                                  \
      jpm::set_error("%s failed: cufft error %d (%s:%d)", #call, (int)r__, __FILE__, __LINE__); \
      return JPM_ERR_CUFFT;                                                                      \
    }                                                                                            \
  } while (0)

// dst[b*dsb + j*dsj + i] = src[b*ssb + i*ssi + j]   (src contiguous in j, dst contiguous in i)
__global__ void __launch_bounds__(256)
transpose_c64_kernel(float2* __restrict__ dst, const float2* __restrict__ src, int ni, int nj,
                     long long ssi, long long ssb, long long dsj, long long dsb) {
  __shared__ float2 tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  const long long b = blockIdx.z;
  const float2* s = src + b * ssb;
  float2* d = dst + b * dsb;
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int i = i0 + r, j = j0 + tx;
    if (i < ni && j < nj) tile[r][tx] = s[(long long)i * ssi + j];
  }
  __syncthreads();
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int j = j0 + r, i = i0 + tx;
    if (i < ni && j < nj) d[(long long)j * dsj + i] = tile[tx][r];
  }
}

// rows of `ncols` contiguous complex numbers: dst[r*drs + c] = src[r*srs + c]
__global__ void __launch_bounds__(256)
copy2d_c64_kernel(float2* __restrict__ dst, const float2* __restrict__ src, long long nrows, int ncols,
                  long long srs, long long drs) {
  for (long long r = blockIdx.x; r < nrows; r += gridDim.x)
    for (int c = threadIdx.x; c < ncols; c += blockDim.x) dst[r * drs + c] = src[r * srs + c];
}

// Fused k-space pass on a local block [n0][n1][n2] of the spectrum whose array axes are an
// arbitrary permutation of (x, y, z): w{0,1,2} / a{0,1,2} are the per-array-axis tables (already
// sliced to the local index range), comp[d] says which array axis is physical direction d.
// KIND 0: out_d = i a_d g delta (3 outputs); 1: shear (6 outputs); 2: transpose of 0 (3 inputs, 1 output).
template <int KIND>
__global__ void __launch_bounds__(256)
kspace_local_kernel(const float2* __restrict__ dk, float2* __restrict__ out, const float* __restrict__ w0,
                    const float* __restrict__ w1, const float* __restrict__ w2,
                    const float* __restrict__ a0, const float* __restrict__ a1,
                    const float* __restrict__ a2, int n0, int n1, int n2, long long nspec, int cx,
                    int cy, int cz, float norm, float r_split2, const float* __restrict__ ftab, int ntab,
                    float fscale) {
  const long long nrows = (long long)n0 * n1;
  for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
    const int i0 = (int)(row / n1), i1 = (int)(row % n1);
    const float k0 = w0[i0], k1 = w1[i1];
    const float k01 = k0 * k0 + k1 * k1;
    const float g0 = a0[i0], g1 = a1[i1];
    for (int i2 = threadIdx.x; i2 < n2; i2 += blockDim.x) {
      const float k2 = w2[i2];
      const float kk = k01 + k2 * k2;
      float g = (kk == 0.f) ? 0.f : (1.0f / kk);
      g *= norm;
      if (KIND != 1) {
        if (r_split2 != 0.f) g *= expf(-kk * r_split2);
        if (ftab) {
          const float t = sqrtf(kk) * fscale;
          const int i = min((int)t, ntab - 2);
          const float fr = fminf(t - (float)i, 1.0f);
          g *= ftab[i] + fr * (ftab[i + 1] - ftab[i]);
        }
      }
      const float g2 = a2[i2];
      const float Ax = (cx == 0) ? g0 : ((cx == 1) ? g1 : g2);  // physical x, y, z
      const float Ay = (cy == 0) ? g0 : ((cy == 1) ? g1 : g2);
      const float Az = (cz == 0) ? g0 : ((cz == 1) ? g1 : g2);
      const float ax = Ax * g, ay = Ay * g, az = Az * g;
      const long long o = row * n2 + i2;
      if (KIND == 2) {
        const float2 d0 = dk[o], d1 = dk[nspec + o], d2 = dk[2 * nspec + o];
        out[o] = make_float2(ax * d0.y + ay * d1.y + az * d2.y, -(ax * d0.x + ay * d1.x + az * d2.x));
        continue;
      }
      const float2 d = dk[o];
      if (KIND == 0) {
        out[o] = make_float2(-ax * d.y, ax * d.x);
        out[nspec + o] = make_float2(-ay * d.y, ay * d.x);
        out[2 * nspec + o] = make_float2(-az * d.y, az * d.x);
      } else {
        const float m[6] = {Ax * Ax * g, Ay * Ay * g, Az * Az * g, Ax * Ay * g, Ax * Az * g, Ay * Az * g};
#pragma unroll
        for (int q = 0; q < 6; ++q) out[q * nspec + o] = make_float2(m[q] * d.x, m[q] * d.y);
      }
    }
  }
}

}  // namespace jpm

using namespace jpm;

extern "C" int32_t jpm_fft1d_create(jpm_fft1d** out, int32_t n, int64_t batch, int32_t kind) {
  JPM_CHECK_ARG(out && n > 0 && batch > 0 && kind >= 0 && kind <= 2, "bad arguments");
  jpm_fft1d* p = new jpm_fft1d();
  p->kind = kind; p->n = n; p->batch = batch;
  const cufftType types[3] = {CUFFT_R2C, CUFFT_C2R, CUFFT_C2C};
  long long nn[1] = {n};
  size_t ws = 0;
  JPM_CUFFT(cufftCreate(&p->h));
  JPM_CUFFT(cufftSetAutoAllocation(p->h, 0));
  // contiguous transforms, batches back to back (R2C/C2R use n/2+1 complex per transform)
  JPM_CUFFT(cufftMakePlanMany64(p->h, 1, nn, nullptr, 1, 0, nullptr, 1, 0, types[kind], batch, &ws));
  if (ws) JPM_CUDA(cudaMalloc(&p->work, ws));
  JPM_CUFFT(cufftSetWorkArea(p->h, p->work));
  *out = p;
  return JPM_OK;
}

extern "C" int32_t jpm_fft1d_destroy(jpm_fft1d* p) {
  if (!p) return JPM_OK;
  if (p->h) cufftDestroy(p->h);
  if (p->work) cudaFree(p->work);
  delete p;
  return JPM_OK;
}

extern "C" int32_t jpm_fft1d_exec(jpm_fft1d* p, void* stream, void* in, void* out, int32_t inverse) {
  JPM_CHECK_ARG(p && in && out, "null pointer");
  JPM_CUFFT(cufftSetStream(p->h, (cudaStream_t)stream));
  if (p->kind == 0) JPM_CUFFT(cufftExecR2C(p->h, (float*)in, (cufftComplex*)out));
  else if (p->kind == 1) JPM_CUFFT(cufftExecC2R(p->h, (cufftComplex*)in, (float*)out));
  else JPM_CUFFT(cufftExecC2C(p->h, (cufftComplex*)in, (cufftComplex*)out, inverse ? CUFFT_INVERSE : CUFFT_FORWARD));
  return JPM_OK;
}

extern "C" int32_t jpm_transpose_c64(void* stream, void* dst, const void* src, int32_t ni, int32_t nj,
                                     int64_t nb, int64_t src_stride_i, int64_t src_stride_b,
                                     int64_t dst_stride_j, int64_t dst_stride_b) {
  JPM_CHECK_ARG(dst && src && ni >= 0 && nj >= 0 && nb >= 0, "bad arguments");
  if (ni == 0 || nj == 0 || nb == 0) return JPM_OK;
  JPM_CHECK_ARG(nb <= 65535, "batch too large for grid.z");
  dim3 grid((nj + 31) / 32, (ni + 31) / 32, (unsigned)nb);
  transpose_c64_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((float2*)dst, (const float2*)src, ni, nj,
                                                              src_stride_i, src_stride_b, dst_stride_j,
                                                              dst_stride_b);
  JPM_LAUNCH_CHECK();
  return JPM_OK;
}

extern "C" int32_t jpm_copy2d_c64(void* stream, void* dst, const void* src, int64_t nrows, int32_t ncols,
                                  int64_t src_row_stride, int64_t dst_row_stride) {
  JPM_CHECK_ARG(dst && src && nrows >= 0 && ncols >= 0, "bad arguments");
  if (nrows == 0 || ncols == 0) return JPM_OK;
  const long long cap = (long long)kNumSMs * 16;
  copy2d_c64_kernel<<<(int)(nrows < cap ? nrows : cap), 256, 0, (cudaStream_t)stream>>>(
      (float2*)dst, (const float2*)src, nrows, ncols, src_row_stride, dst_row_stride);
  JPM_LAUNCH_CHECK();
  return JPM_OK;
}

extern "C" int32_t jpm_kspace_local_c64(void* stream, int32_t kind, const void* in, void* out,
                                        const float* w0, const float* w1, const float* w2,
                                        const float* a0, const float* a1, const float* a2, int32_t n0,
                                        int32_t n1, int32_t n2, int32_t axis_of_x, int32_t axis_of_y,
                                        int32_t axis_of_z, float norm, float r_split,
                                        const float* filter_tab, int32_t n_tab, float filter_kmax) {
  JPM_CHECK_ARG(in && out && w0 && w1 && w2 && a0 && a1 && a2, "null pointer");
  JPM_CHECK_ARG(kind >= 0 && kind <= 2 && n0 >= 0 && n1 >= 0 && n2 >= 0, "bad kind / shape");
  JPM_CHECK_ARG(axis_of_x >= 0 && axis_of_x < 3 && axis_of_y >= 0 && axis_of_y < 3 && axis_of_z >= 0 &&
                    axis_of_z < 3, "bad axis map");
  const long long rows = (long long)n0 * n1, nspec = rows * n2;
  if (nspec == 0) return JPM_OK;
  const float fscale = filter_tab ? (float)(n_tab - 1) / filter_kmax : 0.f;
  const long long cap = (long long)kNumSMs * 8;
  const int grid = (int)(rows < cap ? rows : cap);
  cudaStream_t s = (cudaStream_t)stream;
#define ARGS (const float2*)in, (float2*)out, w0, w1, w2, a0, a1, a2, n0, n1, n2, nspec, axis_of_x, \
             axis_of_y, axis_of_z, norm, r_split * r_split, filter_tab, n_tab, fscale
  if (kind == 0) kspace_local_kernel<0><<<grid, 256, 0, s>>>(ARGS);
  else if (kind == 1) kspace_local_kernel<1><<<grid, 256, 0, s>>>(ARGS);
  else kspace_local_kernel<2><<<grid, 256, 0, s>>>(ARGS);
#undef ARGS
  JPM_LAUNCH_CHECK();
  return JPM_OK;
}
