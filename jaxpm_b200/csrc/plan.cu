// K2/K3/K4 — FFT plan (cuFFT R2C / batched C2R), the fused k-space pass, and the
// composed force / step entry points.
//   reference: jaxpm/distributed.py:37-42 (fft3d/ifft3d), jaxpm/kernels.py:10-23,41-115,139-165,
//              jaxpm/pm.py:12-58 (pm_forces), :88-124 (2LPT source), jaxpm/ode.py:91-117
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"

#include "plan_internal.cuh"

namespace jpm {

#define JPM_CUFFT(call)                                                          \
  do {                                                                           \
    cufftResult r__ = (call);                                                    \
    if (r__ != CUFFT_SUCCESS) {                                                  \
      jpm::set_error("%s failed: cufft error %d (%s:%d)", #call, (int)r__, __FILE__, __LINE__); \
      return JPM_ERR_CUFFT;                                                      \
    }                                                                            \
  } while (0)

// ---------------------------------------------------------------------------------
// K3: delta_k -> NOUT spectra, one HBM pass (read 8 B, write NOUT*8 B per mode).
// KIND 0: force spectra  out_d = i a_d delta / k^2 * G * norm       (3 outputs)
// KIND 1: shear spectra  out_ij = a_i a_j delta / k^2 * norm        (6 outputs: 00 11 22 01 02 12)
// KIND 2: transpose of KIND 0 (its VJP): out = sum_d (-i a_d) in_d / k^2 * G * norm   (3 inputs, 1 output)
// KIND 3: transpose of KIND 1 (the multipliers a_i a_j / k^2 are real and even: self-adjoint):
//         out = sum_q a_i a_j in_q / k^2 * norm                                      (6 inputs, 1 output)
// Threads run along z (fastest axis) so loads/stores of the interleaved complex rows coalesce.
// ---------------------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(256)
kspace_kernel(const float2* __restrict__ dk, float2* __restrict__ out, const float* __restrict__ wx,
              const float* __restrict__ wy, const float* __restrict__ wz, const float* __restrict__ ax,
              const float* __restrict__ ay, const float* __restrict__ az, int nx, int ny, int nzh,
              long long nspec, float norm, float r_split2, const float* __restrict__ ftab, int ntab,
              float fscale) {
  const long long nrows = (long long)nx * ny;
  for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
    const int ix = (int)(row / ny), iy = (int)(row % ny);
    const float kx = wx[ix], ky = wy[iy];
    const float kxy2 = kx * kx + ky * ky;  // ((0 + kx^2) + ky^2), kernels.py:87
    const float a0 = ax[ix], a1 = ay[iy];
    for (int iz = threadIdx.x; iz < nzh; iz += blockDim.x) {
      const float kz = wz[iz];
      const float kk = kxy2 + kz * kz;
      float g = (kk == 0.f) ? 0.f : (1.0f / kk);  // -invlaplace = 1/k^2, 0 at k=0
      g *= norm;
      if (KIND != 1 && KIND != 3) {
        if (r_split2 != 0.f) g *= expf(-kk * r_split2);
        if (ftab) {
          const float t = sqrtf(kk) * fscale;
          const int i = min((int)t, ntab - 2);
          const float fr = fminf(t - (float)i, 1.0f);
          g *= ftab[i] + fr * (ftab[i + 1] - ftab[i]);
        }
      }
      const long long o = row * nzh + iz;
      const float a2 = az[iz];
      if (KIND == 2) {
        const float2 d0 = __ldcs(dk + o), d1 = __ldcs(dk + nspec + o), d2 = __ldcs(dk + 2 * nspec + o);
        const float g0 = a0 * g, g1 = a1 * g, g2 = a2 * g;
        // (-i a g)(re + i im) = (a g im, -a g re)
        __stcs(out + o, make_float2(g0 * d0.y + g1 * d1.y + g2 * d2.y,
                                    -(g0 * d0.x + g1 * d1.x + g2 * d2.x)));
        continue;
      }
      if (KIND == 3) {
        const float m[6] = {a0 * a0 * g, a1 * a1 * g, a2 * a2 * g, a0 * a1 * g, a0 * a2 * g, a1 * a2 * g};
        float2 acc = make_float2(0.f, 0.f);
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          const float2 v = __ldcs(dk + q * nspec + o);
          acc.x = fmaf(m[q], v.x, acc.x);
          acc.y = fmaf(m[q], v.y, acc.y);
        }
        __stcs(out + o, acc);
        continue;
      }
      const float2 d = __ldcs(dk + o);
      if (KIND == 0) {
        // -(i a) * (-1/k^2) * delta = i a g delta : (re,im) -> (-a g im, a g re)
        const float g0 = a0 * g, g1 = a1 * g, g2 = a2 * g;
        __stcs(out + o, make_float2(-g0 * d.y, g0 * d.x));
        __stcs(out + nspec + o, make_float2(-g1 * d.y, g1 * d.x));
        __stcs(out + 2 * nspec + o, make_float2(-g2 * d.y, g2 * d.x));
      } else {
        // (i a_i)(i a_j) * (-1/k^2) * delta = a_i a_j g delta
        const float m[6] = {a0 * a0 * g, a1 * a1 * g, a2 * a2 * g, a0 * a1 * g, a0 * a2 * g, a1 * a2 * g};
#pragma unroll
        for (int q = 0; q < 6; ++q) __stcs(out + q * nspec + o, make_float2(m[q] * d.x, m[q] * d.y));
      }
    }
  }
}

__global__ void __launch_bounds__(256)
kfilter_logtab_kernel(const float2* __restrict__ in, float2* __restrict__ out,
                      const float* __restrict__ wx, const float* __restrict__ wy,
                      const float* __restrict__ wz, int nx, int ny, int nzh, const float* __restrict__ tab,
                      int ntab, float lkmin, float lkmax, float sx, float sy, float sz, float norm) {
  const long long nrows = (long long)nx * ny;
  const float inv = (float)(ntab - 1) / (lkmax - lkmin);
  for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
    const int ix = (int)(row / ny), iy = (int)(row % ny);
    const float kx = wx[ix] * sx, ky = wy[iy] * sy;
    const float kxy2 = kx * kx + ky * ky;
    for (int iz = threadIdx.x; iz < nzh; iz += blockDim.x) {
      const float kz = wz[iz] * sz;
      const float kk = kxy2 + kz * kz;
      float t = (kk > 0.f) ? (0.5f * log10f(kk) - lkmin) * inv : 0.f;
      t = fminf(fmaxf(t, 0.f), (float)(ntab - 1));
      const int i = min((int)t, ntab - 2);
      const float fr = t - (float)i;
      const float m = (tab[i] + fr * (tab[i + 1] - tab[i])) * norm;
      const long long o = row * nzh + iz;
      const float2 d = in[o];
      out[o] = make_float2(m * d.x, m * d.y);
    }
  }
}

__global__ void __launch_bounds__(256)
lpt2_source_kernel(float* __restrict__ d2, const float* __restrict__ s, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float s00 = s[i], s11 = s[n + i], s22 = s[2 * n + i];
    const float s01 = s[3 * n + i], s02 = s[4 * n + i], s12 = s[5 * n + i];
    // pm.py:92-109: delta2 = s11*s00 + s22*(s00+s11) - s01^2 - s02^2 - s12^2
    float v = s11 * s00;
    v -= s01 * s01;
    v -= s02 * s02;
    v += s22 * (s00 + s11);
    v -= s12 * s12;
    d2[i] = v;
  }
}


// VJP of lpt2_source_kernel: t_q = g * d delta2 / d s_q  (6 meshes), pm.py:92-109 differentiated
__global__ void __launch_bounds__(256)
lpt2_source_adj_kernel(float* __restrict__ t, const float* __restrict__ s, const float* __restrict__ g, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float s00 = s[i], s11 = s[n + i], s22 = s[2 * n + i];
    const float s01 = s[3 * n + i], s02 = s[4 * n + i], s12 = s[5 * n + i];
    const float gi = g[i];
    t[i] = gi * (s11 + s22);
    t[n + i] = gi * (s00 + s22);
    t[2 * n + i] = gi * (s00 + s11);
    t[3 * n + i] = -2.0f * gi * s01;
    t[4 * n + i] = -2.0f * gi * s02;
    t[5 * n + i] = -2.0f * gi * s12;
  }
}

static void build_tables(int n, int nh, std::vector<float>& w, std::vector<float>& a) {
  // fftk: w = 2*pi*fftfreq(n) (kernels.py:10-23, [ext] jaxdecomp.fftfreq3d), stored fp32;
  // gradient_kernel order 1 (kernels.py:62-66) evaluated in fp64 at the fp32 frequency.
  w.resize(nh);
  a.resize(nh);
  for (int i = 0; i < nh; ++i) {
    const int f = (i < (n + 1) / 2) ? i : i - n;
    const float wf = (float)(2.0 * M_PI * (double)f / (double)n);
    w[i] = wf;
    const double wd = (double)wf;
    a[i] = (float)((8.0 * std::sin(wd) - std::sin(2.0 * wd)) / 6.0);
    // the reference takes .real of a C2C inverse (distributed.py:41-42): the self-conjugate
    // Nyquist mode of an odd kernel contributes nothing.  a(pi) is 0 up to rounding; make it exact
    // so that the C2R transform sees a Hermitian spectrum.
    if (n % 2 == 0 && i == n / 2) a[i] = 0.f;
  }
}

static int32_t upload(float** dst, const std::vector<float>& v) {
  JPM_CUDA(cudaMalloc(dst, v.size() * sizeof(float)));
  JPM_CUDA(cudaMemcpy(*dst, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice));
  return JPM_OK;
}

static int kspace_grid(const jpm_plan* p) {
  const long long rows = (long long)p->nx * p->ny;
  const long long cap = (long long)kNumSMs * 8;
  return (int)(rows < cap ? rows : cap);
}


// ---------------------------------------------------------------------------------
// Ghost zones of the padded meshes.  Along one axis of interior length n with G ghosts per side
// (padded coordinates): ghost [0, G) is the periodic image of interior [n, n+G), ghost
// [n+G, n+2G) is the image of interior [G, 2G).
//   FOLD (after paint):  interior += ghost, axes in the order x, y, z with shrinking extents of the
//                        other axes, so that edge/corner ghosts end up in the interior;
//   FILL (before read):  ghost = interior, axes in the order z, y, x with growing extents.
// One thread per ghost cell of the slab pair; u/v enumerate the other two axes (v fastest).
// ---------------------------------------------------------------------------------
template <bool FOLD>
__global__ void __launch_bounds__(256)
ghost_kernel(float* __restrict__ a, long long batch_stride, int G, int n, long long sa, int nu, int u0,
             long long su, int nv, int v0, long long sv, bool g_fastest) {
  float* m = a + blockIdx.y * batch_stride;
  const long long total = 2LL * G * nu * nv;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    int g2, u, v;
    if (g_fastest) {  // z-axis slabs: the ghost index is the contiguous one
      g2 = (int)(t % (2 * G));
      const long long r = t / (2 * G);
      v = (int)(r % nv);
      u = (int)(r / nv);
    } else {
      v = (int)(t % nv);
      const long long r = t / nv;
      u = (int)(r % nu);
      g2 = (int)(r / nu);
    }
    const int ghost = g2 < G ? g2 : n + g2;                 // [0,G) or [n+G, n+2G)
    const int inner = g2 < G ? g2 + n : g2;                 // its periodic image inside
    const long long base = (long long)(u + u0) * su + (long long)(v + v0) * sv;
    if (FOLD) m[base + inner * sa] += m[base + ghost * sa];
    else m[base + ghost * sa] = m[base + inner * sa];
  }
}

template <bool FOLD>
static int32_t ghost_pass(jpm_plan* p, cudaStream_t st, float* a, int batch) {
  const int G = p->G;
  const long long sx = (long long)p->nyp * p->nzp, sy = p->nzp, sz = 1;
  // axis, other axes (u, v) with [start, count)
  for (int step = 0; step < 3; ++step) {
    const int axis = FOLD ? step : 2 - step;
    long long total;
    if (axis == 0) {        // x slabs: y, z over the full padded extents
      total = 2LL * G * p->nyp * p->nzp;
      ghost_kernel<FOLD><<<dim3((unsigned)std::min<long long>((total + 255) / 256, kNumSMs * 16), batch), 256, 0, st>>>(
          a, p->npad, G, p->nx, sx, p->nyp, 0, sy, p->nzp, 0, sz, false);
    } else if (axis == 1) { // y slabs: x interior, z full
      total = 2LL * G * p->nx * p->nzp;
      ghost_kernel<FOLD><<<dim3((unsigned)std::min<long long>((total + 255) / 256, kNumSMs * 16), batch), 256, 0, st>>>(
          a, p->npad, G, p->ny, sy, p->nx, G, sx, p->nzp, 0, sz, false);
    } else {                // z slabs: x, y interior
      total = 2LL * G * p->nx * p->ny;
      ghost_kernel<FOLD><<<dim3((unsigned)std::min<long long>((total + 255) / 256, kNumSMs * 16), batch), 256, 0, st>>>(
          a, p->npad, G, p->nz, sz, p->nx, G, sx, p->ny, G, sy, true);
    }
    JPM_LAUNCH_CHECK();
  }
  return JPM_OK;
}

int32_t encode_tensor_map(CUtensorMap* out, float* base, int rank, const unsigned long long* dims,
                          const unsigned long long* strides_bytes, const unsigned* box) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                               const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                               CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    JPM_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q));
    if (q != cudaDriverEntryPointSuccess || !sym) {
      set_error("cuTensorMapEncodeTiled not available from the driver");
      return JPM_ERR_CUDA;
    }
    fn = (EncodeFn)sym;
  }
  cuuint64_t d[5], s[4];
  cuuint32_t b[5], e[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; e[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
  const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, base, d, s, b, e,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: CUresult %d", (int)r);
    return JPM_ERR_CUDA;
  }
  return JPM_OK;
}

int32_t plan_enable_padded(jpm_plan* p) {
  if (p->G || p->density_p) return JPM_OK;
  if (p->nz % 4 != 0) return JPM_OK;  // TMA needs 16-byte row strides
  const int G = kGhost;
  p->nxp = p->nx + 2 * G; p->nyp = p->ny + 2 * G; p->nzp = p->nz + 2 * G;
  p->npad = (long long)p->nxp * p->nyp * p->nzp;
  if (p->npad >= (1ll << 31)) return JPM_OK;
  JPM_CUDA(cudaMalloc(&p->density_p, p->npad * sizeof(float)));
  JPM_CUDA(cudaMalloc(&p->force3_p, 3 * p->npad * sizeof(float)));
  JPM_CUDA(cudaMemset(p->force3_p, 0, 3 * p->npad * sizeof(float)));
  p->G = G;
  // the hand-written chain when the shape allows it (JPM_PMFFT=0 keeps cuFFT for A/B comparisons)
  const char* env = getenv("JPM_PMFFT");
  if (!(env && env[0] == '0')) {
    int32_t rc = pmfft_enable(p);
    if (rc) return rc;
    if (p->fft_on) return JPM_OK;
  }
  int n[3] = {p->nx, p->ny, p->nz};
  int remb[3] = {p->nxp, p->nyp, p->nzp};      // real side: embedded in the padded array
  int cemb[3] = {p->nx, p->ny, p->nzh};        // spectrum side: compact
  size_t ws[2] = {0, 0};
  JPM_CUFFT(cufftCreate(&p->r2c_p));
  JPM_CUFFT(cufftSetAutoAllocation(p->r2c_p, 0));
  JPM_CUFFT(cufftMakePlanMany(p->r2c_p, 3, n, remb, 1, (int)p->npad, cemb, 1, (int)p->nspec, CUFFT_R2C, 1, &ws[0]));
  JPM_CUFFT(cufftCreate(&p->c2r3_p));
  JPM_CUFFT(cufftSetAutoAllocation(p->c2r3_p, 0));
  JPM_CUFFT(cufftMakePlanMany(p->c2r3_p, 3, n, cemb, 1, (int)p->nspec, remb, 1, (int)p->npad, CUFFT_C2R, 3, &ws[1]));
  const size_t need = ws[0] > ws[1] ? ws[0] : ws[1];
  if (need > p->work_bytes) {
    if (p->work) cudaFree(p->work);
    p->work = nullptr;
    JPM_CUDA(cudaMalloc(&p->work, need));
    p->work_bytes = need;
    JPM_CUFFT(cufftSetWorkArea(p->r2c, p->work));
    JPM_CUFFT(cufftSetWorkArea(p->c2r1, p->work));
    JPM_CUFFT(cufftSetWorkArea(p->c2r3, p->work));
  }
  JPM_CUFFT(cufftSetWorkArea(p->r2c_p, p->work));
  JPM_CUFFT(cufftSetWorkArea(p->c2r3_p, p->work));
  return JPM_OK;
}

int32_t plan_padded_forces(jpm_plan* p, cudaStream_t st, float r_split, const float* filter_tab, int n_tab,
                           float filter_kmax) {
  JPM_CHECK_ARG(p->G > 0, "padded meshes not enabled");
  if (p->fft_on) return pmfft_forces(p, st, r_split, filter_tab, n_tab, filter_kmax);
  const long long off = ((long long)p->G * p->nyp + p->G) * p->nzp + p->G;   // interior origin
  int32_t rc;
  if ((rc = ghost_pass<true>(p, st, p->density_p, 1))) return rc;
  if (p->timer) p->timer->mark(st, "ghost_fold");
  JPM_CUFFT(cufftSetStream(p->r2c_p, st));
  JPM_CUFFT(cufftExecR2C(p->r2c_p, p->density_p + off, (cufftComplex*)p->spec));
  if (p->timer) p->timer->mark(st, "fft_r2c(cuFFT)");
  if ((rc = jpm_greens_grad_c64(p, st, p->spec, p->spec3, 1.0f / (float)p->ncell, r_split, filter_tab, n_tab,
                                filter_kmax)))
    return rc;
  if (p->timer) p->timer->mark(st, "greens_grad");
  JPM_CUFFT(cufftSetStream(p->c2r3_p, st));
  JPM_CUFFT(cufftExecC2R(p->c2r3_p, (cufftComplex*)p->spec3, p->force3_p + off));
  if (p->timer) p->timer->mark(st, "ifft_c2r_x3(cuFFT)");
  rc = ghost_pass<false>(p, st, p->force3_p, 3);
  if (p->timer) p->timer->mark(st, "ghost_fill");
  return rc;
}

// compact [nx][ny][nz] <-> interior of the padded [nxp][nyp][nzp] array (float4 along z)
template <bool TO_PADDED>
__global__ void __launch_bounds__(256)
pad_copy_kernel(float* __restrict__ padded, float* __restrict__ compact, int nx, int ny, int nz4, int nyp, int nzp,
                int G, long long pbatch, long long cbatch) {
  const long long total = (long long)nx * ny * nz4;
  float* pp = padded + blockIdx.y * pbatch;
  float* cc = compact + blockIdx.y * cbatch;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int k4 = (int)(t % nz4);
    const long long r = t / nz4;
    const int j = (int)(r % ny), i = (int)(r / ny);
    float4* a = reinterpret_cast<float4*>(pp + ((long long)(i + G) * nyp + (j + G)) * nzp + G) + k4;
    float4* b = reinterpret_cast<float4*>(cc + ((long long)i * ny + j) * (4 * nz4)) + k4;
    if (TO_PADDED) *a = *b;
    else *b = *a;
  }
}

}  // namespace jpm

using namespace jpm;

extern "C" int32_t jpm_density_to_force_meshes_fused(jpm_plan* p, void* stream, const float* density,
                                                     float* force3, float r_split, const float* filter_tab,
                                                     int32_t n_tab, float filter_kmax) {
  JPM_CHECK_ARG(!(p && p->is_slab), "not available on a multi-GPU slab plan");
  JPM_CHECK_ARG(p && density && force3, "null pointer");
  JPM_CHECK_ARG(!filter_tab || (n_tab >= 2 && filter_kmax > 0.f), "bad filter table");
  int32_t rc = plan_enable_padded(p);
  if (rc) return rc;
  JPM_CHECK_ARG(p->G > 0, "shape not supported by the ghost-zone path (nz % 4 != 0 or too large)");
  cudaStream_t st = (cudaStream_t)stream;
  const long long total = p->ncell / 4;
  const unsigned blocks = (unsigned)std::min<long long>((total + 255) / 256, kNumSMs * 16);
  JPM_CUDA(cudaMemsetAsync(p->density_p, 0, p->npad * sizeof(float), st));
  pad_copy_kernel<true><<<dim3(blocks, 1), 256, 0, st>>>(p->density_p, const_cast<float*>(density), p->nx, p->ny,
                                                        p->nz / 4, p->nyp, p->nzp, p->G, p->npad, p->ncell);
  JPM_LAUNCH_CHECK();
  if ((rc = plan_padded_forces(p, st, r_split, filter_tab, n_tab, filter_kmax))) return rc;
  pad_copy_kernel<false><<<dim3(blocks, 3), 256, 0, st>>>(p->force3_p, force3, p->nx, p->ny, p->nz / 4, p->nyp,
                                                         p->nzp, p->G, p->npad, p->ncell);
  JPM_LAUNCH_CHECK();
  return JPM_OK;
}

extern "C" int32_t jpm_density_to_potential_fused(jpm_plan* p, void* stream, const float* density, float* psi,
                                                  float r_split, const float* filter_tab, int32_t n_tab,
                                                  float filter_kmax) {
  JPM_CHECK_ARG(!(p && p->is_slab), "not available on a multi-GPU slab plan");
  JPM_CHECK_ARG(p && density && psi, "null pointer");
  JPM_CHECK_ARG(!filter_tab || (n_tab >= 2 && filter_kmax > 0.f), "bad filter table");
  int32_t rc = plan_enable_padded(p);
  if (rc) return rc;
  JPM_CHECK_ARG(p->G > 0 && p->fft_on, "potential chain needs a power-of-two mesh (fused FFT chain)");
  cudaStream_t st = (cudaStream_t)stream;
  const long long total = p->ncell / 4;
  const unsigned blocks = (unsigned)std::min<long long>((total + 255) / 256, kNumSMs * 16);
  JPM_CUDA(cudaMemsetAsync(p->density_p, 0, p->npad * sizeof(float), st));
  pad_copy_kernel<true><<<dim3(blocks, 1), 256, 0, st>>>(p->density_p, const_cast<float*>(density), p->nx, p->ny,
                                                        p->nz / 4, p->nyp, p->nzp, p->G, p->npad, p->ncell);
  JPM_LAUNCH_CHECK();
  if ((rc = pmfft_potential(p, st, r_split, filter_tab, n_tab, filter_kmax))) return rc;
  pad_copy_kernel<false><<<dim3(blocks, 1), 256, 0, st>>>(p->force3_p, psi, p->nx, p->ny, p->nz / 4, p->nyp,
                                                         p->nzp, p->G, p->npad, p->ncell);
  JPM_LAUNCH_CHECK();
  return JPM_OK;
}

extern "C" int32_t jpm_linear_field_f32(jpm_plan* p, void* stream, const float* white, float* out, const float* tab,
                                        int32_t n_tab, float log10_kmin, float log10_kmax, float kscale_x,
                                        float kscale_y, float kscale_z, float dc_amp) {
  JPM_CHECK_ARG(!(p && p->is_slab), "not available on a multi-GPU slab plan");
  JPM_CHECK_ARG(p && white && out && tab && n_tab >= 2 && log10_kmax > log10_kmin, "bad arguments");
  int32_t rc = plan_enable_padded(p);
  if (rc) return rc;
  JPM_CHECK_ARG(p->G > 0 && p->fft_on, "linear_field on the fused chain needs a power-of-two mesh");
  cudaStream_t st = (cudaStream_t)stream;
  const long long total = p->ncell / 4;
  const unsigned blocks = (unsigned)std::min<long long>((total + 255) / 256, kNumSMs * 16);
  JPM_CUDA(cudaMemsetAsync(p->density_p, 0, p->npad * sizeof(float), st));
  pad_copy_kernel<true><<<dim3(blocks, 1), 256, 0, st>>>(p->density_p, const_cast<float*>(white), p->nx, p->ny,
                                                        p->nz / 4, p->nyp, p->nzp, p->G, p->npad, p->ncell);
  JPM_LAUNCH_CHECK();
  if ((rc = pmfft_linear_field(p, st, tab, n_tab, log10_kmin, log10_kmax, kscale_x, kscale_y, kscale_z, dc_amp)))
    return rc;
  pad_copy_kernel<false><<<dim3(blocks, 1), 256, 0, st>>>(p->psi_p, out, p->nx, p->ny, p->nz / 4, p->nyp, p->nzp,
                                                         p->G, p->npad, p->ncell);
  JPM_LAUNCH_CHECK();
  return JPM_OK;
}

extern "C" int32_t jpm_plan_create(jpm_plan** out, int32_t nx, int32_t ny, int32_t nz) {
  JPM_CHECK_ARG(out, "null plan pointer");
  JPM_CHECK_ARG(nx > 0 && ny > 0 && nz > 0, "bad mesh shape");
  JPM_CHECK_ARG((int64_t)nx * ny * nz < (1ll << 31), "mesh too large for int32 cell ids");
  jpm_plan* p = new jpm_plan();
  p->nx = nx; p->ny = ny; p->nz = nz; p->nzh = nz / 2 + 1;
  p->ncell = (long long)nx * ny * nz;
  p->nspec = (long long)nx * ny * p->nzh;
  int n[3] = {nx, ny, nz};
  size_t ws[3] = {0, 0, 0};
  cufftHandle* hs[3] = {&p->r2c, &p->c2r1, &p->c2r3};
  const cufftType types[3] = {CUFFT_R2C, CUFFT_C2R, CUFFT_C2R};
  const int batches[3] = {1, 1, 3};
  for (int i = 0; i < 3; ++i) {
    JPM_CUFFT(cufftCreate(hs[i]));
    JPM_CUFFT(cufftSetAutoAllocation(*hs[i], 0));
    const long long idist = (types[i] == CUFFT_R2C) ? p->ncell : p->nspec;
    const long long odist = (types[i] == CUFFT_R2C) ? p->nspec : p->ncell;
    JPM_CUFFT(cufftMakePlanMany(*hs[i], 3, n, nullptr, 1, (int)idist, nullptr, 1, (int)odist, types[i],
                                batches[i], &ws[i]));
    if (ws[i] > p->work_bytes) p->work_bytes = ws[i];
  }
  if (p->work_bytes) JPM_CUDA(cudaMalloc(&p->work, p->work_bytes));
  for (int i = 0; i < 3; ++i) JPM_CUFFT(cufftSetWorkArea(*hs[i], p->work));
  std::vector<float> w, a;
  int32_t rc;
  build_tables(nx, nx, w, a);
  if ((rc = upload(&p->wx, w)) || (rc = upload(&p->ax, a))) return rc;
  build_tables(ny, ny, w, a);
  if ((rc = upload(&p->wy, w)) || (rc = upload(&p->ay, a))) return rc;
  build_tables(nz, p->nzh, w, a);
  if ((rc = upload(&p->wz, w)) || (rc = upload(&p->az, a))) return rc;
  JPM_CUDA(cudaMalloc(&p->density, p->ncell * sizeof(float)));
  JPM_CUDA(cudaMalloc(&p->spec, p->nspec * sizeof(float2)));
  JPM_CUDA(cudaMalloc(&p->spec3, 3 * p->nspec * sizeof(float2)));
  JPM_CUDA(cudaMalloc(&p->force3, 3 * p->ncell * sizeof(float)));
  *out = p;
  return JPM_OK;
}

extern "C" int32_t jpm_plan_destroy(jpm_plan* p) {
  if (!p) return JPM_OK;
  if (p->r2c) cufftDestroy(p->r2c);
  if (p->c2r1) cufftDestroy(p->c2r1);
  if (p->c2r3) cufftDestroy(p->c2r3);
  if (p->r2c_p) cufftDestroy(p->r2c_p);
  if (p->c2r3_p) cufftDestroy(p->c2r3_p);
  if (p->is_slab) {
    // every mesh of a slab plan lives inside sym_base; peers must have stopped using it (the caller
    // synchronises the ranks before destroying)
    if (p->ipc_peers)
      for (int r = 0; r < 8; ++r)
        if (p->peer_base[r]) cudaIpcCloseMemHandle(p->peer_base[r]);
    p->fft_at = nullptr; p->fft_b3 = nullptr; p->fft_t01 = nullptr;
    p->density_p = nullptr; p->force3_p = nullptr;
    if (p->sym_base) cudaFree(p->sym_base);
  }
  pmfft_destroy(p);
  if (p->density_p) cudaFree(p->density_p);
  if (p->force3_p) cudaFree(p->force3_p);
  if (p->psi_p) cudaFree(p->psi_p);
  void* bufs[] = {p->work, p->wx, p->wy, p->wz, p->ax, p->ay, p->az, p->density, p->spec, p->spec3, p->force3};
  for (void* b : bufs)
    if (b) cudaFree(b);
  delete p;
  return JPM_OK;
}

extern "C" int32_t jpm_fft3d_r2c(jpm_plan* p, void* stream, const float* in, void* out) {
  JPM_CHECK_ARG(!(p && p->is_slab), "not available on a multi-GPU slab plan");
  JPM_CHECK_ARG(p && in && out, "null pointer");
  JPM_CUFFT(cufftSetStream(p->r2c, (cudaStream_t)stream));
  JPM_CUFFT(cufftExecR2C(p->r2c, const_cast<float*>(in), (cufftComplex*)out));
  return JPM_OK;
}

extern "C" int32_t jpm_ifft3d_c2r(jpm_plan* p, void* stream, void* in, float* out, int32_t batch) {
  JPM_CHECK_ARG(!(p && p->is_slab), "not available on a multi-GPU slab plan");
  JPM_CHECK_ARG(p && in && out, "null pointer");
  JPM_CHECK_ARG(batch == 1 || batch == 3, "batch must be 1 or 3");
  cufftHandle h = (batch == 3) ? p->c2r3 : p->c2r1;
  JPM_CUFFT(cufftSetStream(h, (cudaStream_t)stream));
  JPM_CUFFT(cufftExecC2R(h, (cufftComplex*)in, out));
  return JPM_OK;
}

extern "C" int32_t jpm_greens_grad_c64(jpm_plan* p, void* stream, const void* delta_k, void* out3,
                                       float norm, float r_split, const float* filter_tab,
                                       int32_t n_tab, float filter_kmax) {
  JPM_CHECK_ARG(!(p && p->is_slab), "not available on a multi-GPU slab plan");
  JPM_CHECK_ARG(p && delta_k && out3, "null pointer");
  JPM_CHECK_ARG(!filter_tab || (n_tab >= 2 && filter_kmax > 0.f), "bad filter table");
  const float fscale = filter_tab ? (float)(n_tab - 1) / filter_kmax : 0.f;
  kspace_kernel<0><<<kspace_grid(p), 256, 0, (cudaStream_t)stream>>>(
      (const float2*)delta_k, (float2*)out3, p->wx, p->wy, p->wz, p->ax, p->ay, p->az, p->nx, p->ny,
      p->nzh, p->nspec, norm, r_split * r_split, filter_tab, n_tab, fscale);
  JPM_LAUNCH_CHECK();
  return JPM_OK;
}

extern "C" int32_t jpm_greens_div_c64(jpm_plan* p, void* stream, const void* in3, void* out,
                                      float norm, float r_split, const float* filter_tab,
                                      int32_t n_tab, float filter_kmax) {
  JPM_CHECK_ARG(!(p && p->is_slab), "not available on a multi-GPU slab plan");
  JPM_CHECK_ARG(p && in3 && out, "null pointer");
  JPM_CHECK_ARG(!filter_tab || (n_tab >= 2 && filter_kmax > 0.f), "bad filter table");
  const float fscale = filter_tab ? (float)(n_tab - 1) / filter_kmax : 0.f;
  kspace_kernel<2><<<kspace_grid(p), 256, 0, (cudaStream_t)stream>>>(
      (const float2*)in3, (float2*)out, p->wx, p->wy, p->wz, p->ax, p->ay, p->az, p->nx, p->ny,
      p->nzh, p->nspec, norm, r_split * r_split, filter_tab, n_tab, fscale);
  JPM_LAUNCH_CHECK();
  return JPM_OK;
}

extern "C" int32_t jpm_lpt2_shear_c64(jpm_plan* p, void* stream, const void* delta_k, void* out6,
                                      float norm) {
  JPM_CHECK_ARG(!(p && p->is_slab), "not available on a multi-GPU slab plan");
  JPM_CHECK_ARG(p && delta_k && out6, "null pointer");
  kspace_kernel<1><<<kspace_grid(p), 256, 0, (cudaStream_t)stream>>>(
      (const float2*)delta_k, (float2*)out6, p->wx, p->wy, p->wz, p->ax, p->ay, p->az, p->nx, p->ny,
      p->nzh, p->nspec, norm, 0.f, nullptr, 0, 0.f);
  JPM_LAUNCH_CHECK();
  return JPM_OK;
}

extern "C" int32_t jpm_lpt2_shear_adj_c64(jpm_plan* p, void* stream, const void* in6, void* out, float norm) {
  JPM_CHECK_ARG(!(p && p->is_slab), "not available on a multi-GPU slab plan");
  JPM_CHECK_ARG(p && in6 && out, "null pointer");
  kspace_kernel<3><<<kspace_grid(p), 256, 0, (cudaStream_t)stream>>>(
      (const float2*)in6, (float2*)out, p->wx, p->wy, p->wz, p->ax, p->ay, p->az, p->nx, p->ny,
      p->nzh, p->nspec, norm, 0.f, nullptr, 0, 0.f);
  JPM_LAUNCH_CHECK();
  return JPM_OK;
}

extern "C" int32_t jpm_lpt2_source_adj_f32(void* stream, float* t6, const float* shear6, const float* cot,
                                           int64_t ncell) {
  JPM_CHECK_ARG(t6 && shear6 && cot && ncell >= 0, "null pointer");
  if (ncell == 0) return JPM_OK;
  long long blocks = (ncell + 255) / 256;
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  lpt2_source_adj_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(t6, shear6, cot, ncell);
  JPM_LAUNCH_CHECK();
  return JPM_OK;
}

extern "C" int32_t jpm_lpt2_source_f32(void* stream, float* delta2, const float* shear6, int64_t ncell) {
  JPM_CHECK_ARG(delta2 && shear6 && ncell >= 0, "null pointer");
  if (ncell == 0) return JPM_OK;
  long long blocks = (ncell + 255) / 256;
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  lpt2_source_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(delta2, shear6, ncell);
  JPM_LAUNCH_CHECK();
  return JPM_OK;
}

extern "C" int32_t jpm_kfilter_logtab_c64(jpm_plan* p, void* stream, const void* in, void* out,
                                          const float* tab, int32_t n_tab, float log10_kmin,
                                          float log10_kmax, float kscale_x, float kscale_y,
                                          float kscale_z, float norm) {
  JPM_CHECK_ARG(!(p && p->is_slab), "not available on a multi-GPU slab plan");
  JPM_CHECK_ARG(p && in && out && tab && n_tab >= 2 && log10_kmax > log10_kmin, "bad arguments");
  kfilter_logtab_kernel<<<kspace_grid(p), 256, 0, (cudaStream_t)stream>>>(
      (const float2*)in, (float2*)out, p->wx, p->wy, p->wz, p->nx, p->ny, p->nzh, tab, n_tab,
      log10_kmin, log10_kmax, kscale_x, kscale_y, kscale_z, norm);
  JPM_LAUNCH_CHECK();
  return JPM_OK;
}

extern "C" int32_t jpm_density_to_force_meshes(jpm_plan* p, void* stream, const float* density,
                                               float* force3, float r_split, const float* filter_tab,
                                               int32_t n_tab, float filter_kmax) {
  JPM_CHECK_ARG(!(p && p->is_slab), "not available on a multi-GPU slab plan");
  JPM_CHECK_ARG(p && density && force3, "null pointer");
  int32_t rc;
  cudaStream_t st = (cudaStream_t)stream;
  if ((rc = jpm_fft3d_r2c(p, stream, density, p->spec))) return rc;
  if (p->timer) p->timer->mark(st, "fft_r2c(cuFFT)");
  if ((rc = jpm_greens_grad_c64(p, stream, p->spec, p->spec3, 1.0f / (float)p->ncell, r_split,
                                filter_tab, n_tab, filter_kmax)))
    return rc;
  if (p->timer) p->timer->mark(st, "greens_grad");
  rc = jpm_ifft3d_c2r(p, stream, p->spec3, force3, 3);
  if (p->timer) p->timer->mark(st, "ifft_c2r_x3(cuFFT)");
  return rc;
}

extern "C" int32_t jpm_pm_step_f32(jpm_plan* p, void* stream, float* pos, float* vel, float kick_coef,
                                   float drift_coef, int32_t relative) {
  JPM_CHECK_ARG(!(p && p->is_slab), "not available on a multi-GPU slab plan");
  JPM_CHECK_ARG(p && pos && vel, "null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  int32_t rc;
  JPM_CUDA(cudaMemsetAsync(p->density, 0, p->ncell * sizeof(float), s));
  if (relative)
    rc = jpm_cic_paint_dx_f32(stream, p->density, pos, nullptr, 1.0f, p->nx, p->ny, p->nz, 0, 0);
  else
    rc = jpm_cic_paint_f32(stream, p->density, pos, nullptr, 1.0f, p->ncell, p->nx, p->ny, p->nz,
                           p->nx, p->ny, p->nz);
  if (rc) return rc;
  if ((rc = jpm_density_to_force_meshes(p, stream, p->density, p->force3, 0.f, nullptr, 0, 0.f)))
    return rc;
  return jpm_cic_read3_kick_drift_f32(stream, pos, vel, nullptr, p->force3, p->force3 + p->ncell,
                                      p->force3 + 2 * p->ncell, pos, vel, pos, vel, kick_coef,
                                      drift_coef, 1, p->ncell, p->nx, p->ny, p->nz, 0, 0, relative);
}

extern "C" int32_t jpm_pm_step_host_f32(jpm_plan* p, void* stream, float* pos_host, float* vel_host,
                                        float* pos_dev, float* vel_dev, float kick_coef,
                                        float drift_coef, int32_t relative) {
  JPM_CHECK_ARG(!(p && p->is_slab), "not available on a multi-GPU slab plan");
  JPM_CHECK_ARG(p && pos_host && vel_host && pos_dev && vel_dev, "null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  const size_t bytes = (size_t)p->ncell * 3 * sizeof(float);
  JPM_CUDA(cudaMemcpyAsync(pos_dev, pos_host, bytes, cudaMemcpyHostToDevice, s));
  JPM_CUDA(cudaMemcpyAsync(vel_dev, vel_host, bytes, cudaMemcpyHostToDevice, s));
  int32_t rc = jpm_pm_step_f32(p, stream, pos_dev, vel_dev, kick_coef, drift_coef, relative);
  if (rc) return rc;
  JPM_CUDA(cudaMemcpyAsync(pos_host, pos_dev, bytes, cudaMemcpyDeviceToHost, s));
  JPM_CUDA(cudaMemcpyAsync(vel_host, vel_dev, bytes, cudaMemcpyDeviceToHost, s));
  JPM_CUDA(cudaStreamSynchronize(s));
  return JPM_OK;
}

// ---------------------------------------------------------------------------------
// Multi-GPU x-slab plan: the pmfft chain (csrc/pmfft.cu) over peer-mapped memory.
//   reference: the distributed FFT + halo protocol the reference delegates to [ext] jaxdecomp
//   (jaxpm/distributed.py:37-42 pfft3d/pifft3d, :45-85 halo_exchange + slice_unpad) for pdims = (P, 1).
// Rank r owns global x planes [r lx, (r+1) lx) and the particles whose Lagrangian site lies there.  Its
// real meshes carry gx ghost planes per side in x (what the reference calls the halo) and kGhost periodic
// ghost cells in y and z.  Every array of the rank lives in ONE cudaMalloc block (one IPC handle); the
// kernels read the neighbours' density ghosts, scatter the FFT transposes straight into the owning
// rank's buffers and write the neighbours' force ghosts through NVLink loads / stores.
// ---------------------------------------------------------------------------------
namespace jpm {

struct SlabLayout { size_t dens, force, psi, at, b3, t01, flags, total; };

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static SlabLayout slab_layout(const Slab& sl) {
  SlabLayout L;
  size_t o = 0;
  L.dens = o;  o = align_up(o + (size_t)sl.npad * sizeof(float), 1024);
  L.force = o; o = align_up(o + 3 * (size_t)sl.npad * sizeof(float), 1024);
  L.psi = o;   o = align_up(o + (size_t)sl.npad * sizeof(float), 1024);
  L.at = o;    o = align_up(o + (size_t)sl.nx * sl.ly * sl.nzc * sizeof(float2), 1024);
  L.b3 = o;    o = align_up(o + 3 * (size_t)sl.lx * sl.ny * sl.nzc * sizeof(float2), 1024);
  L.t01 = o;   o = align_up(o + (size_t)sl.lx * sl.ny * sl.nzc * sizeof(float4), 1024);
  L.flags = o; o = align_up(o + 256 * sizeof(unsigned), 1024);
  L.total = o;
  return L;
}

static void slab_point(Slab& sl, int r, void* base_v) {
  char* base = (char*)base_v;
  const SlabLayout L = slab_layout(sl);
  // the real arrays carry kGhost spare planes below / above the [lx + 2 gx] planes the FFT kernels index, so
  // that the tile boxes of the particle kernels never start at a negative x coordinate (a TMA reduce-add
  // with a negative outer coordinate raises "illegal instruction" on sm_100, tools/tma_probe.cu)
  const size_t skip = (size_t)sl.G * sl.nyp * sl.nzp;
  sl.dens[r] = (float*)(base + L.dens) + skip;
  sl.force[r] = (float*)(base + L.force) + skip;
  sl.psi[r] = (float*)(base + L.psi) + skip;
  sl.at[r] = (float2*)(base + L.at);
  sl.b3[r] = (float2*)(base + L.b3);
  sl.t01[r] = (float4*)(base + L.t01);
  sl.flags[r] = (unsigned*)(base + L.flags);
}

}  // namespace jpm

extern "C" int32_t jpm_slab_create(jpm_plan** out, int32_t nx, int32_t ny, int32_t nz, int32_t nranks,
                                   int32_t rank, int32_t gx) {
  return jpm_slab_create_ex(out, nx, ny, nz, nranks, 1, rank, gx, 0);
}

extern "C" int32_t jpm_slab_create_ex(jpm_plan** out, int32_t nx, int32_t ny, int32_t nz, int32_t px, int32_t py,
                                      int32_t rank, int32_t gx, int32_t gy) {
  JPM_CHECK_ARG(out, "null plan pointer");
  JPM_CHECK_ARG(px >= 1 && py >= 1, "bad process grid");
  const int nranks = px * py;
  JPM_CHECK_ARG(nranks >= 1 && nranks <= 8 && rank >= 0 && rank < nranks, "bad rank / nranks (1..8 GPUs of one box)");
  JPM_CHECK_ARG(pmfft_shape_ok(nx, ny, nz), "slab plan: mesh sides must be powers of two in [16, 1024]");
  JPM_CHECK_ARG(nx % nranks == 0 && ny % nranks == 0, "slab plan: nx and ny must divide by the rank count");
  const int lx = nx / nranks, ly = ny / nranks;
  const int Lx = nx / px, Ly = ny / py;
  JPM_CHECK_ARG(gx >= 1 && gx <= Lx, "slab plan: ghost width must be in [1, nx / px]");
  if (py > 1) {
    JPM_CHECK_ARG(gy >= 1 && gy <= Ly, "pencil plan: ghost width in y must be in [1, ny / py]");
    JPM_CHECK_ARG(Ly % 16 == 0, "pencil plan: ny / py must be a multiple of 16 (row tiles of the z passes)");
  } else {
    gy = 0;
  }
  jpm_plan* p = new jpm_plan();
  p->is_slab = true;
  // the particle kernels see the rank's local mesh INCLUDING its ghost planes / rows as their logical mesh
  p->nx = Lx + 2 * gx; p->ny = Ly + 2 * gy; p->nz = nz; p->nzh = nz / 2 + 1;
  p->ncell = (long long)p->nx * p->ny * nz;
  p->nspec = 0;
  p->G = kGhost;
  p->nxp = p->nx + 2 * kGhost; p->nyp = p->ny + 2 * kGhost; p->nzp = nz + 2 * kGhost;
  p->npad = (long long)p->nxp * p->nyp * p->nzp;
  if (p->npad >= (1ll << 31)) {
    delete p;
    set_error("slab plan: local padded mesh too large for int32 cell ids");
    return JPM_ERR_INVALID;
  }
  Slab& sl = p->slab;
  memset(&sl, 0, sizeof(sl));
  sl.P = nranks; sl.rank = rank;
  sl.nx = nx; sl.ny = ny; sl.nz = nz;
  sl.lx = lx; sl.ly = ly; sl.gx = gx; sl.G = kGhost;
  sl.px = px; sl.py = py; sl.Lx = Lx; sl.Ly = Ly; sl.gy = gy;
  sl.nxp = p->nx; sl.nyp = p->nyp; sl.nzp = p->nzp; sl.npad = p->npad;   // npad: component stride of force[]
  sl.nzh = p->nzh; sl.nzc = (p->nzh + 7) & ~7;
  std::vector<float> w, a;
  int32_t rc;
  build_tables(nx, nx, w, a);
  if ((rc = upload(&p->wx, w)) || (rc = upload(&p->ax, a))) return rc;
  build_tables(ny, ny, w, a);
  if ((rc = upload(&p->wy, w)) || (rc = upload(&p->ay, a))) return rc;
  build_tables(nz, p->nzh, w, a);
  if ((rc = upload(&p->wz, w)) || (rc = upload(&p->az, a))) return rc;
  const SlabLayout L = slab_layout(sl);
  p->sym_bytes = L.total;
  JPM_CUDA(cudaMalloc(&p->sym_base, L.total));
  JPM_CUDA(cudaMemset(p->sym_base, 0, L.total));
  slab_point(sl, rank, p->sym_base);
  {
    // touched x / y range unknown, ghost width = all of it
    const int init[5] = {0x7fffffff, (int)0x80000000, std::max(gx, gy), 0x7fffffff, (int)0x80000000};
    JPM_CUDA(cudaMemcpy(sl.flags[rank] + kFlagXmin, init, sizeof(init), cudaMemcpyHostToDevice));
  }
  JPM_CUDA(cudaDeviceSynchronize());
  p->density_p = (float*)((char*)p->sym_base + L.dens);
  p->force3_p = (float*)((char*)p->sym_base + L.force);
  *out = p;
  return JPM_OK;
}

extern "C" int32_t jpm_slab_ipc_handle(jpm_plan* p, void* handle_out, int32_t handle_bytes) {
  JPM_CHECK_ARG(p && p->is_slab && handle_out, "not a slab plan");
  JPM_CHECK_ARG(handle_bytes == (int32_t)sizeof(cudaIpcMemHandle_t), "handle buffer must be 64 bytes");
  cudaIpcMemHandle_t h;
  JPM_CUDA(cudaIpcGetMemHandle(&h, p->sym_base));
  memcpy(handle_out, &h, sizeof(h));
  return JPM_OK;
}

static int32_t slab_finish_attach(jpm_plan* p) {
  for (int r = 0; r < p->slab.P; ++r) slab_point(p->slab, r, r == p->slab.rank ? p->sym_base : p->peer_base[r]);
  int32_t rc = pmfft_setup(p);
  if (rc) return rc;
  p->fft_on = true;
  return JPM_OK;
}

extern "C" int32_t jpm_slab_attach_ipc(jpm_plan* p, const void* handles, int32_t nranks) {
  JPM_CHECK_ARG(p && p->is_slab && handles, "not a slab plan");
  JPM_CHECK_ARG(nranks == p->slab.P, "handle count != rank count of the plan");
  JPM_CHECK_ARG(!p->fft_on, "slab plan already attached");
  const cudaIpcMemHandle_t* h = (const cudaIpcMemHandle_t*)handles;
  for (int r = 0; r < nranks; ++r) {
    if (r == p->slab.rank) continue;
    cudaIpcMemHandle_t hr;
    memcpy(&hr, h + r, sizeof(hr));
    JPM_CUDA(cudaIpcOpenMemHandle(&p->peer_base[r], hr, cudaIpcMemLazyEnablePeerAccess));
  }
  p->ipc_peers = true;
  return slab_finish_attach(p);
}

extern "C" int32_t jpm_slab_attach_ptrs(jpm_plan* p, void* const* bases, int32_t nranks) {
  JPM_CHECK_ARG(p && p->is_slab && bases, "not a slab plan");
  JPM_CHECK_ARG(nranks == p->slab.P, "pointer count != rank count of the plan");
  JPM_CHECK_ARG(!p->fft_on, "slab plan already attached");
  for (int r = 0; r < nranks; ++r) {
    if (r == p->slab.rank) continue;
    JPM_CHECK_ARG(bases[r], "null peer base");
    p->peer_base[r] = bases[r];
  }
  p->ipc_peers = false;
  return slab_finish_attach(p);
}

extern "C" int32_t jpm_slab_base(jpm_plan* p, void** base_out, int64_t* bytes_out) {
  JPM_CHECK_ARG(p && p->is_slab && base_out, "not a slab plan");
  *base_out = p->sym_base;
  if (bytes_out) *bytes_out = (int64_t)p->sym_bytes;
  return JPM_OK;
}

// which = 0: density, 1..3: force component; interior [lx][ny][nz] of this rank <-> compact array
extern "C" int32_t jpm_slab_get_interior_f32(jpm_plan* p, void* stream, int32_t which, float* dst) {
  JPM_CHECK_ARG(p && p->is_slab && dst && which >= 0 && which <= 3, "bad arguments");
  const Slab& sl = p->slab;
  float* src = (which == 0 ? p->density_p : p->force3_p + (long long)(which - 1) * p->npad) +
               (long long)sl.gx * sl.nyp * sl.nzp + (long long)sl.gy * sl.nzp;   // pad_copy skips the kGhost spare planes / rows itself
  const long long total = (long long)sl.Lx * sl.Ly * sl.nz / 4;
  const unsigned blocks = (unsigned)std::min<long long>((total + 255) / 256, kNumSMs * 16);
  pad_copy_kernel<false><<<dim3(blocks, 1), 256, 0, (cudaStream_t)stream>>>(src, dst, sl.Lx, sl.Ly, sl.nz / 4, sl.nyp,
                                                                          sl.nzp, sl.G, 0, 0);
  JPM_LAUNCH_CHECK();
  return JPM_OK;
}

extern "C" int32_t jpm_slab_set_density_f32(jpm_plan* p, void* stream, const float* src) {
  JPM_CHECK_ARG(p && p->is_slab && src, "bad arguments");
  const Slab& sl = p->slab;
  cudaStream_t st = (cudaStream_t)stream;
  JPM_CUDA(cudaMemsetAsync(p->density_p, 0, p->npad * sizeof(float), st));
  JPM_CUDA(cudaMemsetAsync(sl.flags[sl.rank] + kFlagXmin, 0x7f, sizeof(int), st));   // 0x7f7f7f7f > any plane: unknown
  JPM_CUDA(cudaMemsetAsync(sl.flags[sl.rank] + kFlagXmax, 0x80, sizeof(int), st));   // 0x80808080 < 0
  JPM_CUDA(cudaMemsetAsync(sl.flags[sl.rank] + kFlagYmin, 0x7f, sizeof(int), st));
  JPM_CUDA(cudaMemsetAsync(sl.flags[sl.rank] + kFlagYmax, 0x80, sizeof(int), st));
  float* dstp = p->density_p + (long long)sl.gx * sl.nyp * sl.nzp + (long long)sl.gy * sl.nzp;
  const long long total = (long long)sl.Lx * sl.Ly * sl.nz / 4;
  const unsigned blocks = (unsigned)std::min<long long>((total + 255) / 256, kNumSMs * 16);
  pad_copy_kernel<true><<<dim3(blocks, 1), 256, 0, st>>>(dstp, const_cast<float*>(src), sl.Lx, sl.Ly, sl.nz / 4, sl.nyp,
                                                        sl.nzp, sl.G, 0, 0);
  JPM_LAUNCH_CHECK();
  return JPM_OK;
}

// density (painted or set, ghosts not folded) -> force meshes with ghosts filled; collective over the ranks
extern "C" int32_t jpm_slab_forces(jpm_plan* p, void* stream, float r_split) {
  JPM_CHECK_ARG(p && p->is_slab && p->fft_on, "slab plan not attached");
  return pmfft_forces(p, (cudaStream_t)stream, r_split, nullptr, 0, 0.f);
}

// Synchronises `stream`; fails if a flag barrier of this rank timed out (a peer died or never arrived).
extern "C" int32_t jpm_slab_check(jpm_plan* p, void* stream) {
  JPM_CHECK_ARG(p && p->is_slab, "not a slab plan");
  unsigned err = 0;
  JPM_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  JPM_CUDA(cudaMemcpy(&err, p->slab.flags[p->slab.rank] + kFlagErr, sizeof(err), cudaMemcpyDeviceToHost));
  if (err) {
    set_error("slab barrier timed out on rank %d (epoch %u): a peer did not arrive", p->slab.rank, p->epoch);
    return JPM_ERR_CUDA;
  }
  return JPM_OK;
}

// Ghost planes per side the last force evaluation used (<= gx; adaptive when the density was painted by a
// jpm_sim).  Synchronises `stream`.
extern "C" int32_t jpm_slab_ghost_width(jpm_plan* p, void* stream, int32_t* out) {
  JPM_CHECK_ARG(p && p->is_slab && out, "not a slab plan");
  int v = 0;
  JPM_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  JPM_CUDA(cudaMemcpy(&v, p->slab.flags[p->slab.rank] + kFlagGe, sizeof(v), cudaMemcpyDeviceToHost));
  *out = p->slab.P == 1 ? p->slab.gx : v;
  return JPM_OK;
}

// 1 if, since the plan was created, some step saw particles on the outermost ghost plane of this rank: the
// halo (gx) is too small for the displacement field and - exactly like in the reference, whose halo_size has
// the same role (painting.py:192-215) - those particles were painted / read at wrapped positions.
extern "C" int32_t jpm_slab_halo_exceeded(jpm_plan* p, void* stream, int32_t* out) {
  JPM_CHECK_ARG(p && p->is_slab && out, "not a slab plan");
  unsigned v = 0;
  JPM_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  JPM_CUDA(cudaMemcpy(&v, p->slab.flags[p->slab.rank] + kFlagReach, sizeof(v), cudaMemcpyDeviceToHost));
  *out = (int32_t)v;
  return JPM_OK;
}

// Debug / test access to the ghost-zone meshes of a plan: which = 0 the painted density (ghosts NOT folded),
// 1..3 a force component (ghosts filled); dst receives the whole padded array [nxp][nyp][nzp].
extern "C" int32_t jpm_plan_padded_get_f32(jpm_plan* p, void* stream, int32_t which, float* dst, int32_t* dims3) {
  JPM_CHECK_ARG(p && which >= 0 && which <= 3, "bad arguments");
  JPM_CHECK_ARG(p->G > 0 && p->density_p, "plan has no ghost-zone meshes (no resident sim stepped on it yet)");
  if (dims3) { dims3[0] = p->nxp; dims3[1] = p->nyp; dims3[2] = p->nzp; }
  if (!dst) return JPM_OK;
  const float* src = which == 0 ? p->density_p : p->force3_p + (long long)(which - 1) * p->npad;
  JPM_CUDA(cudaMemcpyAsync(dst, src, p->npad * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return JPM_OK;
}
