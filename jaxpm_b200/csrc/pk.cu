// Power-spectrum estimator on the R2C half-spectrum, and its adjoint.
//   reference: jaxpm/utils.py:14-73 (_initialize_pk: |k| mesh, np.digitize on kedges, kcount, kavg, mu mesh),
//              :76-128 (power_spectrum: |FFT_ortho|^2 (auto) or a b* (cross), (2l+1) L_l(mu) weights, bincount,
//              / kcount, [1:-1], * cell volume).
// One pass over the half-spectrum: modes with 0 < kz < nz/2 stand for themselves and their conjugate partner
// (weight 2: |delta|^2, the bin and L_l for even l are the same for k and -k; the imaginary part of a cross
// spectrum cancels between the two), the kz = 0 and Nyquist planes hold both partners (weight 1).
// Bin rule, bit for bit: k^2 = ((kx^2) + ky^2) + kz^2 in float64 from the host's float64 per-axis tables,
// |k| = float32(sqrt(k^2)) (utils.py:54 is a jnp.sqrt with x64 off), bin = #edges <= |k| compared in float64.
// Threads run along kz, where the bin is non-decreasing: a segmented warp reduction leaves one atomic per
// (warp, bin) instead of one per mode; per-CTA partial sums live in shared memory as float64.
#include <algorithm>

#include "common.cuh"

namespace jpm {

constexpr int kPkMaxBins = 1024;   // bins incl. the two overflow bins (len(kedges) + 1)
constexpr int kPkMaxEll = 3;

struct PkGeom {
  int nx, ny, nz, nzh;
  int nb;                 // len(kedges) + 1
  int nl;                 // number of multipoles
  int ell[kPkMaxEll];
  float lx, ly, lz;       // unit line of sight (ignored when all ell == 0)
  int use_mu;
};

__device__ __forceinline__ float legendre_even(int ell, float mu) {
  const float m2 = mu * mu;
  if (ell == 0) return 1.0f;
  if (ell == 2) return 1.5f * m2 - 0.5f;
  return (35.0f * m2 * m2 - 30.0f * m2 + 3.0f) * 0.125f;   // ell == 4
}

__device__ __forceinline__ int digitize(const double* __restrict__ edges, int ne, double x) {
  int lo = 0, hi = ne;          // first index with edges[i] > x  ==  number of edges <= x
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (edges[mid] <= x) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// value of lane `lane` += values of the following lanes of the same bin (bins are contiguous runs)
__device__ __forceinline__ double seg_reduce(double v, int bin, int lane) {
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const double o = __shfl_down_sync(0xffffffffu, v, off);
    const int b = __shfl_down_sync(0xffffffffu, bin, off);
    if (lane + off < 32 && b == bin) v += o;
  }
  return v;
}

// out layout (float64): [nl] re sums | [nl] im sums | count | sum of |k|, each [nb]
template <bool CROSS>
__global__ void __launch_bounds__(256)
pk_bin_kernel(PkGeom g, const float2* __restrict__ a, const float2* __restrict__ b, const double* __restrict__ kx,
              const double* __restrict__ ky, const double* __restrict__ kz, const double* __restrict__ edges,
              float norm, int want_counts, double* __restrict__ out) {
  extern __shared__ double sacc[];              // [(2 nl + 2)][nb]
  const int nb = g.nb, nrow = 2 * g.nl + 2;
  for (int i = threadIdx.x; i < nrow * nb; i += blockDim.x) sacc[i] = 0.0;
  __shared__ double sedges[kPkMaxBins];
  for (int i = threadIdx.x; i < nb - 1; i += blockDim.x) sedges[i] = edges[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long nrows = (long long)g.nx * g.ny;
  const int nzi = (g.nzh + 31) & ~31;           // warp-uniform trip count along kz
  for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
    const int ix = (int)(row / g.ny), iy = (int)(row % g.ny);
    const double kxv = kx[ix], kyv = ky[iy];
    const double kxy2 = kxv * kxv + kyv * kyv;   // (0 + kx^2) + ky^2
    const double dxy = kxv * (double)g.lx + kyv * (double)g.ly;
    for (int iz = threadIdx.x; iz < nzi; iz += blockDim.x) {
      const bool valid = iz < g.nzh;
      int bin = -1 - lane;                       // distinct per lane: never merges
      double w = 0.0, kabs = 0.0;
      double pre[kPkMaxEll], pim[kPkMaxEll];
#pragma unroll
      for (int l = 0; l < kPkMaxEll; ++l) { pre[l] = 0.0; pim[l] = 0.0; }
      if (valid) {
        const double kzv = kz[iz];
        const float kf = (float)sqrt(kxy2 + kzv * kzv);
        bin = digitize(sedges, nb - 1, (double)kf);
        w = (iz == 0 || (2 * iz == g.nz)) ? 1.0 : 2.0;
        kabs = (double)kf;
        const long long o = row * g.nzh + iz;
        const float2 va = a[o];
        float re, im;
        if (CROSS) {
          const float2 vb = b[o];                // a conj(b)
          re = va.x * vb.x + va.y * vb.y;
          im = va.y * vb.x - va.x * vb.y;
        } else {
          re = va.x * va.x + va.y * va.y;
          im = 0.f;
        }
        re *= norm; im *= norm;
        float mu = 0.f;
        if (g.use_mu) mu = (kf == 0.f) ? 0.f : (float)(dxy + kzv * (double)g.lz) / kf;
#pragma unroll
        for (int l = 0; l < kPkMaxEll; ++l) {
          if (l < g.nl) {
            const float wl = (float)(2 * g.ell[l] + 1) * legendre_even(g.ell[l], mu);
            pre[l] = w * (double)(re * wl);
            // the partner -k contributes conj: imaginary parts cancel for interior kz (weight 2 -> 0)
            pim[l] = (w == 1.0) ? (double)(im * wl) : 0.0;
          }
        }
      }
      const int prev = __shfl_up_sync(0xffffffffu, bin, 1);
      const bool head = valid && (lane == 0 || prev != bin);
#pragma unroll
      for (int l = 0; l < kPkMaxEll; ++l) {
        if (l < g.nl) {
          const double r = seg_reduce(pre[l], bin, lane);
          if (head) atomicAdd(&sacc[l * nb + bin], r);
          if (CROSS) {
            const double q = seg_reduce(pim[l], bin, lane);
            if (head) atomicAdd(&sacc[(g.nl + l) * nb + bin], q);
          }
        }
      }
      if (want_counts) {
        const double c = seg_reduce(w, bin, lane);
        const double ks = seg_reduce(w * kabs, bin, lane);
        if (head) {
          atomicAdd(&sacc[(2 * g.nl) * nb + bin], c);
          atomicAdd(&sacc[(2 * g.nl + 1) * nb + bin], ks);
        }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nrow * nb; i += blockDim.x)
    if (sacc[i] != 0.0) atomicAdd(out + i, sacc[i]);
}

// adjoint of the auto spectrum: out_k = delta_k * 2 * norm * sum_l W[l][bin(k)] (2l+1) L_l(mu_k); a C2R of `out`
// is d(sum_lb g_lb pk_l[b]) / d mesh when W[l][b] = g_lb * cell_volume / kcount[b].
__global__ void __launch_bounds__(256)
pk_weight_kernel(PkGeom g, const float2* __restrict__ a, const double* __restrict__ kx, const double* __restrict__ ky,
                 const double* __restrict__ kz, const double* __restrict__ edges, const double* __restrict__ wbin,
                 float norm, float2* __restrict__ out) {
  __shared__ double sedges[kPkMaxBins];
  for (int i = threadIdx.x; i < g.nb - 1; i += blockDim.x) sedges[i] = edges[i];
  __syncthreads();
  const long long nrows = (long long)g.nx * g.ny;
  for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
    const int ix = (int)(row / g.ny), iy = (int)(row % g.ny);
    const double kxv = kx[ix], kyv = ky[iy];
    const double kxy2 = kxv * kxv + kyv * kyv;
    const double dxy = kxv * (double)g.lx + kyv * (double)g.ly;
    for (int iz = threadIdx.x; iz < g.nzh; iz += blockDim.x) {
      const double kzv = kz[iz];
      const float kf = (float)sqrt(kxy2 + kzv * kzv);
      const int bin = digitize(sedges, g.nb - 1, (double)kf);
      float mu = 0.f;
      if (g.use_mu) mu = (kf == 0.f) ? 0.f : (float)(dxy + kzv * (double)g.lz) / kf;
      double wsum = 0.0;
      for (int l = 0; l < g.nl; ++l)
        wsum += wbin[l * g.nb + bin] * (double)((float)(2 * g.ell[l] + 1) * legendre_even(g.ell[l], mu));
      const float f = (float)(2.0 * (double)norm * wsum);
      const long long o = row * g.nzh + iz;
      const float2 v = a[o];
      out[o] = make_float2(f * v.x, f * v.y);
    }
  }
}

static int32_t pk_geom(PkGeom& g, int nx, int ny, int nz, int n_edges, const int32_t* ells, int n_ell, const float* los) {
  JPM_CHECK_ARG(nx > 0 && ny > 0 && nz > 0, "bad mesh shape");
  JPM_CHECK_ARG(n_edges >= 1 && n_edges + 1 <= kPkMaxBins, "number of k edges must be in [1, 1023]");
  JPM_CHECK_ARG(n_ell >= 1 && n_ell <= kPkMaxEll && ells, "1 to 3 multipoles");
  g.nx = nx; g.ny = ny; g.nz = nz; g.nzh = nz / 2 + 1;
  g.nb = n_edges + 1; g.nl = n_ell;
  g.use_mu = 0;
  for (int l = 0; l < kPkMaxEll; ++l) g.ell[l] = 0;
  for (int l = 0; l < n_ell; ++l) {
    JPM_CHECK_ARG(ells[l] == 0 || ells[l] == 2 || ells[l] == 4, "multipoles must be 0, 2 or 4 (odd ones vanish for a real field)");
    g.ell[l] = ells[l];
    if (ells[l] != 0) g.use_mu = 1;
  }
  g.lx = los ? los[0] : 0.f; g.ly = los ? los[1] : 0.f; g.lz = los ? los[2] : 1.f;
  return JPM_OK;
}

}  // namespace jpm

using namespace jpm;

extern "C" int32_t jpm_pk_bin_c64(void* stream, const void* spec_a, const void* spec_b, int32_t nx, int32_t ny,
                                  int32_t nz, const double* kx, const double* ky, const double* kz,
                                  const double* kedges, int32_t n_edges, const int32_t* ells, int32_t n_ell,
                                  const float* los3, float norm, int32_t want_counts, double* out) {
  JPM_CHECK_ARG(spec_a && kx && ky && kz && kedges && out, "null pointer");
  PkGeom g;
  int32_t rc = pk_geom(g, nx, ny, nz, n_edges, ells, n_ell, los3);
  if (rc) return rc;
  const size_t smem = (size_t)(2 * g.nl + 2) * g.nb * sizeof(double);
  JPM_CHECK_ARG(smem <= 36 * 1024, "too many bins x multipoles for the shared-memory partial sums");
  const long long rows = (long long)nx * ny;
  const int blocks = (int)std::min<long long>(rows, (long long)kNumSMs * 4);
  cudaStream_t st = (cudaStream_t)stream;
  JPM_CUDA(cudaMemsetAsync(out, 0, (size_t)(2 * g.nl + 2) * g.nb * sizeof(double), st));
  if (spec_b)
    pk_bin_kernel<true><<<blocks, 256, smem, st>>>(g, (const float2*)spec_a, (const float2*)spec_b, kx, ky, kz, kedges,
                                                   norm, want_counts, out);
  else
    pk_bin_kernel<false><<<blocks, 256, smem, st>>>(g, (const float2*)spec_a, nullptr, kx, ky, kz, kedges, norm,
                                                    want_counts, out);
  JPM_LAUNCH_CHECK();
  return JPM_OK;
}

extern "C" int32_t jpm_pk_weight_c64(void* stream, const void* spec_a, void* out, int32_t nx, int32_t ny, int32_t nz,
                                     const double* kx, const double* ky, const double* kz, const double* kedges,
                                     int32_t n_edges, const int32_t* ells, int32_t n_ell, const float* los3,
                                     const double* wbin, float norm) {
  JPM_CHECK_ARG(spec_a && out && kx && ky && kz && kedges && wbin, "null pointer");
  PkGeom g;
  int32_t rc = pk_geom(g, nx, ny, nz, n_edges, ells, n_ell, los3);
  if (rc) return rc;
  const long long rows = (long long)nx * ny;
  const int blocks = (int)std::min<long long>(rows, (long long)kNumSMs * 8);
  pk_weight_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(g, (const float2*)spec_a, kx, ky, kz, kedges, wbin, norm,
                                                             (float2*)out);
  JPM_LAUNCH_CHECK();
  return JPM_OK;
}

// spectrum times a separable real filter: out_k = in_k * norm * tx[ix] ty[iy] tz[iz]  (one pass; in place allowed).
//   used by compensate_cic (jaxpm/painting.py:263-275 with kernels.py:118-136: the filter of that function is
//   prod_d sinc(k_d / 2 pi)^-2, a product of per-axis factors)
namespace jpm {
__global__ void __launch_bounds__(256)
kseparable_kernel(const float2* __restrict__ in, float2* __restrict__ out, const float* __restrict__ tx,
                  const float* __restrict__ ty, const float* __restrict__ tz, int nx, int ny, int nzh, float norm) {
  const long long nrows = (long long)nx * ny;
  for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
    const float fxy = norm * (tx[row / ny] * ty[row % ny]);
    for (int iz = threadIdx.x; iz < nzh; iz += blockDim.x) {
      const long long o = row * nzh + iz;
      const float f = fxy * tz[iz];
      const float2 v = in[o];
      out[o] = make_float2(f * v.x, f * v.y);
    }
  }
}
}  // namespace jpm

extern "C" int32_t jpm_kseparable_c64(void* stream, const void* in, void* out, const float* tx, const float* ty,
                                      const float* tz, int32_t nx, int32_t ny, int32_t nz, float norm) {
  JPM_CHECK_ARG(in && out && tx && ty && tz && nx > 0 && ny > 0 && nz > 0, "bad arguments");
  const long long rows = (long long)nx * ny;
  const int blocks = (int)std::min<long long>(rows, (long long)jpm::kNumSMs * 8);
  jpm::kseparable_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const float2*)in, (float2*)out, tx, ty, tz, nx, ny,
                                                                  nz / 2 + 1, norm);
  JPM_LAUNCH_CHECK();
  return JPM_OK;
}
