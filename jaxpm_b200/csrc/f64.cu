// float64 variants of the four CIC primitives ("_f64 variants where x64 parity is wanted" - the
// reference's distributed tests run with jax_enable_x64, tests/test_distributed_pm.py:30, where positions, weights and
// meshes are float64 and the index / weight rules below run in double).
//   reference: jaxpm/painting.py:15-45 (_cic_paint_impl), :78-106 (_cic_read_impl), :161-189 / :218-236 (the _dx forms),
//              jaxpm/painting_utils.py:28-131, :144-187 (enmesh / scatter / gather, relative rule)
// One particle per thread, 8 double atomics (RED.E.ADD.F64, native on sm_60+) / 8 gathers; the fp32 hot path
// (csrc/sim.cu, csrc/pmfft.cu) is untouched.  Same rules as common.cuh, in double.
#include "common.cuh"

namespace jpm {

struct Cic1D {
  int i0, i1;       // wrapped cell indices; -1 == dropped by the reference (relative rule only)
  double w0, w1;    // 1 - |x - corner|
};

// absolute rule, painting.py:22-37
__device__ __forceinline__ Cic1D cic_abs_d(double p, int n) {
  Cic1D c;
  const double f = floor(p);
  c.w0 = 1.0 - fabs(p - f);
  c.w1 = 1.0 - fabs(p - (f + 1.0));
  c.i0 = pymod((int)f, n);
  c.i1 = pymod((int)(f + 1.0), n);
  return c;
}

// relative rule, painting_utils.py:48-65 (cell_size 1, offset 0)
__device__ __forceinline__ void cic_rel_corner_d(double pp, double o, double L, int n, int& idx, double& w) {
  const double x = pp + o;
  double r = x;
  if (!(x >= 0.0 && x < L)) {
    r = fmod(x, L);
    if (r != 0.0 && r < 0.0) r = r + L;
  }
  const double fi = floor(r);
  double nd = pp - fi;
  nd = nd - rint(nd / L) * L;
  w = 1.0 - fabs(nd);
  const int i = (int)fi;
  idx = (i >= 0 && i < n) ? i : -1;
}

__device__ __forceinline__ Cic1D cic_rel_d(int base, double d, int n) {
  Cic1D c;
  const double pp = (double)base + d, L = (double)n;
  cic_rel_corner_d(pp, 0.0, L, n, c.i0, c.w0);
  cic_rel_corner_d(pp, 1.0, L, n, c.i1, c.w1);
  return c;
}

template <bool REL>
__device__ __forceinline__ void stencil_d(long long p, const double* __restrict__ pos, int nx, int ny, int nz, int pny,
                                          int pnz, int hx, int hy, Cic1D& cx, Cic1D& cy, Cic1D& cz) {
  const double px = pos[3 * p], py = pos[3 * p + 1], pz = pos[3 * p + 2];
  if (REL) {
    const int bk = (int)(p % pnz);
    const long long t = p / pnz;
    cx = cic_rel_d((int)(t / pny) + hx, px, nx);
    cy = cic_rel_d((int)(t % pny) + hy, py, ny);
    cz = cic_rel_d(bk, pz, nz);
  } else {
    cx = cic_abs_d(px, nx);
    cy = cic_abs_d(py, ny);
    cz = cic_abs_d(pz, nz);
  }
}

template <bool REL>
__global__ void __launch_bounds__(256)
paint_f64_kernel(double* __restrict__ mesh, const double* __restrict__ pos, const double* __restrict__ weight,
                 double wscalar, long long np, int nx, int ny, int nz, int pny, int pnz, int hx, int hy) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < np; p += stride) {
    Cic1D cx, cy, cz;
    stencil_d<REL>(p, pos, nx, ny, nz, pny, pnz, hx, hy, cx, cy, cz);
    const double w = weight ? weight[p] : wscalar;
    const int ix[2] = {cx.i0, cx.i1}, iy[2] = {cy.i0, cy.i1}, iz[2] = {cz.i0, cz.i1};
    const double wx[2] = {cx.w0, cx.w1}, wy[2] = {cy.w0, cy.w1}, wz[2] = {cz.w0, cz.w1};
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          if (REL && (ix[a] < 0 || iy[b] < 0 || iz[c] < 0)) continue;
          // reference order: (kx * ky) * kz, then * weight (painting.py:29-33)
          atomicAdd(mesh + ((long long)ix[a] * ny + iy[b]) * nz + iz[c], w * ((wx[a] * wy[b]) * wz[c]));
        }
  }
}

template <bool REL>
__global__ void __launch_bounds__(256)
read_f64_kernel(double* __restrict__ out, const double* __restrict__ mesh, const double* __restrict__ pos, long long np,
                int nx, int ny, int nz, int pny, int pnz, int hx, int hy) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < np; p += stride) {
    Cic1D cx, cy, cz;
    stencil_d<REL>(p, pos, nx, ny, nz, pny, pnz, hx, hy, cx, cy, cz);
    const int ix[2] = {cx.i0, cx.i1}, iy[2] = {cy.i0, cy.i1}, iz[2] = {cz.i0, cz.i1};
    const double wx[2] = {cx.w0, cx.w1}, wy[2] = {cy.w0, cy.w1}, wz[2] = {cz.w0, cz.w1};
    double acc = 0.0;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          if (REL && (ix[a] < 0 || iy[b] < 0 || iz[c] < 0)) continue;
          acc += mesh[((long long)ix[a] * ny + iy[b]) * nz + iz[c]] * ((wx[a] * wy[b]) * wz[c]);
        }
    out[p] = acc;
  }
}

static int grid_f64(long long np) {
  const long long blocks = (np + 255) / 256, cap = (long long)kNumSMs * 32;
  return (int)(blocks > cap ? cap : blocks);
}

}  // namespace jpm

using namespace jpm;

#define JPM_CHECK_MESH64(nx, ny, nz)                                           \
  JPM_CHECK_ARG((nx) > 0 && (ny) > 0 && (nz) > 0, "bad mesh shape");           \
  JPM_CHECK_ARG((int64_t)(nx) * (ny) * (nz) < (1ll << 31), "mesh too large for int32 cell ids")

extern "C" int32_t jpm_cic_paint_f64(void* stream, double* mesh, const double* positions, const double* weight,
                                     double weight_scalar, int64_t np, int32_t nx, int32_t ny, int32_t nz) {
  JPM_CHECK_ARG(mesh && np >= 0 && (np == 0 || positions), "null pointer");
  JPM_CHECK_MESH64(nx, ny, nz);
  if (np == 0) return JPM_OK;
  paint_f64_kernel<false><<<grid_f64(np), 256, 0, (cudaStream_t)stream>>>(mesh, positions, weight, weight_scalar, np, nx,
                                                                         ny, nz, ny, nz, 0, 0);
  JPM_LAUNCH_CHECK();
  return JPM_OK;
}

extern "C" int32_t jpm_cic_paint_dx_f64(void* stream, double* mesh, const double* disp, const double* weight,
                                        double weight_scalar, int32_t nx, int32_t ny, int32_t nz, int32_t hx,
                                        int32_t hy) {
  JPM_CHECK_ARG(mesh && disp && hx >= 0 && hy >= 0, "null pointer / bad halo");
  const int mx = nx + 2 * hx, my = ny + 2 * hy;
  JPM_CHECK_MESH64(mx, my, nz);
  const long long np = (long long)nx * ny * nz;
  paint_f64_kernel<true><<<grid_f64(np), 256, 0, (cudaStream_t)stream>>>(mesh, disp, weight, weight_scalar, np, mx, my, nz,
                                                                        ny, nz, hx, hy);
  JPM_LAUNCH_CHECK();
  return JPM_OK;
}

extern "C" int32_t jpm_cic_read_f64(void* stream, double* out, const double* mesh, const double* positions, int64_t np,
                                    int32_t nx, int32_t ny, int32_t nz) {
  JPM_CHECK_ARG(np >= 0 && mesh && (np == 0 || (out && positions)), "null pointer");
  JPM_CHECK_MESH64(nx, ny, nz);
  if (np == 0) return JPM_OK;
  read_f64_kernel<false><<<grid_f64(np), 256, 0, (cudaStream_t)stream>>>(out, mesh, positions, np, nx, ny, nz, ny, nz, 0, 0);
  JPM_LAUNCH_CHECK();
  return JPM_OK;
}

extern "C" int32_t jpm_cic_read_dx_f64(void* stream, double* out, const double* mesh, const double* disp, int32_t nx,
                                       int32_t ny, int32_t nz, int32_t hx, int32_t hy) {
  JPM_CHECK_ARG(out && mesh && disp && hx >= 0 && hy >= 0, "null pointer / bad halo");
  const int mx = nx + 2 * hx, my = ny + 2 * hy;
  JPM_CHECK_MESH64(mx, my, nz);
  const long long np = (long long)nx * ny * nz;
  read_f64_kernel<true><<<grid_f64(np), 256, 0, (cudaStream_t)stream>>>(out, mesh, disp, np, mx, my, nz, ny, nz, hx, hy);
  JPM_LAUNCH_CHECK();
  return JPM_OK;
}
