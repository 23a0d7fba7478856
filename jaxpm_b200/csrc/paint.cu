// K1 — CIC paint (scatter-add), absolute and relative rules.
//   reference: jaxpm/painting.py:15-45 (_cic_paint_impl), :161-189 (_cic_paint_dx_impl),
//              jaxpm/painting_utils.py:28-112 (enmesh + _scatter_chunk)
#include "common.cuh"

namespace jpm {

// ---------------------------------------------------------------------------------
// Direct path: one particle per thread, 8 REDG.E.ADD.F32 into the global mesh.  Correct
// for any particle order; used for small meshes and as the fallback of the tiled path.
// ---------------------------------------------------------------------------------
template <bool REL>
__global__ void __launch_bounds__(256)
paint_direct_kernel(float* __restrict__ mesh, const float* __restrict__ pos,
                    const float* __restrict__ weight, float wscalar, long long np, int nx, int ny,
                    int nz, int pnx, int pny, int pnz, int hx, int hy) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < np; p += stride) {
    const float px = ld_stream(pos + 3 * p + 0);
    const float py = ld_stream(pos + 3 * p + 1);
    const float pz = ld_stream(pos + 3 * p + 2);
    int bi = 0, bj = 0, bk = 0;
    if (REL) {
      bk = (int)(p % pnz);
      const long long t = p / pnz;
      bj = (int)(t % pny) + hy;
      bi = (int)(t / pny) + hx;
    }
    const Cic1 cx = cic_1d<REL, false>(bi, px, nx);
    const Cic1 cy = cic_1d<REL, false>(bj, py, ny);
    const Cic1 cz = cic_1d<REL, false>(bk, pz, nz);
    const float w = weight ? weight[p] : wscalar;
    const int ix[2] = {cx.i0, cx.i1}, iy[2] = {cy.i0, cy.i1}, iz[2] = {cz.i0, cz.i1};
    const float wx[2] = {cx.w0, cx.w1}, wy[2] = {cy.w0, cy.w1}, wz[2] = {cz.w0, cz.w1};
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          if (REL && (ix[a] < 0 || iy[b] < 0 || iz[c] < 0)) continue;
          // reference order: (kx*ky)*kz, then * weight (painting.py:29-33)
          const float k = w * ((wx[a] * wy[b]) * wz[c]);
          atomicAdd(mesh + ((long long)ix[a] * ny + iy[b]) * nz + iz[c], k);
        }
  }
}

// Three weighted paints in one pass over the particles (the read adjoint of pm.py:54-56: G_d = paint(weight = u_d)):
// mesh3[d][c] += u[p][d] * K(x_p, c), stencil and weights computed once, 24 REDG.E.ADD.F32.
template <bool REL>
__global__ void __launch_bounds__(256)
paint3_kernel(float* __restrict__ mesh3, long long mstride, const float* __restrict__ pos, const float* __restrict__ u,
              float scale, long long np, int nx, int ny, int nz, int pny, int pnz, int hx, int hy) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < np; p += stride) {
    const float px = ld_stream(pos + 3 * p + 0);
    const float py = ld_stream(pos + 3 * p + 1);
    const float pz = ld_stream(pos + 3 * p + 2);
    const float u0 = scale * ld_stream(u + 3 * p + 0), u1 = scale * ld_stream(u + 3 * p + 1),
                u2 = scale * ld_stream(u + 3 * p + 2);
    int bi = 0, bj = 0, bk = 0;
    if (REL) {
      bk = (int)(p % pnz);
      const long long t = p / pnz;
      bj = (int)(t % pny) + hy;
      bi = (int)(t / pny) + hx;
    }
    const Cic1 cx = cic_1d<REL, false>(bi, px, nx);
    const Cic1 cy = cic_1d<REL, false>(bj, py, ny);
    const Cic1 cz = cic_1d<REL, false>(bk, pz, nz);
    const int ix[2] = {cx.i0, cx.i1}, iy[2] = {cy.i0, cy.i1}, iz[2] = {cz.i0, cz.i1};
    const float wx[2] = {cx.w0, cx.w1}, wy[2] = {cy.w0, cy.w1}, wz[2] = {cz.w0, cz.w1};
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          if (REL && (ix[a] < 0 || iy[b] < 0 || iz[c] < 0)) continue;
          const float k = (wx[a] * wy[b]) * wz[c];
          float* m = mesh3 + ((long long)ix[a] * ny + iy[b]) * nz + iz[c];
          atomicAdd(m, u0 * k);
          atomicAdd(m + mstride, u1 * k);
          atomicAdd(m + 2 * mstride, u2 * k);
        }
  }
}

// Forward mode of the paint with respect to the positions (what jax.jvp / jacfwd of painting.py:15-45 produces,
// the transpose of readgrad_kernel): mesh[c] += w_p * sum_d v_{p,d} * dK/dx_d(p, c), dK/dx_d = -sign(x_d - c_d)
// prod_{e != d} (1 - |x_e - c_e|), sign(0) = 0.  One particle per thread, 8 REDG.E.ADD.F32.
template <bool REL>
__global__ void __launch_bounds__(256)
paintgrad_kernel(float* __restrict__ mesh, const float* __restrict__ pos, const float* __restrict__ tangent,
                 const float* __restrict__ weight, float wscalar, long long np, int nx, int ny, int nz, int pny,
                 int pnz, int hx, int hy) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < np; p += stride) {
    int bi = 0, bj = 0, bk = 0;
    if (REL) {
      bk = (int)(p % pnz);
      const long long t = p / pnz;
      bj = (int)(t % pny) + hy;
      bi = (int)(t / pny) + hx;
    }
    const Cic1 cx = cic_1d<REL, true>(bi, ld_stream(pos + 3 * p + 0), nx);
    const Cic1 cy = cic_1d<REL, true>(bj, ld_stream(pos + 3 * p + 1), ny);
    const Cic1 cz = cic_1d<REL, true>(bk, ld_stream(pos + 3 * p + 2), nz);
    const float w = weight ? weight[p] : wscalar;
    const float vx = w * ld_stream(tangent + 3 * p + 0), vy = w * ld_stream(tangent + 3 * p + 1),
                vz = w * ld_stream(tangent + 3 * p + 2);
    const int ix[2] = {cx.i0, cx.i1}, iy[2] = {cy.i0, cy.i1}, iz[2] = {cz.i0, cz.i1};
    const float wx[2] = {cx.w0, cx.w1}, wy[2] = {cy.w0, cy.w1}, wz[2] = {cz.w0, cz.w1};
    const float sx[2] = {cx.s0, cx.s1}, sy[2] = {cy.s0, cy.s1}, sz[2] = {cz.s0, cz.s1};
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          if (REL && (ix[a] < 0 || iy[b] < 0 || iz[c] < 0)) continue;
          const float k = vx * ((sx[a] * wy[b]) * wz[c]) + vy * ((wx[a] * sy[b]) * wz[c]) +
                          vz * ((wx[a] * wy[b]) * sz[c]);
          atomicAdd(mesh + ((long long)ix[a] * ny + iy[b]) * nz + iz[c], k);
        }
  }
}

// 2-D CIC paint of a projected particle set (the density planes of a light cone): 4 REDG.E.ADD.F32 per particle.
//   reference: jaxpm/painting.py:131-158 (cic_paint_2d): floor, +{0,1}, kernel = (1-|dx|)(1-|dy|) * weight,
//   int32 cast, python mod.  A plane (<= a few MB) lives in L2, so the global reductions stay on chip.
__global__ void __launch_bounds__(256)
paint2d_kernel(float* __restrict__ mesh, const float* __restrict__ pos, const float* __restrict__ weight, long long np,
               int nx, int ny) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < np; p += stride) {
    const Cic1 cx = cic_abs<false>(ld_stream(pos + 2 * p), nx);
    const Cic1 cy = cic_abs<false>(ld_stream(pos + 2 * p + 1), ny);
    const float w = weight ? weight[p] : 1.0f;
    const int ix[2] = {cx.i0, cx.i1}, iy[2] = {cy.i0, cy.i1};
    const float wx[2] = {cx.w0, cx.w1}, wy[2] = {cy.w0, cy.w1};
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const float k = (wx[a] * wy[b]) * w;      // kernel[...,0] * kernel[...,1], then * weight (:143-145)
        if (k != 0.f) atomicAdd(mesh + (long long)ix[a] * ny + iy[b], k);
      }
  }
}

// Density plane of a light cone in ONE pass over the particles (no xy / weight arrays are materialised):
//   reference: jaxpm/lensing.py:11-44 (density_plane): xy = mod(pos[:2], nx); xy = xy / nx * res; weight = 1 where
//   center - width/2 < pos[2] <= center + width/2; cic_paint_2d(zeros(res, res), xy, weight).  The normalisation
//   (:37-38) and the optional smoothing are applied by the caller.
__device__ __forceinline__ float pymod_f(float x, float n) {   // jnp.mod on floats: sign of the divisor
  float r = fmodf(x, n);
  if (r != 0.0f && r < 0.0f) r += n;
  return r;
}

__global__ void __launch_bounds__(256)
density_plane_kernel(float* __restrict__ plane, const float* __restrict__ pos, long long np, float nx, float lo,
                     float hi, int res) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const float fres = (float)res;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < np; p += stride) {
    const float d = ld_stream(pos + 3 * p + 2);
    if (!(d > lo && d <= hi)) continue;                        // weight 0: contributes exactly nothing
    const float x = pymod_f(ld_stream(pos + 3 * p), nx) / nx * fres;
    const float y = pymod_f(ld_stream(pos + 3 * p + 1), nx) / nx * fres;
    const Cic1 cx = cic_abs<false>(x, res);
    const Cic1 cy = cic_abs<false>(y, res);
    const int ix[2] = {cx.i0, cx.i1}, iy[2] = {cy.i0, cy.i1};
    const float wx[2] = {cx.w0, cx.w1}, wy[2] = {cy.w0, cy.w1};
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const float k = wx[a] * wy[b];
        if (k != 0.f) atomicAdd(plane + (long long)ix[a] * res + iy[b], k);
      }
  }
}

__global__ void __launch_bounds__(256)
cell_index_kernel(int* __restrict__ out, const float* __restrict__ pos, long long np, int nx, int ny,
                  int nz, int hx, int hy, int rel) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= np) return;
  const float px = pos[3 * p], py = pos[3 * p + 1], pz = pos[3 * p + 2];
  int i, j, k;
  if (rel) {
    const int pnz = nz, pny = ny - 2 * hy;
    const int bk = (int)(p % pnz);
    const long long t = p / pnz;
    const int bj = (int)(t % pny) + hy, bi = (int)(t / pny) + hx;
    i = cic_rel<false>(bi, px, nx).i0;
    j = cic_rel<false>(bj, py, ny).i0;
    k = cic_rel<false>(bk, pz, nz).i0;
  } else {
    i = cic_abs<false>(px, nx).i0;
    j = cic_abs<false>(py, ny).i0;
    k = cic_abs<false>(pz, nz).i0;
  }
  out[p] = (i < 0 || j < 0 || k < 0) ? -1 : (i * ny + j) * nz + k;
}

template <bool REL>
static int32_t launch_paint(cudaStream_t s, float* mesh, const float* pos, const float* weight,
                            float wscalar, long long np, int nx, int ny, int nz, int pnx, int pny,
                            int pnz, int hx, int hy) {
  if (np == 0) return JPM_OK;
  const int threads = 256;
  long long blocks = (np + threads - 1) / threads;
  const long long cap = (long long)kNumSMs * 32;
  if (blocks > cap) blocks = cap;
  paint_direct_kernel<REL><<<(int)blocks, threads, 0, s>>>(mesh, pos, weight, wscalar, np, nx, ny,
                                                           nz, pnx, pny, pnz, hx, hy);
  JPM_LAUNCH_CHECK();
  return JPM_OK;
}

}  // namespace jpm

using namespace jpm;

extern "C" int32_t jpm_cic_paint_f32(void* stream, float* mesh, const float* positions,
                                     const float* weight, float weight_scalar, int64_t np,
                                     int32_t nx, int32_t ny, int32_t nz, int32_t pgx, int32_t pgy,
                                     int32_t pgz) {
  JPM_CHECK_ARG(mesh && (positions || np == 0), "null mesh/positions");
  JPM_CHECK_ARG(nx > 0 && ny > 0 && nz > 0 && np >= 0, "bad mesh shape / np");
  JPM_CHECK_ARG((int64_t)nx * ny * nz < (1ll << 31), "mesh too large for int32 cell ids");
  JPM_CHECK_ARG((int64_t)pgx * pgy * pgz == np, "particle grid does not match np");
  return launch_paint<false>((cudaStream_t)stream, mesh, positions, weight, weight_scalar, np, nx,
                             ny, nz, pgx, pgy, pgz, 0, 0);
}

extern "C" int32_t jpm_cic_paint_dx_f32(void* stream, float* mesh, const float* disp,
                                        const float* weight, float weight_scalar, int32_t nx,
                                        int32_t ny, int32_t nz, int32_t hx, int32_t hy) {
  JPM_CHECK_ARG(mesh && disp, "null mesh/disp");
  JPM_CHECK_ARG(nx > 0 && ny > 0 && nz > 0 && hx >= 0 && hy >= 0, "bad shape / halo");
  const int mx = nx + 2 * hx, my = ny + 2 * hy;
  JPM_CHECK_ARG((int64_t)mx * my * nz < (1ll << 31), "mesh too large for int32 cell ids");
  return launch_paint<true>((cudaStream_t)stream, mesh, disp, weight, weight_scalar,
                            (long long)nx * ny * nz, mx, my, nz, nx, ny, nz, hx, hy);
}

extern "C" int32_t jpm_cic_paintgrad_f32(void* stream, float* mesh, const float* pos_or_disp, const float* tangent,
                                         const float* weight, float weight_scalar, int64_t np, int32_t nx,
                                         int32_t ny, int32_t nz, int32_t hx, int32_t hy, int32_t relative) {
  JPM_CHECK_ARG(mesh && pos_or_disp && tangent && np >= 0, "null pointer");
  JPM_CHECK_ARG(nx > 0 && ny > 0 && nz > 0 && hx >= 0 && hy >= 0, "bad shape / halo");
  if (relative) JPM_CHECK_ARG((int64_t)(nx - 2 * hx) * (ny - 2 * hy) * nz == np, "np != particle grid");
  if (np == 0) return JPM_OK;
  long long blocks = (np + 255) / 256;
  if (blocks > kNumSMs * 32) blocks = kNumSMs * 32;
  if (relative)
    paintgrad_kernel<true><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(mesh, pos_or_disp, tangent, weight, weight_scalar,
                                                                         np, nx, ny, nz, ny - 2 * hy, nz, hx, hy);
  else
    paintgrad_kernel<false><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(mesh, pos_or_disp, tangent, weight, weight_scalar,
                                                                          np, nx, ny, nz, ny, nz, hx, hy);
  JPM_LAUNCH_CHECK();
  return JPM_OK;
}

extern "C" int32_t jpm_cic_cell_index_i32(void* stream, int32_t* out, const float* pos_or_disp,
                                          int64_t np, int32_t nx, int32_t ny, int32_t nz,
                                          int32_t hx, int32_t hy, int32_t mode) {
  JPM_CHECK_ARG(out && pos_or_disp && np >= 0, "null pointer");
  JPM_CHECK_ARG(nx > 0 && ny > 0 && nz > 0, "bad mesh shape");
  if (mode) JPM_CHECK_ARG((int64_t)(nx - 2 * hx) * (ny - 2 * hy) * nz == np, "np != particle grid");
  if (np == 0) return JPM_OK;
  cell_index_kernel<<<div_up(np, 256), 256, 0, (cudaStream_t)stream>>>(out, pos_or_disp, np, nx, ny,
                                                                       nz, hx, hy, mode);
  JPM_LAUNCH_CHECK();
  return JPM_OK;
}

extern "C" int32_t jpm_cic_paint_2d_f32(void* stream, float* mesh, const float* pos2, const float* weight, int64_t np,
                                        int32_t nx, int32_t ny) {
  JPM_CHECK_ARG(mesh && (pos2 || np == 0) && np >= 0 && nx > 0 && ny > 0, "bad arguments");
  if (np == 0) return JPM_OK;
  long long blocks = (np + 255) / 256;
  if (blocks > (long long)jpm::kNumSMs * 16) blocks = (long long)jpm::kNumSMs * 16;
  jpm::paint2d_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(mesh, pos2, weight, np, nx, ny);
  JPM_LAUNCH_CHECK();
  return JPM_OK;
}

extern "C" int32_t jpm_density_plane_f32(void* stream, float* plane, const float* pos3, int64_t np, float box_nx,
                                         double center, double width, int32_t plane_resolution) {
  JPM_CHECK_ARG(plane && (pos3 || np == 0) && np >= 0 && box_nx > 0.f && plane_resolution > 0, "bad arguments");
  if (np == 0) return JPM_OK;
  long long blocks = (np + 255) / 256;
  if (blocks > (long long)jpm::kNumSMs * 16) blocks = (long long)jpm::kNumSMs * 16;
  // the comparison bounds as the reference forms them: Python floats (float64) center -+ width / 2, compared with
  // the float32 coordinates as weak-typed scalars, i.e. rounded to float32
  const float lo = (float)(center - width / 2), hi = (float)(center + width / 2);
  jpm::density_plane_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(plane, pos3, np, box_nx, lo, hi,
                                                                          plane_resolution);
  JPM_LAUNCH_CHECK();
  return JPM_OK;
}

extern "C" int32_t jpm_cic_paint3_f32(void* stream, float* mesh3, const float* pos_or_disp, const float* weights3,
                                      float scale, int64_t np, int32_t nx, int32_t ny, int32_t nz, int32_t hx,
                                      int32_t hy, int32_t relative) {
  JPM_CHECK_ARG(mesh3 && pos_or_disp && weights3 && np >= 0, "null pointer");
  JPM_CHECK_ARG(nx > 0 && ny > 0 && nz > 0 && (int64_t)nx * ny * nz < (1ll << 31), "bad mesh shape");
  if (relative)
    JPM_CHECK_ARG((int64_t)(nx - 2 * hx) * (ny - 2 * hy) * nz == np, "np != particle grid");
  if (np == 0) return JPM_OK;
  long long blocks = (np + 255) / 256;
  const long long cap = (long long)jpm::kNumSMs * 32;
  if (blocks > cap) blocks = cap;
  const long long ms = (long long)nx * ny * nz;
  if (relative)
    jpm::paint3_kernel<true><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(mesh3, ms, pos_or_disp, weights3, scale, np, nx,
                                                                       ny, nz, ny - 2 * hy, nz, hx, hy);
  else
    jpm::paint3_kernel<false><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(mesh3, ms, pos_or_disp, weights3, scale, np, nx,
                                                                        ny, nz, ny - 2 * hy, nz, hx, hy);
  JPM_LAUNCH_CHECK();
  return JPM_OK;
}
