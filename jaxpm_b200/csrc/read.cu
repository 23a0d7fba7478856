// K5 / K6 — CIC read (8-corner gather), read3 + kick + drift, read-with-gradient.
//   reference: jaxpm/painting.py:78-106 (_cic_read_impl), :218-236 (_cic_read_dx_impl),
//              jaxpm/painting_utils.py:144-187 (gather), jaxpm/pm.py:54-56 (3 reads + stack),
//              jaxpm/ode.py:91-117 (drift / kick)
#include "common.cuh"

namespace jpm {

template <bool REL, bool GRAD>
__device__ __forceinline__ void make_stencil(long long p, float px, float py, float pz, int nx,
                                             int ny, int nz, int pny, int pnz, int hx, int hy,
                                             Cic1& cx, Cic1& cy, Cic1& cz) {
  int bi = 0, bj = 0, bk = 0;
  if (REL) {
    bk = (int)(p % pnz);
    const long long t = p / pnz;
    bj = (int)(t % pny) + hy;
    bi = (int)(t / pny) + hx;
  }
  cx = cic_1d<REL, GRAD>(bi, px, nx);
  cy = cic_1d<REL, GRAD>(bj, py, ny);
  cz = cic_1d<REL, GRAD>(bk, pz, nz);
}

// NF force meshes gathered at once; MODE 0: out[np][NF] = scale*read ; MODE 1: kick+drift.
template <bool REL, int NF, int MODE>
__global__ void __launch_bounds__(256)
read_kernel(float* out, float* pos_out, float* vel_out,  // may alias pos_prev / vel_prev
            const float* __restrict__ f0, const float* __restrict__ f1, const float* __restrict__ f2,
            const float* pos_in, const float* vel_in, const float* pos_prev, const float* vel_prev,
            float scale,
            float kick, float drift, int use_new_vel, long long np, int nx, int ny, int nz, int pny,
            int pnz, int hx, int hy) {
  const float* fm[3] = {f0, f1, f2};
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < np; p += stride) {
    const float px = ld_stream(pos_in + 3 * p + 0);
    const float py = ld_stream(pos_in + 3 * p + 1);
    const float pz = ld_stream(pos_in + 3 * p + 2);
    Cic1 cx, cy, cz;
    make_stencil<REL, false>(p, px, py, pz, nx, ny, nz, pny, pnz, hx, hy, cx, cy, cz);
    const int ix[2] = {cx.i0, cx.i1}, iy[2] = {cy.i0, cy.i1}, iz[2] = {cz.i0, cz.i1};
    const float wx[2] = {cx.w0, cx.w1}, wy[2] = {cy.w0, cy.w1}, wz[2] = {cz.w0, cz.w1};
    float acc[NF];
#pragma unroll
    for (int f = 0; f < NF; ++f) acc[f] = 0.f;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          if (REL && (ix[a] < 0 || iy[b] < 0 || iz[c] < 0)) continue;
          const float k = (wx[a] * wy[b]) * wz[c];
          const long long o = ((long long)ix[a] * ny + iy[b]) * nz + iz[c];
#pragma unroll
          for (int f = 0; f < NF; ++f) acc[f] = fmaf(__ldg(fm[f] + o), k, acc[f]);
        }
    if (MODE == 0) {
#pragma unroll
      for (int f = 0; f < NF; ++f) st_stream(out + NF * p + f, scale * acc[f]);
    } else {
      float v[3], x[3];
#pragma unroll
      for (int f = 0; f < 3; ++f) {
        const float vin = ld_stream(vel_in + 3 * p + f);
        const float vp = (vel_prev == vel_in) ? vin : ld_stream(vel_prev + 3 * p + f);
        v[f] = fmaf(kick, acc[f], vp);
        const float pin = (f == 0) ? px : ((f == 1) ? py : pz);
        const float pp = (pos_prev == pos_in) ? pin : ld_stream(pos_prev + 3 * p + f);
        x[f] = fmaf(drift, use_new_vel ? v[f] : vin, pp);
      }
#pragma unroll
      for (int f = 0; f < 3; ++f) {
        st_stream(vel_out + 3 * p + f, v[f]);
        st_stream(pos_out + 3 * p + f, x[f]);
        if (out) st_stream(out + 3 * p + f, acc[f]);
      }
    }
  }
}

template <bool REL>
__global__ void __launch_bounds__(256)
readgrad_kernel(float* __restrict__ value, float* __restrict__ grad, const float* __restrict__ mesh,
                const float* __restrict__ pos, const float* __restrict__ pscale, float scale, long long np, int nx, int ny, int nz, int pny, int pnz,
                int hx, int hy) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < np; p += stride) {
    const float px = ld_stream(pos + 3 * p + 0);
    const float py = ld_stream(pos + 3 * p + 1);
    const float pz = ld_stream(pos + 3 * p + 2);
    Cic1 cx, cy, cz;
    make_stencil<REL, true>(p, px, py, pz, nx, ny, nz, pny, pnz, hx, hy, cx, cy, cz);
    const int ix[2] = {cx.i0, cx.i1}, iy[2] = {cy.i0, cy.i1}, iz[2] = {cz.i0, cz.i1};
    const float wx[2] = {cx.w0, cx.w1}, wy[2] = {cy.w0, cy.w1}, wz[2] = {cz.w0, cz.w1};
    const float sx[2] = {cx.s0, cx.s1}, sy[2] = {cy.s0, cy.s1}, sz[2] = {cz.s0, cz.s1};
    float v = 0.f, gx = 0.f, gy = 0.f, gz = 0.f;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          if (REL && (ix[a] < 0 || iy[b] < 0 || iz[c] < 0)) continue;
          const float m = __ldg(mesh + ((long long)ix[a] * ny + iy[b]) * nz + iz[c]);
          v = fmaf(m, (wx[a] * wy[b]) * wz[c], v);
          gx = fmaf(m, (sx[a] * wy[b]) * wz[c], gx);
          gy = fmaf(m, (wx[a] * sy[b]) * wz[c], gy);
          gz = fmaf(m, (wx[a] * wy[b]) * sz[c], gz);
        }
    if (value) st_stream(value + p, v);
    if (grad) {
      const float sc = pscale ? scale * pscale[p] : scale;
      st_stream(grad + 3 * p + 0, sc * gx);
      st_stream(grad + 3 * p + 1, sc * gy);
      st_stream(grad + 3 * p + 2, sc * gz);
    }
  }
}

// The adjoint of the three reads of pm.py:54-56 with respect to the positions, in ONE pass over the particles:
//   grad[p] (+)= sum_d u[p][d] * d read(F_d)(x_p) / d x_p      (NM == 3, cotangent u[np][3])
//   grad[p] (+)= d read(mesh)(x_p) / d x_p                      (NM == 1, the paint adjoint: u == nullptr)
// (three readgrad_kernel launches + three axpby before).  ACC: accumulate into grad.
template <bool REL, int NM, bool ACC>
__global__ void __launch_bounds__(256)
readgradn_kernel(float* __restrict__ grad, const float* __restrict__ m0, const float* __restrict__ m1,
                 const float* __restrict__ m2, const float* __restrict__ pos, const float* __restrict__ u, float scale,
                 long long np, int nx, int ny, int nz, int pny, int pnz, int hx, int hy) {
  const float* mm[3] = {m0, m1, m2};
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < np; p += stride) {
    const float px = ld_stream(pos + 3 * p + 0);
    const float py = ld_stream(pos + 3 * p + 1);
    const float pz = ld_stream(pos + 3 * p + 2);
    float w[NM];
#pragma unroll
    for (int d = 0; d < NM; ++d) w[d] = u ? scale * ld_stream(u + NM * p + d) : scale;
    Cic1 cx, cy, cz;
    make_stencil<REL, true>(p, px, py, pz, nx, ny, nz, pny, pnz, hx, hy, cx, cy, cz);
    const int ix[2] = {cx.i0, cx.i1}, iy[2] = {cy.i0, cy.i1}, iz[2] = {cz.i0, cz.i1};
    const float wx[2] = {cx.w0, cx.w1}, wy[2] = {cy.w0, cy.w1}, wz[2] = {cz.w0, cz.w1};
    const float sx[2] = {cx.s0, cx.s1}, sy[2] = {cy.s0, cy.s1}, sz[2] = {cz.s0, cz.s1};
    float gx = 0.f, gy = 0.f, gz = 0.f;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          if (REL && (ix[a] < 0 || iy[b] < 0 || iz[c] < 0)) continue;
          const long long o = ((long long)ix[a] * ny + iy[b]) * nz + iz[c];
          float m = 0.f;
#pragma unroll
          for (int d = 0; d < NM; ++d) m = fmaf(w[d], __ldg(mm[d] + o), m);
          gx = fmaf(m, (sx[a] * wy[b]) * wz[c], gx);
          gy = fmaf(m, (wx[a] * sy[b]) * wz[c], gy);
          gz = fmaf(m, (wx[a] * wy[b]) * sz[c], gz);
        }
    if (ACC) {
      gx += grad[3 * p + 0]; gy += grad[3 * p + 1]; gz += grad[3 * p + 2];
    }
    st_stream(grad + 3 * p + 0, gx);
    st_stream(grad + 3 * p + 1, gy);
    st_stream(grad + 3 * p + 2, gz);
  }
}

static int grid_for(long long np) {
  long long blocks = (np + 255) / 256;
  const long long cap = (long long)kNumSMs * 32;
  return (int)(blocks > cap ? cap : blocks);
}

template <int NF, int MODE>
static int32_t launch_read(cudaStream_t s, bool rel, float* out, float* pos_out, float* vel_out,
                           const float* f0, const float* f1, const float* f2, const float* pos_in,
                           const float* vel_in, const float* pos_prev, const float* vel_prev,
                           float scale, float kick, float drift, int use_new_vel, long long np,
                           int nx, int ny, int nz, int hx, int hy) {
  if (np == 0) return JPM_OK;
  const int pny = ny - 2 * hy, pnz = nz;
  if (rel)
    read_kernel<true, NF, MODE><<<grid_for(np), 256, 0, s>>>(
        out, pos_out, vel_out, f0, f1, f2, pos_in, vel_in, pos_prev, vel_prev, scale, kick, drift,
        use_new_vel, np, nx, ny, nz, pny, pnz, hx, hy);
  else
    read_kernel<false, NF, MODE><<<grid_for(np), 256, 0, s>>>(
        out, pos_out, vel_out, f0, f1, f2, pos_in, vel_in, pos_prev, vel_prev, scale, kick, drift,
        use_new_vel, np, nx, ny, nz, pny, pnz, hx, hy);
  JPM_LAUNCH_CHECK();
  return JPM_OK;
}

}  // namespace jpm

using namespace jpm;

#define JPM_CHECK_MESH(nx, ny, nz)                                             \
  JPM_CHECK_ARG((nx) > 0 && (ny) > 0 && (nz) > 0, "bad mesh shape");           \
  JPM_CHECK_ARG((int64_t)(nx) * (ny) * (nz) < (1ll << 31), "mesh too large for int32 cell ids")

extern "C" int32_t jpm_cic_read_f32(void* stream, float* out, const float* mesh,
                                    const float* positions, int64_t np, int32_t nx, int32_t ny,
                                    int32_t nz) {
  JPM_CHECK_ARG(np >= 0 && mesh && (np == 0 || (out && positions)), "null pointer");
  JPM_CHECK_MESH(nx, ny, nz);
  return launch_read<1, 0>((cudaStream_t)stream, false, out, nullptr, nullptr, mesh, nullptr,
                           nullptr, positions, nullptr, nullptr, nullptr, 1.0f, 0.f, 0.f, 0, np, nx,
                           ny, nz, 0, 0);
}

extern "C" int32_t jpm_cic_read_dx_f32(void* stream, float* out, const float* mesh,
                                       const float* disp, int32_t nx, int32_t ny, int32_t nz,
                                       int32_t hx, int32_t hy) {
  JPM_CHECK_ARG(out && mesh && disp && hx >= 0 && hy >= 0, "null pointer / bad halo");
  const int mx = nx + 2 * hx, my = ny + 2 * hy;
  JPM_CHECK_MESH(mx, my, nz);
  return launch_read<1, 0>((cudaStream_t)stream, true, out, nullptr, nullptr, mesh, nullptr,
                           nullptr, disp, nullptr, nullptr, nullptr, 1.0f, 0.f, 0.f, 0,
                           (long long)nx * ny * nz, mx, my, nz, hx, hy);
}

extern "C" int32_t jpm_cic_read3_f32(void* stream, float* out, const float* fx, const float* fy,
                                     const float* fz, const float* pos_or_disp, float scale,
                                     int64_t np, int32_t nx, int32_t ny, int32_t nz, int32_t hx,
                                     int32_t hy, int32_t relative) {
  JPM_CHECK_ARG(out && fx && fy && fz && pos_or_disp && np >= 0, "null pointer");
  JPM_CHECK_MESH(nx, ny, nz);
  if (relative)
    JPM_CHECK_ARG((int64_t)(nx - 2 * hx) * (ny - 2 * hy) * nz == np, "np != particle grid");
  return launch_read<3, 0>((cudaStream_t)stream, relative != 0, out, nullptr, nullptr, fx, fy, fz,
                           pos_or_disp, nullptr, nullptr, nullptr, scale, 0.f, 0.f, 0, np, nx, ny,
                           nz, hx, hy);
}

extern "C" int32_t jpm_cic_read3_kick_drift_f32(void* stream, float* pos_out, float* vel_out,
                                                float* forces_out, const float* fx, const float* fy,
                                                const float* fz, const float* pos_in,
                                                const float* vel_in, const float* pos_prev,
                                                const float* vel_prev, float kick_coef,
                                                float drift_coef, int32_t use_new_vel, int64_t np,
                                                int32_t nx, int32_t ny, int32_t nz, int32_t hx,
                                                int32_t hy, int32_t relative) {
  JPM_CHECK_ARG(pos_out && vel_out && fx && fy && fz && pos_in && vel_in && pos_prev && vel_prev,
                "null pointer");
  JPM_CHECK_ARG(np >= 0, "np < 0");
  JPM_CHECK_MESH(nx, ny, nz);
  if (relative)
    JPM_CHECK_ARG((int64_t)(nx - 2 * hx) * (ny - 2 * hy) * nz == np, "np != particle grid");
  return launch_read<3, 1>((cudaStream_t)stream, relative != 0, forces_out, pos_out, vel_out, fx,
                           fy, fz, pos_in, vel_in, pos_prev, vel_prev, 1.0f, kick_coef, drift_coef,
                           use_new_vel, np, nx, ny, nz, hx, hy);
}

extern "C" int32_t jpm_cic_readgrad_f32(void* stream, float* value, float* grad, const float* mesh,
                                        const float* pos_or_disp, const float* grad_scale,
                                        float grad_scale_scalar, int64_t np, int32_t nx,
                                        int32_t ny, int32_t nz, int32_t hx, int32_t hy,
                                        int32_t relative) {
  JPM_CHECK_ARG((value || grad) && mesh && pos_or_disp && np >= 0, "null pointer");
  JPM_CHECK_MESH(nx, ny, nz);
  if (relative)
    JPM_CHECK_ARG((int64_t)(nx - 2 * hx) * (ny - 2 * hy) * nz == np, "np != particle grid");
  if (np == 0) return JPM_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int pny = ny - 2 * hy;
  if (relative)
    readgrad_kernel<true><<<grid_for(np), 256, 0, s>>>(value, grad, mesh, pos_or_disp, grad_scale,
                                                       grad_scale_scalar, np, nx, ny, nz, pny, nz, hx, hy);
  else
    readgrad_kernel<false><<<grid_for(np), 256, 0, s>>>(value, grad, mesh, pos_or_disp, grad_scale,
                                                        grad_scale_scalar, np, nx, ny, nz, pny, nz, hx, hy);
  JPM_LAUNCH_CHECK();
  return JPM_OK;
}

extern "C" int32_t jpm_cic_readgrad3_f32(void* stream, float* grad, const float* m0, const float* m1, const float* m2,
                                         const float* pos_or_disp, const float* cotangent, float scale, int64_t np,
                                         int32_t nx, int32_t ny, int32_t nz, int32_t hx, int32_t hy, int32_t relative,
                                         int32_t accumulate) {
  JPM_CHECK_ARG(grad && m0 && pos_or_disp && np >= 0, "null pointer");
  JPM_CHECK_ARG((m1 != nullptr) == (m2 != nullptr), "pass one mesh or three");
  JPM_CHECK_ARG(!(m1 && !cotangent), "three meshes need the cotangent u[np][3]");
  JPM_CHECK_ARG(!(!m1 && cotangent), "one mesh takes no per-particle cotangent (use jpm_cic_readgrad_f32)");
  JPM_CHECK_MESH(nx, ny, nz);
  if (relative)
    JPM_CHECK_ARG((int64_t)(nx - 2 * hx) * (ny - 2 * hy) * nz == np, "np != particle grid");
  if (np == 0) return JPM_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int pny = ny - 2 * hy;
#define RG(REL_, NM_, ACC_)                                                                                  \
  readgradn_kernel<REL_, NM_, ACC_><<<grid_for(np), 256, 0, s>>>(grad, m0, m1, m2, pos_or_disp, cotangent, scale, np, \
                                                                 nx, ny, nz, pny, nz, hx, hy)
  if (m1) {
    if (relative) { if (accumulate) RG(true, 3, true); else RG(true, 3, false); }
    else { if (accumulate) RG(false, 3, true); else RG(false, 3, false); }
  } else {
    if (relative) { if (accumulate) RG(true, 1, true); else RG(true, 1, false); }
    else { if (accumulate) RG(false, 1, true); else RG(false, 1, false); }
  }
#undef RG
  JPM_LAUNCH_CHECK();
  return JPM_OK;
}
