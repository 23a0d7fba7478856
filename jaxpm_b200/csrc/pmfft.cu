// pmfft — the fused k-space chain of the PM force evaluation as five hand-written FFT passes
// (replaces cuFFT R2C + the Green's/gradient pass + 3x cuFFT C2R for power-of-two meshes):
//
//   Z-fwd   density_p rows (ghost zones folded on the fly)  --R2C along z-->  A[x][y][kz]
//   Y-fwd   A columns along y, in place
//   X-fused A columns along x: forward FFT, delta_k stays in REGISTERS, times
//           i a_d(k) / k^2 * G(k) * norm for d = x, y, z, three inverse FFTs along x  -->  B3[d][x][y][kz]
//   Y-inv   B3 columns along y, in place (3 components)
//   Z-inv   B3 rows --C2R along z--> force3_p rows, ghost zones filled on the fly
//
//   reference: jaxpm/pm.py:41-56 (fft3d, invlaplace * longrange, -gradient_kernel, ifft3d x3),
//              jaxpm/kernels.py:10-23,41-115, jaxpm/distributed.py:37-42.
//
// HBM traffic per force evaluation (Nc cells, fp32), measured with ncu at 512^3 (profiles/traffic_512.json):
//   three-transform chain: 8 + 8 + 12 + 20 + 25 = 73 B/cell (9.8 GB) against 8 + 16 + 24 = 48 B/cell algorithmic and
//                          ~130 B/cell for cuFFT (3 passes per 3-D transform) + the k-space pass;
//   potential chain:       8 + 8 + 8 + 8 + 8 = 40 B/cell (5.4 GB) + the gradient pass 4 + 12 = 16 B/cell (2.2 GB):
//                          ONE inverse transform psi = IFFT(delta_k / k^2), the three force meshes by the 4th-order
//                          difference stencil the reference's gradient kernel is the symbol of (X-pot, Y-pot, gradient
//                          pass below; chosen per step by the measured fp32 error bound, csrc/sim.cu).
//
// Each 1-D FFT is a Stockham autosort transform in shared memory, radix 8/4, twiddles from a table
// computed in double precision.  Column passes keep a [N][C] tile (C consecutive kz = one 128-byte or
// 64-byte segment per row, so every global access is a full-sector segment); the first stage loads
// straight from global memory and the last stage stores straight to it, so an N = 512 transform makes
// 4 shared-memory passes instead of 8.  Row passes (z) stage 16 rows with a pitch of N/2+1 complex, lanes
// across rows: conflict-free for every stage permutation.
#include <cmath>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "plan_internal.cuh"

namespace jpm {
namespace fft {

// ---- complex helpers --------------------------------------------------------------------------
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
// multiply by -i (forward transforms) or +i (inverse)
template <bool INV>
__device__ __forceinline__ float2 mul_mi(float2 a) {
  return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}

template <bool INV>
__device__ __forceinline__ void dft4(float2& v0, float2& v1, float2& v2, float2& v3) {
  const float2 t0 = cadd(v0, v2), t1 = csub(v0, v2), t2 = cadd(v1, v3), t3 = mul_mi<INV>(csub(v1, v3));
  v0 = cadd(t0, t2); v1 = cadd(t1, t3); v2 = csub(t0, t2); v3 = csub(t1, t3);
}

template <int R, bool INV>
__device__ __forceinline__ void dft(float2* v) {
  if (R == 2) {
    const float2 a = v[0], b = v[1];
    v[0] = cadd(a, b); v[1] = csub(a, b);
  } else if (R == 4) {
    dft4<INV>(v[0], v[1], v[2], v[3]);
  } else {  // R == 8
    float2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6], o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];
    dft4<INV>(e0, e1, e2, e3);
    dft4<INV>(o0, o1, o2, o3);
    const float h = 0.70710678118654752440f;
    // o1 *= W8, o2 *= -+i, o3 *= W8^3
    o1 = INV ? make_float2((o1.x - o1.y) * h, (o1.x + o1.y) * h) : make_float2((o1.x + o1.y) * h, (o1.y - o1.x) * h);
    o2 = mul_mi<INV>(o2);
    o3 = INV ? make_float2((-o3.x - o3.y) * h, (o3.x - o3.y) * h) : make_float2((o3.y - o3.x) * h, (-o3.x - o3.y) * h);
    v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
    v[1] = cadd(e1, o1); v[5] = csub(e1, o1);
    v[2] = cadd(e2, o2); v[6] = csub(e2, o2);
    v[3] = cadd(e3, o3); v[7] = csub(e3, o3);
  }
}

// ---- radix schedule ---------------------------------------------------------------------------
__host__ __device__ constexpr int pick_radix(int rem) { return rem == 16 ? 4 : (rem >= 8 ? 8 : rem); }
__host__ __device__ constexpr int radix_count(int N) {
  int c = 0, ns = 1;
  while (ns < N) { ns *= pick_radix(N / ns); ++c; }
  return c;
}
__host__ __device__ constexpr int radix_fwd(int N, int i) {
  int ns = 1, r = 1;
  for (int s = 0; s <= i; ++s) { r = pick_radix(N / ns); ns *= r; }
  return r;
}
// stage i of the schedule, forward order or reversed
__host__ __device__ constexpr int radix_at(int N, int i, bool rev) {
  return rev ? radix_fwd(N, radix_count(N) - 1 - i) : radix_fwd(N, i);
}
__host__ __device__ constexpr int ns_at(int N, int i, bool rev) {
  int ns = 1;
  for (int s = 0; s < i; ++s) ns *= radix_at(N, s, rev);
  return ns;
}

// ---- shared-memory layouts --------------------------------------------------------------------
template <int C>
struct LayCols {   // element n of column c
  __device__ __forceinline__ static int idx(int n, int c) { return n * C + c; }
};
template <int N>
struct LayRows {   // element n of row c, pitch N + 1 (odd number of complex words: conflict-free across rows)
  __device__ __forceinline__ static int idx(int n, int c) { return c * (N + 1) + n; }
};

template <int C>
struct LayPair {   // element n of column c in a [N][C][2] tile (T0 and T1 of a mode side by side; base + 0 | 1)
  __device__ __forceinline__ static int idx(int n, int c) { return (n * C + c) * 2; }
};

template <class LAY>
struct SmemIO {
  static constexpr bool kSmem = true;
  float2* s;
  __device__ __forceinline__ float2 operator()(int n, int c) const { return s[LAY::idx(n, c)]; }
  __device__ __forceinline__ void operator()(int n, int c, float2 v) const { s[LAY::idx(n, c)] = v; }
};
struct GlobalIO {   // element (n, c) at base[n * stride + c]
  static constexpr bool kSmem = false;
  float2* base;
  long long stride;
  __device__ __forceinline__ float2 operator()(int n, int c) const { return __ldcs(base + n * stride + c); }
  __device__ __forceinline__ void operator()(int n, int c, float2 v) const { __stcs(base + n * stride + c, v); }
};
// keeps the values in the caller's registers (X-fused pass): slot = task-local index
struct NullIO {
  static constexpr bool kSmem = false;
};

// One Stockham stage over the tile: N points, radix R, Ns = product of the radices already applied.
// Tasks (j, c): butterfly j of column c; task t -> c = t % C, j = t / C; TPT tasks per thread.
// Inputs n = j + r N/R, outputs n = (j / Ns) Ns R + j % Ns + r Ns.
template <int N, int Ns, int R, int C, int NT, bool INV, bool KEEP_IN, bool KEEP_OUT, class LD, class ST>
__device__ __forceinline__ void stage(LD ld, ST st, const float2* __restrict__ tw, int ncol,
                                      float2 (*keep)[8]) {
  constexpr int TASKS = (N / R) * C;
  constexpr int TPT = (TASKS + NT - 1) / NT;
  float2 v[TPT][R];
#pragma unroll
  for (int i = 0; i < TPT; ++i) {
    const int task = threadIdx.x + i * NT;
    const int c = task % C, j = task / C;
    if ((TASKS % NT == 0 || task < TASKS) && c < ncol) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if constexpr (KEEP_IN) v[i][r] = keep[i][r];
        else v[i][r] = ld(j + r * (N / R), c);
      }
    }
  }
  if constexpr (!KEEP_IN && LD::kSmem) __syncthreads();
#pragma unroll
  for (int i = 0; i < TPT; ++i) {
    const int task = threadIdx.x + i * NT;
    const int c = task % C, j = task / C;
    if ((TASKS % NT == 0 || task < TASKS) && c < ncol) {
      const int k = j & (Ns - 1);
      if (Ns > 1) {
        const int t = k * (N / (Ns * R));
#pragma unroll
        for (int r = 1; r < R; ++r) {
          float2 w = tw[t * r];
          if (INV) w.y = -w.y;
          v[i][r] = cmul(v[i][r], w);
        }
      }
      dft<R, INV>(v[i]);
      const int j0 = (j - k) * R + k;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if constexpr (KEEP_OUT) keep[i][r] = v[i][r];
        else st(j0 + r * Ns, c, v[i][r]);
      }
    }
  }
  if constexpr (!KEEP_OUT && ST::kSmem) __syncthreads();
}

// All stages of an N-point transform.  Stage 0 reads through `ld0` (or the kept registers), the last
// stage writes through `stl` (or into the kept registers); everything in between lives in `s`.
template <int N, int C, int NT, bool INV, bool REV, bool KEEP_IN, bool KEEP_OUT, class LAY, int I, class LD0, class STL>
__device__ __forceinline__ void run_stages(float2* s, const float2* __restrict__ tw, int ncol, LD0 ld0, STL stl,
                                           float2 (*keep)[8]) {
  constexpr int L = radix_count(N);
  if constexpr (I < L) {
    constexpr int R = radix_at(N, I, REV), Ns = ns_at(N, I, REV);
    constexpr bool first = (I == 0), last = (I == L - 1);
    SmemIO<LAY> sm{s};
    if constexpr (first && last) {
      stage<N, Ns, R, C, NT, INV, KEEP_IN, KEEP_OUT>(ld0, stl, tw, ncol, keep);
    } else if constexpr (first) {
      stage<N, Ns, R, C, NT, INV, KEEP_IN, false>(ld0, sm, tw, ncol, keep);
    } else if constexpr (last) {
      stage<N, Ns, R, C, NT, INV, false, KEEP_OUT>(sm, stl, tw, ncol, keep);
    } else {
      stage<N, Ns, R, C, NT, INV, false, false>(sm, sm, tw, ncol, keep);
    }
    run_stages<N, C, NT, INV, REV, KEEP_IN, KEEP_OUT, LAY, I + 1>(s, tw, ncol, ld0, stl, keep);
  }
}

// threads per CTA: N * C / D (D = points of the tile per thread in a radix-8 stage), clamped to [64, 1024]
template <int N, int C, int D = 16>
__host__ __device__ constexpr int threads_for() {
  return (N * C / D) < 64 ? 64 : ((N * C / D) > 1024 ? 1024 : (N * C / D));
}

// ---- TMA stores (shared -> global tensor tile; SASS UTMASTG) ------------------------------------------
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tm, int c0, int c1, int c2, const void* src) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];"
               ::"l"((unsigned long long)tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_addr(src)) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* tm, int c0, int c1, int c2, int c3, const void* src) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3, %4}], [%5];"
               ::"l"((unsigned long long)tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_addr(src)) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the shared-memory sources of all committed stores have been read (the buffers may be reused / the CTA may exit)
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// ---- peer-aware global I/O of the slab decomposition -------------------------------------------------
// y-transformed element (y, c) of x plane `xg` goes to the rank that owns y: AT[xg][y % ly][kz] there
struct ScatterYIO {
  static constexpr bool kSmem = false;
  const Slab* sl;
  long long xoff;   // xg * ly * nzc + kz0
  __device__ __forceinline__ void operator()(int n, int c, float2 v) const {
    const int d = n / sl->ly, yl = n - d * sl->ly;
    __stcs(sl->at[d] + xoff + (long long)yl * sl->nzc + c, v);
  }
};
// x-inverse-transformed element (x, c) of row yg goes to the rank that owns x: T01[x % lx][yg][kz][comp] there
struct ScatterXIO {
  static constexpr bool kSmem = false;
  const Slab* sl;
  long long off;    // 2 * (yg * nzc + kz0) + comp   (float2 units)
  __device__ __forceinline__ void operator()(int n, int c, float2 v) const {
    const int d = n / sl->lx, xl = n - d * sl->lx;
    __stcs(reinterpret_cast<float2*>(sl->t01[d]) + off + 2 * ((long long)xl * sl->ny * sl->nzc + c), v);
  }
};

// x-inverse-transformed element (x, c) of row yg goes to the rank that owns x: B[comp][x % lx][yg][kz] there
struct ScatterXPlanarIO {
  static constexpr bool kSmem = false;
  const Slab* sl;
  long long off;    // comp * lx * ny * nzc + yg * nzc + kz0
  __device__ __forceinline__ void operator()(int n, int c, float2 v) const {
    const int d = n / sl->lx, xl = n - d * sl->lx;
    __stcs(sl->b3[d] + off + (long long)xl * sl->ny * sl->nzc + c, v);
  }
};

// ---- Y-fwd: FFT along y of the local x planes, result transposed onto the y-owning ranks ---------------
// grid: (ntile, lx).  in: A_loc = B3[2] region of this rank [lx][ny][nzc]; out: AT[nx][ly][nzc] of every rank.
template <int N, int C, bool TMAST>
__global__ void __launch_bounds__(threads_for<N, C>())
yfwd_kernel(const __grid_constant__ Slab sl, const float2* __restrict__ twg, const __grid_constant__ TmapPack tp,
            int x0) {
  constexpr int NT = threads_for<N, C>();
  extern __shared__ __align__(128) float2 sm[];
  float2* tw = sm;            // [N]
  float2* s = sm + N;         // [N][C]
  for (int i = threadIdx.x; i < N; i += NT) tw[i] = twg[i];
  const int kz0 = blockIdx.x * C, xl = x0 + blockIdx.y;
  const int ncol = min(C, sl.nzh - kz0);
  const long long cs = (long long)sl.lx * sl.ny * sl.nzc;
  GlobalIO gin{sl.b3[sl.rank] + 2 * cs + (long long)xl * sl.ny * sl.nzc + kz0, sl.nzc};
  const int xg = sl.rank * sl.lx + xl;
  if constexpr (TMAST) {
    // the finished tile stays in shared memory; rows [d ly, (d+1) ly) leave as one tensor store per
    // destination rank (<= 256 rows per box)
    SmemIO<LayCols<C>> so{s};
    run_stages<N, C, NT, false, false, false, false, LayCols<C>, 0>(s, tw, ncol, gin, so, nullptr);
    fence_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
      const int rows = min(sl.ly, 256);
      for (int d = 0; d < sl.P; ++d)
        for (int r0 = 0; r0 < sl.ly; r0 += rows)
          tma_store_3d(&tp.m[d], 2 * kz0, r0, xg, s + (size_t)(d * sl.ly + r0) * C);
      tma_store_commit();
      tma_store_wait_read();
    }
  } else {
    ScatterYIO gout{&sl, (long long)xg * sl.ly * sl.nzc + kz0};
    run_stages<N, C, NT, false, false, false, false, LayCols<C>, 0>(s, tw, ncol, gin, gout, nullptr);
  }
}

// ---- X-fused pass ---------------------------------------------------------------------------------
// grid: (ntile, ly).  Forward FFT along x; delta_k stays in registers and is scaled by norm * G(k) / k^2
// (operation order of kspace_kernel<0>, plan.cu); then TWO inverse FFTs along x:
//   T0 = IFFT_x(i a_x(kx) g delta)  -> B[0]   (x force; a_y, a_z do not depend on kx, so the y and z
//   T1 = IFFT_x(g delta)            -> B[1]    forces share T1: their factors are applied in Y-inv / Z-inv)
template <int N, int C, bool TMAST, bool PAIR, bool SUMSQ = false>
__global__ void __launch_bounds__(threads_for<N, C, 8>(), (threads_for<N, C, 8>() <= 512 ? 2 : 1))
xfused_kernel(const __grid_constant__ Slab sl, const float2* __restrict__ twg, const float* __restrict__ wx,
              const float* __restrict__ wy, const float* __restrict__ wz, const float* __restrict__ ax,
              float norm, float r_split2, const float* __restrict__ ftab, int ntab, float fscale,
              const __grid_constant__ TmapPack tp, double* __restrict__ sumsq) {
  constexpr int NT = threads_for<N, C, 8>();
  constexpr int L = radix_count(N);
  constexpr int RL = radix_at(N, L - 1, false);        // radix of the last forward == first inverse stage
  constexpr int TASKS = (N / RL) * C;
  constexpr int TPT = (TASKS + NT - 1) / NT;
  extern __shared__ __align__(128) float2 sm[];
  float2* tw = sm;            // [N]
  float2* s = sm + N;         // [N][C]
  float* swx = reinterpret_cast<float*>(s + N * C);   // [N] k_x table
  float* sax = swx + N;                               // [N] gradient table
  float2* so = reinterpret_cast<float2*>(sax + N);    // TMAST: [N][C][2] finished T0 | T1 of every mode, side by side
  for (int i = threadIdx.x; i < N; i += NT) { tw[i] = twg[i]; swx[i] = wx[i]; sax[i] = ax[i]; }
  const int kz0 = blockIdx.x * C;
  const int ncol = min(C, sl.nzh - kz0);
  const int yl = blockIdx.y, yg = sl.rank * sl.ly + yl;
  GlobalIO gin{sl.at[sl.rank] + (long long)yl * sl.nzc + kz0, (long long)sl.ly * sl.nzc};
  float2 keep[TPT][8];
  run_stages<N, C, NT, false, false, false, true, LayCols<C>, 0>(s, tw, ncol, gin, NullIO{}, keep);
  // delta_k(kx = j + r N/RL, yg, kz0 + c) is in keep[i][r]; scale it by norm * G(k) / k^2 in place
  const float ky = wy[yg];
#pragma unroll
  for (int i = 0; i < TPT; ++i) {
    const int task = threadIdx.x + i * NT;
    const int c = task % C, j = task / C;
    if ((TASKS % NT == 0 || task < TASKS) && c < ncol) {
      const float kz = wz[kz0 + c];
#pragma unroll
      for (int r = 0; r < RL; ++r) {
        const float kx = swx[j + r * (N / RL)];
        const float kxy2 = kx * kx + ky * ky;
        const float kk = kxy2 + kz * kz;
        float g = (kk == 0.f) ? 0.f : __frcp_rn(kk);   // correctly rounded, == 1.0f / kk
        g *= norm;
        if (r_split2 != 0.f) g *= expf(-kk * r_split2);
        if (ftab) {
          const float t = sqrtf(kk) * fscale;
          const int ti = min((int)t, ntab - 2);
          const float fr = fminf(t - (float)ti, 1.0f);
          g *= ftab[ti] + fr * (ftab[ti + 1] - ftab[ti]);
        }
        keep[i][r].x *= g;
        keep[i][r].y *= g;
      }
    }
  }
  if (SUMSQ && sumsq) {
    // sum_k |g delta_k|^2 over the full spectrum = mean_x psi^2 (error bound of the potential chain, csrc/sim.cu)
    float part = 0.f;
#pragma unroll
    for (int i = 0; i < TPT; ++i) {
      const int task = threadIdx.x + i * NT;
      const int c = task % C;
      if ((TASKS % NT == 0 || task < TASKS) && c < ncol) {
        const int kzi = kz0 + c;
        const float herm = (kzi == 0 || 2 * kzi == sl.nz) ? 1.f : 2.f;
#pragma unroll
        for (int r = 0; r < RL; ++r) part = fmaf(herm, keep[i][r].x * keep[i][r].x + keep[i][r].y * keep[i][r].y, part);
      }
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
    if ((threadIdx.x & 31) == 0) atomicAdd(sumsq, (double)part);
  }
#pragma unroll 1
  for (int d = 0; d < 2; ++d) {
    float2 w[TPT][8];
#pragma unroll
    for (int i = 0; i < TPT; ++i) {
      const int task = threadIdx.x + i * NT;
      const int j = task / C;
#pragma unroll
      for (int r = 0; r < RL; ++r) {
        if (d == 0) {
          const float ad = sax[(j + r * (N / RL)) & (N - 1)];
          w[i][r] = make_float2(-ad * keep[i][r].y, ad * keep[i][r].x);   // i a_x (g delta)
        } else {
          w[i][r] = keep[i][r];
        }
      }
    }
    if constexpr (TMAST && PAIR) {
      // the last stage writes the finished column into its half of the pair tile
      SmemIO<LayPair<C>> stp{so + d};
      run_stages<N, C, NT, true, true, true, false, LayCols<C>, 0>(s, tw, ncol, NullIO{}, stp, w);
    } else if constexpr (TMAST) {
      // planar (one GPU): T0 -> B[0], T1 -> B[1]; each finished tile drains while the next one is built
      float2* sb = d == 0 ? s : so;
      SmemIO<LayCols<C>> stl{sb};
      run_stages<N, C, NT, true, true, true, false, LayCols<C>, 0>(sb, tw, ncol, NullIO{}, stl, w);
      fence_async_smem();
      __syncthreads();
      if (threadIdx.x == 0) {
        const int rows = min(sl.lx, 256);
        for (int dst = 0; dst < sl.P; ++dst)
          for (int r0 = 0; r0 < sl.lx; r0 += rows)
            tma_store_4d(&tp.m[dst], 2 * kz0, yg, r0, d, sb + (size_t)(dst * sl.lx + r0) * C);
        tma_store_commit();
        if (d == 1) tma_store_wait_read();
      }
    } else if constexpr (PAIR) {
      ScatterXIO gout{&sl, 2 * ((long long)yg * sl.nzc + kz0) + d};
      run_stages<N, C, NT, true, true, true, false, LayCols<C>, 0>(s, tw, ncol, NullIO{}, gout, w);
    } else {
      const long long cs = (long long)sl.lx * sl.ny * sl.nzc;
      ScatterXPlanarIO gout{&sl, d * cs + (long long)yg * sl.nzc + kz0};
      run_stages<N, C, NT, true, true, true, false, LayCols<C>, 0>(s, tw, ncol, NullIO{}, gout, w);
    }
  }
  if constexpr (TMAST && PAIR) {
    // rows [dst lx, (dst+1) lx) of the pair tile (128 bytes each) leave as one tensor store per destination rank
    fence_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
      const int rows = min(sl.lx, 256);
      for (int dst = 0; dst < sl.P; ++dst)
        for (int r0 = 0; r0 < sl.lx; r0 += rows)
          tma_store_3d(&tp.m[dst], 4 * kz0, yg, r0, so + (size_t)(dst * sl.lx + r0) * C * 2);
      tma_store_commit();
      tma_store_wait_read();
    }
  }
}

// ---- X-pot: the x pass of the POTENTIAL chain -----------------------------------------------------------
// grid: (ntile, ly).  Forward FFT along x, delta_k (registers) times norm * G(k) / k^2, ONE inverse FFT along x:
//   psi = IFFT(g delta) = -phi  (pm.py:51-52 with the sign folded)  ->  B[0][x][yg][kz] of the rank that owns x.
// The reference's gradient kernel i (8 sin w - sin 2w) / 6 (kernels.py:62-66) is EXACTLY the symbol of the
// 4th-order central difference D f = [8 (f(x+1) - f(x-1)) - (f(x+2) - f(x-2))] / 12, so the three force meshes
// F_d = IFFT(i a_d g delta) = D_d psi are formed in real space by the read kernel (csrc/sim.cu) while it stages
// the tile box: one inverse transform instead of three, 40 instead of 72 B/cell through the five passes.
// `sumsq` (nullable) accumulates sum_k |psi_k|^2 over the full spectrum = mean_x psi^2 (Parseval), the
// numerator of the fp32 cancellation bound that decides between this chain and the three-transform one.
// KMODE 1 (linear_field, pm.py:134-143): the multiplier is amp(|k_phys|) * norm instead, k_phys^2 = sum_d (w_d s_d)^2,
// amp tabulated linearly in log10 k (`ftab`, `ntab` entries from col.lkmin in steps of 1 / col.inv), k = 0 -> col.dc.
struct KColour { float lkmin, inv, sx, sy, sz, dc; };
template <int N, int C, bool TMAST, int KMODE = 0>
__global__ void __launch_bounds__(threads_for<N, C, 8>(), (threads_for<N, C, 8>() <= 512 ? 2 : 1))
xpot_kernel(const __grid_constant__ Slab sl, const float2* __restrict__ twg, const float* __restrict__ wx,
            const float* __restrict__ wy, const float* __restrict__ wz, float norm, float r_split2,
            const float* __restrict__ ftab, int ntab, float fscale, const __grid_constant__ TmapPack tp,
            double* __restrict__ sumsq, const KColour col) {
  constexpr int NT = threads_for<N, C, 8>();
  constexpr int L = radix_count(N);
  constexpr int RL = radix_at(N, L - 1, false);
  constexpr int TASKS = (N / RL) * C;
  constexpr int TPT = (TASKS + NT - 1) / NT;
  extern __shared__ __align__(128) float2 sm[];
  float2* tw = sm;            // [N]
  float2* s = sm + N;         // [N][C]
  float* swx = reinterpret_cast<float*>(s + N * C);   // [N] k_x table
  __shared__ float s_part[32];
  for (int i = threadIdx.x; i < N; i += NT) { tw[i] = twg[i]; swx[i] = wx[i]; }
  const int kz0 = blockIdx.x * C;
  const int ncol = min(C, sl.nzh - kz0);
  const int yl = blockIdx.y, yg = sl.rank * sl.ly + yl;
  GlobalIO gin{sl.at[sl.rank] + (long long)yl * sl.nzc + kz0, (long long)sl.ly * sl.nzc};
  float2 keep[TPT][8];
  run_stages<N, C, NT, false, false, false, true, LayCols<C>, 0>(s, tw, ncol, gin, NullIO{}, keep);
  const float ky = wy[yg];
  float part = 0.f;
#pragma unroll
  for (int i = 0; i < TPT; ++i) {
    const int task = threadIdx.x + i * NT;
    const int c = task % C, j = task / C;
    if ((TASKS % NT == 0 || task < TASKS) && c < ncol) {
      const int kzi = kz0 + c;
      const float kz = wz[kzi];
      const float herm = (kzi == 0 || 2 * kzi == sl.nz) ? 1.f : 2.f;   // modes the half-spectrum stands for
#pragma unroll
      for (int r = 0; r < RL; ++r) {
        const float kx = swx[j + r * (N / RL)];
        float g;
        if constexpr (KMODE == 1) {
          const float px = kx * col.sx, py = ky * col.sy, pz = kz * col.sz;
          const float kk = (px * px + py * py) + pz * pz;
          float t = (kk > 0.f) ? (0.5f * log10f(kk) - col.lkmin) * col.inv : 0.f;
          t = fminf(fmaxf(t, 0.f), (float)(ntab - 1));
          const int ti = min((int)t, ntab - 2);
          const float fr = t - (float)ti;
          g = (kk > 0.f) ? (ftab[ti] + fr * (ftab[ti + 1] - ftab[ti])) * norm : col.dc * norm;
        } else {
          const float kxy2 = kx * kx + ky * ky;
          const float kk = kxy2 + kz * kz;
          g = (kk == 0.f) ? 0.f : __frcp_rn(kk);
          g *= norm;
          if (r_split2 != 0.f) g *= expf(-kk * r_split2);
          if (ftab) {
            const float t = sqrtf(kk) * fscale;
            const int ti = min((int)t, ntab - 2);
            const float fr = fminf(t - (float)ti, 1.0f);
            g *= ftab[ti] + fr * (ftab[ti + 1] - ftab[ti]);
          }
        }
        keep[i][r].x *= g;
        keep[i][r].y *= g;
        part = fmaf(herm, keep[i][r].x * keep[i][r].x + keep[i][r].y * keep[i][r].y, part);
      }
    }
  }
  if (sumsq) {
#pragma unroll
    for (int off = 16; off; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = part;
  }
  if constexpr (TMAST) {
    SmemIO<LayCols<C>> stl{s};
    run_stages<N, C, NT, true, true, true, false, LayCols<C>, 0>(s, tw, ncol, NullIO{}, stl, keep);
    fence_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
      const int rows = min(sl.lx, 256);
      for (int dst = 0; dst < sl.P; ++dst)
        for (int r0 = 0; r0 < sl.lx; r0 += rows)
          tma_store_4d(&tp.m[dst], 2 * kz0, yg, r0, 0, s + (size_t)(dst * sl.lx + r0) * C);
      tma_store_commit();
      tma_store_wait_read();
    }
  } else {
    ScatterXPlanarIO gout{&sl, (long long)yg * sl.nzc + kz0};
    run_stages<N, C, NT, true, true, true, false, LayCols<C>, 0>(s, tw, ncol, NullIO{}, gout, keep);
    __syncthreads();
  }
  if (sumsq && threadIdx.x == 0) {
    float tot = 0.f;
    for (int w = 0; w < NT / 32; ++w) tot += s_part[w];
    atomicAdd(sumsq, (double)tot);
  }
}

// ---- Y-pot: inverse FFT along y of ONE spectrum (B[0]) in place, the y pass of the potential chain -----------
// grid: (ntile, lx).  Same tiling as Y-fwd ([N][16] column tiles, 128-byte segments): the first stage loads from
// global memory, the last one stores to it; every CTA owns its columns of the plane, so in place is race free.
template <int N, int C>
__global__ void __launch_bounds__(threads_for<N, C>())
ypot_kernel(const __grid_constant__ Slab sl, const float2* __restrict__ twg, int x0) {
  constexpr int NT = threads_for<N, C>();
  static_assert(radix_count(N) >= 2, "in-place column pass needs the loads of a tile to precede its stores");
  extern __shared__ __align__(128) float2 sm[];
  float2* tw = sm;            // [N]
  float2* s = sm + N;         // [N][C]
  for (int i = threadIdx.x; i < N; i += NT) tw[i] = twg[i];
  const int kz0 = blockIdx.x * C, xl = x0 + blockIdx.y;
  const int ncol = min(C, sl.nzh - kz0);
  GlobalIO gio{sl.b3[sl.rank] + (long long)xl * sl.ny * sl.nzc + kz0, sl.nzc};
  run_stages<N, C, NT, true, false, false, false, LayCols<C>, 0>(s, tw, ncol, gio, gio, nullptr);
}

// ---- Y-inv: inverse FFT along y of the local x planes ---------------------------------------------------
// grid: (ntile, lx).  One 16-byte load per mode brings T0 and T1 (T01, written by X-fused); T0 -> F_x = B[0];
// T1 is transformed twice: as it is -> B[2] (F_z up to the factor i a_z(kz), applied by Z-inv) and times
// i a_y(ky) -> B[1] (F_y).  All loads of the tile are in flight before the first transform starts.
template <int N, int C, bool PAIR>
__global__ void __launch_bounds__(threads_for<N, C, 8>(), (threads_for<N, C, 8>() <= 512 ? 2 : 1))
yinv_kernel(const __grid_constant__ Slab sl, const float2* __restrict__ twg, const float* __restrict__ ay, int x0) {
  constexpr int NT = threads_for<N, C, 8>();
  constexpr int R0 = radix_at(N, 0, false);           // first inverse stage (forward radix order)
  constexpr int TASKS = (N / R0) * C;
  constexpr int TPT = (TASKS + NT - 1) / NT;
  extern __shared__ __align__(16) float2 sm[];
  float2* tw = sm;            // [N]
  float2* s = sm + N;         // [N][C]
  float* say = reinterpret_cast<float*>(s + N * C);   // [N] gradient table along y
  for (int i = threadIdx.x; i < N; i += NT) { tw[i] = twg[i]; say[i] = ay[i]; }
  const int kz0 = blockIdx.x * C, xl = x0 + blockIdx.y;
  const int ncol = min(C, sl.nzh - kz0);
  const long long cs = (long long)sl.lx * sl.ny * sl.nzc;
  const float4* in = PAIR ? sl.t01[sl.rank] + (long long)xl * sl.ny * sl.nzc + kz0 : nullptr;
  float2* base = sl.b3[sl.rank] + (long long)xl * sl.ny * sl.nzc + kz0;
  GlobalIO g0{base, sl.nzc}, g1{base + cs, sl.nzc}, g2{base + 2 * cs, sl.nzc};
  float2 keep[TPT][8];
  if constexpr (PAIR) {
    float2 w[TPT][8];
#pragma unroll
    for (int i = 0; i < TPT; ++i) {
      const int task = threadIdx.x + i * NT;
      const int c = task % C, j = task / C;
      if ((TASKS % NT == 0 || task < TASKS) && c < ncol) {
#pragma unroll
        for (int r = 0; r < R0; ++r) {
          const float4 v = __ldcs(in + (long long)(j + r * (N / R0)) * sl.nzc + c);
          w[i][r] = make_float2(v.x, v.y);
          keep[i][r] = make_float2(v.z, v.w);
        }
      }
    }
    run_stages<N, C, NT, true, false, true, false, LayCols<C>, 0>(s, tw, ncol, NullIO{}, g0, w);
  } else {
    // planar (one GPU): T0 = B[0] is transformed in place, T1 = B[1] is requested BEFORE that transform starts
#pragma unroll
    for (int i = 0; i < TPT; ++i) {
      const int task = threadIdx.x + i * NT;
      const int c = task % C, j = task / C;
      if ((TASKS % NT == 0 || task < TASKS) && c < ncol) {
#pragma unroll
        for (int r = 0; r < R0; ++r) keep[i][r] = g1(j + r * (N / R0), c);
      }
    }
    run_stages<N, C, NT, true, false, false, false, LayCols<C>, 0>(s, tw, ncol, g0, g0, nullptr);
  }
  __syncthreads();   // the T0 transform is done with `s` (its last stage reads it)
  {
    float2 w[TPT][8];
#pragma unroll
    for (int i = 0; i < TPT; ++i)
#pragma unroll
      for (int r = 0; r < R0; ++r) w[i][r] = keep[i][r];
    run_stages<N, C, NT, true, false, true, false, LayCols<C>, 0>(s, tw, ncol, NullIO{}, g2, w);
  }
  __syncthreads();
  {
    float2 w[TPT][8];
#pragma unroll
    for (int i = 0; i < TPT; ++i) {
      const int task = threadIdx.x + i * NT;
      const int j = task / C;
#pragma unroll
      for (int r = 0; r < R0; ++r) {
        const float a = say[(j + r * (N / R0)) & (N - 1)];
        w[i][r] = make_float2(-a * keep[i][r].y, a * keep[i][r].x);       // i a_y T1
      }
    }
    run_stages<N, C, NT, true, false, true, false, LayCols<C>, 0>(s, tw, ncol, NullIO{}, g1, w);
  }
}

// ---- Z-fwd: R2C along z of 16 rows, ghost zones folded while loading -------------------------------
// grid: (ny / 16, lx).  Real arrays are [lx + 2 gx][nyp][nzp]: gx ghost planes per side in x (images of the
// neighbour slabs; of this slab itself when P == 1), G ghost cells per side in y and z (periodic images).
constexpr int kRows = 16;
#ifndef JPM_ZPASS_CTAS
#define JPM_ZPASS_CTAS 5
#endif
constexpr int kZPassCtas = JPM_ZPASS_CTAS;   // resident CTAs per SM the z-FORWARD pass is register-capped for (48 regs:
                                             // 0.328 -> 0.316 ms; the same cap on the z-inverse pass loses 0.015 ms)
#ifndef JPM_ZINV_BATCH
#define JPM_ZINV_BATCH 16
#endif
constexpr int kZinvBatch = JPM_ZINV_BATCH;   // spectrum loads in flight per thread in the z-inverse pass
__device__ __forceinline__ int ghost_width(const Slab& sl);
// x plane of a z-pass block.  P > 1: planes next to the slab faces exchange ghosts with the neighbours over
// NVLink; an odd stride (lx is a power of two) spreads them over the whole launch, so that at any time most
// resident CTAs run at HBM speed while a few wait on the link, instead of one NVLink-bound phase.
__device__ __forceinline__ int z_pass_plane(const Slab& sl, int by) {
  return (sl.P > 1 && sl.lx >= 16) ? ((by * 37) & (sl.lx - 1)) : by;
}
// The (up to 3 x 3) arrays a row tile of a z pass touches on a pencil grid: the pencil that owns the rows, the x
// neighbours whose ghost planes image the plane, the y neighbours whose ghost rows image the rows, and the corners.
struct PencilTargets {
  int nxs, nys;
  int xr[3], xp[3];                 // grid row of the array, plane inside it
  int yc[3], yr[3], r0[3], r1[3];   // grid column, array row of tile row 0, tile rows [r0, r1) that exist there
};
// plane xl of this rank's slab, rows y0 .. y0 + 15 (global); ge = ghost planes / rows in use this step
__device__ __forceinline__ PencilTargets pencil_targets(const Slab& sl, int xl, int y0, int ge) {
  PencilTargets t;
  const int a = sl.rank / sl.py, b = sl.rank - a * sl.py;
  const int xq = b * sl.lx + xl;                          // plane inside pencil row a
  const int by = y0 / sl.Ly, yl0 = y0 - by * sl.Ly;       // column that owns the rows, first row inside it
  const int gex = min(sl.gx, ge), gey = min(sl.gy, ge);
  t.nxs = 0;
  t.xr[t.nxs] = a; t.xp[t.nxs++] = sl.gx + xq;
  if (xq < gex) { t.xr[t.nxs] = (a + sl.px - 1) % sl.px; t.xp[t.nxs++] = sl.gx + sl.Lx + xq; }            // its high ghost
  if (xq >= sl.Lx - gex) { t.xr[t.nxs] = (a + 1) % sl.px; t.xp[t.nxs++] = xq - (sl.Lx - sl.gx); }         // its low ghost
  t.nys = 0;
  t.yc[t.nys] = by; t.yr[t.nys] = sl.G + sl.gy + yl0; t.r0[t.nys] = 0; t.r1[t.nys++] = kRows;
  if (yl0 < gey) {
    t.yc[t.nys] = (by + sl.py - 1) % sl.py; t.yr[t.nys] = sl.G + sl.gy + sl.Ly + yl0;
    t.r0[t.nys] = 0; t.r1[t.nys++] = min(kRows, gey - yl0);
  }
  if (yl0 + kRows > sl.Ly - gey) {
    t.yc[t.nys] = (by + 1) % sl.py; t.yr[t.nys] = sl.G + sl.gy + yl0 - sl.Ly;
    t.r0[t.nys] = max(0, sl.Ly - gey - yl0); t.r1[t.nys++] = kRows;
  }
  return t;
}

template <int NZ, bool PEN = false>
__global__ void __launch_bounds__(threads_for<NZ / 2, kRows>(), (threads_for<NZ / 2, kRows>() <= 256 && !PEN ? kZPassCtas : 1))
zfwd_kernel(const __grid_constant__ Slab sl, const float2* __restrict__ twh, const float2* __restrict__ twfull, int x0) {
  constexpr int NH = NZ / 2, NT = threads_for<NH, kRows>();
  extern __shared__ __align__(16) float2 sm[];
  float2* tw = sm;            // [NH]
  float2* s = sm + NH;        // [16][NH + 1]
  for (int i = threadIdx.x; i < NH; i += NT) tw[i] = twh[i];
  const int xl = z_pass_plane(sl, x0 + blockIdx.y), y0 = blockIdx.x * kRows;
  const int G = sl.G, gx = sl.gx, lx = sl.lx, ny = sl.ny, nyp = sl.nyp, nzp = sl.nzp;
  const int GH = G / 2;       // ghost width in float2 units
  if constexpr (PEN) {
    // pencil grid: the rows live in the arrays of the row group (the row-group transpose happens in this load), ghost
    // planes / rows of the x / y / corner neighbours fold in as further sources; every load is branch-free per source
    constexpr int ITER = kRows * NH / NT, ZB = ITER < 8 ? ITER : 8;
    static_assert(kRows * NH % NT == 0 && ITER % ZB == 0, "row tile must divide evenly over the threads");
    const PencilTargets t = pencil_targets(sl, xl, y0, ghost_width(sl));
    const int rowp = nzp / 2;
    bool first = true;
    for (int ix = 0; ix < t.nxs; ++ix)
      for (int iy = 0; iy < t.nys; ++iy) {
        const float* plane = sl.dens[t.xr[ix] * sl.py + t.yc[iy]] + (long long)t.xp[ix] * nyp * nzp;
        const float2* base = reinterpret_cast<const float2*>(plane + (long long)t.yr[iy] * nzp) + GH;
        const int ra = t.r0[iy], rb = t.r1[iy];
#pragma unroll 1
        for (int b = 0; b < ITER; b += ZB) {
          float2 acc[ZB];
#pragma unroll
          for (int i = 0; i < ZB; ++i) {
            const int e = threadIdx.x + (b + i) * NT;
            const int r = e / NH, m = e - r * NH;
            acc[i] = (r >= ra && r < rb) ? __ldcs(base + r * rowp + m) : make_float2(0.f, 0.f);
          }
#pragma unroll
          for (int i = 0; i < ZB; ++i) {
            const int e = threadIdx.x + (b + i) * NT;
            const int r = e / NH, m = e - r * NH;
            float2 v = acc[i];
            if (!first) v = cadd(v, s[LayRows<NH>::idx(m, r)]);
            s[LayRows<NH>::idx(m, r)] = v;
          }
        }
        first = false;
      }
    __syncthreads();
    // z ghost cells of every source: high ghost -> first cells, low ghost -> last cells (2 GH pairs per row)
    for (int e = threadIdx.x; e < kRows * 2 * GH; e += NT) {
      const int r = e / (2 * GH), g = e - r * (2 * GH);
      const int m = g < GH ? g : NH - 2 * GH + g;
      const int off = g < GH ? NH : -NH;
      float2 v = s[LayRows<NH>::idx(m, r)];
      for (int ix = 0; ix < t.nxs; ++ix)
        for (int iy = 0; iy < t.nys; ++iy) {
          if (r < t.r0[iy] || r >= t.r1[iy]) continue;
          const float* plane = sl.dens[t.xr[ix] * sl.py + t.yc[iy]] + (long long)t.xp[ix] * nyp * nzp;
          const float2* base = reinterpret_cast<const float2*>(plane + (long long)t.yr[iy] * nzp) + GH;
          v = cadd(v, __ldcs(base + r * rowp + m + off));
        }
      s[LayRows<NH>::idx(m, r)] = v;
    }
  } else {
  // x planes that fold onto interior plane xl: this rank's own, the left neighbour's high ghost, the right
  // neighbour's low ghost
  const float* src[3] = {sl.dens[sl.rank] + (long long)(gx + xl) * nyp * nzp, nullptr, nullptr};
  const int ge = ghost_width(sl);   // planes xl < ge / xl >= lx - ge receive neighbour ghosts
  if (xl < ge) src[1] = sl.dens[(sl.rank + sl.P - 1) % sl.P] + (long long)(gx + lx + xl) * nyp * nzp;
  if (xl >= lx - ge) src[2] = sl.dens[(sl.rank + 1) % sl.P] + (long long)(xl - (lx - gx)) * nyp * nzp;
  constexpr int ITER = kRows * NH / NT, BATCH = ITER < 8 ? ITER : 8;
  static_assert(kRows * NH % NT == 0 && ITER % BATCH == 0, "row tile must divide evenly over the threads");
  // blocks away from the y faces: per source plane (own, left neighbour's high ghost, right neighbour's low
  // ghost) branch-free loads with BATCH in flight per thread - the neighbour planes are NVLink reads with
  // ~2 us latency, so the number of loads in flight is what sets their rate.  Every thread owns the same
  // elements in every pass: the accumulation into the tile needs no barrier.
  const bool y_interior = y0 >= G && y0 + kRows <= ny - G;
  if (y_interior) {
    const int rowp = nzp / 2;   // row pitch in float2
    constexpr int ZB = ITER < 16 ? ITER : 16;
    static_assert(ITER % ZB == 0, "row tile must divide evenly over the load batches");
#pragma unroll
    for (int ix = 0; ix < 3; ++ix) {
      if (!src[ix]) continue;
      const float2* base = reinterpret_cast<const float2*>(src[ix] + (long long)(y0 + G) * nzp) + GH;
#pragma unroll 1
      for (int b = 0; b < ITER; b += ZB) {
        float2 acc[ZB];
#pragma unroll
        for (int i = 0; i < ZB; ++i) {
          const int e = threadIdx.x + (b + i) * NT;
          const int r = e / NH, m = e - r * NH;
          acc[i] = __ldcs(base + r * rowp + m);
        }
#pragma unroll
        for (int i = 0; i < ZB; ++i) {
          const int e = threadIdx.x + (b + i) * NT;
          const int r = e / NH, m = e - r * NH;
          float2 v = acc[i];
          if (ix > 0) v = cadd(v, s[LayRows<NH>::idx(m, r)]);
          s[LayRows<NH>::idx(m, r)] = v;
        }
      }
    }
    __syncthreads();
    // z ghost cells of every source plane: high ghost -> first cells, low ghost -> last cells (2 GH pairs per row)
    for (int e = threadIdx.x; e < kRows * 2 * GH; e += NT) {
      const int r = e / (2 * GH), g = e - r * (2 * GH);
      const int m = g < GH ? g : NH - 2 * GH + g;            // destination pair
      const int off = g < GH ? NH : -NH;                      // its ghost image
      float2 v = s[LayRows<NH>::idx(m, r)];
#pragma unroll
      for (int ix = 0; ix < 3; ++ix) {
        if (!src[ix]) continue;
        const float2* base = reinterpret_cast<const float2*>(src[ix] + (long long)(y0 + G) * nzp) + GH;
        v = cadd(v, __ldcs(base + r * rowp + m + off));
      }
      s[LayRows<NH>::idx(m, r)] = v;
    }
  } else {
#pragma unroll 2
    for (int e = threadIdx.x; e < kRows * NH; e += NT) {
      const int r = e / NH, m = e - r * NH;
      const int y = y0 + r;
      int ys[2] = {y + G, -1};
      if (y < G) ys[1] = y + ny + G;
      else if (y >= ny - G) ys[1] = y - ny + G;
      float2 acc = make_float2(0.f, 0.f);
#pragma unroll
      for (int ix = 0; ix < 3; ++ix) {
        if (!src[ix]) continue;
#pragma unroll
        for (int iy = 0; iy < 2; ++iy) {
          if (ys[iy] < 0) continue;
          const float2* row = reinterpret_cast<const float2*>(src[ix] + (long long)ys[iy] * nzp);
          acc = cadd(acc, row[GH + m]);
          if (m < GH) acc = cadd(acc, row[NH + GH + m]);
          if (m >= NH - GH) acc = cadd(acc, row[m - (NH - GH)]);
        }
      }
      s[LayRows<NH>::idx(m, r)] = acc;
    }
  }
  }   // !PEN
  __syncthreads();
  SmemIO<LayRows<NH>> io{s};
  run_stages<NH, kRows, NT, false, false, false, false, LayRows<NH>, 0>(s, tw, kRows, io, io, nullptr);
  // untangle: X[k] = (Z[k] + conj Z[NH-k]) / 2 + W_NZ^k (Z[k] - conj Z[NH-k]) / (2i)   -> A_loc (= B[2] region)
  float2* aloc = sl.b3[sl.rank] + 2ll * sl.lx * ny * sl.nzc + ((long long)xl * ny + y0) * sl.nzc;
  for (int e = threadIdx.x; e < kRows * NH; e += NT) {
    const int r = e / NH, k = e - r * NH;
    const float2 zk = s[LayRows<NH>::idx(k, r)];
    const float2 zc = cconj(s[LayRows<NH>::idx((NH - k) & (NH - 1), r)]);
    const float2 ev = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y + zc.y));
    const float2 df = csub(zk, zc);
    const float2 od = make_float2(0.5f * df.y, -0.5f * df.x);       // (zk - zc) / (2i)
    const float2 w = __ldg(twfull + k);
    float2* dst = aloc + (long long)r * sl.nzc;
    __stcs(dst + k, cadd(ev, cmul(w, od)));
    if (k == 0) __stcs(dst + NH, make_float2(zk.x - zk.y, 0.f));
  }
}

// ---- Z-inv: C2R along z of 16 rows, ghost zones filled while storing -------------------------------
// grid: (ny / 16, lx, 3).  Output is unnormalised (the 1/Nc lives in the k-space factor).  Component 2
// still carries the factor i a_z(kz) (see X-fused / Y-inv).  Every row is written to this rank's interior
// plane, to its y / z ghost images and to the x ghost planes of the neighbour slabs that image it.
template <int NZ, bool PEN = false>
__global__ void __launch_bounds__(threads_for<NZ / 2, kRows>())
zinv_kernel(const __grid_constant__ Slab sl, const float2* __restrict__ twh, const float2* __restrict__ twfull,
            const float* __restrict__ az, int x0, int variant, int ge_extra, int to_psi) {
  constexpr int NH = NZ / 2, NT = threads_for<NH, kRows>();
  extern __shared__ __align__(16) float2 sm[];
  float2* tw = sm;            // [NH]
  float2* s = sm + NH;        // [16][NH + 1]
  for (int i = threadIdx.x; i < NH; i += NT) tw[i] = twh[i];
  const int xl = z_pass_plane(sl, x0 + blockIdx.y), y0 = blockIdx.x * kRows, comp = blockIdx.z;
  const int G = sl.G, gx = sl.gx, lx = sl.lx, ny = sl.ny, nyp = sl.nyp, nzp = sl.nzp, nzc = sl.nzc;
  const float2* src = sl.b3[sl.rank] + (long long)comp * lx * ny * nzc + ((long long)xl * ny + y0) * nzc;
  constexpr int ITER = kRows * NH / NT, BATCH = ITER < kZinvBatch ? ITER : kZinvBatch;
  static_assert(kRows * NH % NT == 0 && ITER % BATCH == 0, "row tile must divide evenly over the threads");
#pragma unroll 1
  for (int b = 0; b < ITER; b += BATCH) {
    float2 v[BATCH];
#pragma unroll
    for (int i = 0; i < BATCH; ++i) {
      const int e = threadIdx.x + (b + i) * NT;
      const int r = e / NH, k = e - r * NH;
      v[i] = __ldcs(src + (long long)r * nzc + k);
    }
#pragma unroll
    for (int i = 0; i < BATCH; ++i) {
      const int e = threadIdx.x + (b + i) * NT;
      const int r = e / NH, k = e - r * NH;
      if (comp == 2) {
        const float a = __ldg(az + k);
        v[i] = make_float2(-a * v[i].y, a * v[i].x);
      }
      s[LayRows<NH>::idx(k, r)] = v[i];
    }
  }
  if (threadIdx.x < kRows) {
    float2 v = __ldcs(src + (long long)threadIdx.x * nzc + NH);
    if (comp == 2) {
      const float a = __ldg(az + NH);
      v = make_float2(-a * v.y, a * v.x);
    }
    s[LayRows<NH>::idx(NH, threadIdx.x)] = v;
  }
  __syncthreads();
  // tangle pairs (k, NH - k): Z'[k] = (X[k] + conj X[NH-k]) + i conj(W^k) (X[k] - conj X[NH-k])
  for (int e = threadIdx.x; e < kRows * (NH / 2 + 1); e += NT) {
    const int r = e / (NH / 2 + 1), k = e - r * (NH / 2 + 1);
    const int kc = NH - k;
    const float2 xk = s[LayRows<NH>::idx(k, r)], xc = s[LayRows<NH>::idx(kc, r)];
    const float2 wk = cconj(__ldg(twfull + k));
    {
      const float2 sum = cadd(xk, cconj(xc)), dif = csub(xk, cconj(xc));
      const float2 t = cmul(wk, dif);
      s[LayRows<NH>::idx(k, r)] = make_float2(sum.x - t.y, sum.y + t.x);     // sum + i t
    }
    if (k != 0 && k != kc) {
      const float2 wc = cconj(__ldg(twfull + kc));
      const float2 sum = cadd(xc, cconj(xk)), dif = csub(xc, cconj(xk));
      const float2 t = cmul(wc, dif);
      s[LayRows<NH>::idx(kc, r)] = make_float2(sum.x - t.y, sum.y + t.x);
    }
  }
  __syncthreads();
  SmemIO<LayRows<NH>> io{s};
  run_stages<NH, kRows, NT, true, false, false, false, LayRows<NH>, 0>(s, tw, kRows, io, io, nullptr);
  // destinations in x: own interior plane, left neighbour's high ghost, right neighbour's low ghost
  const long long cofs = (long long)comp * sl.npad;
  float* const* const dmesh = to_psi ? sl.psi : sl.force;      // potential chain: the psi mesh (one component)
  if constexpr (PEN) {
    // pencil grid: the pencil that owns the rows (row-group transpose in this store) + the ghost images of the
    // x / y / corner neighbours
    const PencilTargets t = pencil_targets(sl, xl, y0, ghost_width(sl) + ge_extra);
    const int GHp = G / 2;
    for (int ix = 0; ix < t.nxs; ++ix)
      for (int iy = 0; iy < t.nys; ++iy) {
        float* plane = dmesh[t.xr[ix] * sl.py + t.yc[iy]] + cofs + (long long)t.xp[ix] * nyp * nzp;
        const int ra = t.r0[iy], rb = t.r1[iy];
        for (int e = threadIdx.x + ra * NH; e < rb * NH; e += NT) {
          const int r = e / NH, m = e - r * NH;
          const float2 v = s[LayRows<NH>::idx(m, r)];
          float2* row = reinterpret_cast<float2*>(plane + (long long)(t.yr[iy] + r) * nzp);
          row[GHp + m] = v;
          if (m < GHp) row[NH + GHp + m] = v;
          if (m >= NH - GHp) row[m - (NH - GHp)] = v;
        }
      }
    return;
  }
  float* dstp[3] = {dmesh[sl.rank] + cofs + (long long)(gx + xl) * nyp * nzp, nullptr, nullptr};
  // ge_extra: the potential chain's mesh is differentiated by a +-2 stencil after the read box is staged
  const int ge = min(gx, ghost_width(sl) + ge_extra);
  if (xl < ge) dstp[1] = dmesh[(sl.rank + sl.P - 1) % sl.P] + cofs + (long long)(gx + lx + xl) * nyp * nzp;
  if (xl >= lx - ge) dstp[2] = dmesh[(sl.rank + 1) % sl.P] + cofs + (long long)(xl - (lx - gx)) * nyp * nzp;
  const int GH = G / 2;
  if ((variant & 1) && !dstp[1] && !dstp[2] && y0 >= G && y0 + kRows <= ny - G) {
    // block without x / y images to write: straight-line, 8 shared-memory loads in flight, then 8 row stores
    static_assert(NT % NH == 0, "threads per CTA must be a multiple of the row length");
    constexpr int RPP = NT / NH, RPT = kRows / RPP, RB = RPT < 8 ? RPT : 8;
    const int m = threadIdx.x % NH, rr = threadIdx.x / NH;
    const bool zhi = m < GH, zlo = m >= NH - GH;
#pragma unroll 1
    for (int b = 0; b < RPT; b += RB) {
      float2 v[RB];
#pragma unroll
      for (int i = 0; i < RB; ++i) v[i] = s[LayRows<NH>::idx(m, rr + (b + i) * RPP)];
#pragma unroll
      for (int i = 0; i < RB; ++i) {
        float2* row = reinterpret_cast<float2*>(dstp[0] + (long long)(y0 + rr + (b + i) * RPP + G) * nzp);
        row[GH + m] = v[i];
        if (zhi) row[NH + GH + m] = v[i];
        if (zlo) row[m - (NH - GH)] = v[i];
      }
    }
    return;
  }
  for (int e = threadIdx.x; e < kRows * NH; e += NT) {
    const int r = e / NH, m = e - r * NH;
    const int y = y0 + r;
    int ys[2] = {y + G, -1};
    if (y < G) ys[1] = y + ny + G;
    else if (y >= ny - G) ys[1] = y - ny + G;
    const float2 v = s[LayRows<NH>::idx(m, r)];
#pragma unroll
    for (int ix = 0; ix < 3; ++ix) {
      if (!dstp[ix]) continue;
#pragma unroll
      for (int iy = 0; iy < 2; ++iy) {
        if (ys[iy] < 0) continue;
        float2* row = reinterpret_cast<float2*>(dstp[ix] + (long long)ys[iy] * nzp);
        row[GH + m] = v;
        if (m < GH) row[NH + GH + m] = v;
        if (m >= NH - GH) row[m - (NH - GH)] = v;
      }
    }
  }
}


// ---- gradient pass of the potential chain -----------------------------------------------------------------
// F_d = D_d psi with the 4th-order central difference the reference's gradient kernel is the symbol of
// (kernels.py:62-66; see X-pot above): psi mesh (ghosts filled) -> the three force meshes INCLUDING their ghost
// cells, so that the tile boxes of the read kernel never wrap.  One thread per 4 consecutive z cells; neighbours
// in y and z are periodic images inside the padded array (index -/+ n), in x too when P == 1; the x planes of a
// slab beyond what the neighbours filled hold nothing a particle within the halo reach reads.
// HBM: 4 B/cell read + 12 B/cell written; the +-2 planes / rows are L2 / L1 hits.
#ifndef JPM_GRAD_PLANES
#define JPM_GRAD_PLANES 8
#endif
#ifndef JPM_GRAD_MINCTAS
#define JPM_GRAD_MINCTAS 4   // register cap for 4 resident CTAs per SM (64 registers instead of 78): 0.489 -> 0.414 ms
#endif                       // at 512^3; 5 (48 registers) 0.423, 6 (40, spills) 0.61; 16 planes per thread: no gain
constexpr int kGradPlanes = JPM_GRAD_PLANES;      // x planes per thread: the +-2 x neighbours slide through registers
__global__ void __launch_bounds__(256, JPM_GRAD_MINCTAS)
fdgrad_kernel(const __grid_constant__ Slab sl, int x_lo, int x_hi, unsigned* __restrict__ fmax_bits) {
  const int nzp = sl.nzp, nyp = sl.nyp, nz4 = nzp / 4;
  const int z4 = blockIdx.x * 32 + (threadIdx.x & 31);
  const int yp = blockIdx.y * 8 + (threadIdx.x >> 5);
  int xb = x_lo + blockIdx.z * kGradPlanes, xe = min(xb + kGradPlanes, x_hi);
  const int nxl = sl.Lx + 2 * sl.gx;                 // planes the FFT kernels index (P == 1: nx + 2 G)
  const bool pen = sl.py > 1;
  if (sl.P > 1) {
    // planes the particles of this step reach: the slab + ghost_width planes per side (+-2 more hold psi)
    const int ge = ghost_width(sl);
    xb = max(xb, sl.gx - min(ge, sl.gx));
    xe = min(xe, sl.gx + sl.Lx + min(ge, sl.gx));
    if (pen) {   // same for the rows of a pencil
      const int gey = min(ge, sl.gy);
      if (yp < sl.G + sl.gy - gey || yp >= sl.G + sl.gy + sl.Ly + gey) return;
    }
  }
  if (z4 >= nz4 || yp >= nyp || xb >= xe) return;   // lanes run along z: a warp exits (or shrinks) as a unit per row
  const float4* __restrict__ ps = reinterpret_cast<const float4*>(sl.psi[sl.rank]);
  const long long sx4 = (long long)nyp * nz4;        // plane stride in float4
  // periodic images inside the padded array: index -/+ n (x too when P == 1; a slab's outermost planes hold
  // nothing a particle within the halo reach reads, clamp there)
  // (rows of a pencil: ghost rows come from the y neighbours, nothing wraps - clamp like the x planes of a slab)
  const int ym1 = pen ? max(yp - 1, 0) : (yp >= 1 ? yp - 1 : yp - 1 + sl.ny);
  const int ym2 = pen ? max(yp - 2, 0) : (yp >= 2 ? yp - 2 : yp - 2 + sl.ny);
  const int yp1 = pen ? min(yp + 1, nyp - 1) : (yp + 1 < nyp ? yp + 1 : yp + 1 - sl.ny);
  const int yp2 = pen ? min(yp + 2, nyp - 1) : (yp + 2 < nyp ? yp + 2 : yp + 2 - sl.ny);
  const int zl = z4 > 0 ? z4 - 1 : z4 - 1 + sl.nz / 4, zh = z4 + 1 < nz4 ? z4 + 1 : z4 + 1 - sl.nz / 4;
  auto wrapx = [&](int i) {
    if (sl.P == 1) return i < 0 ? i + sl.nx : (i >= nxl ? i - sl.nx : i);
    return min(max(i, 0), nxl - 1);
  };
  const long long o_c = (long long)yp * nz4 + z4;
  const long long o_ym1 = (long long)ym1 * nz4 + z4, o_ym2 = (long long)ym2 * nz4 + z4;
  const long long o_yp1 = (long long)yp1 * nz4 + z4, o_yp2 = (long long)yp2 * nz4 + z4;
  const long long o_zl = (long long)yp * nz4 + zl, o_zh = (long long)yp * nz4 + zh;
  constexpr float c8 = 2.0f / 3.0f, c1 = 1.0f / 12.0f;
  float4 x_2 = ps[wrapx(xb - 2) * sx4 + o_c], x_1 = ps[wrapx(xb - 1) * sx4 + o_c];
  float4 c = ps[xb * sx4 + o_c], x1 = ps[wrapx(xb + 1) * sx4 + o_c];
  float m = 0.f;
  float* outp = sl.force[sl.rank] + 4 * ((long long)xb * sx4 + o_c);
  for (int xp = xb; xp < xe; ++xp) {
    const float4* pl = ps + xp * sx4;
    const float4 x2 = ps[wrapx(xp + 2) * sx4 + o_c];
    const float4 y1 = pl[o_yp1], y_1 = pl[o_ym1], y2 = pl[o_yp2], y_2 = pl[o_ym2];
    const float4 a = pl[o_zl], b = pl[o_zh];
    float4 fx, fy, fz;
    fx.x = c8 * (x1.x - x_1.x) - c1 * (x2.x - x_2.x); fx.y = c8 * (x1.y - x_1.y) - c1 * (x2.y - x_2.y);
    fx.z = c8 * (x1.z - x_1.z) - c1 * (x2.z - x_2.z); fx.w = c8 * (x1.w - x_1.w) - c1 * (x2.w - x_2.w);
    fy.x = c8 * (y1.x - y_1.x) - c1 * (y2.x - y_2.x); fy.y = c8 * (y1.y - y_1.y) - c1 * (y2.y - y_2.y);
    fy.z = c8 * (y1.z - y_1.z) - c1 * (y2.z - y_2.z); fy.w = c8 * (y1.w - y_1.w) - c1 * (y2.w - y_2.w);
    fz.x = c8 * (c.y - a.w) - c1 * (c.z - a.z);
    fz.y = c8 * (c.z - c.x) - c1 * (c.w - a.w);
    fz.z = c8 * (c.w - c.y) - c1 * (b.x - c.x);
    fz.w = c8 * (b.x - c.z) - c1 * (b.y - c.y);
    __stcs(reinterpret_cast<float4*>(outp), fx);
    __stcs(reinterpret_cast<float4*>(outp + sl.npad), fy);
    __stcs(reinterpret_cast<float4*>(outp + 2 * sl.npad), fz);
    outp += 4 * sx4;
    m = fmaxf(m, fmaxf(fmaxf(fmaxf(fabsf(fx.x), fabsf(fx.y)), fmaxf(fabsf(fx.z), fabsf(fx.w))),
                       fmaxf(fmaxf(fabsf(fy.x), fabsf(fy.y)), fmaxf(fabsf(fy.z), fabsf(fy.w)))));
    m = fmaxf(m, fmaxf(fmaxf(fabsf(fz.x), fabsf(fz.y)), fmaxf(fabsf(fz.z), fabsf(fz.w))));
    x_2 = x_1; x_1 = c; c = x1; x1 = x2;
  }
  if (fmax_bits) {
    // largest force component on the mesh: denominator of the fp32 cancellation bound (csrc/sim.cu, AUTO mode)
    const unsigned act = __activemask();
    const unsigned mb = __reduce_max_sync(act, __float_as_uint(m));
    if ((threadIdx.x & 31) == (__ffs(act) - 1) && mb > *fmax_bits) atomicMax(fmax_bits, mb);
  }
}

// ---- inter-GPU barrier over peer-mapped flags ------------------------------------------------------------
// flags[r] (on every rank) has one slot per peer; rank `me` writes `epoch` into slot `me` of every peer and
// waits until its own slots all reach `epoch`.  One CTA, one thread per peer.  A lost peer trips the
// timeout (~4 s) and raises the error word instead of hanging the GPU.
template <bool GHOSTW>
__global__ void slab_barrier_kernel(const __grid_constant__ Slab sl, unsigned epoch, int reach_extra,
                                    long long timeout_cycles) {
  const int t = threadIdx.x;
  __threadfence_system();
  if (t < sl.P) {
    if (GHOSTW) {
      // ghost planes this rank's particles reach (sim_paint_kernel recorded the touched x range); every
      // peer gets the number, the step then uses the maximum over the ranks.  Unknown range -> all gx.
      const int* mine = reinterpret_cast<const int*>(sl.flags[sl.rank]);
      const int xmin = mine[kFlagXmin], xmax = mine[kFlagXmax];
      int need = max(sl.gx, sl.gy);
      if (reach_extra <= -1000) {                    // JPM_SLAB_GHOST_OVERRIDE=n (timing experiments only): exchange n planes
        need = min(need, -reach_extra - 1000);
      } else if (xmin <= xmax && reach_extra < 1000) {   // reach_extra >= 1000 (JPM_SLAB_FULL_GHOST=1): every ghost plane
        const int raw = max(0, max(sl.gx - xmin, xmax - (sl.gx + sl.Lx - 1)));
        need = min(sl.gx, raw);
        // the outermost ghost plane was touched: some particle is at (or wrapped past) the reach of the halo
        if (raw + reach_extra >= sl.gx && t == 0) sl.flags[sl.rank][kFlagReach] = 1u;
        if (sl.py > 1) {   // pencil grid: the rows too; one width serves both axes (clamped per axis by its users)
          const int ymin = mine[kFlagYmin], ymax = mine[kFlagYmax];
          const int rawy = (ymin <= ymax) ? max(0, max(sl.gy - ymin, ymax - (sl.gy + sl.Ly - 1))) : sl.gy;
          need = max(need, min(sl.gy, rawy));
          if (rawy + reach_extra >= sl.gy && t == 0) sl.flags[sl.rank][kFlagReach] = 1u;
        }
      }
      asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(sl.flags[t] + kFlagGeSlots + sl.rank), "r"(need) : "memory");
    }
    unsigned* remote = sl.flags[t] + sl.rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(epoch) : "memory");
    const unsigned* mine = sl.flags[sl.rank] + t;
    unsigned v;
    long long t0 = clock64();
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
      if ((int)(v - epoch) >= 0) break;
      if (clock64() - t0 > timeout_cycles) {   // JPM_SLAB_TIMEOUT_S (default 4 s): a lost peer must not hang the GPU
        sl.flags[sl.rank][kFlagErr] = 1u;    // error word
        break;
      }
    } while (true);
  }
  __syncthreads();
  if (GHOSTW && t == 0) {
    int ge = 0;
    for (int r = 0; r < sl.P; ++r) {
      unsigned v;
      asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(sl.flags[sl.rank] + kFlagGeSlots + r) : "memory");
      ge = max(ge, (int)v);
    }
    int* mine = reinterpret_cast<int*>(sl.flags[sl.rank]);
    mine[kFlagGe] = min(ge, max(sl.gx, sl.gy));
    mine[kFlagXmin] = 0x7fffffff;          // the next paint starts a new range
    mine[kFlagXmax] = (int)0x80000000;
    mine[kFlagYmin] = 0x7fffffff;
    mine[kFlagYmax] = (int)0x80000000;
  }
  __threadfence_system();
}

// this rank's force statistics -> slot `slot` of every rank's flag block (exact: fixed-point u64 add, u32 max)
__global__ void slab_stats_kernel(const __grid_constant__ Slab sl, const double* __restrict__ stats, int slot) {
  const int r = threadIdx.x;
  if (r >= sl.P) return;
  const double q = fmin(fmax(stats[0] * kStatsFix, 0.0), 9.0e18);
  const unsigned fb = reinterpret_cast<const unsigned*>(stats + 1)[0];
  unsigned* dst = sl.flags[r] + kFlagStats + 4 * slot;
  atomicAdd_system(reinterpret_cast<unsigned long long*>(dst), (unsigned long long)q);
  atomicMax_system(dst + 2, fb);
}

// ghost planes per side in use this step (see slab_barrier_kernel); P == 1: the periodic images, always gx
__device__ __forceinline__ int ghost_width(const Slab& sl) {
  if (sl.P == 1) return sl.gx;
  return min(max(sl.gx, sl.gy), reinterpret_cast<const int*>(sl.flags[sl.rank])[kFlagGe]);
}

static void make_twiddles(int n, int count, std::vector<float2>& out) {
  out.resize(count);
  for (int k = 0; k < count; ++k) {
    const double a = -2.0 * M_PI * (double)k / (double)n;
    out[k] = make_float2((float)std::cos(a), (float)std::sin(a));
  }
}

static int32_t upload2(float2** dst, const std::vector<float2>& v) {
  JPM_CUDA(cudaMalloc(dst, v.size() * sizeof(float2)));
  JPM_CUDA(cudaMemcpy(*dst, v.data(), v.size() * sizeof(float2), cudaMemcpyHostToDevice));
  return JPM_OK;
}

static bool pow2_in_range(int n) { return n >= 16 && n <= 1024 && (n & (n - 1)) == 0; }

template <int N, int C> constexpr size_t cols_smem() { return (size_t)(N + N * C) * sizeof(float2); }
template <int N, int C> constexpr size_t xfused_smem() { return (size_t)(N + N * C) * sizeof(float2) + 2 * N * sizeof(float); }
template <int N, int C> constexpr size_t xpot_smem() { return (size_t)(N + N * C) * sizeof(float2) + N * sizeof(float); }
template <int N, int C> constexpr size_t xfused_smem_tma() { return xfused_smem<N, C>() + 2 * (size_t)N * C * sizeof(float2); }
template <int N, int C> constexpr size_t xfused_smem_tma_planar() { return xfused_smem<N, C>() + (size_t)N * C * sizeof(float2); }
template <int NZ> constexpr size_t z_smem() { return (size_t)(NZ / 2 + kRows * (NZ / 2 + 1)) * sizeof(float2); }

constexpr int kColsC = 16;   // kz columns per tile of the Y-fwd pass across NVLink (P > 1): 128-byte remote rows
constexpr int kColsC1 = 8;   // ... on one GPU: [N][8] tiles (32 KB, 256 threads) interleave better, 0.258 -> 0.237 ms at 512^3
constexpr int kXC = 8;       // ... of the X-fused and Y-inv passes (64-byte segments; data held in registers)
constexpr int kXCW = 16;     // ... of the potential chain's x pass across NVLink (P > 1): 128-byte remote rows
constexpr int kYPC = 8;      // ... of the potential chain's y-inverse pass (local on every rank): 0.244 -> 0.224 ms; 4: 0.318, 32: 0.280

#define JPM_FFT_SWITCH(n, MACRO)          \
  switch (n) {                            \
    case 16: MACRO(16); break;            \
    case 32: MACRO(32); break;            \
    case 64: MACRO(64); break;            \
    case 128: MACRO(128); break;          \
    case 256: MACRO(256); break;          \
    case 512: MACRO(512); break;          \
    case 1024: MACRO(1024); break;        \
    default: set_error("pmfft: unsupported size %d", n); return JPM_ERR_INVALID; \
  }

static int32_t set_attrs(const Slab& sl) {
#define ATTR_Y(N_)                                                                                              \
  JPM_CUDA(cudaFuncSetAttribute(yfwd_kernel<N_, kColsC, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                (int)cols_smem<N_, kColsC>()));                                                 \
  JPM_CUDA(cudaFuncSetAttribute(yfwd_kernel<N_, kColsC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                (int)cols_smem<N_, kColsC>()));                                                 \
  JPM_CUDA(cudaFuncSetAttribute(yfwd_kernel<N_, kColsC1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                (int)cols_smem<N_, kColsC1>()));                                                \
  JPM_CUDA(cudaFuncSetAttribute(yfwd_kernel<N_, kColsC1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                (int)cols_smem<N_, kColsC1>()));                                                \
  JPM_CUDA(cudaFuncSetAttribute(yinv_kernel<N_, kXC, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                                (int)xfused_smem<N_, kXC>()));                                                  \
  JPM_CUDA(cudaFuncSetAttribute(yinv_kernel<N_, kXC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                                (int)xfused_smem<N_, kXC>()));
  JPM_FFT_SWITCH(sl.ny, ATTR_Y)
#undef ATTR_Y
#define ATTR_YP(N_)                                                                                             \
  JPM_CUDA(cudaFuncSetAttribute(ypot_kernel<N_, kYPC>, cudaFuncAttributeMaxDynamicSharedMemorySize,             \
                                (int)cols_smem<N_, kYPC>()));
  JPM_FFT_SWITCH(sl.ny, ATTR_YP)
#undef ATTR_YP
#define ATTR_XP(N_)                                                                                             \
  JPM_CUDA(cudaFuncSetAttribute(xpot_kernel<N_, kXC, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                                (int)xpot_smem<N_, kXC>()));                                                    \
  JPM_CUDA(cudaFuncSetAttribute(xpot_kernel<N_, kXC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                                (int)xpot_smem<N_, kXC>()));                                                    \
  JPM_CUDA(cudaFuncSetAttribute(xpot_kernel<N_, kXC, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                (int)xpot_smem<N_, kXC>()));
  JPM_FFT_SWITCH(sl.nx, ATTR_XP)
#undef ATTR_XP
  if (sl.nx <= 512) {
#define ATTR_XPW(N_)                                                                                            \
  if constexpr (N_ <= 512)                                                                                      \
    JPM_CUDA(cudaFuncSetAttribute(xpot_kernel<N_, kXCW, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                  (int)xpot_smem<N_, kXCW>()));
    JPM_FFT_SWITCH(sl.nx, ATTR_XPW)
#undef ATTR_XPW
  }
#define ATTR_X(N_)                                                                                              \
  JPM_CUDA(cudaFuncSetAttribute(xfused_kernel<N_, kXC, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                (int)xfused_smem<N_, kXC>()));                                                  \
  JPM_CUDA(cudaFuncSetAttribute(xfused_kernel<N_, kXC, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                (int)xfused_smem<N_, kXC>()));                                                  \
  JPM_CUDA(cudaFuncSetAttribute(xfused_kernel<N_, kXC, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                (int)xfused_smem_tma_planar<N_, kXC>()));                                              \
  JPM_CUDA(cudaFuncSetAttribute(xfused_kernel<N_, kXC, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                (int)xfused_smem_tma<N_, kXC>()));                                              \
  JPM_CUDA(cudaFuncSetAttribute(xfused_kernel<N_, kXC, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                (int)xfused_smem_tma_planar<N_, kXC>()));                                       \
  JPM_CUDA(cudaFuncSetAttribute(xfused_kernel<N_, kXC, true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                (int)xfused_smem_tma<N_, kXC>()));
  JPM_FFT_SWITCH(sl.nx, ATTR_X)
#undef ATTR_X
#define ATTR_Z(N_)                                                                                              \
  JPM_CUDA(cudaFuncSetAttribute(zfwd_kernel<N_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)z_smem<N_>())); \
  JPM_CUDA(cudaFuncSetAttribute(zinv_kernel<N_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)z_smem<N_>())); \
  JPM_CUDA(cudaFuncSetAttribute(zfwd_kernel<N_, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)z_smem<N_>())); \
  JPM_CUDA(cudaFuncSetAttribute(zinv_kernel<N_, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)z_smem<N_>()));
  JPM_FFT_SWITCH(sl.nz, ATTR_Z)
#undef ATTR_Z
  return JPM_OK;
}

}  // namespace fft

bool pmfft_shape_ok(int nx, int ny, int nz) {
  return fft::pow2_in_range(nx) && fft::pow2_in_range(ny) && fft::pow2_in_range(nz);
}

// Twiddle tables + kernel attributes for the slab geometry in p->slab (arrays already allocated).
int32_t pmfft_setup(jpm_plan* p) {
  const Slab& sl = p->slab;
  std::vector<float2> t;
  int32_t rc;
  fft::make_twiddles(sl.nx, sl.nx, t);
  if ((rc = fft::upload2(&p->tw_x, t))) return rc;
  fft::make_twiddles(sl.ny, sl.ny, t);
  if ((rc = fft::upload2(&p->tw_y, t))) return rc;
  fft::make_twiddles(sl.nz / 2, sl.nz / 2, t);
  if ((rc = fft::upload2(&p->tw_zh, t))) return rc;
  fft::make_twiddles(sl.nz, sl.nz / 2 + 1, t);
  if ((rc = fft::upload2(&p->tw_zfull, t))) return rc;
  // load both barrier kernels NOW: with lazy module loading the first launch of a kernel can block the host
  // until the device is idle, which never happens while this rank's previous barrier is still spinning for a
  // peer driven by the same host thread (several ranks in one process)
  // tensor maps of every rank's AT / B3 for the TMA-store flavour (JPM_FFT_TMASTORE=0 keeps plain stores)
  {
    const char* env = getenv("JPM_FFT_TMASTORE");
    p->fft_tma_store = !(env && env[0] == '0') && sl.lx >= 2 && sl.ly >= 2 &&
                       (sl.lx <= 256 || sl.lx % 256 == 0) && (sl.ly <= 256 || sl.ly % 256 == 0);
    // interleaved (T0, T1) output of the x pass when it crosses NVLink (128-byte rows, one store per mode pair);
    // planar on one GPU, where the per-transform drain overlaps better (JPM_FFT_PAIR=0|1 overrides)
    p->fft_pair = sl.P > 1;
    if (const char* e = getenv("JPM_FFT_PAIR")) p->fft_pair = e[0] == '1';
    if (p->fft_pair && !sl.t01[sl.rank]) {
      set_error("pmfft: the pair layout needs the T01 buffer");
      return JPM_ERR_INVALID;
    }
    if (const char* e = getenv("JPM_FFT_CHUNK")) p->fft_chunk = atoi(e);
    if (const char* e = getenv("JPM_FFT_ZVAR")) p->fft_zvariant = atoi(e);
    if (p->fft_tma_store) {
      if (!p->tm_at) p->tm_at = new TmapPack();
      if (!p->tm_t01) p->tm_t01 = new TmapPack();
      if (!p->tm_b3) p->tm_b3 = new TmapPack();
      if (!p->tm_b3w) p->tm_b3w = new TmapPack();
      memset(p->tm_at, 0, sizeof(TmapPack));
      memset(p->tm_t01, 0, sizeof(TmapPack));
      memset(p->tm_b3, 0, sizeof(TmapPack));
      memset(p->tm_b3w, 0, sizeof(TmapPack));
      const unsigned long long row = (unsigned long long)sl.nzc * sizeof(float2);
      for (int d = 0; d < sl.P; ++d) {
        const unsigned long long da[3] = {2ull * sl.nzc, (unsigned long long)sl.ly, (unsigned long long)sl.nx};
        const unsigned long long sa[2] = {row, row * sl.ly};
        const unsigned ba[3] = {2u * (sl.P == 1 ? fft::kColsC1 : fft::kColsC), (unsigned)std::min(sl.ly, 256), 1u};
        if ((rc = encode_tensor_map(&p->tm_at->m[d], reinterpret_cast<float*>(sl.at[d]), 3, da, sa, ba))) return rc;
        {   // planar B3 of rank d (the potential chain stores its single spectrum into component 0)
          const unsigned long long db[4] = {2ull * sl.nzc, (unsigned long long)sl.ny, (unsigned long long)sl.lx, 3ull};
          const unsigned long long sb[3] = {row, row * sl.ny, row * sl.ny * sl.lx};
          const unsigned bb[4] = {2u * fft::kXC, 1u, (unsigned)std::min(sl.lx, 256), 1u};
          if ((rc = encode_tensor_map(&p->tm_b3->m[d], reinterpret_cast<float*>(sl.b3[d]), 4, db, sb, bb))) return rc;
          const unsigned bw[4] = {2u * fft::kXCW, 1u, (unsigned)std::min(sl.lx, 256), 1u};
          if ((rc = encode_tensor_map(&p->tm_b3w->m[d], reinterpret_cast<float*>(sl.b3[d]), 4, db, sb, bw))) return rc;
        }
        if (p->fft_pair) {
          const unsigned long long dt[3] = {4ull * sl.nzc, (unsigned long long)sl.ny, (unsigned long long)sl.lx};
          const unsigned long long st[2] = {2 * row, 2 * row * sl.ny};
          const unsigned bt[3] = {4u * fft::kXC, 1u, (unsigned)std::min(sl.lx, 256)};
          if ((rc = encode_tensor_map(&p->tm_t01->m[d], reinterpret_cast<float*>(sl.t01[d]), 3, dt, st, bt))) return rc;
        } else {
          const unsigned long long db[4] = {2ull * sl.nzc, (unsigned long long)sl.ny, (unsigned long long)sl.lx, 3ull};
          const unsigned long long sb[3] = {row, row * sl.ny, row * sl.ny * sl.lx};
          const unsigned bb[4] = {2u * fft::kXC, 1u, (unsigned)std::min(sl.lx, 256), 1u};
          if ((rc = encode_tensor_map(&p->tm_t01->m[d], reinterpret_cast<float*>(sl.b3[d]), 4, db, sb, bb))) return rc;
        }
      }
    }
  }
  cudaFuncAttributes fa;
  JPM_CUDA(cudaFuncGetAttributes(&fa, fft::slab_barrier_kernel<true>));
  JPM_CUDA(cudaFuncGetAttributes(&fa, fft::slab_barrier_kernel<false>));
  JPM_CUDA(cudaFuncGetAttributes(&fa, fft::slab_stats_kernel));
  JPM_CUDA(cudaFuncGetAttributes(&fa, fft::fdgrad_kernel));
  return fft::set_attrs(sl);
}

// Single-GPU plan: allocate AT / B3 and describe the plan's ghost-zone meshes as a one-rank slab.
int32_t pmfft_enable(jpm_plan* p) {
  if (p->fft_on) return JPM_OK;
  if (!(pmfft_shape_ok(p->nx, p->ny, p->nz) && p->G > 0 && (p->G % 2) == 0)) return JPM_OK;
  Slab& sl = p->slab;
  memset(&sl, 0, sizeof(sl));
  sl.P = 1; sl.rank = 0;
  sl.nx = p->nx; sl.ny = p->ny; sl.nz = p->nz;
  sl.lx = p->nx; sl.ly = p->ny; sl.gx = p->G; sl.G = p->G;
  sl.px = sl.py = 1; sl.Lx = p->nx; sl.Ly = p->ny; sl.gy = 0;
  sl.nxp = p->nxp; sl.nyp = p->nyp; sl.nzp = p->nzp; sl.npad = p->npad;
  sl.nzh = p->nzh; sl.nzc = (p->nzh + 7) & ~7;
  const long long na = (long long)sl.nx * sl.ny * sl.nzc;
  JPM_CUDA(cudaMalloc(&p->fft_at, na * sizeof(float2)));
  JPM_CUDA(cudaMalloc(&p->fft_b3, 3 * na * sizeof(float2)));
  JPM_CUDA(cudaMemset(p->fft_at, 0, na * sizeof(float2)));
  JPM_CUDA(cudaMemset(p->fft_b3, 0, 3 * na * sizeof(float2)));
  {
    const char* e = getenv("JPM_FFT_PAIR");     // one GPU uses the planar layout unless asked otherwise
    if (e && e[0] == '1') {
      JPM_CUDA(cudaMalloc(&p->fft_t01, na * sizeof(float4)));
      JPM_CUDA(cudaMemset(p->fft_t01, 0, na * sizeof(float4)));
    }
  }
  sl.dens[0] = p->density_p; sl.force[0] = p->force3_p; sl.at[0] = p->fft_at; sl.b3[0] = p->fft_b3; sl.t01[0] = p->fft_t01;
  int32_t rc = pmfft_setup(p);
  if (rc) return rc;
  p->fft_on = true;
  return JPM_OK;
}

void pmfft_destroy(jpm_plan* p) {
  delete p->tm_at; p->tm_at = nullptr;
  delete p->tm_t01; p->tm_t01 = nullptr;
  delete p->tm_b3; p->tm_b3 = nullptr;
  delete p->tm_b3w; p->tm_b3w = nullptr;
  if (p->pot_stats) cudaFree(p->pot_stats);
  p->pot_stats = nullptr;
  void* bufs[] = {p->fft_at, p->fft_b3, p->fft_t01, p->tw_x, p->tw_y, p->tw_zh, p->tw_zfull};
  for (void* b : bufs)
    if (b) cudaFree(b);
  p->fft_at = nullptr; p->fft_b3 = nullptr; p->fft_t01 = nullptr;
  p->fft_on = false;
}

int32_t slab_barrier(jpm_plan* p, cudaStream_t st, bool exchange_ghost_width, int reach_extra) {
  if (p->slab.P == 1) return JPM_OK;
  // a rank may legitimately lag (snapshot I/O in a callback, a slow first launch): the watchdog is configurable, and
  // jpm_slab_check reports a trip instead of letting later kernels consume incomplete peer data silently
  static const long long timeout_cycles =
      (long long)((getenv("JPM_SLAB_TIMEOUT_S") ? std::max(0.1, atof(getenv("JPM_SLAB_TIMEOUT_S"))) : 4.0) * 2.0e9);
  static const bool full_ghost = getenv("JPM_SLAB_FULL_GHOST") && getenv("JPM_SLAB_FULL_GHOST")[0] == '1';
  static const int ghost_override = getenv("JPM_SLAB_GHOST_OVERRIDE") ? atoi(getenv("JPM_SLAB_GHOST_OVERRIDE")) : -1;
  if (ghost_override >= 0) reach_extra = -1000 - ghost_override;
  if (exchange_ghost_width)
    fft::slab_barrier_kernel<true><<<1, 32, 0, st>>>(p->slab, ++p->epoch, full_ghost ? 1000 : reach_extra, timeout_cycles);
  else
    fft::slab_barrier_kernel<false><<<1, 32, 0, st>>>(p->slab, ++p->epoch, 0, timeout_cycles);
  JPM_LAUNCH_CHECK();
  return JPM_OK;
}

// density_p (painted, ghosts NOT folded) -> force3_p (ghosts filled).  P == 1: five kernels on `st`;
// P > 1: the same five kernels, each rank on its own stream, with four flag barriers between them.
int32_t pmfft_forces(jpm_plan* p, cudaStream_t st, float r_split, const float* filter_tab, int n_tab,
                     float filter_kmax, bool skip_first_barrier) {
  using namespace fft;
  JPM_CHECK_ARG(p->fft_on, "pmfft not enabled for this plan");
  const Slab& sl = p->slab;
  const int nzh = sl.nzh;
  const float norm = 1.0f / ((float)sl.nx * (float)sl.ny * (float)sl.nz);
  const float fscale = filter_tab ? (float)(n_tab - 1) / filter_kmax : 0.f;
  const bool ycols1 = sl.P == 1;     // one GPU: 8-column tiles in the forward y pass
  const int nty = ycols1 ? (nzh + kColsC1 - 1) / kColsC1 : (nzh + kColsC - 1) / kColsC, ntx = (nzh + kXC - 1) / kXC;
  static const TmapPack kNoMaps{};
  const TmapPack& tat = p->fft_tma_store ? *p->tm_at : kNoMaps;
  const bool pair = p->fft_pair;
  const TmapPack& tb3 = p->fft_tma_store ? *p->tm_t01 : kNoMaps;   // T01 maps (pair) or planar B3 maps
  // P == 1: the passes that hand x planes to each other (z-fwd -> y-fwd, y-inv -> z-inv) can run as launch pairs
  // over chunks of planes, so that the consumer finds the producer's output in L2 (126 MB) instead of HBM
  const bool chunked = sl.P == 1 && p->fft_chunk > 0 && p->fft_chunk < sl.lx;
  const int cx = chunked ? p->fft_chunk : sl.lx;
  double* sumsq_ptr = (p->want_sumsq && p->pot_stats) ? p->pot_stats : nullptr;
  int32_t rc;
  // every rank has painted: neighbours' ghost planes are final; agree on the ghost width of this step
  if (!skip_first_barrier && (rc = slab_barrier(p, st, true))) return rc;
  for (int x0 = 0; x0 < sl.lx; x0 += cx) {
    const int nxl = std::min(cx, sl.lx - x0);
#define RUN_ZF(N_)                                                                                             \
  if (sl.py > 1)                                                                                               \
    zfwd_kernel<N_, true><<<dim3(sl.ny / kRows, nxl, 1), threads_for<N_ / 2, kRows>(), z_smem<N_>(), st>>>(    \
        sl, p->tw_zh, p->tw_zfull, x0);                                                                        \
  else                                                                                                         \
    zfwd_kernel<N_><<<dim3(sl.ny / kRows, nxl, 1), threads_for<N_ / 2, kRows>(), z_smem<N_>(), st>>>(          \
        sl, p->tw_zh, p->tw_zfull, x0);
    JPM_FFT_SWITCH(sl.nz, RUN_ZF)
#undef RUN_ZF
    JPM_LAUNCH_CHECK();
    if (p->timer && !chunked) p->timer->mark(st, "fft_z_r2c+ghost_fold");
#define RUN_YF(N_)                                                                                             \
  if (ycols1 && p->fft_tma_store)                                                                              \
    yfwd_kernel<N_, kColsC1, true><<<dim3(nty, nxl, 1), threads_for<N_, kColsC1>(), cols_smem<N_, kColsC1>(), st>>>( \
        sl, p->tw_y, tat, x0);                                                                                 \
  else if (ycols1)                                                                                             \
    yfwd_kernel<N_, kColsC1, false><<<dim3(nty, nxl, 1), threads_for<N_, kColsC1>(), cols_smem<N_, kColsC1>(), st>>>( \
        sl, p->tw_y, tat, x0);                                                                                 \
  else if (p->fft_tma_store)                                                                                   \
    yfwd_kernel<N_, kColsC, true><<<dim3(nty, nxl, 1), threads_for<N_, kColsC>(), cols_smem<N_, kColsC>(), st>>>( \
        sl, p->tw_y, tat, x0);                                                                                 \
  else                                                                                                         \
    yfwd_kernel<N_, kColsC, false><<<dim3(nty, nxl, 1), threads_for<N_, kColsC>(), cols_smem<N_, kColsC>(), st>>>( \
        sl, p->tw_y, tat, x0);
    JPM_FFT_SWITCH(sl.ny, RUN_YF)
#undef RUN_YF
    JPM_LAUNCH_CHECK();
  }
  if ((rc = slab_barrier(p, st))) return rc;      // AT complete on every rank
  if (p->timer) p->timer->mark(st, chunked ? "fft_z_r2c+ghost_fold|fft_y_fwd (chunked pairs)" : "fft_y_fwd+transpose");
#define RUN_X(N_)                                                                                              \
  if (p->fft_tma_store && pair && sumsq_ptr)                                                                   \
    xfused_kernel<N_, kXC, true, true, true><<<dim3(ntx, sl.ly, 1), threads_for<N_, kXC, 8>(), xfused_smem_tma<N_, kXC>(), st>>>( \
        sl, p->tw_x, p->wx, p->wy, p->wz, p->ax, norm, r_split * r_split, filter_tab, n_tab, fscale, tb3, sumsq_ptr);     \
  else if (p->fft_tma_store && sumsq_ptr)                                                                      \
    xfused_kernel<N_, kXC, true, false, true><<<dim3(ntx, sl.ly, 1), threads_for<N_, kXC, 8>(), xfused_smem_tma_planar<N_, kXC>(), st>>>( \
        sl, p->tw_x, p->wx, p->wy, p->wz, p->ax, norm, r_split * r_split, filter_tab, n_tab, fscale, tb3, sumsq_ptr);     \
  else if (p->fft_tma_store && pair)                                                                           \
    xfused_kernel<N_, kXC, true, true><<<dim3(ntx, sl.ly, 1), threads_for<N_, kXC, 8>(), xfused_smem_tma<N_, kXC>(), st>>>( \
        sl, p->tw_x, p->wx, p->wy, p->wz, p->ax, norm, r_split * r_split, filter_tab, n_tab, fscale, tb3, sumsq_ptr);     \
  else if (p->fft_tma_store)                                                                                   \
    xfused_kernel<N_, kXC, true, false><<<dim3(ntx, sl.ly, 1), threads_for<N_, kXC, 8>(), xfused_smem_tma_planar<N_, kXC>(), st>>>( \
        sl, p->tw_x, p->wx, p->wy, p->wz, p->ax, norm, r_split * r_split, filter_tab, n_tab, fscale, tb3, sumsq_ptr);     \
  else if (pair)                                                                                               \
    xfused_kernel<N_, kXC, false, true><<<dim3(ntx, sl.ly, 1), threads_for<N_, kXC, 8>(), xfused_smem<N_, kXC>(), st>>>( \
        sl, p->tw_x, p->wx, p->wy, p->wz, p->ax, norm, r_split * r_split, filter_tab, n_tab, fscale, tb3, sumsq_ptr);     \
  else                                                                                                         \
    xfused_kernel<N_, kXC, false, false><<<dim3(ntx, sl.ly, 1), threads_for<N_, kXC, 8>(), xfused_smem<N_, kXC>(), st>>>( \
        sl, p->tw_x, p->wx, p->wy, p->wz, p->ax, norm, r_split * r_split, filter_tab, n_tab, fscale, tb3, sumsq_ptr);
  JPM_FFT_SWITCH(sl.nx, RUN_X)
#undef RUN_X
  JPM_LAUNCH_CHECK();
  if ((rc = slab_barrier(p, st))) return rc;      // B[0], B[1] complete on every rank
  if (p->timer) p->timer->mark(st, "fft_x_fwd+greens_grad+ifft_x_x2+transpose");
  for (int x0 = 0; x0 < sl.lx; x0 += cx) {
    const int nxl = std::min(cx, sl.lx - x0);
#define RUN_YI(N_)                                                                                             \
  if (pair)                                                                                                    \
    yinv_kernel<N_, kXC, true><<<dim3(ntx, nxl, 1), threads_for<N_, kXC, 8>(), xfused_smem<N_, kXC>(), st>>>(  \
        sl, p->tw_y, p->ay, x0);                                                                               \
  else                                                                                                         \
    yinv_kernel<N_, kXC, false><<<dim3(ntx, nxl, 1), threads_for<N_, kXC, 8>(), xfused_smem<N_, kXC>(), st>>>( \
        sl, p->tw_y, p->ay, x0);
    JPM_FFT_SWITCH(sl.ny, RUN_YI)
#undef RUN_YI
    JPM_LAUNCH_CHECK();
    if (p->timer && !chunked) p->timer->mark(st, "ifft_y_x3");
#define RUN_ZI(N_)                                                                                             \
  if (sl.py > 1)                                                                                               \
    zinv_kernel<N_, true><<<dim3(sl.ny / kRows, nxl, 3), threads_for<N_ / 2, kRows>(), z_smem<N_>(), st>>>(    \
        sl, p->tw_zh, p->tw_zfull, p->az, x0, p->fft_zvariant, 0, 0);                                          \
  else                                                                                                         \
    zinv_kernel<N_><<<dim3(sl.ny / kRows, nxl, 3), threads_for<N_ / 2, kRows>(), z_smem<N_>(), st>>>(          \
        sl, p->tw_zh, p->tw_zfull, p->az, x0, p->fft_zvariant, 0, 0);
    JPM_FFT_SWITCH(sl.nz, RUN_ZI)
#undef RUN_ZI
    JPM_LAUNCH_CHECK();
  }
  if ((rc = slab_barrier(p, st))) return rc;      // force ghost planes written by the neighbours are final
  if (p->timer) p->timer->mark(st, chunked ? "ifft_y_x3|ifft_z_c2r_x3 (chunked pairs)" : "ifft_z_c2r_x3+ghost_fill");
  return JPM_OK;
}

int32_t slab_stats_share(jpm_plan* p, cudaStream_t st, int slot) {
  if (p->slab.P == 1 || !p->pot_stats) return JPM_OK;
  fft::slab_stats_kernel<<<1, 32, 0, st>>>(p->slab, p->pot_stats, slot);
  JPM_LAUNCH_CHECK();
  return JPM_OK;
}

// density_p (painted, ghosts NOT folded) -> psi = IFFT(G delta / k^2) in force3_p component 0, ghosts filled
// (+2 planes for the read kernel's difference stencil).  Same passes / barriers as pmfft_forces with ONE
// spectrum through the inverse half: 8 + 8 + 8 + 8 + 8 = 40 B/cell instead of 72.
int32_t pmfft_potential(jpm_plan* p, cudaStream_t st, float r_split, const float* filter_tab, int n_tab,
                        float filter_kmax, bool to_psi, bool skip_first_barrier, const fft::KColour* colour) {
  using namespace fft;
  JPM_CHECK_ARG(p->fft_on, "pmfft not enabled for this plan");
  if (to_psi && !p->slab.psi[p->slab.rank]) {
    JPM_CHECK_ARG(p->slab.P == 1, "slab plan without a psi mesh");
    JPM_CUDA(cudaMalloc(&p->psi_p, p->npad * sizeof(float)));
    JPM_CUDA(cudaMemsetAsync(p->psi_p, 0, p->npad * sizeof(float), st));
    p->slab.psi[0] = p->psi_p;
  }
  const Slab& sl = p->slab;
  const int nzh = sl.nzh;
  const float norm = 1.0f / ((float)sl.nx * (float)sl.ny * (float)sl.nz);
  const float fscale = filter_tab ? (float)(n_tab - 1) / filter_kmax : 0.f;
  const bool ycols1 = sl.P == 1;     // one GPU: 8-column tiles in the forward y pass
  const int nty = ycols1 ? (nzh + kColsC1 - 1) / kColsC1 : (nzh + kColsC - 1) / kColsC, ntx = (nzh + kXC - 1) / kXC;
  static const TmapPack kNoMaps{};
  const TmapPack& tat = p->fft_tma_store ? *p->tm_at : kNoMaps;
  const TmapPack& tb3 = p->fft_tma_store ? *p->tm_b3 : kNoMaps;
  if (!p->pot_stats) JPM_CUDA(cudaMalloc(&p->pot_stats, 4 * sizeof(double)));
  JPM_CUDA(cudaMemsetAsync(p->pot_stats, 0, 4 * sizeof(double), st));   // [0] sum |psi_k|^2, [1] max |F| bits of this step
  int32_t rc;
  if (!skip_first_barrier && (rc = slab_barrier(p, st, true, 2))) return rc;
#define RUN_ZF(N_)                                                                                             \
  if (sl.py > 1)                                                                                               \
    zfwd_kernel<N_, true><<<dim3(sl.ny / kRows, sl.lx, 1), threads_for<N_ / 2, kRows>(), z_smem<N_>(), st>>>(  \
        sl, p->tw_zh, p->tw_zfull, 0);                                                                         \
  else                                                                                                         \
    zfwd_kernel<N_><<<dim3(sl.ny / kRows, sl.lx, 1), threads_for<N_ / 2, kRows>(), z_smem<N_>(), st>>>(        \
        sl, p->tw_zh, p->tw_zfull, 0);
  JPM_FFT_SWITCH(sl.nz, RUN_ZF)
#undef RUN_ZF
  JPM_LAUNCH_CHECK();
  if (p->timer) p->timer->mark(st, "fft_z_r2c+ghost_fold");
#define RUN_YF(N_)                                                                                             \
  if (ycols1 && p->fft_tma_store)                                                                              \
    yfwd_kernel<N_, kColsC1, true><<<dim3(nty, sl.lx, 1), threads_for<N_, kColsC1>(), cols_smem<N_, kColsC1>(), st>>>( \
        sl, p->tw_y, tat, 0);                                                                                  \
  else if (ycols1)                                                                                             \
    yfwd_kernel<N_, kColsC1, false><<<dim3(nty, sl.lx, 1), threads_for<N_, kColsC1>(), cols_smem<N_, kColsC1>(), st>>>( \
        sl, p->tw_y, tat, 0);                                                                                  \
  else if (p->fft_tma_store)                                                                                   \
    yfwd_kernel<N_, kColsC, true><<<dim3(nty, sl.lx, 1), threads_for<N_, kColsC>(), cols_smem<N_, kColsC>(), st>>>( \
        sl, p->tw_y, tat, 0);                                                                                  \
  else                                                                                                         \
    yfwd_kernel<N_, kColsC, false><<<dim3(nty, sl.lx, 1), threads_for<N_, kColsC>(), cols_smem<N_, kColsC>(), st>>>( \
        sl, p->tw_y, tat, 0);
  JPM_FFT_SWITCH(sl.ny, RUN_YF)
#undef RUN_YF
  JPM_LAUNCH_CHECK();
  if ((rc = slab_barrier(p, st))) return rc;
  if (p->timer) p->timer->mark(st, "fft_y_fwd+transpose");
  // across NVLink 64-byte rows run at about half the link rate: 16 columns per tile there (128-byte rows)
  static const bool wide_env = !(getenv("JPM_XPOT_WIDE") && getenv("JPM_XPOT_WIDE")[0] == '0');
  const bool wide = sl.P > 1 && p->fft_tma_store && sl.nx <= 512 && wide_env && !colour;
  const KColour kc = colour ? *colour : KColour{0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const int ntxw = (nzh + kXCW - 1) / kXCW;
#define RUN_XP(N_)                                                                                             \
  if (wide) {                                                                                                  \
    if constexpr (N_ <= 512)                                                                                   \
      xpot_kernel<N_, kXCW, true><<<dim3(ntxw, sl.ly, 1), threads_for<N_, kXCW, 8>(), xpot_smem<N_, kXCW>(), st>>>( \
          sl, p->tw_x, p->wx, p->wy, p->wz, norm, r_split * r_split, filter_tab, n_tab, fscale, *p->tm_b3w, p->pot_stats, kc); \
  } else if (colour)                                                                                           \
    xpot_kernel<N_, kXC, false, 1><<<dim3(ntx, sl.ly, 1), threads_for<N_, kXC, 8>(), xpot_smem<N_, kXC>(), st>>>( \
        sl, p->tw_x, p->wx, p->wy, p->wz, norm, 0.f, filter_tab, n_tab, 0.f, tb3, p->pot_stats, kc);           \
  else if (p->fft_tma_store)                                                                                   \
    xpot_kernel<N_, kXC, true><<<dim3(ntx, sl.ly, 1), threads_for<N_, kXC, 8>(), xpot_smem<N_, kXC>(), st>>>(  \
        sl, p->tw_x, p->wx, p->wy, p->wz, norm, r_split * r_split, filter_tab, n_tab, fscale, tb3, p->pot_stats, kc); \
  else                                                                                                         \
    xpot_kernel<N_, kXC, false><<<dim3(ntx, sl.ly, 1), threads_for<N_, kXC, 8>(), xpot_smem<N_, kXC>(), st>>>( \
        sl, p->tw_x, p->wx, p->wy, p->wz, norm, r_split * r_split, filter_tab, n_tab, fscale, tb3, p->pot_stats, kc);
  JPM_FFT_SWITCH(sl.nx, RUN_XP)
#undef RUN_XP
  JPM_LAUNCH_CHECK();
  if ((rc = slab_barrier(p, st))) return rc;
  if (p->timer) p->timer->mark(st, "fft_x_fwd+greens+ifft_x+transpose");
#define RUN_YP(N_)                                                                                             \
  ypot_kernel<N_, kYPC><<<dim3((nzh + kYPC - 1) / kYPC, sl.lx, 1), threads_for<N_, kYPC>(), cols_smem<N_, kYPC>(), st>>>( \
      sl, p->tw_y, 0);
  JPM_FFT_SWITCH(sl.ny, RUN_YP)
#undef RUN_YP
  JPM_LAUNCH_CHECK();
  if (p->timer) p->timer->mark(st, "ifft_y");
#define RUN_ZP(N_)                                                                                             \
  if (sl.py > 1)                                                                                               \
    zinv_kernel<N_, true><<<dim3(sl.ny / kRows, sl.lx, 1), threads_for<N_ / 2, kRows>(), z_smem<N_>(), st>>>(  \
        sl, p->tw_zh, p->tw_zfull, p->az, 0, p->fft_zvariant, 2, to_psi ? 1 : 0);                              \
  else                                                                                                         \
    zinv_kernel<N_><<<dim3(sl.ny / kRows, sl.lx, 1), threads_for<N_ / 2, kRows>(), z_smem<N_>(), st>>>(        \
        sl, p->tw_zh, p->tw_zfull, p->az, 0, p->fft_zvariant, 2, to_psi ? 1 : 0);
  JPM_FFT_SWITCH(sl.nz, RUN_ZP)
#undef RUN_ZP
  JPM_LAUNCH_CHECK();
  if ((rc = slab_barrier(p, st))) return rc;
  if (p->timer) p->timer->mark(st, "ifft_z_c2r+ghost_fill");
  return JPM_OK;
}

// density_p (holding white noise) -> psi mesh = IFFT(FFT(white) amp(|k_phys|)) / Nc: linear_field on the chain
int32_t pmfft_linear_field(jpm_plan* p, cudaStream_t st, const float* tab, int n_tab, float lkmin, float lkmax,
                           float sx, float sy, float sz, float dc_amp) {
  const fft::KColour kc{lkmin, (float)(n_tab - 1) / (lkmax - lkmin), sx, sy, sz, dc_amp};
  return pmfft_potential(p, st, 0.f, tab, n_tab, 0.f, true, false, &kc);
}

// psi mesh (ghosts filled by pmfft_potential(to_psi)) -> force3_p, ghost cells included.
int32_t pmfft_gradient(jpm_plan* p, cudaStream_t st) {
  const Slab& sl = p->slab;
  JPM_CHECK_ARG(sl.psi[sl.rank], "no psi mesh (run pmfft_potential(to_psi) first)");
  const int nxl = sl.Lx + 2 * sl.gx;
  unsigned* fmax_bits = p->pot_stats ? reinterpret_cast<unsigned*>(p->pot_stats + 1) : nullptr;
  fft::fdgrad_kernel<<<dim3((sl.nzp / 4 + 31) / 32, (sl.nyp + 7) / 8, (nxl + fft::kGradPlanes - 1) / fft::kGradPlanes), 256, 0, st>>>(
      sl, 0, nxl, fmax_bits);
  JPM_LAUNCH_CHECK();
  if (p->timer) p->timer->mark(st, "fd_gradient");
  return JPM_OK;
}

}  // namespace jpm
