// Tile-sorted resident particle state ("sim"): the B200 design of the PM step for scattered
// (late-time) particle distributions.
//
//   reference semantics: jaxpm/painting.py:15-45 / painting_utils.py:28-112 (paint),
//   jaxpm/painting.py:78-106 / painting_utils.py:144-187 (read), jaxpm/pm.py:54-56 (3 reads),
//   jaxpm/ode.py:91-117 (drift, kick).
//
// Why: with 1 Mpc/h cells particles wander ~10 cells from their Lagrangian site; a warp of
// Lagrangian neighbours then touches ~32 different cache lines per gather/atomic and both
// kernels become L1-wavefront / L2-atomic bound (measured: 6.1 ms paint, 13.9 ms read3 at 512^3).
// Shared memory does the same random accesses ~10x faster (measured 2.9 cyc/particle/SM for 8
// CAS float atomics, 0.85 for 8 LDS gathers).  So particles are kept SORTED BY MESH TILE:
//   - state = pos4[np] (x, y, z | displacement, particle id) + vel SoA [3][np], grouped by tile,
//     start[t] = first slot of tile t;
//   - paint:  one CTA per tile, (T+2m+1)^3 box accumulated in shared memory, flushed with REDG;
//             particles outside the box (drifted > m cells) fall back to global atomics;
//             the same pass histograms the tile every particle is in NOW (count[]);
//   - read3 + kick + drift: one CTA per tile, the three force boxes staged in shared memory,
//             8-corner gathers from smem, and the updated particle is written straight into
//             its slot of the NEXT ordering (exclusive scan of count[] -> cursor[]), so the
//             re-sort costs no extra pass over the particles.
// The user-visible order is restored by jpm_sim_store (scatter by id).
//
// Two flavours of the tile kernels (template flag TMA):
//   TMA = false: the mesh is a plain compact [nx][ny][nz] array (public jpm_sim_paint /
//                jpm_sim_read_kick_drift on caller-owned meshes); boxes are staged with cp.async rows and
//                flushed with vector red.global.add, wrapping periodically cell by cell;
//   TMA = true : the mesh is the plan's ghost-zone array (plan_internal.cuh), where no box ever wraps:
//                the read box (3 force meshes) arrives as ONE cp.async.bulk.tensor.4d signalled on an
//                mbarrier, the paint box leaves as ONE cp.reduce.async.bulk.tensor.3d (.add.f32).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "plan_internal.cuh"

struct SimGeom {
  int nx, ny, nz;            // mesh painted into / read from (logical, periodic)
  int pny, pnz, hx, hy;      // particle grid (relative rule) and halo offsets
  int tshift, T, m;          // tile edge = 1 << tshift, margin
  int ntx, nty, ntz, nt;     // tile grid
  int BX, BY, BZ;            // shared-memory box = T + 2m + 1 per axis
  // storage of the mesh the kernels address directly (fallback atomics / gathers, non-TMA rows):
  // element (i, j, k) lives at ((i + mo) * msx + (j + mo) * msy + (k + mo)); batch stride mb
  long long msx, msy, mb;
  int mo, mox;               // storage offset of cell 0 along y/z (mo) and along x (mox)
};

struct jpm_sim {
  jpm_plan* plan = nullptr;
  SimGeom g;                           // compact-mesh geometry
  SimGeom gp;                          // same tiles on the plan's ghost-zone meshes (TMA path)
  bool tma = false;
  CUtensorMap tm_rho, tm_f3;           // box maps of density_p (3-D) and force3_p (4-D)
  int relative = 0;
  long long np = 0;
  float4* pos[2] = {nullptr, nullptr};
  float* vel[2] = {nullptr, nullptr};  // SoA [3][np]
  int* start[2] = {nullptr, nullptr};  // [nt+1]
  int* count = nullptr;                // [nt]  occupancy of the next ordering
  int* cursor = nullptr;               // [nt]  slot cursors while scattering
  unsigned long long* stats = nullptr; // [0]/[1] paint/read global-memory fallbacks, [2]/[3] generic-stencil particles
  int cur = 0;
  bool painted = false, loaded = false;
  bool pos_only = false;               // JPM_SIM_POSITIONS_ONLY: no velocities, no second ordering (paint / forces only)
  // ---- potential force path: ONE inverse transform (psi mesh) + the gradient pass of csrc/pmfft.cu ----
  bool pot_ok = false;                 // potential chain available: TMA tile path + fused FFT chain (power-of-two mesh)
  int force_mode = 0;                  // JPM_FORCE_SPECTRAL / JPM_FORCE_POTENTIAL / JPM_FORCE_AUTO
  int cur_mode = 0;                    // what the next step runs (auto switches it from the measured error bound)
  // AUTO: the statistics of step n (device doubles, plan->pot_stats) are copied to pinned slot n % 3 behind the step;
  // the mode of step n is decided from the slot of step n - 2 after waiting for ITS event - a fixed lag, so the
  // sequence of modes does not depend on how far the host runs ahead of the device (reproducible runs, and the
  // same decision on every rank of a multi-GPU run)
  double* stats_host = nullptr;        // pinned [3][4]
  cudaEvent_t stats_ev[3] = {nullptr, nullptr, nullptr};
  bool stats_pending[3] = {false, false, false};
  bool stats_fixed[3] = {false, false, false};   // slot holds the slab ranks' global statistics (u64 fixed point)
  long long nstep = 0;
  double last_bound = -1.0;            // last evaluated error bound (auto), < 0 = none yet
  long long mode_steps[2] = {0, 0};    // steps run in spectral / potential mode
  // ---- jpm_sim_steps_host_f32: copy streams, double-buffered device staging (allocated at the first call) ----
  cudaStream_t st_h2d = nullptr, st_d2h = nullptr;
  float* stage_in[2] = {nullptr, nullptr};    // [2][np][3]: pos | vel
  float* stage_out[2] = {nullptr, nullptr};
  cudaEvent_t ev_in_ready[2] = {nullptr, nullptr}, ev_in_free[2] = {nullptr, nullptr};
  cudaEvent_t ev_out_ready[2] = {nullptr, nullptr}, ev_out_free[2] = {nullptr, nullptr};
};

namespace jpm {

#ifndef JPM_PAINT_CTAS
#define JPM_PAINT_CTAS 4     // resident paint CTAs per SM (64 registers, 4 x 52 KB of shared memory): 1.36 -> 1.21 ms vs 3
#endif
#ifndef JPM_L2_AHEAD
#define JPM_L2_AHEAD 1
#endif
constexpr int kL2Ahead = JPM_L2_AHEAD;   // iterations of particle stream requested from L2 ahead of the register prefetch
constexpr int kTmaMz = 4;   // z margin below a tile in the TMA flavour (== kGhost: box z origin = tile origin in padded coordinates)
static_assert(kTmaMz == kGhost, "TMA box z origin must coincide with a 16-byte aligned padded coordinate");

__device__ __forceinline__ int wrap_local(int i, int o, int n) {
  int a = i - o;
  if (a < 0) a += n;
  else if (a >= n) a -= n;
  return a;
}

__device__ __forceinline__ int wrap_global(int a, int n) {  // a in [-n, 2n)
  if (a < 0) a += n;
  else if (a >= n) a -= n;
  return (a >= 0 && a < n) ? a : pymod(a, n);
}

// relative mode keeps the Lagrangian site packed as (i << 20 | j << 10 | k) in pos.w, so the hot
// kernels decode it with shifts; absolute mode keeps the plain particle id there.
__device__ __forceinline__ int pack_ijk(int i, int j, int k) { return (i << 20) | (j << 10) | k; }

template <bool REL>
__device__ __forceinline__ void sim_stencil(const SimGeom& g, float x, float y, float z, int w,
                                            Cic1& cx, Cic1& cy, Cic1& cz) {
  int bi = 0, bj = 0, bk = 0;
  if (REL) {
    bk = w & 1023;
    bj = ((w >> 10) & 1023) + g.hy;
    bi = (w >> 20) + g.hx;
  }
  cx = cic_1d<REL, false>(bi, x, g.nx);
  cy = cic_1d<REL, false>(bj, y, g.ny);
  cz = cic_1d<REL, false>(bk, z, g.nz);
}

template <int TS>
__device__ __forceinline__ int tile_of(const SimGeom& g, int i0, int j0, int k0) {
  i0 = max(i0, 0); j0 = max(j0, 0); k0 = max(k0, 0);  // dropped corner (-1) -> tile of cell 0
  return ((i0 >> TS) * g.nty + (j0 >> TS)) * g.ntz + (k0 >> TS);
}

// Warp-aggregated "add n to counter[key]" (RET=false) or slot claim (RET=true: returns this lane's
// slot).  All 32 lanes must call; invalid lanes pass valid=false.
template <bool RET>
__device__ __forceinline__ int warp_claim(int* counter, int key, bool valid) {
  const int lane = threadIdx.x & 31;
  const unsigned peers = __match_any_sync(0xffffffffu, valid ? key : (0x40000000 | lane));
  const int leader = __ffs(peers) - 1;
  int base = 0;
  if (valid && lane == leader) {
    if (RET) base = atomicAdd(counter + key, __popc(peers));
    else atomicAdd(counter + key, __popc(peers));
  }
  if (!RET) return 0;
  base = __shfl_sync(0xffffffffu, base, leader);
  return base + __popc(peers & ((1u << lane) - 1u));
}

// ---- build / re-sort from user arrays -------------------------------------------------------
template <bool REL, int TS>
__global__ void __launch_bounds__(256)
sim_count_kernel(SimGeom g, const float* __restrict__ pos, long long np, int* __restrict__ count) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = p < np;
  int tt = 0;
  if (valid) {
    int w = (int)p;
    if (REL) {
      const int k = (int)(p % g.pnz);
      const long long t = p / g.pnz;
      w = pack_ijk((int)(t / g.pny), (int)(t % g.pny), k);
    }
    Cic1 cx, cy, cz;
    sim_stencil<REL>(g, pos[3 * p], pos[3 * p + 1], pos[3 * p + 2], w, cx, cy, cz);
    tt = tile_of<TS>(g, cx.i0, cy.i0, cz.i0);
  }
  warp_claim<false>(count, tt, valid);
}

template <bool REL, int TS>
__global__ void __launch_bounds__(256)
sim_fill_kernel(SimGeom g, const float* __restrict__ pos, const float* __restrict__ vel, long long np,
                int* __restrict__ cursor, float4* __restrict__ spos, float* __restrict__ svel) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = p < np;
  int tt = 0, w = (int)p;
  float x = 0, y = 0, z = 0;
  if (valid) {
    x = pos[3 * p]; y = pos[3 * p + 1]; z = pos[3 * p + 2];
    if (REL) {
      const int k = (int)(p % g.pnz);
      const long long t = p / g.pnz;
      w = pack_ijk((int)(t / g.pny), (int)(t % g.pny), k);
    }
    Cic1 cx, cy, cz;
    sim_stencil<REL>(g, x, y, z, w, cx, cy, cz);
    tt = tile_of<TS>(g, cx.i0, cy.i0, cz.i0);
  }
  const int slot = warp_claim<true>(cursor, tt, valid);
  if (valid) {
    spos[slot] = make_float4(x, y, z, __int_as_float(w));
    if (vel) {
      svel[slot] = vel[3 * p];
      svel[np + slot] = vel[3 * p + 1];
      svel[2 * np + slot] = vel[3 * p + 2];
    }
  }
}

template <bool REL>
__global__ void __launch_bounds__(256)
sim_store_kernel(SimGeom g, const float4* __restrict__ spos, const float* __restrict__ svel, long long np,
                 float* __restrict__ pos, float* __restrict__ vel) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= np) return;
  const float4 p = spos[q];
  const int w = __float_as_int(p.w);
  const long long id = REL ? ((long long)(w >> 20) * g.pny + ((w >> 10) & 1023)) * g.pnz + (w & 1023) : w;
  if (pos) { pos[3 * id] = p.x; pos[3 * id + 1] = p.y; pos[3 * id + 2] = p.z; }
  if (vel) { vel[3 * id] = svel[q]; vel[3 * id + 1] = svel[np + q]; vel[3 * id + 2] = svel[2 * np + q]; }
}

// exclusive scan of count[nt] -> start[nt+1]; cursor = start; count = 0.  One CTA of 32 warps; warp w
// owns the contiguous segment [w*seg, (w+1)*seg) and walks it 32 entries at a time (coalesced).
__global__ void __launch_bounds__(1024)
sim_scan_kernel(int* __restrict__ count, int* __restrict__ start, int* __restrict__ cursor, int nt) {
  __shared__ int wsum[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int seg = ((nt + 31) / 32 + 31) & ~31;
  const int b = warp * seg, e = min(b + seg, nt);
  int s = 0;
  for (int i = b + lane; i < e; i += 32) s += count[i];
#pragma unroll
  for (int off = 16; off; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (lane == 0) wsum[warp] = s;
  __syncthreads();
  int run = 0;
  for (int w = 0; w < warp; ++w) run += wsum[w];
  for (int i0 = b; i0 < e; i0 += 32) {
    const int i = i0 + lane;
    const int c = (i < e) ? count[i] : 0;
    int incl = c;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += v;
    }
    if (i < e) {
      start[i] = run + incl - c;
      cursor[i] = run + incl - c;
      count[i] = 0;
    }
    run += __shfl_sync(0xffffffffu, incl, 31);
  }
  if (threadIdx.x == 1023) start[nt] = run;   // the last warp's running total is the grand total
}

// pull a line of the particle stream into L2 ahead of its use (no register held, unlike a real load)
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ void red_add_v2(float* addr, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}

// Generic (slow-path) view of the 8 corners: wrapped global indices, box-local coordinates, weights.
struct Corners {
  int ix[2], iy[2], iz[2];     // wrapped global cell indices (-1 = dropped)
  int lx[2], ly[2], lz[2];     // box-local coordinates
  float wx[2], wy[2], wz[2];
  bool inside;
};

template <int B, int BZV>
__device__ __forceinline__ void make_corners(const SimGeom& g, const Cic1& cx, const Cic1& cy,
                                             const Cic1& cz, int ox, int oy, int oz, Corners& c) {
  c.ix[0] = cx.i0; c.ix[1] = cx.i1; c.iy[0] = cy.i0; c.iy[1] = cy.i1; c.iz[0] = cz.i0; c.iz[1] = cz.i1;
  c.wx[0] = cx.w0; c.wx[1] = cx.w1; c.wy[0] = cy.w0; c.wy[1] = cy.w1; c.wz[0] = cz.w0; c.wz[1] = cz.w1;
  c.inside = true;
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    c.lx[a] = wrap_local(max(c.ix[a], 0), ox, g.nx);
    c.ly[a] = wrap_local(max(c.iy[a], 0), oy, g.ny);
    c.lz[a] = wrap_local(max(c.iz[a], 0), oz, g.nz);
    c.inside = c.inside && c.lx[a] < B && c.ly[a] < B && c.lz[a] < BZV;
  }
}

// ---- fast stencil ---------------------------------------------------------------------------------
// Per axis: cell of corner 0 and the two weights, plus "the two corners are periodic neighbours
// (i1 == i0 + 1 mod n) and neither is dropped".  The common case (no wrap on this axis) reduces the
// generic rules of common.cuh to a handful of operations with bit-identical results:
//   relative (painting_utils.py:48-65): pp = base + d; corner c: r = pp + c (in [0, L), no mod),
//       idx = floor(r), nd = pp - idx (|nd| <= 1 < L/4, no rint correction), w = 1 - |nd|;
//   absolute (painting.py:22-37): idx = floor(p) + c, w = 1 - |p - idx|, 0 <= idx < n (no mod).
// Lanes at the periodic edge of an axis take the generic path for the whole particle.
template <bool REL>
__device__ __forceinline__ bool axis_fast(int base, float v, int n, int& i0, float& w0, float& w1) {
  // straight-line code, one predicate out: lanes that fail (periodic edge, dropped corner, huge / NaN
  // coordinate) are re-done by the caller with the generic rules of common.cuh
  if (REL) {
    const float pp = (float)base + v;
    const float x1 = pp + 1.0f;
    const float f0 = floorf(pp), f1 = floorf(x1);
    w0 = 1.0f - fabsf(pp - f0);
    w1 = 1.0f - fabsf(pp - f1);
    i0 = (int)f0;
    return pp >= 0.0f && x1 < (float)n && f1 == f0 + 1.0f;
  } else {
    const float f = floorf(v);
    w0 = 1.0f - fabsf(v - f);
    w1 = 1.0f - fabsf(v - (f + 1.0f));
    i0 = (int)f;
    return (unsigned)i0 < (unsigned)(n - 1) && fabsf(f) < 1.0e9f;
  }
}

// box-local coordinate of global cell i for a box starting at o (o in [-M, n)), periodic
__device__ __forceinline__ int local_wrap(int i, int o, int n) {
  int l = i - o;
  if (l < 0) l += n;
  else if (l >= n) l -= n;
  return l;
}

struct FastStencil {
  int i0, j0, k0;               // global cell of corner (0,0,0)
  float wx0, wx1, wy0, wy1, wz0, wz1;
};

// One axis by the generic rules of common.cuh (periodic edge).  Out of line and returning BY VALUE: the fast
// path's stencil must stay in registers (an address-taken struct would live in local memory).
struct AxisG { int i0; float w0, w1; int ok; };
template <bool REL>
__device__ __noinline__ AxisG axis_generic(int base, float v, int n) {
  const Cic1 c = cic_1d<REL, false>(base, v, n);
  AxisG r;
  r.i0 = c.i0; r.w0 = c.w0; r.w1 = c.w1;
  // "the two corners are periodic neighbours and neither is dropped"
  r.ok = c.i0 >= 0 && c.i1 >= 0 && (c.i1 == c.i0 + 1 || (c.i0 == n - 1 && c.i1 == 0));
  return r;
}

// Stencil of one particle relative to the box starting at (ox, oy, oz): straight-line code for all three axes,
// then - only in warps that hold a lane at a periodic edge (~2 % of the particles late in a run: relative
// coordinates base + disp are not wrapped) - the generic rule for the axes that need it.  Returns whether the
// 8 corners are (lx..lx+1, ly..ly+1, lz..lz+1) inside the box; s.i0/j0/k0 = wrapped cell of corner 0.
template <bool REL, int BX, int BY, int BZ>
__device__ __forceinline__ bool stencil_box(const SimGeom& g, const float4& p, int ox, int oy, int oz,
                                            FastStencil& s, int& lx, int& ly, int& lz) {
  int bi = 0, bj = 0, bk = 0;
  if (REL) {
    const int w = __float_as_int(p.w);
    bk = w & 1023;
    bj = ((w >> 10) & 1023) + g.hy;
    bi = (w >> 20) + g.hx;
  }
  bool fx = axis_fast<REL>(bi, p.x, g.nx, s.i0, s.wx0, s.wx1);
  bool fy = axis_fast<REL>(bj, p.y, g.ny, s.j0, s.wy0, s.wy1);
  bool fz = axis_fast<REL>(bk, p.z, g.nz, s.k0, s.wz0, s.wz1);
  lx = s.i0 - ox; ly = s.j0 - oy; lz = s.k0 - oz;
  if (!(fx & fy & fz)) {
    if (!fx) {
      const AxisG a = axis_generic<REL>(bi, p.x, g.nx);
      s.i0 = a.i0; s.wx0 = a.w0; s.wx1 = a.w1; fx = a.ok; lx = local_wrap(max(a.i0, 0), ox, g.nx);
    }
    if (!fy) {
      const AxisG a = axis_generic<REL>(bj, p.y, g.ny);
      s.j0 = a.i0; s.wy0 = a.w0; s.wy1 = a.w1; fy = a.ok; ly = local_wrap(max(a.i0, 0), oy, g.ny);
    }
    if (!fz) {
      const AxisG a = axis_generic<REL>(bk, p.z, g.nz);
      s.k0 = a.i0; s.wz0 = a.w0; s.wz1 = a.w1; fz = a.ok; lz = local_wrap(max(a.i0, 0), oz, g.nz);
    }
  }
  return fx & fy & fz & ((unsigned)lx < (unsigned)(BX - 1)) & ((unsigned)ly < (unsigned)(BY - 1)) &
         ((unsigned)lz < (unsigned)(BZ - 1));
}

// ---- TMA / mbarrier helpers (sm_90+ PTX; SASS: UTMALDG / UTMAREDG, SYNCS) ------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
// global (4-D tensor map: z, y, x, component) -> shared box, completion counted on `bar`
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3,
                                            unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(smem_u32(dst)), "l"((unsigned long long)tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
      : "memory");
}
// shared box -> global += (3-D tensor map), f32 add performed by the TMA unit at L2
__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap* tm, int c0, int c1, int c2, const void* src) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2, %3}], [%4];"
               ::"l"((unsigned long long)tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(src))
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ long long mesh_index(const SimGeom& g, int i, int j, int k) {
  return (long long)(i + g.mox) * g.msx + (long long)(j + g.mo) * g.msy + (k + g.mo);
}

// ---- paint -------------------------------------------------------------------------------------
// One CTA per tile.  Box = (T+2M+1)^3 cells, rows padded to an even length BZ so that the flush
// can use 8-byte vector reductions (REDG.E.ADD.F32x2).
//
// Accumulation is 40.24 fixed point on NATIVE 32-bit shared-memory integer atomics: fp32
// atomicAdd on shared memory is a compare-and-swap loop that retries whenever two lanes of a warp
// hit the same cell (3.4 passes on average on a clustered 512^3 set, profiles/r01a), integer ATOMS.ADD
// resolves the conflict in hardware.  lo[] holds the low 32 bits (weights scaled by 2^24, rounded to
// nearest: |error| <= 3e-8 per corner, below fp32 resolution of the weights themselves), a carry
// out of lo bumps a 16-bit field of hi[] (two cells per word), so a cell can take 2^24 particles.
// The sum inside a tile is therefore order-independent (bit-reproducible); only the fp32 merge of
// overlapping tile margins into the global mesh is not.
constexpr float kFixScale = 16777216.0f;           // 2^24
constexpr float kFixInv = 1.0f / 16777216.0f;

__device__ __forceinline__ void fixed_carry(unsigned* hi, int idx) {
  atomicAdd(hi + (idx >> 1), 1u << ((idx & 1) * 16));
}

__device__ __forceinline__ float fixed_to_float(unsigned lo, unsigned hi16) {
  return (hi16 ? __ull2float_rn(((unsigned long long)hi16 << 32) | lo) : __uint2float_rn(lo)) * kFixInv;
}

template <bool REL, int TS, int M, bool TMA>
__global__ void __launch_bounds__(256, JPM_PAINT_CTAS)
sim_paint_kernel(const __grid_constant__ CUtensorMap tm, SimGeom g, const float4* __restrict__ spos,
                 const int* __restrict__ start, float* __restrict__ mesh, int* __restrict__ count,
                 unsigned long long* __restrict__ stats, int* __restrict__ xrange, int track_y) {
  // TMA boxes must start on a 16-byte boundary of the innermost (z) axis (misaligned coordinates raise
  // "illegal instruction", tools/tma_probe.cu): the z margin below the tile is kTmaMz = 4 cells there.
  constexpr int T = 1 << TS, B = T + 2 * M + 1, MZ = TMA ? kTmaMz : M;
  constexpr int BZ = TMA ? ((T + MZ + M + 1 + 3) & ~3) : ((B + 1) & ~1), NBOX = B * B * BZ;
  constexpr int BZV = TMA ? BZ : B;          // z cells of the box that may be touched
  constexpr int SX = B * BZ, SY = BZ;
  extern __shared__ __align__(128) unsigned sbox[];    // lo[NBOX] | hi[NBOX / 2]
  unsigned* const lo = sbox;
  unsigned* const hi = sbox + NBOX;
  __shared__ int scnt[28];                             // 27 neighbour tiles + generic-stencil count
  __shared__ int sxr[4];                               // lowest / highest x plane (and y row) touched (slab plans: ghost width)
  int xlo = 0x7fffffff, xhi = (int)0x80000000;
  int ylo = 0x7fffffff, yhi = (int)0x80000000;        // pencil grids: the rows too (xrange[3], xrange[4])
  const int t = blockIdx.x;
  const int beg = start[t], end = start[t + 1];
  if (beg == end) return;
  const int tz = t % g.ntz, ty = (t / g.ntz) % g.nty, tx = t / (g.ntz * g.nty);
  const int ox = (tx << TS) - M, oy = (ty << TS) - M, oz = (tz << TS) - MZ;
  const int lane = threadIdx.x & 31;
  int nslow = 0;
  // first particle of this thread is in flight while the box is being zeroed
  // the first two particles of this thread are in flight while the box is being zeroed; the loop
  // keeps two loads per thread outstanding (the kernel is otherwise latency-bound on this stream)
  int q = beg + threadIdx.x;
  float4 p = make_float4(0.f, 0.f, 0.f, 0.f), pn = p;
  if (q < end) p = __ldcs(spos + q);
  if (q + (int)blockDim.x < end) pn = __ldcs(spos + q + blockDim.x);
  constexpr int NW = NBOX + NBOX / 2, NW4 = NW / 4;
  for (int i = threadIdx.x; i < NW4; i += blockDim.x) reinterpret_cast<uint4*>(sbox)[i] = make_uint4(0u, 0u, 0u, 0u);
  if constexpr (NW > 4 * NW4) {
    if (threadIdx.x < NW - 4 * NW4) sbox[4 * NW4 + threadIdx.x] = 0u;
  }
  if (threadIdx.x < 28) scnt[threadIdx.x] = 0;
  if (threadIdx.x < 4) sxr[threadIdx.x] = (threadIdx.x & 1) ? (int)0x80000000 : 0x7fffffff;
  __syncthreads();
  for (int qb = beg + (threadIdx.x & ~31); qb < end; qb += blockDim.x) {   // warp-uniform trip count
    const bool valid = q < end;
    const int qn = q + blockDim.x;
    float4 pnn = make_float4(0.f, 0.f, 0.f, 0.f);
    if (qn + (int)blockDim.x < end) pnn = __ldcs(spos + qn + blockDim.x);  // prefetch, two ahead
    int ti = 0, tj = 0, tk = 0;            // tile the particle is in now (next ordering)
    if (valid) {
      FastStencil s;
      int lx, ly, lz;
      const bool fast = stencil_box<REL, B, B, BZV>(g, p, ox, oy, oz, s, lx, ly, lz);
      if (fast) {
        const int o = lx * SX + ly * SY + lz;
        const float w00 = s.wx0 * s.wy0, w10 = s.wx1 * s.wy0, w01 = s.wx0 * s.wy1, w11 = s.wx1 * s.wy1;
        // reference order (kx*ky)*kz, weight 1; corner offsets are compile-time constants
        // (kx ky) kz 2^24: scaling kz by a power of two first gives the same bits as scaling the product
        const float z0 = s.wz0 * kFixScale, z1 = s.wz1 * kFixScale;
        const unsigned v[8] = {__float2uint_rn(w00 * z0), __float2uint_rn(w00 * z1), __float2uint_rn(w01 * z0),
                               __float2uint_rn(w01 * z1), __float2uint_rn(w10 * z0), __float2uint_rn(w10 * z1),
                               __float2uint_rn(w11 * z0), __float2uint_rn(w11 * z1)};
        constexpr int off[8] = {0, 1, SY, SY + 1, SX, SX + 1, SX + SY, SX + SY + 1};
        unsigned old[8];
#pragma unroll
#pragma unroll
        for (int c = 0; c < 8; ++c) old[c] = atomicAdd(lo + o + off[c], v[c]);
        // carries out of the low words are rare: count them with the carry flag, branch once
        unsigned ncarry = 0;
#pragma unroll
        for (int c = 0; c < 8; ++c)
          asm("{\n\t.reg .u32 t;\n\tadd.cc.u32 t, %1, %2;\n\taddc.u32 %0, %0, 0;\n\t}" : "+r"(ncarry) : "r"(old[c]), "r"(v[c]));
        if (ncarry) {
#pragma unroll
          for (int c = 0; c < 8; ++c)
            if (old[c] + v[c] < old[c]) fixed_carry(hi, o + off[c]);
        }
        ti = s.i0 >> TS; tj = s.j0 >> TS; tk = s.k0 >> TS;
        xlo = min(xlo, s.i0); xhi = max(xhi, s.i0 + 1);
        if (track_y) { ylo = min(ylo, s.j0); yhi = max(yhi, s.j0 + 1); }
      } else {
        Cic1 cx, cy, cz;
        sim_stencil<REL>(g, p.x, p.y, p.z, __float_as_int(p.w), cx, cy, cz);
        Corners c;
        make_corners<B, BZV>(g, cx, cy, cz, ox, oy, oz, c);
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) {
          const int a = cc & 1, b = (cc >> 1) & 1, d = cc >> 2;
          if (REL && (c.ix[a] < 0 || c.iy[b] < 0 || c.iz[d] < 0)) continue;   // dropped by the reference
          const float w = (c.wx[a] * c.wy[b]) * c.wz[d];
          if (c.inside) {
            const int idx = c.lx[a] * SX + c.ly[b] * SY + c.lz[d];
            const unsigned v = __float2uint_rn(w * kFixScale);
            const unsigned old = atomicAdd(lo + idx, v);
            if (old + v < old) fixed_carry(hi, idx);
          } else {
            atomicAdd(mesh + mesh_index(g, c.ix[a], c.iy[b], c.iz[d]), w);
          }
        }
        if (c.inside) ++nslow; else atomicAdd(stats, 1ull);
        if (cx.i0 >= 0) { xlo = min(xlo, cx.i0); xhi = max(xhi, cx.i0); }
        if (cx.i1 >= 0) { xlo = min(xlo, cx.i1); xhi = max(xhi, cx.i1); }
        if (track_y) {
          if (cy.i0 >= 0) { ylo = min(ylo, cy.i0); yhi = max(yhi, cy.i0); }
          if (cy.i1 >= 0) { ylo = min(ylo, cy.i1); yhi = max(yhi, cy.i1); }
        }
        ti = max(cx.i0, 0) >> TS; tj = max(cy.i0, 0) >> TS; tk = max(cz.i0, 0) >> TS;
      }
    }
    // occupancy of the next ordering: the home tile is counted once per warp, the 26 neighbours in
    // shared counters, anything further away straight in global memory
    const bool home = valid && ti == tx && tj == ty && tk == tz;
    const unsigned mh = __ballot_sync(0xffffffffu, home);
    if (lane == 0 && mh) atomicAdd(scnt + 13, __popc(mh));
    if (valid && !home) {
      int dx = ti - tx, dy = tj - ty, dz = tk - tz;
      if (dx > 1) dx -= g.ntx; else if (dx < -1) dx += g.ntx;
      if (dy > 1) dy -= g.nty; else if (dy < -1) dy += g.nty;
      if (dz > 1) dz -= g.ntz; else if (dz < -1) dz += g.ntz;
      if (dx >= -1 && dx <= 1 && dy >= -1 && dy <= 1 && dz >= -1 && dz <= 1)
        atomicAdd(scnt + (dx + 1) * 9 + (dy + 1) * 3 + (dz + 1), 1);
      else
        atomicAdd(count + (ti * g.nty + tj) * g.ntz + tk, 1);
    }
    p = pn;
    pn = pnn;
    q = qn;
  }
  if (nslow) atomicAdd(scnt + 27, nslow);
  if (xrange) {
    xlo = __reduce_min_sync(0xffffffffu, xlo);
    xhi = __reduce_max_sync(0xffffffffu, xhi);
    if (lane == 0) { atomicMin(sxr, xlo); atomicMax(sxr + 1, xhi); }
    if (track_y) {
      ylo = __reduce_min_sync(0xffffffffu, ylo);
      yhi = __reduce_max_sync(0xffffffffu, yhi);
      if (lane == 0) { atomicMin(sxr + 2, ylo); atomicMax(sxr + 3, yhi); }
    }
  }
  __syncthreads();
  if (xrange && threadIdx.x < 4) {
    if (threadIdx.x == 0) atomicMin(xrange, sxr[0]);
    else if (threadIdx.x == 1) atomicMax(xrange + 1, sxr[1]);
    else if (track_y && threadIdx.x == 2) atomicMin(xrange + (kFlagYmin - kFlagXmin), sxr[2]);
    else if (track_y) atomicMax(xrange + (kFlagYmax - kFlagXmin), sxr[3]);
  }
  if (threadIdx.x == 27 && scnt[27]) atomicAdd(stats + 2, (unsigned long long)scnt[27]);
  if (threadIdx.x < 27 && scnt[threadIdx.x]) {
    const int dx = threadIdx.x / 9 - 1, dy = (threadIdx.x / 3) % 3 - 1, dz = threadIdx.x % 3 - 1;
    const int i0 = pymod(tx + dx, g.ntx), j0 = pymod(ty + dy, g.nty), k0 = pymod(tz + dz, g.ntz);
    atomicAdd(count + (i0 * g.nty + j0) * g.ntz + k0, scnt[threadIdx.x]);
  }
  if constexpr (TMA) {
    // convert the fixed-point box to fp32 in place, then ONE tensor reduce-add moves it into the
    // ghost-zone mesh (the box never wraps there; cells past the array edge are clipped by the TMA unit)
    static_assert(NBOX % 4 == 0, "box rows are multiples of four cells in the TMA flavour");
    for (int i = threadIdx.x; i < NBOX / 4; i += blockDim.x) {
      const uint4 l = reinterpret_cast<const uint4*>(lo)[i];
      const uint2 h = reinterpret_cast<const uint2*>(hi)[i];
      float4 f;
      if ((h.x | h.y) == 0u) {   // no cell of the four holds more than 256 particles (the usual case)
        f = make_float4(__uint2float_rn(l.x) * kFixInv, __uint2float_rn(l.y) * kFixInv,
                        __uint2float_rn(l.z) * kFixInv, __uint2float_rn(l.w) * kFixInv);
      } else {
        f.x = fixed_to_float(l.x, h.x & 0xffffu);
        f.y = fixed_to_float(l.y, h.x >> 16);
        f.z = fixed_to_float(l.z, h.y & 0xffffu);
        f.w = fixed_to_float(l.w, h.y >> 16);
      }
      reinterpret_cast<float4*>(lo)[i] = f;
    }
    fence_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) tma_reduce_add_3d(&tm, oz + g.mo, oy + g.mo, ox + g.mox, lo);
    return;
  }
  // flush: two box rows per warp pass, one z-pair per lane (8-byte vector reductions when the pair
  // cannot straddle the periodic wrap); the z index of a lane is loop invariant.
  constexpr int HP = BZ / 2;                 // pairs per row (<= 16)
  const int warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int sub = lane / HP, zp = lane - sub * HP;     // sub-row 0/1 (lanes >= 2*HP idle)
  const bool vec = ((oz & 1) == 0) && ((g.nz & 1) == 0) && g.nz >= BZ;
  const int gz0 = wrap_global(oz + 2 * zp, g.nz), gz1 = wrap_global(oz + 2 * zp + 1, g.nz);
  if (sub < 2) {
    for (int r = 2 * warp + sub; r < B * B; r += 2 * nwarp) {
      const uint2 l = reinterpret_cast<const uint2*>(lo)[r * HP + zp];
      const unsigned h = hi[r * HP + zp];
      if ((l.x | l.y | h) == 0u) continue;
      const float vx = fixed_to_float(l.x, h & 0xffffu), vy = fixed_to_float(l.y, h >> 16);
      const int lx = r / B, ly = r - lx * B;
      const int gx = wrap_global(ox + lx, g.nx), gy = wrap_global(oy + ly, g.ny);
      float* row = mesh + mesh_index(g, gx, gy, 0);
      if (vec) {
        red_add_v2(row + gz0, vx, vy);
      } else {
        if (vx != 0.f) atomicAdd(row + gz0, vx);
        if (vy != 0.f && 2 * zp + 1 < B) atomicAdd(row + gz1, vy);
      }
    }
  }
}

__device__ __forceinline__ void cp_async4(float* smem, const float* gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gmem) : "memory");
}

// ---- read3 + kick + drift + scatter into the next ordering --------------------------------------
template <bool REL, int TS, int M, bool TMA, bool FMAX = false>
__global__ void __launch_bounds__(512, 2)
sim_read_kernel(const __grid_constant__ CUtensorMap tm, SimGeom g, const float4* __restrict__ spos,
                const float* __restrict__ svel, const int* __restrict__ start, const float* __restrict__ f0,
                const float* __restrict__ f1, const float* __restrict__ f2, float kick, float drift,
                long long np, int* __restrict__ cursor, float4* __restrict__ npos,
                float* __restrict__ nvel, unsigned long long* __restrict__ stats, int l2ahead,
                unsigned* __restrict__ fmax_bits) {
  constexpr int T = 1 << TS, B = T + 2 * M + 1, MZ = TMA ? kTmaMz : M;
  constexpr int BZ = TMA ? ((T + MZ + M + 1 + 3) & ~3) : B, NBOX = B * B * BZ;
  extern __shared__ __align__(128) float box[];        // [3][B][B][BZ]
  float fmax = 0.f;
  __shared__ __align__(8) unsigned long long mbar;
  const int t = blockIdx.x;
  const int beg = start[t], end = start[t + 1];
  if (beg == end) return;
  const int tz = t % g.ntz, ty = (t / g.ntz) % g.nty, tx = t / (g.ntz * g.nty);
  const int ox = (tx << TS) - M, oy = (ty << TS) - M, oz = (tz << TS) - MZ;
  const float* fm[3] = {f0, f1, f2};
  int nslow = 0;
  if (TMA) {
    // one 4-D tensor load brings the (z, y, x, component) box of all three force meshes
    if (threadIdx.x == 0) {
      mbar_init(&mbar, 1);
      fence_async_smem();
      mbar_expect_tx(&mbar, 3u * NBOX * (unsigned)sizeof(float));
      tma_load_4d(box, &tm, oz + g.mo, oy + g.mo, ox + g.mox, 0, &mbar);
    }
  } else {
    // stage the three force boxes with cp.async (LDGSTS): one box row (B contiguous floats) per
    // warp pass, no registers held, all rows of a warp in flight at once
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    if (lane < B) {
      const int gz = wrap_global(oz + lane, g.nz);
      for (int r = warp; r < B * B; r += nwarp) {
        const int lx = r / B, ly = r - lx * B;
        const int gx = wrap_global(ox + lx, g.nx), gy = wrap_global(oy + ly, g.ny);
        const long long o = mesh_index(g, gx, gy, gz);
#pragma unroll
        for (int f = 0; f < 3; ++f) cp_async4(box + f * NBOX + r * B + lane, fm[f] + o);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  // first particle of this thread streams in while the boxes land
  int q = beg + threadIdx.x;
  float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
  float vin[3] = {0.f, 0.f, 0.f};
  if (q < end) {
    p = __ldcs(spos + q);
#pragma unroll
    for (int f = 0; f < 3; ++f) vin[f] = __ldcs(svel + f * np + q);
  }
  if (TMA) {
    __syncthreads();           // the barrier initialisation by thread 0 is visible
    mbar_wait(&mbar, 0);
  } else {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
  }
  for (int qb = beg; qb < end; qb += blockDim.x) {
    const bool valid = q < end;
    const int qn = q + blockDim.x;
    float4 pn = make_float4(0.f, 0.f, 0.f, 0.f);
    float vn[3] = {0.f, 0.f, 0.f};
    if (qn < end) {  // prefetch the next particle of this thread
      pn = __ldcs(spos + qn);
#pragma unroll
      for (int f = 0; f < 3; ++f) vn[f] = __ldcs(svel + f * np + qn);
    }
    if (l2ahead > 0) {   // and ask L2 for the lines of the iterations after that
      const int qf = qn + l2ahead * (int)blockDim.x;
      if (qf < end) {
        prefetch_l2(spos + qf);
        if ((threadIdx.x & 3) == 0) {   // one request per 16 bytes of each velocity row is plenty
#pragma unroll
          for (int f = 0; f < 3; ++f) prefetch_l2(svel + f * np + qf);
        }
      }
    }
    // (1) stencil and destination tile; the slot claim (a global atomic with ~1 us round trip) is
    //     issued right away so that its latency is covered by the gather below
    int tt = 0, lx = 0, ly = 0, lz = 0;
    bool fast = false;
    FastStencil s;
    if (valid) {
      fast = stencil_box<REL, B, B, BZ>(g, p, ox, oy, oz, s, lx, ly, lz);
      if (fast) {
        tt = ((s.i0 >> TS) * g.nty + (s.j0 >> TS)) * g.ntz + (s.k0 >> TS);
      } else {
        Cic1 cx, cy, cz;
        sim_stencil<REL>(g, p.x, p.y, p.z, __float_as_int(p.w), cx, cy, cz);
        tt = tile_of<TS>(g, cx.i0, cy.i0, cz.i0);
      }
    }
    const int lane = threadIdx.x & 31;
    const unsigned peers = __match_any_sync(0xffffffffu, valid ? tt : (0x40000000 | lane));
    const int leader = __ffs(peers) - 1;
    int slot = 0;
    if (valid && lane == leader) slot = atomicAdd(cursor + tt, __popc(peers));
    // (2) gather + kick + drift
    float v[3] = {0, 0, 0};
    if (valid) {
      float acc[3] = {0.f, 0.f, 0.f};
      if (fast) {
        const float* b0 = box + (lx * B + ly) * BZ + lz;
        const float w00 = s.wx0 * s.wy0, w10 = s.wx1 * s.wy0, w01 = s.wx0 * s.wy1, w11 = s.wx1 * s.wy1;
        const float kk[8] = {w00 * s.wz0, w00 * s.wz1, w01 * s.wz0, w01 * s.wz1,
                             w10 * s.wz0, w10 * s.wz1, w11 * s.wz0, w11 * s.wz1};
        constexpr int off[8] = {0, 1, BZ, BZ + 1, B * BZ, B * BZ + 1, B * BZ + BZ, B * BZ + BZ + 1};
        float mv[3][8];
#pragma unroll
        for (int f = 0; f < 3; ++f)
#pragma unroll
          for (int c = 0; c < 8; ++c) mv[f][c] = b0[f * NBOX + off[c]];   // 24 LDS, immediate offsets
#pragma unroll
        for (int c = 0; c < 8; ++c)
#pragma unroll
          for (int f = 0; f < 3; ++f) acc[f] = fmaf(mv[f][c], kk[c], acc[f]);
      } else {
        Cic1 cx, cy, cz;
        sim_stencil<REL>(g, p.x, p.y, p.z, __float_as_int(p.w), cx, cy, cz);
        Corners c;
        make_corners<B, BZ>(g, cx, cy, cz, ox, oy, oz, c);
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) {
          const int a = cc & 1, b = (cc >> 1) & 1, d = cc >> 2;
          if (REL && (c.ix[a] < 0 || c.iy[b] < 0 || c.iz[d] < 0)) continue;
          const float k = (c.wx[a] * c.wy[b]) * c.wz[d];
          if (c.inside) {
            const int o = (c.lx[a] * B + c.ly[b]) * BZ + c.lz[d];
#pragma unroll
            for (int f = 0; f < 3; ++f) acc[f] = fmaf(box[f * NBOX + o], k, acc[f]);
          } else {
            const long long o = mesh_index(g, c.ix[a], c.iy[b], c.iz[d]);
#pragma unroll
            for (int f = 0; f < 3; ++f) acc[f] = fmaf(__ldg(fm[f] + o), k, acc[f]);
          }
        }
        if (c.inside) ++nslow; else atomicAdd(stats + 1, 1ull);
      }
      if (FMAX) fmax = fmaxf(fmax, fmaxf(fabsf(acc[0]), fmaxf(fabsf(acc[1]), fabsf(acc[2]))));
#pragma unroll
      for (int f = 0; f < 3; ++f) v[f] = fmaf(kick, acc[f], vin[f]);
      p.x = fmaf(drift, v[0], p.x);
      p.y = fmaf(drift, v[1], p.y);
      p.z = fmaf(drift, v[2], p.z);
    }
    // (3) finish the claim: broadcast the group's base slot, rank within the group
    slot = __shfl_sync(0xffffffffu, slot, leader) + __popc(peers & ((1u << lane) - 1u));
    if (valid) {
      npos[slot] = p;
#pragma unroll
      for (int f = 0; f < 3; ++f) nvel[f * np + slot] = v[f];
    }
    p = pn;
    q = qn;
#pragma unroll
    for (int f = 0; f < 3; ++f) vin[f] = vn[f];
  }
  nslow = __reduce_add_sync(0xffffffffu, nslow);
  if ((threadIdx.x & 31) == 0 && nslow) atomicAdd(stats + 3, (unsigned long long)nslow);
  if (FMAX && fmax_bits) {   // largest force component seen (non-negative floats order like their bit patterns)
    const unsigned m = __reduce_max_sync(0xffffffffu, __float_as_uint(fmax));
    if ((threadIdx.x & 31) == 0 && m) atomicMax(fmax_bits, m);
  }
}

// ---- forces only: the three reads + stack of pm_forces (pm.py:54-56) from the tile-sorted state ------------
// Same staging / gather as sim_read_kernel (TMA flavour), no kick / drift / re-sort: the interpolated force of
// every particle goes to out[id][3] in the CALLER's particle order (id = Lagrangian site or list index).
template <bool REL, int TS, int M>
__global__ void __launch_bounds__(512, 2)
sim_forces_kernel(const __grid_constant__ CUtensorMap tm, SimGeom g, const float4* __restrict__ spos,
                  const int* __restrict__ start, const float* __restrict__ f0, const float* __restrict__ f1,
                  const float* __restrict__ f2, float scale, float* __restrict__ out,
                  unsigned long long* __restrict__ stats) {
  constexpr int T = 1 << TS, B = T + 2 * M + 1, MZ = kTmaMz;
  constexpr int BZ = (T + MZ + M + 1 + 3) & ~3, NBOX = B * B * BZ;
  extern __shared__ __align__(128) float box[];        // [3][B][B][BZ]
  __shared__ __align__(8) unsigned long long mbar;
  const int t = blockIdx.x;
  const int beg = start[t], end = start[t + 1];
  if (beg == end) return;
  const int tz = t % g.ntz, ty = (t / g.ntz) % g.nty, tx = t / (g.ntz * g.nty);
  const int ox = (tx << TS) - M, oy = (ty << TS) - M, oz = (tz << TS) - MZ;
  const float* fm[3] = {f0, f1, f2};
  if (threadIdx.x == 0) {
    mbar_init(&mbar, 1);
    fence_async_smem();
    mbar_expect_tx(&mbar, 3u * NBOX * (unsigned)sizeof(float));
    tma_load_4d(box, &tm, oz + g.mo, oy + g.mo, ox + g.mox, 0, &mbar);
  }
  int q = beg + threadIdx.x;
  float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
  if (q < end) p = __ldcs(spos + q);
  __syncthreads();
  mbar_wait(&mbar, 0);
  int nslow = 0;
  for (; q < end; q += blockDim.x) {
    float4 pn = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q + (int)blockDim.x < end) pn = __ldcs(spos + q + blockDim.x);
    FastStencil s;
    int lx, ly, lz;
    float acc[3] = {0.f, 0.f, 0.f};
    if (stencil_box<REL, B, B, BZ>(g, p, ox, oy, oz, s, lx, ly, lz)) {
      const float* b0 = box + (lx * B + ly) * BZ + lz;
      const float w00 = s.wx0 * s.wy0, w10 = s.wx1 * s.wy0, w01 = s.wx0 * s.wy1, w11 = s.wx1 * s.wy1;
      const float kk[8] = {w00 * s.wz0, w00 * s.wz1, w01 * s.wz0, w01 * s.wz1,
                           w10 * s.wz0, w10 * s.wz1, w11 * s.wz0, w11 * s.wz1};
      constexpr int off[8] = {0, 1, BZ, BZ + 1, B * BZ, B * BZ + 1, B * BZ + BZ, B * BZ + BZ + 1};
#pragma unroll
      for (int c = 0; c < 8; ++c)
#pragma unroll
        for (int f = 0; f < 3; ++f) acc[f] = fmaf(b0[f * NBOX + off[c]], kk[c], acc[f]);
    } else {
      Cic1 cx, cy, cz;
      sim_stencil<REL>(g, p.x, p.y, p.z, __float_as_int(p.w), cx, cy, cz);
      Corners c;
      make_corners<B, BZ>(g, cx, cy, cz, ox, oy, oz, c);
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) {
        const int a = cc & 1, b = (cc >> 1) & 1, d = cc >> 2;
        if (REL && (c.ix[a] < 0 || c.iy[b] < 0 || c.iz[d] < 0)) continue;
        const float k = (c.wx[a] * c.wy[b]) * c.wz[d];
        if (c.inside) {
          const int o = (c.lx[a] * B + c.ly[b]) * BZ + c.lz[d];
#pragma unroll
          for (int f = 0; f < 3; ++f) acc[f] = fmaf(box[f * NBOX + o], k, acc[f]);
        } else {
          const long long o = mesh_index(g, c.ix[a], c.iy[b], c.iz[d]);
#pragma unroll
          for (int f = 0; f < 3; ++f) acc[f] = fmaf(__ldg(fm[f] + o), k, acc[f]);
        }
      }
      if (c.inside) ++nslow; else atomicAdd(stats + 1, 1ull);
    }
    const int w = __float_as_int(p.w);
    const long long id = REL ? ((long long)(w >> 20) * g.pny + ((w >> 10) & 1023)) * g.pnz + (w & 1023) : (long long)w;
    out[3 * id] = acc[0] * scale;
    out[3 * id + 1] = acc[1] * scale;
    out[3 * id + 2] = acc[2] * scale;
    p = pn;
  }
  nslow = __reduce_add_sync(0xffffffffu, nslow);
  if ((threadIdx.x & 31) == 0 && nslow) atomicAdd(stats + 3, (unsigned long long)nslow);
}

template <int TS, int M, bool TMA> constexpr int paint_smem() {
  constexpr int B = (1 << TS) + 2 * M + 1;
  constexpr int NBOX = B * B * (TMA ? (((1 << TS) + kTmaMz + M + 1 + 3) & ~3) : ((B + 1) & ~1));
  return (NBOX + NBOX / 2) * (int)sizeof(unsigned);   // lo[NBOX] + hi[NBOX / 2]
}
template <int TS, int M, bool TMA> constexpr int read_smem() {
  constexpr int B = (1 << TS) + 2 * M + 1;
  return 3 * B * B * (TMA ? (((1 << TS) + kTmaMz + M + 1 + 3) & ~3) : B) * (int)sizeof(float);
}

// (tile shift, margin) instantiations
#define JPM_SIM_DISPATCH(ts, m, MACRO)                        \
  do {                                                        \
    if (ts == 3 && m == 0) { MACRO(3, 0); }                   \
    else if (ts == 3 && m == 1) { MACRO(3, 1); }              \
    else if (ts == 3 && m == 2) { MACRO(3, 2); }              \
    else if (ts == 3 && m == 3) { MACRO(3, 3); }              \
    else if (ts == 4 && m == 0) { MACRO(4, 0); }              \
    else if (ts == 4 && m == 1) { MACRO(4, 1); }              \
    else if (ts == 4 && m == 2) { MACRO(4, 2); }              \
    else if (ts == 4 && m == 3) { MACRO(4, 3); }              \
    else { set_error("unsupported tile/margin"); return JPM_ERR_INVALID; } \
  } while (0)
// the TMA flavour is instantiated for the tile/margin pairs the step driver uses
#define JPM_SIM_DISPATCH_TMA(ts, m, MACRO)                    \
  do {                                                        \
    if (ts == 3 && m == 1) { MACRO(3, 1); }                   \
    else if (ts == 3 && m == 2) { MACRO(3, 2); }              \
    else if (ts == 4 && m == 1) { MACRO(4, 1); }              \
    else if (ts == 4 && m == 2) { MACRO(4, 2); }              \
    else { set_error("unsupported tile/margin for the TMA path"); return JPM_ERR_INVALID; } \
  } while (0)
static bool tma_pair(int ts, int m) { return (ts == 3 || ts == 4) && (m == 1 || m == 2); }

static SimGeom make_geom(int nx, int ny, int nz, int pny, int pnz, int hx, int hy, int tile, int m) {
  SimGeom g;
  g.nx = nx; g.ny = ny; g.nz = nz; g.pny = pny; g.pnz = pnz; g.hx = hx; g.hy = hy;
  g.tshift = 0;
  while ((1 << g.tshift) < tile) ++g.tshift;
  g.T = 1 << g.tshift;
  g.m = m;
  g.ntx = (nx + g.T - 1) / g.T; g.nty = (ny + g.T - 1) / g.T; g.ntz = (nz + g.T - 1) / g.T;
  g.nt = g.ntx * g.nty * g.ntz;
  g.BX = g.BY = g.BZ = g.T + 2 * m + 1;
  g.msx = (long long)ny * nz; g.msy = nz; g.mb = (long long)nx * ny * nz; g.mo = 0; g.mox = 0;   // compact mesh
  return g;
}

}  // namespace jpm

using namespace jpm;

extern "C" int32_t jpm_sim_create(jpm_sim** out, jpm_plan* plan, int32_t nx, int32_t ny, int32_t nz,
                                  int32_t pnx, int32_t pny, int32_t pnz, int32_t hx, int32_t hy,
                                  int32_t relative, int32_t tile, int32_t margin) {
  return jpm_sim_create_ex(out, plan, nx, ny, nz, pnx, pny, pnz, hx, hy, relative, tile, margin, 0);
}

extern "C" int32_t jpm_sim_create_ex(jpm_sim** out, jpm_plan* plan, int32_t nx, int32_t ny, int32_t nz,
                                     int32_t pnx, int32_t pny, int32_t pnz, int32_t hx, int32_t hy,
                                     int32_t relative, int32_t tile, int32_t margin, int32_t flags) {
  JPM_CHECK_ARG(out, "null sim pointer");
  JPM_CHECK_ARG(nx > 0 && ny > 0 && nz > 0 && pnx > 0 && pny > 0 && pnz > 0, "bad shape");
  JPM_CHECK_ARG((int64_t)nx * ny * nz < (1ll << 31), "mesh too large for int32 cell ids");
  JPM_CHECK_ARG((int64_t)pnx * pny * pnz < (1ll << 31), "too many particles for int32 ids");
  JPM_CHECK_ARG(tile == 8 || tile == 16, "tile must be 8 or 16");
  JPM_CHECK_ARG(margin >= 0 && margin <= 3, "margin must be in [0, 3]");
  JPM_CHECK_ARG(hx >= 0 && hy >= 0, "bad halo");
  if (relative) {
    JPM_CHECK_ARG(pnx + 2 * hx == nx && pny + 2 * hy == ny && pnz == nz,
                  "relative mode: mesh must be the particle grid padded by the halo");
    JPM_CHECK_ARG(pnx <= 1024 && pny <= 1024 && pnz <= 1024,
                  "relative mode: local particle grid limited to 1024 per axis (packed ids)");
  }
  if (plan) JPM_CHECK_ARG(plan->nx == nx && plan->ny == ny && plan->nz == nz, "plan shape != sim mesh shape");
  jpm_sim* s = new jpm_sim();
  s->plan = plan;
  s->pos_only = (flags & JPM_SIM_POSITIONS_ONLY) != 0;
  s->relative = relative;
  s->np = (long long)pnx * pny * pnz;
  s->g = make_geom(nx, ny, nz, pny, pnz, hx, hy, tile, margin);
  const int ts = s->g.tshift, m = margin;
#define SET_ATTR(TS_, M_)                                                                                   \
  {                                                                                                         \
    JPM_CUDA(cudaFuncSetAttribute(sim_paint_kernel<false, TS_, M_, false>,                                  \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, paint_smem<TS_, M_, false>())); \
    JPM_CUDA(cudaFuncSetAttribute(sim_paint_kernel<true, TS_, M_, false>,                                   \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, paint_smem<TS_, M_, false>())); \
    JPM_CUDA(cudaFuncSetAttribute(sim_read_kernel<false, TS_, M_, false>,                                   \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, read_smem<TS_, M_, false>()));  \
    JPM_CUDA(cudaFuncSetAttribute(sim_read_kernel<true, TS_, M_, false>,                                    \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, read_smem<TS_, M_, false>()));  \
  }
  JPM_SIM_DISPATCH(ts, m, SET_ATTR);
#undef SET_ATTR
  // TMA path: needs the plan's ghost-zone meshes (jpm_sim_step only) and a supported tile/margin pair
  memset(&s->tm_rho, 0, sizeof(CUtensorMap));
  memset(&s->tm_f3, 0, sizeof(CUtensorMap));
  if (plan && tma_pair(ts, m) && nx >= s->g.T && ny >= s->g.T && nz >= s->g.T) {
    int32_t rc = plan_enable_padded(plan);
    if (rc) return rc;
    if (plan->G > 0) {
      const int B = s->g.T + 2 * m + 1, BZ = (s->g.T + kTmaMz + m + 1 + 3) & ~3;
      const unsigned long long d3[3] = {(unsigned long long)plan->nzp, (unsigned long long)plan->nyp,
                                        (unsigned long long)plan->nxp};
      const unsigned long long st3[2] = {(unsigned long long)plan->nzp * 4, (unsigned long long)plan->nyp * plan->nzp * 4};
      const unsigned b3[3] = {(unsigned)BZ, (unsigned)B, (unsigned)B};
      if ((rc = encode_tensor_map(&s->tm_rho, plan->density_p, 3, d3, st3, b3))) return rc;
      const unsigned long long d4[4] = {d3[0], d3[1], d3[2], 3ull};
      const unsigned long long st4[3] = {st3[0], st3[1], (unsigned long long)plan->npad * 4};
      const unsigned b4[4] = {(unsigned)BZ, (unsigned)B, (unsigned)B, 3u};
      if ((rc = encode_tensor_map(&s->tm_f3, plan->force3_p, 4, d4, st4, b4))) return rc;
      s->gp = s->g;
      s->gp.msx = (long long)plan->nyp * plan->nzp; s->gp.msy = plan->nzp; s->gp.mb = plan->npad; s->gp.mo = plan->G;
      s->gp.mox = plan->G;
      s->tma = true;
#define SET_ATTR_TMA(TS_, M_)                                                                              \
  {                                                                                                         \
    JPM_CUDA(cudaFuncSetAttribute(sim_paint_kernel<false, TS_, M_, true>,                                   \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, paint_smem<TS_, M_, true>())); \
    JPM_CUDA(cudaFuncSetAttribute(sim_paint_kernel<true, TS_, M_, true>,                                    \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, paint_smem<TS_, M_, true>())); \
    JPM_CUDA(cudaFuncSetAttribute(sim_read_kernel<false, TS_, M_, true>,                                    \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, read_smem<TS_, M_, true>()));  \
    JPM_CUDA(cudaFuncSetAttribute(sim_read_kernel<true, TS_, M_, true>,                                     \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, read_smem<TS_, M_, true>()));  \
    JPM_CUDA(cudaFuncSetAttribute(sim_read_kernel<false, TS_, M_, true, true>,                              \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, read_smem<TS_, M_, true>()));  \
    JPM_CUDA(cudaFuncSetAttribute(sim_read_kernel<true, TS_, M_, true, true>,                               \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, read_smem<TS_, M_, true>()));  \
    JPM_CUDA(cudaFuncSetAttribute(sim_forces_kernel<false, TS_, M_>,                                        \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, read_smem<TS_, M_, true>()));  \
    JPM_CUDA(cudaFuncSetAttribute(sim_forces_kernel<true, TS_, M_>,                                         \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, read_smem<TS_, M_, true>()));  \
  }
      JPM_SIM_DISPATCH_TMA(ts, m, SET_ATTR_TMA);
#undef SET_ATTR_TMA
      s->pot_ok = plan->fft_on;      // potential chain: fused FFT chain (power-of-two mesh) + gradient pass
    }
  }
  JPM_CUDA(cudaMallocHost(&s->stats_host, 3 * 4 * sizeof(double)));
  for (int i = 0; i < 3; ++i) JPM_CUDA(cudaEventCreateWithFlags(&s->stats_ev[i], cudaEventDisableTiming));
  if (const char* e = getenv("JPM_FORCE_MODE")) {   // default force path of new sims (tests / A-B runs)
    s->force_mode = atoi(e);
    s->cur_mode = s->force_mode == 1 ? 1 : 0;
  }
  for (int i = 0; i < (s->pos_only ? 1 : 2); ++i) {
    JPM_CUDA(cudaMalloc(&s->pos[i], s->np * sizeof(float4)));
    if (!s->pos_only) JPM_CUDA(cudaMalloc(&s->vel[i], 3 * s->np * sizeof(float)));
    JPM_CUDA(cudaMalloc(&s->start[i], (s->g.nt + 1) * sizeof(int)));
  }
  JPM_CUDA(cudaMalloc(&s->count, s->g.nt * sizeof(int)));
  JPM_CUDA(cudaMalloc(&s->cursor, s->g.nt * sizeof(int)));
  JPM_CUDA(cudaMalloc(&s->stats, 4 * sizeof(unsigned long long)));
  JPM_CUDA(cudaMemset(s->stats, 0, 4 * sizeof(unsigned long long)));
  JPM_CUDA(cudaMemset(s->count, 0, s->g.nt * sizeof(int)));
  *out = s;
  return JPM_OK;
}

extern "C" int32_t jpm_sim_destroy(jpm_sim* s) {
  if (!s) return JPM_OK;
  for (int i = 0; i < 2; ++i) {
    if (s->pos[i]) cudaFree(s->pos[i]);
    if (s->vel[i]) cudaFree(s->vel[i]);
    if (s->start[i]) cudaFree(s->start[i]);
  }
  if (s->count) cudaFree(s->count);
  if (s->cursor) cudaFree(s->cursor);
  if (s->stats) cudaFree(s->stats);
  if (s->stats_host) cudaFreeHost(s->stats_host);
  for (int i = 0; i < 3; ++i)
    if (s->stats_ev[i]) cudaEventDestroy(s->stats_ev[i]);
  for (int i = 0; i < 2; ++i) {
    if (s->stage_in[i]) cudaFree(s->stage_in[i]);
    if (s->stage_out[i]) cudaFree(s->stage_out[i]);
    for (cudaEvent_t e : {s->ev_in_ready[i], s->ev_in_free[i], s->ev_out_ready[i], s->ev_out_free[i]})
      if (e) cudaEventDestroy(e);
  }
  if (s->st_h2d) cudaStreamDestroy(s->st_h2d);
  if (s->st_d2h) cudaStreamDestroy(s->st_d2h);
  delete s;
  return JPM_OK;
}

extern "C" int32_t jpm_sim_load(jpm_sim* s, void* stream, const float* pos, const float* vel) {
  JPM_CHECK_ARG(s && pos, "null pointer");
  JPM_CHECK_ARG(s->pos_only ? vel == nullptr : vel != nullptr,
                "velocities are required, except for a JPM_SIM_POSITIONS_ONLY sim, which takes none");
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = div_up(s->np, 256);
  JPM_CUDA(cudaMemsetAsync(s->count, 0, s->g.nt * sizeof(int), st));
  const bool t8 = s->g.tshift == 3;
  if (s->relative) {
    if (t8) sim_count_kernel<true, 3><<<blocks, 256, 0, st>>>(s->g, pos, s->np, s->count);
    else sim_count_kernel<true, 4><<<blocks, 256, 0, st>>>(s->g, pos, s->np, s->count);
  } else {
    if (t8) sim_count_kernel<false, 3><<<blocks, 256, 0, st>>>(s->g, pos, s->np, s->count);
    else sim_count_kernel<false, 4><<<blocks, 256, 0, st>>>(s->g, pos, s->np, s->count);
  }
  JPM_LAUNCH_CHECK();
  s->cur = 0;
  sim_scan_kernel<<<1, 1024, 0, st>>>(s->count, s->start[0], s->cursor, s->g.nt);
  JPM_LAUNCH_CHECK();
  if (s->relative) {
    if (t8) sim_fill_kernel<true, 3><<<blocks, 256, 0, st>>>(s->g, pos, vel, s->np, s->cursor, s->pos[0], s->vel[0]);
    else sim_fill_kernel<true, 4><<<blocks, 256, 0, st>>>(s->g, pos, vel, s->np, s->cursor, s->pos[0], s->vel[0]);
  } else {
    if (t8) sim_fill_kernel<false, 3><<<blocks, 256, 0, st>>>(s->g, pos, vel, s->np, s->cursor, s->pos[0], s->vel[0]);
    else sim_fill_kernel<false, 4><<<blocks, 256, 0, st>>>(s->g, pos, vel, s->np, s->cursor, s->pos[0], s->vel[0]);
  }
  JPM_LAUNCH_CHECK();
  s->loaded = true;
  s->painted = false;
  return JPM_OK;
}

extern "C" int32_t jpm_sim_store(jpm_sim* s, void* stream, float* pos, float* vel) {
  JPM_CHECK_ARG(s && (pos || vel), "null pointer");
  JPM_CHECK_ARG(!(s->pos_only && vel), "a JPM_SIM_POSITIONS_ONLY sim holds no velocities");
  JPM_CHECK_ARG(s->loaded, "sim has no particles loaded");
  cudaStream_t st = (cudaStream_t)stream;
  if (s->relative)
    sim_store_kernel<true><<<div_up(s->np, 256), 256, 0, st>>>(s->g, s->pos[s->cur], s->vel[s->cur], s->np, pos, vel);
  else
    sim_store_kernel<false><<<div_up(s->np, 256), 256, 0, st>>>(s->g, s->pos[s->cur], s->vel[s->cur], s->np, pos, vel);
  JPM_LAUNCH_CHECK();
  return JPM_OK;
}

// paint into a compact caller mesh (TMA = false) or into the plan's ghost-zone density (TMA = true)
static int32_t sim_paint_impl(jpm_sim* s, cudaStream_t st, float* mesh, bool tma) {
  JPM_CUDA(cudaMemsetAsync(s->count, 0, s->g.nt * sizeof(int), st));
  const int ts = s->g.tshift, m = s->g.m;
  const SimGeom& g = tma ? s->gp : s->g;
  // slab plans: record the x planes the particles touch, the first barrier of the step turns it into the ghost width
  int* xrange = (tma && s->plan && s->plan->is_slab && s->plan->slab.P > 1)
                    ? reinterpret_cast<int*>(s->plan->slab.flags[s->plan->slab.rank]) + kFlagXmin : nullptr;
  const int track_y = (xrange && s->plan->slab.py > 1) ? 1 : 0;
#define LAUNCH_PAINT_T(TS_, M_, TMA_)                                                                    \
  {                                                                                                      \
    if (s->relative)                                                                                     \
      sim_paint_kernel<true, TS_, M_, TMA_><<<g.nt, 256, paint_smem<TS_, M_, TMA_>(), st>>>(             \
          s->tm_rho, g, s->pos[s->cur], s->start[s->cur], mesh, s->count, s->stats, xrange, track_y);    \
    else                                                                                                 \
      sim_paint_kernel<false, TS_, M_, TMA_><<<g.nt, 256, paint_smem<TS_, M_, TMA_>(), st>>>(            \
          s->tm_rho, g, s->pos[s->cur], s->start[s->cur], mesh, s->count, s->stats, xrange, track_y);    \
  }
#define LAUNCH_PAINT(TS_, M_) LAUNCH_PAINT_T(TS_, M_, false)
#define LAUNCH_PAINT_TMA(TS_, M_) LAUNCH_PAINT_T(TS_, M_, true)
  if (tma) JPM_SIM_DISPATCH_TMA(ts, m, LAUNCH_PAINT_TMA);
  else JPM_SIM_DISPATCH(ts, m, LAUNCH_PAINT);
#undef LAUNCH_PAINT
#undef LAUNCH_PAINT_TMA
#undef LAUNCH_PAINT_T
  JPM_LAUNCH_CHECK();
  s->painted = true;
  return JPM_OK;
}

static int32_t sim_read_impl(jpm_sim* s, cudaStream_t st, const float* fx, const float* fy, const float* fz,
                             float kick_coef, float drift_coef, bool tma) {
  const int nxt = s->cur ^ 1;
  sim_scan_kernel<<<1, 1024, 0, st>>>(s->count, s->start[nxt], s->cursor, s->g.nt);
  JPM_LAUNCH_CHECK();
  const int ts = s->g.tshift, m = s->g.m;
  const SimGeom& g = tma ? s->gp : s->g;
  // AUTO, spectral step: the read tracks max |F| (a potential step gets it from the gradient pass instead)
  const bool want_fmax = tma && s->force_mode == 2 && s->cur_mode == 0 && s->plan && s->plan->pot_stats;
  unsigned* fmax_bits = want_fmax ? reinterpret_cast<unsigned*>(s->plan->pot_stats + 1) : nullptr;
  static const int l2ahead = getenv("JPM_L2_AHEAD") ? atoi(getenv("JPM_L2_AHEAD")) : kL2Ahead;
#define LAUNCH_READ_T(TS_, M_, TMA_)                                                                     \
  if (TMA_ && want_fmax) {                                                                               \
    if (s->relative)                                                                                     \
      sim_read_kernel<true, TS_, M_, TMA_, TMA_><<<g.nt, 512, read_smem<TS_, M_, TMA_>(), st>>>(         \
          s->tm_f3, g, s->pos[s->cur], s->vel[s->cur], s->start[s->cur], fx, fy, fz, kick_coef,          \
          drift_coef, s->np, s->cursor, s->pos[nxt], s->vel[nxt], s->stats, l2ahead, fmax_bits);         \
    else                                                                                                 \
      sim_read_kernel<false, TS_, M_, TMA_, TMA_><<<g.nt, 512, read_smem<TS_, M_, TMA_>(), st>>>(        \
          s->tm_f3, g, s->pos[s->cur], s->vel[s->cur], s->start[s->cur], fx, fy, fz, kick_coef,          \
          drift_coef, s->np, s->cursor, s->pos[nxt], s->vel[nxt], s->stats, l2ahead, fmax_bits);         \
  } else {                                                                                               \
    if (s->relative)                                                                                     \
      sim_read_kernel<true, TS_, M_, TMA_><<<g.nt, 512, read_smem<TS_, M_, TMA_>(), st>>>(               \
          s->tm_f3, g, s->pos[s->cur], s->vel[s->cur], s->start[s->cur], fx, fy, fz, kick_coef,          \
          drift_coef, s->np, s->cursor, s->pos[nxt], s->vel[nxt], s->stats, l2ahead, fmax_bits);         \
    else                                                                                                 \
      sim_read_kernel<false, TS_, M_, TMA_><<<g.nt, 512, read_smem<TS_, M_, TMA_>(), st>>>(              \
          s->tm_f3, g, s->pos[s->cur], s->vel[s->cur], s->start[s->cur], fx, fy, fz, kick_coef,          \
          drift_coef, s->np, s->cursor, s->pos[nxt], s->vel[nxt], s->stats, l2ahead, fmax_bits);         \
  }
#define LAUNCH_READ(TS_, M_) LAUNCH_READ_T(TS_, M_, false)
#define LAUNCH_READ_TMA(TS_, M_) LAUNCH_READ_T(TS_, M_, true)
  if (tma) JPM_SIM_DISPATCH_TMA(ts, m, LAUNCH_READ_TMA);
  else JPM_SIM_DISPATCH(ts, m, LAUNCH_READ);
#undef LAUNCH_READ
#undef LAUNCH_READ_TMA
#undef LAUNCH_READ_T
  JPM_LAUNCH_CHECK();
  s->cur = nxt;
  s->painted = false;
  return JPM_OK;
}

// fp32 cancellation bound of the potential path.  Differencing psi (rms psi_rms, FFT round-off ~ 5e-7 of its
// maximum) leaves an absolute force error ~ 5.4e-7 * max|psi|; measured against the float64 oracle on 256^3 and
// 512^3 LCDM fields (linear a = 0.1 and clustered a = 1, tools/phi_fd_precision.py): max|psi| <= 5.3 psi_rms, so
//     max|dF| / max|F|  <~  2.7e-6 * psi_rms / max|F|.
// AUTO runs the three-transform chain while that bound exceeds kPotSwitchUp and the potential chain below
// kPotSwitchDown (hysteresis), so every step stays inside BASELINE.json's 1e-5 force tolerance.
constexpr double kPotErrCoef = 2.7e-6, kPotSwitchDown = 4.0e-6, kPotSwitchUp = 6.0e-6;

constexpr int kStatsLag = 2;

// evaluate the statistics of step `step` (waits for that step to finish on the device)
static void sim_eval_stats(jpm_sim* s, long long step) {
  if (step < 0) return;
  const int i = (int)(step % 3);
  if (!s->stats_pending[i]) return;
  if (cudaEventSynchronize(s->stats_ev[i]) != cudaSuccess) return;
  s->stats_pending[i] = false;
  const double* h = s->stats_host + 4 * i;
  double sumsq = h[0];
  if (s->stats_fixed[i]) {
    unsigned long long q;
    memcpy(&q, &h[0], sizeof(q));
    sumsq = (double)q / kStatsFix;
  }
  unsigned long long bits;
  memcpy(&bits, &h[1], sizeof(bits));
  const unsigned fb = (unsigned)(bits & 0xffffffffull);
  float fmax;
  memcpy(&fmax, &fb, sizeof(fmax));
  if (!(sumsq > 0.0) || !(fmax > 0.f)) return;
  s->last_bound = kPotErrCoef * std::sqrt(sumsq) / (double)fmax;
  if (s->force_mode == 2) {
    if (s->last_bound < kPotSwitchDown) s->cur_mode = 1;
    else if (s->last_bound > kPotSwitchUp) s->cur_mode = 0;
  }
}

extern "C" int32_t jpm_sim_set_force_mode(jpm_sim* s, int32_t mode) {
  JPM_CHECK_ARG(s, "null sim");
  JPM_CHECK_ARG(mode >= 0 && mode <= 2, "force mode must be 0 (spectral), 1 (potential) or 2 (auto)");
  JPM_CHECK_ARG(mode == 0 || s->pot_ok,
                "potential force path needs the TMA tile path on a power-of-two mesh (fused FFT chain), one GPU");
  s->force_mode = mode;
  s->cur_mode = mode == 1 ? 1 : 0;
  return JPM_OK;
}

extern "C" int32_t jpm_sim_force_info(jpm_sim* s, void* stream, double* out6_host) {
  JPM_CHECK_ARG(s && out6_host, "null pointer");
  JPM_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  for (long long n = s->nstep - 3; n < s->nstep; ++n) sim_eval_stats(s, n);     // in step order
  out6_host[0] = (double)s->force_mode;
  out6_host[1] = (double)s->cur_mode;
  out6_host[2] = s->last_bound;
  out6_host[3] = (double)s->mode_steps[0];
  out6_host[4] = (double)s->mode_steps[1];
  out6_host[5] = s->pot_ok ? 1.0 : 0.0;
  return JPM_OK;
}

extern "C" int32_t jpm_sim_paint(jpm_sim* s, void* stream, float* mesh) {
  JPM_CHECK_ARG(s && mesh, "null pointer");
  JPM_CHECK_ARG(s->loaded, "sim has no particles loaded");
  return sim_paint_impl(s, (cudaStream_t)stream, mesh, false);
}

extern "C" int32_t jpm_sim_read_kick_drift(jpm_sim* s, void* stream, const float* fx, const float* fy,
                                           const float* fz, float kick_coef, float drift_coef) {
  JPM_CHECK_ARG(s && fx && fy && fz, "null pointer");
  JPM_CHECK_ARG(s->painted, "jpm_sim_read_kick_drift must follow jpm_sim_paint (tile occupancy)");
  JPM_CHECK_ARG(!s->pos_only, "a JPM_SIM_POSITIONS_ONLY sim cannot step");
  return sim_read_impl(s, (cudaStream_t)stream, fx, fy, fz, kick_coef, drift_coef, false);
}

// pm_forces on the fast kernels: paint the loaded state (TMA reduce-add), fused FFT chain, gather the three
// force components of every particle into out[np][3] in the caller's order.
extern "C" int32_t jpm_sim_forces(jpm_sim* s, void* stream, float* out, float scale, float r_split,
                                  const float* filter_tab, int32_t n_tab, float filter_kmax) {
  JPM_CHECK_ARG(s && s->plan && out, "null pointer / sim has no FFT plan attached");
  JPM_CHECK_ARG(s->loaded, "sim has no particles loaded");
  JPM_CHECK_ARG(s->tma, "jpm_sim_forces needs the TMA tile path (tile 8/16, margin 1/2, nz % 4 == 0)");
  JPM_CHECK_ARG(!filter_tab || (n_tab >= 2 && filter_kmax > 0.f), "bad filter table");
  cudaStream_t st = (cudaStream_t)stream;
  jpm_plan* p = s->plan;
  int32_t rc;
  JPM_CUDA(cudaMemsetAsync(p->density_p, 0, p->npad * sizeof(float), st));
  if ((rc = sim_paint_impl(s, st, p->density_p, true))) return rc;
  if ((rc = plan_padded_forces(p, st, r_split, filter_tab, n_tab, filter_kmax))) return rc;
  const SimGeom& g = s->gp;
  const int ts = s->g.tshift, m = s->g.m;
  const float *fx = p->force3_p, *fy = p->force3_p + p->npad, *fz = p->force3_p + 2 * p->npad;
#define LAUNCH_FORCES(TS_, M_)                                                                           \
  {                                                                                                      \
    if (s->relative)                                                                                     \
      sim_forces_kernel<true, TS_, M_><<<g.nt, 512, read_smem<TS_, M_, true>(), st>>>(                   \
          s->tm_f3, g, s->pos[s->cur], s->start[s->cur], fx, fy, fz, scale, out, s->stats);              \
    else                                                                                                 \
      sim_forces_kernel<false, TS_, M_><<<g.nt, 512, read_smem<TS_, M_, true>(), st>>>(                  \
          s->tm_f3, g, s->pos[s->cur], s->start[s->cur], fx, fy, fz, scale, out, s->stats);              \
  }
  JPM_SIM_DISPATCH_TMA(ts, m, LAUNCH_FORCES);
#undef LAUNCH_FORCES
  JPM_LAUNCH_CHECK();
  return JPM_OK;
}

extern "C" int32_t jpm_sim_forces_batched(jpm_sim* s, void* stream, const float* positions, float* out,
                                          int32_t nbatch, float scale, float r_split, const float* filter_tab,
                                          int32_t n_tab, float filter_kmax) {
  JPM_CHECK_ARG(s && positions && out && nbatch >= 0, "bad arguments");
  JPM_CHECK_ARG(s->pos_only, "jpm_sim_forces_batched takes a JPM_SIM_POSITIONS_ONLY sim");
  for (int b = 0; b < nbatch; ++b) {
    int32_t rc = jpm_sim_load(s, stream, positions + (size_t)b * 3 * s->np, nullptr);
    if (rc) return rc;
    if ((rc = jpm_sim_forces(s, stream, out + (size_t)b * 3 * s->np, scale, r_split, filter_tab, n_tab, filter_kmax)))
      return rc;
  }
  return JPM_OK;
}

extern "C" int32_t jpm_sim_step(jpm_sim* s, void* stream, float kick_coef, float drift_coef) {
  JPM_CHECK_ARG(s && s->plan, "sim has no FFT plan attached");
  JPM_CHECK_ARG(!s->pos_only, "a JPM_SIM_POSITIONS_ONLY sim cannot step");
  JPM_CHECK_ARG(s->loaded, "sim has no particles loaded");
  cudaStream_t st = (cudaStream_t)stream;
  jpm_plan* p = s->plan;
  StageTimer* tm = p->timer;
  int32_t rc;
  if (tm) tm->mark(st, "start");
  if (s->tma) {
    // ghost-zone meshes: TMA reduce-add paint -> fused FFT chain (or ghost fold, cuFFT, ghost fill) -> TMA read
    JPM_CUDA(cudaMemsetAsync(p->density_p, 0, p->npad * sizeof(float), st));
    if (tm) tm->mark(st, "mesh_memset");
    if ((rc = sim_paint_impl(s, st, p->density_p, true))) return rc;
    if (tm) tm->mark(st, "sim_paint");
    if (s->force_mode != 0 && !s->pot_ok) {
      set_error("potential force path not available for this sim (needs the TMA tile path and the fused FFT chain)");
      return JPM_ERR_INVALID;
    }
    const bool multi = p->is_slab && p->slab.P > 1;
    // AUTO: fixed-lag decision (one GPU: statistics of step n - 2; slab ranks: the GLOBAL statistics of step n - 3,
    // which every rank receives in its own flag block, so that all ranks switch at the same step)
    if (s->force_mode == 2) sim_eval_stats(s, s->nstep - (multi ? kStatsLag + 1 : kStatsLag));
    const bool want_stats = s->force_mode == 2;
    if (want_stats && !p->pot_stats) {
      JPM_CUDA(cudaMalloc(&p->pot_stats, 4 * sizeof(double)));
    }
    const bool pot = s->force_mode != 0 && s->cur_mode == 1;
    if (multi) {
      // first barrier of the step (every rank has painted; ghost width agreed).  Behind it the global statistics of
      // the previous step are complete in this rank's flag block: copy them out, then clear this step's slot
      if ((rc = slab_barrier(p, st, true, pot ? 2 : 0))) return rc;
      if (want_stats) {
        unsigned* fl = p->slab.flags[p->slab.rank] + kFlagStats;
        if (s->nstep >= 1) {
          const int i = (int)((s->nstep - 1) % 3);
          JPM_CUDA(cudaMemcpyAsync(s->stats_host + 4 * i, fl + 4 * ((s->nstep - 1) & 1), 16, cudaMemcpyDeviceToHost, st));
          JPM_CUDA(cudaEventRecord(s->stats_ev[i], st));
          s->stats_pending[i] = true;
          s->stats_fixed[i] = true;
        }
        JPM_CUDA(cudaMemsetAsync(fl + 4 * (s->nstep & 1), 0, 16, st));
      }
    }
    if (pot) {
      // potential chain: ONE inverse transform, then one real-space pass psi -> three force meshes
      if ((rc = pmfft_potential(p, st, 0.f, nullptr, 0, 0.f, true, multi))) return rc;
      if ((rc = pmfft_gradient(p, st))) return rc;
      rc = sim_read_impl(s, st, p->force3_p, p->force3_p + p->npad, p->force3_p + 2 * p->npad, kick_coef,
                         drift_coef, true);
      if (tm) tm->mark(st, "tile_scan+sim_read3_kick_drift");
      ++s->mode_steps[1];
    } else {
      if (want_stats) JPM_CUDA(cudaMemsetAsync(p->pot_stats, 0, 4 * sizeof(double), st));
      p->want_sumsq = want_stats;
      rc = p->fft_on ? pmfft_forces(p, st, 0.f, nullptr, 0, 0.f, multi) : plan_padded_forces(p, st, 0.f, nullptr, 0, 0.f);
      p->want_sumsq = false;
      if (rc) return rc;
      rc = sim_read_impl(s, st, p->force3_p, p->force3_p + p->npad, p->force3_p + 2 * p->npad, kick_coef,
                         drift_coef, true);
      if (tm) tm->mark(st, "tile_scan+sim_read3_kick_drift");
      ++s->mode_steps[0];
    }
    if (rc == JPM_OK && want_stats) {
      if (multi) {
        rc = slab_stats_share(p, st, (int)(s->nstep & 1));
      } else {
        const int i = (int)(s->nstep % 3);
        JPM_CUDA(cudaMemcpyAsync(s->stats_host + 4 * i, p->pot_stats, 4 * sizeof(double), cudaMemcpyDeviceToHost, st));
        JPM_CUDA(cudaEventRecord(s->stats_ev[i], st));
        s->stats_pending[i] = true;
        s->stats_fixed[i] = false;
      }
    }
    ++s->nstep;
    return rc;
  }
  JPM_CHECK_ARG(!p->is_slab, "slab plan: the tile/margin pair must be one the TMA path supports");
  JPM_CUDA(cudaMemsetAsync(p->density, 0, p->ncell * sizeof(float), st));
  if (tm) tm->mark(st, "mesh_memset");
  if ((rc = sim_paint_impl(s, st, p->density, false))) return rc;
  if (tm) tm->mark(st, "sim_paint");
  if ((rc = jpm_density_to_force_meshes(p, stream, p->density, p->force3, 0.f, nullptr, 0, 0.f))) return rc;
  rc = sim_read_impl(s, st, p->force3, p->force3 + p->ncell, p->force3 + 2 * p->ncell, kick_coef, drift_coef,
                     false);
  if (tm) tm->mark(st, "tile_scan+sim_read3_kick_drift");
  return rc;
}

// End to end through HOST buffers on the fast kernels: H2D pos / vel, tile sort (jpm_sim_load), one resident step,
// un-sort (jpm_sim_store), D2H.  Synchronises `stream`.
extern "C" int32_t jpm_sim_step_host_f32(jpm_sim* s, void* stream, float* pos_host, float* vel_host, float* pos_dev,
                                         float* vel_dev, float kick_coef, float drift_coef) {
  JPM_CHECK_ARG(s && pos_host && vel_host && pos_dev && vel_dev, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t bytes = (size_t)s->np * 3 * sizeof(float);
  JPM_CUDA(cudaMemcpyAsync(pos_dev, pos_host, bytes, cudaMemcpyHostToDevice, st));
  JPM_CUDA(cudaMemcpyAsync(vel_dev, vel_host, bytes, cudaMemcpyHostToDevice, st));
  int32_t rc;
  if ((rc = jpm_sim_load(s, stream, pos_dev, vel_dev))) return rc;
  if ((rc = jpm_sim_step(s, stream, kick_coef, drift_coef))) return rc;
  if ((rc = jpm_sim_store(s, stream, pos_dev, vel_dev))) return rc;
  JPM_CUDA(cudaMemcpyAsync(pos_host, pos_dev, bytes, cudaMemcpyDeviceToHost, st));
  JPM_CUDA(cudaMemcpyAsync(vel_host, vel_dev, bytes, cudaMemcpyDeviceToHost, st));
  JPM_CUDA(cudaStreamSynchronize(st));
  return JPM_OK;
}

// A batch of independent particle states, each taken through jpm_sim_step_host_f32's sequence (H2D, tile sort, one
// resident step, un-sort, D2H) - what a caller holding its states in host memory (or jax.vmap over host-resident
// states) issues.  The three legs of consecutive batch elements overlap: element b + 1 uploads on a copy stream while
// element b computes on `stream` and element b - 1 downloads on a second copy stream (PCIe is full duplex), through
// double-buffered device staging owned by the sim.  Per element the bytes are the same as the unpipelined entry; the
// rate tends to max(H2D, compute, D2H) instead of their sum.  Host buffers may repeat in the lists (a ring): an upload
// from a buffer waits for the last download into it.  Synchronises all three streams before returning.
extern "C" int32_t jpm_sim_steps_host_f32(jpm_sim* s, void* stream, int32_t nbatch, float* const* pos_hosts,
                                          float* const* vel_hosts, const float* kick_coefs,
                                          const float* drift_coefs) {
  JPM_CHECK_ARG(s && pos_hosts && vel_hosts && kick_coefs && drift_coefs && nbatch >= 0, "bad arguments");
  JPM_CHECK_ARG(s->plan && !s->pos_only, "sim has no FFT plan attached / holds no velocities");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n3 = (size_t)s->np * 3, bytes = n3 * sizeof(float);
  if (!s->st_h2d) {
    JPM_CUDA(cudaStreamCreateWithFlags(&s->st_h2d, cudaStreamNonBlocking));
    JPM_CUDA(cudaStreamCreateWithFlags(&s->st_d2h, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      JPM_CUDA(cudaMalloc(&s->stage_in[i], 2 * bytes));
      JPM_CUDA(cudaMalloc(&s->stage_out[i], 2 * bytes));
      JPM_CUDA(cudaEventCreateWithFlags(&s->ev_in_ready[i], cudaEventDisableTiming));
      JPM_CUDA(cudaEventCreateWithFlags(&s->ev_in_free[i], cudaEventDisableTiming));
      JPM_CUDA(cudaEventCreateWithFlags(&s->ev_out_ready[i], cudaEventDisableTiming));
      JPM_CUDA(cudaEventCreateWithFlags(&s->ev_out_free[i], cudaEventDisableTiming));
    }
  }
  std::vector<cudaEvent_t> done(nbatch, nullptr);     // download of element b finished (host ring reuse)
  int32_t rc = JPM_OK;
  auto fail = [&](cudaError_t e, const char* what) {
    if (e != cudaSuccess && rc == JPM_OK) {
      set_error("%s failed: %s", what, cudaGetErrorString(e));
      rc = JPM_ERR_CUDA;
    }
    return e != cudaSuccess;
  };
  for (int b = 0; b < nbatch && rc == JPM_OK; ++b) {
    const int i = b & 1;
    if (!pos_hosts[b] || !vel_hosts[b]) { set_error("null host buffer in batch element %d", b); rc = JPM_ERR_INVALID; break; }
    // ---- upload (copy stream 1): staging slot free (element b - 2 sorted out of it), host buffer not being written
    if (b >= 2 && fail(cudaStreamWaitEvent(s->st_h2d, s->ev_in_free[i], 0), "cudaStreamWaitEvent")) break;
    for (int c = b - 1; c >= 0; --c)
      if (pos_hosts[c] == pos_hosts[b] || vel_hosts[c] == vel_hosts[b] || pos_hosts[c] == vel_hosts[b] ||
          vel_hosts[c] == pos_hosts[b]) {
        if (fail(cudaStreamWaitEvent(s->st_h2d, done[c], 0), "cudaStreamWaitEvent")) break;
        break;
      }
    if (rc) break;
    if (fail(cudaMemcpyAsync(s->stage_in[i], pos_hosts[b], bytes, cudaMemcpyHostToDevice, s->st_h2d), "H2D")) break;
    if (fail(cudaMemcpyAsync(s->stage_in[i] + n3, vel_hosts[b], bytes, cudaMemcpyHostToDevice, s->st_h2d), "H2D")) break;
    if (fail(cudaEventRecord(s->ev_in_ready[i], s->st_h2d), "cudaEventRecord")) break;
    // ---- compute (caller's stream)
    if (fail(cudaStreamWaitEvent(st, s->ev_in_ready[i], 0), "cudaStreamWaitEvent")) break;
    if ((rc = jpm_sim_load(s, stream, s->stage_in[i], s->stage_in[i] + n3))) break;
    if (fail(cudaEventRecord(s->ev_in_free[i], st), "cudaEventRecord")) break;
    if ((rc = jpm_sim_step(s, stream, kick_coefs[b], drift_coefs[b]))) break;
    if (b >= 2 && fail(cudaStreamWaitEvent(st, s->ev_out_free[i], 0), "cudaStreamWaitEvent")) break;
    if ((rc = jpm_sim_store(s, stream, s->stage_out[i], s->stage_out[i] + n3))) break;
    if (fail(cudaEventRecord(s->ev_out_ready[i], st), "cudaEventRecord")) break;
    // ---- download (copy stream 2)
    if (fail(cudaStreamWaitEvent(s->st_d2h, s->ev_out_ready[i], 0), "cudaStreamWaitEvent")) break;
    if (fail(cudaMemcpyAsync(pos_hosts[b], s->stage_out[i], bytes, cudaMemcpyDeviceToHost, s->st_d2h), "D2H")) break;
    if (fail(cudaMemcpyAsync(vel_hosts[b], s->stage_out[i] + n3, bytes, cudaMemcpyDeviceToHost, s->st_d2h), "D2H")) break;
    if (fail(cudaEventRecord(s->ev_out_free[i], s->st_d2h), "cudaEventRecord")) break;
    if (fail(cudaEventCreateWithFlags(&done[b], cudaEventDisableTiming), "cudaEventCreate")) break;
    if (fail(cudaEventRecord(done[b], s->st_d2h), "cudaEventRecord")) break;
  }
  fail(cudaStreamSynchronize(s->st_h2d), "cudaStreamSynchronize");
  fail(cudaStreamSynchronize(st), "cudaStreamSynchronize");
  fail(cudaStreamSynchronize(s->st_d2h), "cudaStreamSynchronize");
  for (cudaEvent_t e : done)
    if (e) cudaEventDestroy(e);
  return rc;
}

// One jpm_sim_step with a CUDA event at every stage boundary.  Synchronises the stream.  names_out
// receives up to `cap` pointers to static strings, ms_out the stage durations; returns the stage count
// in *n_out.
extern "C" int32_t jpm_sim_step_profile(jpm_sim* s, void* stream, float kick_coef, float drift_coef,
                                        const char** names_out, float* ms_out, int32_t cap, int32_t* n_out) {
  JPM_CHECK_ARG(s && s->plan && names_out && ms_out && n_out && cap > 0, "bad arguments");
  StageTimer tm;
  for (int i = 0; i < StageTimer::kMax; ++i) JPM_CUDA(cudaEventCreate(&tm.ev[i]));
  s->plan->timer = &tm;
  int32_t rc = jpm_sim_step(s, stream, kick_coef, drift_coef);
  s->plan->timer = nullptr;
  if (rc == JPM_OK) {
    cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
    if (e != cudaSuccess) {
      set_error("cudaStreamSynchronize failed: %s", cudaGetErrorString(e));
      rc = JPM_ERR_CUDA;
    }
  }
  int n = 0;
  if (rc == JPM_OK) {
    for (int i = 1; i < tm.n && n < cap; ++i, ++n) {
      names_out[n] = tm.name[i];
      cudaEventElapsedTime(&ms_out[n], tm.ev[i - 1], tm.ev[i]);
    }
  }
  *n_out = n;
  for (int i = 0; i < StageTimer::kMax; ++i) cudaEventDestroy(tm.ev[i]);
  return rc;
}

extern "C" int32_t jpm_sim_stats_host(jpm_sim* s, void* stream, int64_t* out4_host) {
  JPM_CHECK_ARG(s && out4_host, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long h[4];
  JPM_CUDA(cudaMemcpyAsync(h, s->stats, sizeof(h), cudaMemcpyDeviceToHost, st));
  JPM_CUDA(cudaStreamSynchronize(st));
  for (int i = 0; i < 4; ++i) out4_host[i] = (int64_t)h[i];
  return JPM_OK;
}
