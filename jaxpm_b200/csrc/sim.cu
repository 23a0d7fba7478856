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
#include "common.cuh"

struct jpm_plan;
extern "C" int32_t jpm_density_to_force_meshes(jpm_plan*, void*, const float*, float*, float,
                                               const float*, int32_t, float);
namespace jpm {
float* plan_density(jpm_plan* p);
float* plan_force3(jpm_plan* p);
long long plan_ncell(jpm_plan* p);
void plan_dims(jpm_plan* p, int* nx, int* ny, int* nz);
}  // namespace jpm

struct SimGeom {
  int nx, ny, nz;            // mesh painted into / read from
  int pny, pnz, hx, hy;      // particle grid (relative rule) and halo offsets
  int tshift, T, m;          // tile edge = 1 << tshift, margin
  int ntx, nty, ntz, nt;     // tile grid
  int BX, BY, BZ;            // shared-memory box = T + 2m + 1 per axis
};

struct jpm_sim {
  jpm_plan* plan = nullptr;
  SimGeom g;
  int relative = 0;
  long long np = 0;
  float4* pos[2] = {nullptr, nullptr};
  float* vel[2] = {nullptr, nullptr};  // SoA [3][np]
  int* start[2] = {nullptr, nullptr};  // [nt+1]
  int* count = nullptr;                // [nt]  occupancy of the next ordering
  int* cursor = nullptr;               // [nt]  slot cursors while scattering
  unsigned long long* stats = nullptr; // [0] paint fallbacks, [1] read fallbacks
  int cur = 0;
  bool painted = false, loaded = false;
};

namespace jpm {

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
    svel[slot] = vel[3 * p];
    svel[np + slot] = vel[3 * p + 1];
    svel[2 * np + slot] = vel[3 * p + 2];
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

// exclusive scan of count[nt] -> start[nt+1]; cursor = start; count = 0.  One CTA.
__global__ void __launch_bounds__(1024)
sim_scan_kernel(int* __restrict__ count, int* __restrict__ start, int* __restrict__ cursor, int nt) {
  __shared__ int part[1024];
  const int tid = threadIdx.x;
  const int chunk = (nt + 1023) / 1024;
  const int b = tid * chunk, e = min(b + chunk, nt);
  int s = 0;
  for (int i = b; i < e; ++i) s += count[i];
  part[tid] = s;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {
    const int v = (tid >= off) ? part[tid - off] : 0;
    __syncthreads();
    part[tid] += v;
    __syncthreads();
  }
  int run = part[tid] - s;  // exclusive prefix of this chunk
  for (int i = b; i < e; ++i) {
    const int c = count[i];
    start[i] = run;
    cursor[i] = run;
    count[i] = 0;
    run += c;
  }
  if (tid == 1023) start[nt] = part[1023];
}

__device__ __forceinline__ void red_add_v2(float* addr, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}

// Per-lane view of the 8 corners, selectable at run time so that each lane can walk the corners
// in a different order (lane-rotated): particles that share a cell then update 8 DIFFERENT
// addresses at any instant and the shared-memory CAS loops rarely retry.
struct Corners {
  int ix[2], iy[2], iz[2];     // wrapped global cell indices (-1 = dropped)
  int lx[2], ly[2], lz[2];     // box-local coordinates
  float wx[2], wy[2], wz[2];
  bool inside;
};

template <int B>
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
    c.inside = c.inside && c.lx[a] < B && c.ly[a] < B && c.lz[a] < B;
  }
}

// ---- paint -------------------------------------------------------------------------------------
// One CTA per tile.  Box = (T+2M+1)^3 cells, rows padded to an even length BZ so that the flush
// can use 8-byte vector reductions (REDG.E.ADD.F32x2).
template <bool REL, int TS, int M>
__global__ void __launch_bounds__(256, 4)
sim_paint_kernel(SimGeom g, const float4* __restrict__ spos, const int* __restrict__ start,
                 float* __restrict__ mesh, int* __restrict__ count, unsigned long long* __restrict__ stats) {
  constexpr int T = 1 << TS, B = T + 2 * M + 1, BZ = (B + 1) & ~1, NBOX = B * B * BZ;
  extern __shared__ __align__(16) float box[];
  __shared__ int scnt[27];
  const int t = blockIdx.x;
  const int beg = start[t], end = start[t + 1];
  if (beg == end) return;
  const int tz = t % g.ntz, ty = (t / g.ntz) % g.nty, tx = t / (g.ntz * g.nty);
  const int ox = (tx << TS) - M, oy = (ty << TS) - M, oz = (tz << TS) - M;
  // first particle of this thread is in flight while the box is being zeroed
  int q = beg + threadIdx.x;
  float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
  if (q < end) p = __ldcs(spos + q);
  for (int i = threadIdx.x; i < NBOX / 2; i += blockDim.x) reinterpret_cast<float2*>(box)[i] = make_float2(0.f, 0.f);
  if (threadIdx.x < 27) scnt[threadIdx.x] = 0;
  __syncthreads();
  // Lanes walk the 8 corners in lane-dependent (XOR-permuted) order: particles sharing a cell then
  // update 8 different addresses at any instant, so the shared-memory CAS loops rarely retry.
  const bool sx = threadIdx.x & 1, sy = threadIdx.x & 2, sz = threadIdx.x & 4;
  while (q < end) {
    const int qn = q + blockDim.x;
    float4 pn = make_float4(0.f, 0.f, 0.f, 0.f);
    if (qn < end) pn = __ldcs(spos + qn);  // prefetch
    Cic1 cx, cy, cz;
    sim_stencil<REL>(g, p.x, p.y, p.z, __float_as_int(p.w), cx, cy, cz);
    Corners c;
    make_corners<B>(g, cx, cy, cz, ox, oy, oz, c);
    if (c.inside) {
      // per-axis swap (once per particle) instead of per-corner selects
      const int lxa = sx ? c.lx[1] : c.lx[0], lxb = sx ? c.lx[0] : c.lx[1];
      const int lya = sy ? c.ly[1] : c.ly[0], lyb = sy ? c.ly[0] : c.ly[1];
      const int lza = sz ? c.lz[1] : c.lz[0], lzb = sz ? c.lz[0] : c.lz[1];
      float wxa = sx ? c.wx[1] : c.wx[0], wxb = sx ? c.wx[0] : c.wx[1];
      float wya = sy ? c.wy[1] : c.wy[0], wyb = sy ? c.wy[0] : c.wy[1];
      float wza = sz ? c.wz[1] : c.wz[0], wzb = sz ? c.wz[0] : c.wz[1];
      if (REL) {  // a dropped corner (index -1, relative rule only) contributes nothing
        if ((sx ? c.ix[1] : c.ix[0]) < 0) wxa = 0.f;
        if ((sx ? c.ix[0] : c.ix[1]) < 0) wxb = 0.f;
        if ((sy ? c.iy[1] : c.iy[0]) < 0) wya = 0.f;
        if ((sy ? c.iy[0] : c.iy[1]) < 0) wyb = 0.f;
        if ((sz ? c.iz[1] : c.iz[0]) < 0) wza = 0.f;
        if ((sz ? c.iz[0] : c.iz[1]) < 0) wzb = 0.f;
      }
      const int lxs[2] = {lxa * (B * BZ), lxb * (B * BZ)}, lys[2] = {lya * BZ, lyb * BZ}, lzs[2] = {lza, lzb};
      const float wxs[2] = {wxa, wxb}, wys[2] = {wya, wyb}, wzs[2] = {wza, wzb};
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
          for (int d = 0; d < 2; ++d)  // reference order (kx*ky)*kz, weight 1
            atomicAdd(box + lxs[a] + lys[b] + lzs[d], (wxs[a] * wys[b]) * wzs[d]);
    } else {
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) {
        const int a = cc & 1, b = (cc >> 1) & 1, d = cc >> 2;
        if (REL && (c.ix[a] < 0 || c.iy[b] < 0 || c.iz[d] < 0)) continue;
        atomicAdd(mesh + ((long long)c.ix[a] * g.ny + c.iy[b]) * g.nz + c.iz[d],
                  (c.wx[a] * c.wy[b]) * c.wz[d]);
      }
      atomicAdd(stats, 1ull);
    }
    // occupancy of the next ordering: shared counters for the 27 surrounding tiles
    const int i0 = max(cx.i0, 0) >> TS, j0 = max(cy.i0, 0) >> TS, k0 = max(cz.i0, 0) >> TS;
    int dx = i0 - tx, dy = j0 - ty, dz = k0 - tz;
    if (dx > 1) dx -= g.ntx; else if (dx < -1) dx += g.ntx;
    if (dy > 1) dy -= g.nty; else if (dy < -1) dy += g.nty;
    if (dz > 1) dz -= g.ntz; else if (dz < -1) dz += g.ntz;
    if (dx >= -1 && dx <= 1 && dy >= -1 && dy <= 1 && dz >= -1 && dz <= 1)
      atomicAdd(scnt + (dx + 1) * 9 + (dy + 1) * 3 + (dz + 1), 1);
    else
      atomicAdd(count + (i0 * g.nty + j0) * g.ntz + k0, 1);
    p = pn;
    q = qn;
  }
  __syncthreads();
  if (threadIdx.x < 27 && scnt[threadIdx.x]) {
    const int dx = threadIdx.x / 9 - 1, dy = (threadIdx.x / 3) % 3 - 1, dz = threadIdx.x % 3 - 1;
    const int i0 = pymod(tx + dx, g.ntx), j0 = pymod(ty + dy, g.nty), k0 = pymod(tz + dz, g.ntz);
    atomicAdd(count + (i0 * g.nty + j0) * g.ntz + k0, scnt[threadIdx.x]);
  }
  // flush: two box rows per warp pass, one z-pair per lane (8-byte vector reductions when the pair
  // cannot straddle the periodic wrap); the z index of a lane is loop invariant.
  constexpr int HP = BZ / 2;                 // pairs per row (<= 16)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int sub = lane / HP, zp = lane - sub * HP;     // sub-row 0/1 (lanes >= 2*HP idle)
  const bool vec = ((oz & 1) == 0) && ((g.nz & 1) == 0) && g.nz >= BZ;
  const int gz0 = wrap_global(oz + 2 * zp, g.nz), gz1 = wrap_global(oz + 2 * zp + 1, g.nz);
  if (sub < 2) {
    for (int r = 2 * warp + sub; r < B * B; r += 2 * nwarp) {
      const float2 v = reinterpret_cast<const float2*>(box)[r * HP + zp];
      if (v.x == 0.f && v.y == 0.f) continue;
      const int lx = r / B, ly = r - lx * B;
      const int gx = wrap_global(ox + lx, g.nx), gy = wrap_global(oy + ly, g.ny);
      float* row = mesh + ((long long)gx * g.ny + gy) * g.nz;
      if (vec) {
        red_add_v2(row + gz0, v.x, v.y);
      } else {
        if (v.x != 0.f) atomicAdd(row + gz0, v.x);
        if (v.y != 0.f && 2 * zp + 1 < B) atomicAdd(row + gz1, v.y);
      }
    }
  }
}

__device__ __forceinline__ void cp_async4(float* smem, const float* gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gmem) : "memory");
}

// ---- read3 + kick + drift + scatter into the next ordering --------------------------------------
template <bool REL, int TS, int M>
__global__ void __launch_bounds__(512)
sim_read_kernel(SimGeom g, const float4* __restrict__ spos, const float* __restrict__ svel,
                const int* __restrict__ start, const float* __restrict__ f0,
                const float* __restrict__ f1, const float* __restrict__ f2, float kick, float drift,
                long long np, int* __restrict__ cursor, float4* __restrict__ npos,
                float* __restrict__ nvel, unsigned long long* __restrict__ stats) {
  constexpr int T = 1 << TS, B = T + 2 * M + 1, NBOX = B * B * B;
  extern __shared__ __align__(16) float box[];
  const int t = blockIdx.x;
  const int beg = start[t], end = start[t + 1];
  if (beg == end) return;
  const int tz = t % g.ntz, ty = (t / g.ntz) % g.nty, tx = t / (g.ntz * g.nty);
  const int ox = (tx << TS) - M, oy = (ty << TS) - M, oz = (tz << TS) - M;
  const float* fm[3] = {f0, f1, f2};
  // stage the three force boxes with cp.async (LDGSTS): one box row (B contiguous floats) per
  // warp pass, no registers held, all rows of a warp in flight at once
  {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    if (lane < B) {
      const int gz = wrap_global(oz + lane, g.nz);
      for (int r = warp; r < B * B; r += nwarp) {
        const int lx = r / B, ly = r - lx * B;
        const int gx = wrap_global(ox + lx, g.nx), gy = wrap_global(oy + ly, g.ny);
        const long long o = ((long long)gx * g.ny + gy) * g.nz + gz;
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
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
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
    int tt = 0;
    float v[3] = {0, 0, 0};
    if (valid) {
      Cic1 cx, cy, cz;
      sim_stencil<REL>(g, p.x, p.y, p.z, __float_as_int(p.w), cx, cy, cz);
      tt = tile_of<TS>(g, cx.i0, cy.i0, cz.i0);
      Corners c;
      make_corners<B>(g, cx, cy, cz, ox, oy, oz, c);
      float acc[3] = {0.f, 0.f, 0.f};
      if (c.inside) {
        float mv[3][8], kk[8];
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) {   // all 24 shared-memory gathers issued back to back
          const int a = cc & 1, b = (cc >> 1) & 1, d = cc >> 2;
          const int o = (c.lx[a] * B + c.ly[b]) * B + c.lz[d];
#pragma unroll
          for (int f = 0; f < 3; ++f) mv[f][cc] = box[f * NBOX + o];
          kk[cc] = (c.wx[a] * c.wy[b]) * c.wz[d];
          if (REL && (c.ix[a] < 0 || c.iy[b] < 0 || c.iz[d] < 0)) kk[cc] = 0.f;
        }
#pragma unroll
        for (int cc = 0; cc < 8; ++cc)
#pragma unroll
          for (int f = 0; f < 3; ++f) acc[f] = fmaf(mv[f][cc], kk[cc], acc[f]);
      } else {
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) {
          const int a = cc & 1, b = (cc >> 1) & 1, d = cc >> 2;
          if (REL && (c.ix[a] < 0 || c.iy[b] < 0 || c.iz[d] < 0)) continue;
          const float k = (c.wx[a] * c.wy[b]) * c.wz[d];
          const long long o = ((long long)c.ix[a] * g.ny + c.iy[b]) * g.nz + c.iz[d];
#pragma unroll
          for (int f = 0; f < 3; ++f) acc[f] = fmaf(__ldg(fm[f] + o), k, acc[f]);
        }
        atomicAdd(stats + 1, 1ull);
      }
#pragma unroll
      for (int f = 0; f < 3; ++f) v[f] = fmaf(kick, acc[f], vin[f]);
      p.x = fmaf(drift, v[0], p.x);
      p.y = fmaf(drift, v[1], p.y);
      p.z = fmaf(drift, v[2], p.z);
    }
    const int slot = warp_claim<true>(cursor, tt, valid);
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
}

template <int TS, int M> constexpr int paint_smem() {
  constexpr int B = (1 << TS) + 2 * M + 1;
  return B * B * ((B + 1) & ~1) * (int)sizeof(float);
}
template <int TS, int M> constexpr int read_smem() {
  constexpr int B = (1 << TS) + 2 * M + 1;
  return 3 * B * B * B * (int)sizeof(float);
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
  return g;
}

}  // namespace jpm

using namespace jpm;

extern "C" int32_t jpm_sim_create(jpm_sim** out, jpm_plan* plan, int32_t nx, int32_t ny, int32_t nz,
                                  int32_t pnx, int32_t pny, int32_t pnz, int32_t hx, int32_t hy,
                                  int32_t relative, int32_t tile, int32_t margin) {
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
  if (plan) {
    int a, b, c;
    plan_dims(plan, &a, &b, &c);
    JPM_CHECK_ARG(a == nx && b == ny && c == nz, "plan shape != sim mesh shape");
  }
  jpm_sim* s = new jpm_sim();
  s->plan = plan;
  s->relative = relative;
  s->np = (long long)pnx * pny * pnz;
  s->g = make_geom(nx, ny, nz, pny, pnz, hx, hy, tile, margin);
  const int ts = s->g.tshift, m = margin;
#define SET_ATTR(TS_, M_)                                                                          \
  {                                                                                                \
    JPM_CUDA(cudaFuncSetAttribute(sim_paint_kernel<false, TS_, M_>,                                \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, paint_smem<TS_, M_>())); \
    JPM_CUDA(cudaFuncSetAttribute(sim_paint_kernel<true, TS_, M_>,                                 \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, paint_smem<TS_, M_>())); \
    JPM_CUDA(cudaFuncSetAttribute(sim_read_kernel<false, TS_, M_>,                                 \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, read_smem<TS_, M_>()));  \
    JPM_CUDA(cudaFuncSetAttribute(sim_read_kernel<true, TS_, M_>,                                  \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, read_smem<TS_, M_>()));  \
  }
  JPM_SIM_DISPATCH(ts, m, SET_ATTR);
#undef SET_ATTR
  for (int i = 0; i < 2; ++i) {
    JPM_CUDA(cudaMalloc(&s->pos[i], s->np * sizeof(float4)));
    JPM_CUDA(cudaMalloc(&s->vel[i], 3 * s->np * sizeof(float)));
    JPM_CUDA(cudaMalloc(&s->start[i], (s->g.nt + 1) * sizeof(int)));
  }
  JPM_CUDA(cudaMalloc(&s->count, s->g.nt * sizeof(int)));
  JPM_CUDA(cudaMalloc(&s->cursor, s->g.nt * sizeof(int)));
  JPM_CUDA(cudaMalloc(&s->stats, 2 * sizeof(unsigned long long)));
  JPM_CUDA(cudaMemset(s->stats, 0, 2 * sizeof(unsigned long long)));
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
  delete s;
  return JPM_OK;
}

extern "C" int32_t jpm_sim_load(jpm_sim* s, void* stream, const float* pos, const float* vel) {
  JPM_CHECK_ARG(s && pos && vel, "null pointer");
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
  JPM_CHECK_ARG(s->loaded, "sim has no particles loaded");
  cudaStream_t st = (cudaStream_t)stream;
  if (s->relative)
    sim_store_kernel<true><<<div_up(s->np, 256), 256, 0, st>>>(s->g, s->pos[s->cur], s->vel[s->cur], s->np, pos, vel);
  else
    sim_store_kernel<false><<<div_up(s->np, 256), 256, 0, st>>>(s->g, s->pos[s->cur], s->vel[s->cur], s->np, pos, vel);
  JPM_LAUNCH_CHECK();
  return JPM_OK;
}

extern "C" int32_t jpm_sim_paint(jpm_sim* s, void* stream, float* mesh) {
  JPM_CHECK_ARG(s && mesh, "null pointer");
  JPM_CHECK_ARG(s->loaded, "sim has no particles loaded");
  cudaStream_t st = (cudaStream_t)stream;
  JPM_CUDA(cudaMemsetAsync(s->count, 0, s->g.nt * sizeof(int), st));
  const int ts = s->g.tshift, m = s->g.m;
#define LAUNCH_PAINT(TS_, M_)                                                                        \
  {                                                                                                  \
    if (s->relative)                                                                                 \
      sim_paint_kernel<true, TS_, M_><<<s->g.nt, 256, paint_smem<TS_, M_>(), st>>>(                  \
          s->g, s->pos[s->cur], s->start[s->cur], mesh, s->count, s->stats);                         \
    else                                                                                             \
      sim_paint_kernel<false, TS_, M_><<<s->g.nt, 256, paint_smem<TS_, M_>(), st>>>(                 \
          s->g, s->pos[s->cur], s->start[s->cur], mesh, s->count, s->stats);                         \
  }
  JPM_SIM_DISPATCH(ts, m, LAUNCH_PAINT);
#undef LAUNCH_PAINT
  JPM_LAUNCH_CHECK();
  s->painted = true;
  return JPM_OK;
}

extern "C" int32_t jpm_sim_read_kick_drift(jpm_sim* s, void* stream, const float* fx, const float* fy,
                                           const float* fz, float kick_coef, float drift_coef) {
  JPM_CHECK_ARG(s && fx && fy && fz, "null pointer");
  JPM_CHECK_ARG(s->painted, "jpm_sim_read_kick_drift must follow jpm_sim_paint (tile occupancy)");
  cudaStream_t st = (cudaStream_t)stream;
  const int nxt = s->cur ^ 1;
  sim_scan_kernel<<<1, 1024, 0, st>>>(s->count, s->start[nxt], s->cursor, s->g.nt);
  JPM_LAUNCH_CHECK();
  const int ts = s->g.tshift, m = s->g.m;
#define LAUNCH_READ(TS_, M_)                                                                         \
  {                                                                                                  \
    if (s->relative)                                                                                 \
      sim_read_kernel<true, TS_, M_><<<s->g.nt, 512, read_smem<TS_, M_>(), st>>>(                    \
          s->g, s->pos[s->cur], s->vel[s->cur], s->start[s->cur], fx, fy, fz, kick_coef, drift_coef, \
          s->np, s->cursor, s->pos[nxt], s->vel[nxt], s->stats);                                     \
    else                                                                                             \
      sim_read_kernel<false, TS_, M_><<<s->g.nt, 512, read_smem<TS_, M_>(), st>>>(                   \
          s->g, s->pos[s->cur], s->vel[s->cur], s->start[s->cur], fx, fy, fz, kick_coef, drift_coef, \
          s->np, s->cursor, s->pos[nxt], s->vel[nxt], s->stats);                                     \
  }
  JPM_SIM_DISPATCH(ts, m, LAUNCH_READ);
#undef LAUNCH_READ
  JPM_LAUNCH_CHECK();
  s->cur = nxt;
  s->painted = false;
  return JPM_OK;
}

extern "C" int32_t jpm_sim_step(jpm_sim* s, void* stream, float kick_coef, float drift_coef) {
  JPM_CHECK_ARG(s && s->plan, "sim has no FFT plan attached");
  cudaStream_t st = (cudaStream_t)stream;
  float* rho = plan_density(s->plan);
  float* f3 = plan_force3(s->plan);
  const long long nc = plan_ncell(s->plan);
  JPM_CUDA(cudaMemsetAsync(rho, 0, nc * sizeof(float), st));
  int32_t rc;
  if ((rc = jpm_sim_paint(s, stream, rho))) return rc;
  if ((rc = jpm_density_to_force_meshes(s->plan, stream, rho, f3, 0.f, nullptr, 0, 0.f))) return rc;
  return jpm_sim_read_kick_drift(s, stream, f3, f3 + nc, f3 + 2 * nc, kick_coef, drift_coef);
}

extern "C" int32_t jpm_sim_stats_host(jpm_sim* s, void* stream, int64_t* out2_host) {
  JPM_CHECK_ARG(s && out2_host, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long h[2];
  JPM_CUDA(cudaMemcpyAsync(h, s->stats, sizeof(h), cudaMemcpyDeviceToHost, st));
  JPM_CUDA(cudaStreamSynchronize(st));
  out2_host[0] = (int64_t)h[0];
  out2_host[1] = (int64_t)h[1];
  return JPM_OK;
}
