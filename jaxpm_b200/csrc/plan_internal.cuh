// Internal layout of the opaque jpm_plan handle (shared by plan.cu and sim.cu).
#pragma once
#include <cuda.h>
#include <cufft.h>

#include "common.cuh"

// Optional per-stage CUDA-event timer of the composed step (jpm_sim_step_profile): every stage boundary
// records an event on the launching stream; read back after a stream synchronise.
struct StageTimer {
  static constexpr int kMax = 24;
  cudaEvent_t ev[kMax];
  const char* name[kMax];
  int n = 0;
  void mark(cudaStream_t st, const char* label) {
    if (n < kMax) {
      cudaEventRecord(ev[n], st);
      name[n++] = label;
    }
  }
};

// Geometry + (peer) pointers of the x-slab decomposition the pmfft kernels work on.  One rank = one GPU of
// the box; rank r owns global x planes [r lx, (r+1) lx).  P == 1 describes the ordinary single-GPU plan.
// Passed to the kernels by value (__grid_constant__).
//
// Pencil process grids (px, py), py > 1 (jaxpm/distributed.py:116-129, rank = a py + b): the PARTICLE domain of
// rank (a, b) - its real arrays - is the pencil x in [a Lx, (a+1) Lx), y in [b Ly, (b+1) Ly) with gx / gy ghost planes
// / rows per side, while the FFT chain keeps the x slabs of the P = px py ranks: slab r = a py + b lies inside pencil
// row a, so the z passes do the row-group transpose while they load / store (the rows of slab r at y in column b'
// live in the arrays of rank (a, b')) and every other pass is unchanged.
struct Slab {
  int P, rank;
  int nx, ny, nz;        // GLOBAL mesh
  int lx, ly;            // nx / P, ny / P (FFT slabs / transposed rows)
  int px, py;            // process grid; slabs: (P, 1)
  int Lx, Ly;            // particle-domain block: nx / px, ny / py (slabs: lx, ny)
  int gx, gy, G;         // ghost planes per side in x, ghost rows per side in y (0 for slabs: the G periodic images
                         // serve), spare / periodic ghost cells in y, z
  int nxp, nyp, nzp;     // local padded real arrays: Lx + 2 gx, Ly + 2 gy + 2 G, nz + 2 G
  long long npad;        // nxp * nyp * nzp
  int nzh, nzc;          // nz / 2 + 1 and its pitch (multiple of 8)
  float* dens[8];        // [nxp][nyp][nzp]       density (painted into, ghosts not folded)
  float* force[8];       // [3][nxp][nyp][nzp]    force meshes, ghosts filled
  float* psi[8];         // [nxp][nyp][nzp]       potential chain: psi = IFFT(G delta / k^2), ghosts filled
  float2* at[8];         // [nx][ly][nzc]         z,y-transformed density, transposed: all x of the rank's y rows
  float2* b3[8];         // [3][lx][ny][nzc]      y-inverse-transformed spectra of the rank's x planes ([2] doubles as
                         //                       the z-transformed density the forward y pass reads)
  float4* t01[8];        // [lx][ny][nzc]         the two x-inverse-transformed spectra (T0, T1) of a mode side by side:
                         //                       16 bytes per mode, 128-byte rows per 8-column tile (NVLink / DRAM friendly)
  unsigned* flags[8];    // [65] barrier slots (one per peer) + error word
};

// word offsets inside a rank's flag block (flags[r], 256 unsigned words)
constexpr int kFlagErr = 64;       // barrier timeout marker
constexpr int kFlagReach = 65;     // set when a rank's particles touched its outermost ghost plane (halo too small)
constexpr int kFlagXmin = 96;      // int: lowest / highest local x plane touched by the last paint (atomicMin / Max by
constexpr int kFlagXmax = 97;      //      sim_paint_kernel; INT_MAX / INT_MIN = unknown)
constexpr int kFlagGe = 98;        // int: ghost planes per side in use this step = max over ranks of what each needs
constexpr int kFlagYmin = 99;      // int: lowest / highest local y row touched by the last paint (pencil grids)
constexpr int kFlagYmax = 100;
constexpr int kFlagGeSlots = 128;  // [P] the ranks' needs, written by the peers in the first barrier of a step
// global force statistics of a step (AUTO force mode, csrc/sim.cu): two slots (step parity) of 4 words each, every
// rank adds its share into EVERY rank's block (system-scope atomics over NVLink), so that all ranks read the same
// numbers and take the same decision: [0..1] u64 fixed point (2^-24) of sum_k |psi_k|^2, [2] bits of max |F|
constexpr int kFlagStats = 160;
constexpr double kStatsFix = 16777216.0;

// one tensor map per destination rank (TMA stores of the transposing FFT passes)
struct alignas(64) TmapPack { CUtensorMap m[8]; };

struct jpm_plan {
  StageTimer* timer = nullptr;
  Slab slab;                  // valid when fft_on
  bool fft_on = false;        // pmfft chain available (power-of-two shape)
  bool is_slab = false;       // created by jpm_slab_create: one rank of a multi-GPU x-slab decomposition
  bool ipc_peers = false;     // peer_base[] opened with cudaIpcOpenMemHandle (else borrowed pointers)
  unsigned epoch = 0;         // barrier epoch (P > 1)
  void* sym_base = nullptr;   // P > 1: the IPC-shared allocation every array of `slab` lives in
  void* peer_base[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  size_t sym_bytes = 0;
  int nx, ny, nz, nzh;
  long long ncell, nspec;
  cufftHandle r2c = 0, c2r1 = 0, c2r3 = 0;
  void* work = nullptr;
  size_t work_bytes = 0;
  // per-axis tables (device): w_d (rad/cell, fp32) and a_d = (8 sin w - sin 2w)/6
  float *wx = nullptr, *wy = nullptr, *wz = nullptr, *ax = nullptr, *ay = nullptr, *az = nullptr;
  // scratch owned by the plan, used by the composed entry points
  float* density = nullptr;   // [ncell]
  float2* spec = nullptr;     // [nspec]
  float2* spec3 = nullptr;    // [3*nspec]
  float* force3 = nullptr;    // [3*ncell]
  // ---- ghost-zone ("padded") meshes for the TMA tile path of csrc/sim.cu ------------------------
  // Every axis carries G ghost cells on both sides: array dims (n+2G)^3, interior starts at index G.
  // Ghost cells are periodic images: paint accumulates into them and jpm::ghost_fold adds them back
  // onto the interior; jpm::ghost_fill copies interior faces out before the read.  With ghosts no
  // tile box ever wraps, so a box is ONE TMA tensor op (cp.async.bulk.tensor / cp.reduce.async.bulk.tensor).
  int G = 0;                  // 0 = padded path not available (nz % 4 != 0, or no driver entry point)
  int nxp = 0, nyp = 0, nzp = 0;
  long long npad = 0;
  float* density_p = nullptr; // [nxp][nyp][nzp]
  float* force3_p = nullptr;  // [3][nxp][nyp][nzp]
  float* psi_p = nullptr;     // [nxp][nyp][nzp]  potential chain + gradient pass (allocated on first use; P == 1)
  cufftHandle r2c_p = 0, c2r3_p = 0;
  // ---- pmfft (csrc/pmfft.cu): hand-written fused FFT chain on the padded meshes, power-of-two shapes --
  float2* fft_at = nullptr;   // P == 1: the AT buffer   [nx][ny][nzc]
  float2* fft_b3 = nullptr;   // P == 1: the B3 buffer   [3][nx][ny][nzc]
  float4* fft_t01 = nullptr;  // P == 1: the T01 buffer  [nx][ny][nzc]
  float2 *tw_x = nullptr, *tw_y = nullptr, *tw_zh = nullptr, *tw_zfull = nullptr;   // exp(-2 pi i k / n) tables
  // TMA-store flavour of the transposing passes: finished tiles leave shared memory as cp.async.bulk.tensor
  // stores (one box per destination rank), so remote (NVLink) stores do not stall the SM
  bool fft_tma_store = false;
  bool fft_pair = false;         // x pass writes (T0, T1) interleaved into T01 (P > 1) instead of planar B[0], B[1]
  int fft_chunk = 0;            // P == 1: x planes per (z-fwd, y-fwd) / (y-inv, z-inv) launch pair, so that the second
                                // pass of a pair finds the first one's output in L2 (0 = whole mesh per launch)
  int fft_zvariant = 1;            // experiments on the z-inverse pass (JPM_FFT_ZVAR bit 0: batched plain epilogue)
  TmapPack* tm_at = nullptr;    // [d]: AT of rank d as {2 nzc, ly, nx} floats, box {32, min(ly,256), 1}
  TmapPack* tm_t01 = nullptr;   // [d]: T01 of rank d as {4 nzc, ny, lx} floats, box {32, 1, min(lx,256)}
  TmapPack* tm_b3 = nullptr;    // [d]: planar B3 of rank d as {2 nzc, ny, lx, 3} floats, box {16, 1, min(lx,256), 1}
  TmapPack* tm_b3w = nullptr;   // same with 16-column boxes {32, 1, min(lx,256), 1}
  // potential chain (pmfft_potential): [0] sum_k |psi_k|^2 of the last evaluation (= mean_x psi^2), [1] bit pattern
  // of max |F| seen by the last read (atomicMax on the float bits), [2..3] spare.  Device doubles.
  double* pot_stats = nullptr;
  bool want_sumsq = false;      // the next pmfft_forces also accumulates pot_stats[0] (auto force mode of csrc/sim.cu)
};

namespace jpm {
constexpr int kGhost = 4;
// Lazily allocate the padded meshes and their cuFFT plans.  Returns JPM_OK and leaves p->G == 0 when
// the shape does not qualify.
int32_t plan_enable_padded(jpm_plan* p);
// density_p (painted, ghosts not yet folded) -> force3_p (ghosts filled), all on `stream`.
int32_t plan_padded_forces(jpm_plan* p, cudaStream_t stream, float r_split, const float* filter_tab,
                           int n_tab, float filter_kmax);
// pmfft: enable (allocates; no-op when the shape is unsupported), run density_p -> force3_p, free.
int32_t pmfft_enable(jpm_plan* p);
int32_t pmfft_setup(jpm_plan* p);
bool pmfft_shape_ok(int nx, int ny, int nz);
// reach_extra: planes beyond the particles' reach the step still needs (2 for the potential chain's stencil)
int32_t slab_barrier(jpm_plan* p, cudaStream_t stream, bool exchange_ghost_width = false, int reach_extra = 0);
// skip_first_barrier: the caller has already run slab_barrier(p, stream, true, ...) for this evaluation
int32_t pmfft_forces(jpm_plan* p, cudaStream_t stream, float r_split, const float* filter_tab, int n_tab,
                     float filter_kmax, bool skip_first_barrier = false);
// P > 1: add this rank's pot_stats of the finished step into slot `slot` of every rank's flag block
int32_t slab_stats_share(jpm_plan* p, cudaStream_t stream, int slot);
// to_psi = false: psi lands in force3_p component 0 (read by sim_readpot_kernel); true: in the separate psi mesh,
// from which pmfft_gradient forms the three force meshes (4th-order differences) in force3_p.
namespace fft { struct KColour; }
// colour != nullptr: the x pass multiplies by the tabulated amplitude of linear_field instead of the Green's function
int32_t pmfft_potential(jpm_plan* p, cudaStream_t stream, float r_split, const float* filter_tab, int n_tab,
                        float filter_kmax, bool to_psi = false, bool skip_first_barrier = false,
                        const fft::KColour* colour = nullptr);
int32_t pmfft_linear_field(jpm_plan* p, cudaStream_t stream, const float* tab, int n_tab, float lkmin, float lkmax,
                           float sx, float sy, float sz, float dc_amp);
int32_t pmfft_gradient(jpm_plan* p, cudaStream_t stream);
void pmfft_destroy(jpm_plan* p);
// cuTensorMapEncodeTiled through the runtime's driver entry point (no libcuda link dependency).
// dims/strides innermost first, rank 3 or 4, fp32, no swizzle/interleave, OOB -> zero fill.
int32_t encode_tensor_map(CUtensorMap* out, float* base, int rank, const unsigned long long* dims,
                          const unsigned long long* strides_bytes, const unsigned* box);
}  // namespace jpm
