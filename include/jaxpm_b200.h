/*
 * jaxpm_b200 — C ABI of the B200-native particle-mesh force loop.
 *
 * This is the drop-in boundary.  The reference (JaxPM) has no
 * FFI of its own: its operator API is the Python functions cited next to each
 * entry point below; an XLA FFI handler (jaxpm_b200/csrc/xla_ffi.cc) or the
 * ctypes host layer (jaxpm_b200/_lib.py) binds exactly these symbols.
 *
 * Conventions
 *   - every function returns 0 on success or a negative jpm_status; the message
 *     is available from jpm_last_error_string() (thread-local);
 *   - `stream` is a cudaStream_t passed as void*; kernels are only enqueued on
 *     it, nothing synchronises or allocates (CUDA-graph / XLA safe), except the
 *     *_create / *_destroy / *_host entry points, which say so;
 *   - all pointers are DEVICE pointers unless the name ends in `_host`;
 *   - the caller owns every buffer; outputs are pre-allocated; in-place
 *     behaviour is documented per function;
 *   - meshes are row-major [nx][ny][nz] float32 (z fastest); particle arrays are
 *     [np][3] float32 in CELL units; spectra are cuFFT R2C half-spectra
 *     [nx][ny][nz/2+1] complex64 (interleaved re,im).
 */
#ifndef JAXPM_B200_H_
#define JAXPM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  JPM_OK = 0,
  JPM_ERR_INVALID = -1, /* bad argument */
  JPM_ERR_CUDA = -2,    /* CUDA runtime error */
  JPM_ERR_CUFFT = -3,   /* cuFFT error */
  JPM_ERR_NOGPU = -4    /* no usable sm_100 device */
} jpm_status;

#define JPM_ABI_VERSION 1

int32_t jpm_abi_version(void);
const char* jpm_last_error_string(void);
/* Name, SM count and compute capability of the current device (host call). */
int32_t jpm_device_info(char* name, int32_t name_len, int32_t* sm_count, int32_t* cc_major,
                        int32_t* cc_minor);

/* ------------------------------------------------------------------------
 * K1  CIC paint (scatter-add)
 * ---------------------------------------------------------------------- */

/* jaxpm/painting.py:15-45 `_cic_paint_impl` (wrapper `cic_paint` :48-75).
 * mesh[(floor(x)+c) mod N] += w * prod_d (1-|x_d-(floor(x_d)+c_d)|), 8 corners.
 * ACCUMULATES into `mesh` (callers pass zeros, pm.py:28-30).  `weight` is a
 * per-particle array [np] or NULL (then `weight_scalar` is used).
 * (pgx,pgy,pgz) is the particle-grid shape used only to form cache-friendly
 * bricks (pgx*pgy*pgz == np; pass (1,1,np) for an unstructured list). */
int32_t jpm_cic_paint_f32(void* stream, float* mesh, const float* positions, const float* weight,
                          float weight_scalar, int64_t np, int32_t nx, int32_t ny, int32_t nz,
                          int32_t pgx, int32_t pgy, int32_t pgz);

/* jaxpm/painting.py:161-189 `_cic_paint_dx_impl` + jaxpm/painting_utils.py:28-96
 * `enmesh` (relative mode).  Particle (i,j,k) of the local [nx][ny][nz] grid sits
 * at (i+hx, j+hy, k) + disp in a mesh of shape [nx+2hx][ny+2hy][nz]; the float
 * mod / floor-div / rint wrap rule of enmesh is reproduced bit-for-bit,
 * including the dropped out-of-range index.  ACCUMULATES into `mesh`. */
int32_t jpm_cic_paint_dx_f32(void* stream, float* mesh, const float* disp, const float* weight,
                             float weight_scalar, int32_t nx, int32_t ny, int32_t nz, int32_t hx,
                             int32_t hy);

/* Debug/parity helper: the int32 flat cell index (ix*ny+iy)*nz+iz of the
 * (0,0,0) corner of every particle, -1 where the reference drops it.
 * mode 0 = absolute rule (painting.py:35-37), 1 = relative rule (painting_utils.py:53-65)
 * with mesh shape [nx][ny][nz], halo offsets (hx,hy) and particle grid (nx-2hx,ny-2hy,nz). */
int32_t jpm_cic_cell_index_i32(void* stream, int32_t* out, const float* pos_or_disp, int64_t np,
                               int32_t nx, int32_t ny, int32_t nz, int32_t hx, int32_t hy,
                               int32_t mode);

/* ------------------------------------------------------------------------
 * K5  CIC read (gather) and the fused read3 + kick + drift
 * ---------------------------------------------------------------------- */

/* jaxpm/painting.py:78-106 `_cic_read_impl` (wrapper `cic_read` :109-128). out[np]. */
int32_t jpm_cic_read_f32(void* stream, float* out, const float* mesh, const float* positions,
                         int64_t np, int32_t nx, int32_t ny, int32_t nz);

/* jaxpm/painting.py:218-236 `_cic_read_dx_impl`.  `mesh` is the padded, halo-filled
 * local mesh [nx+2hx][ny+2hy][nz]; out[nx][ny][nz]. */
int32_t jpm_cic_read_dx_f32(void* stream, float* out, const float* mesh, const float* disp,
                            int32_t nx, int32_t ny, int32_t nz, int32_t hx, int32_t hy);

/* float64 forms of the four CIC primitives above (what the reference computes under jax_enable_x64, the mode of
 * its distributed tests, tests/test_distributed_pm.py:30): positions / displacements, weights and meshes are double and
 * the index / weight rules run in double.  Same argument meaning as the _f32 entries (paint accumulates into `mesh`;
 * weight: per-particle array or NULL + scalar; the _dx forms take the padded local mesh [nx+2hx][ny+2hy][nz]). */
int32_t jpm_cic_paint_f64(void* stream, double* mesh, const double* positions, const double* weight,
                          double weight_scalar, int64_t np, int32_t nx, int32_t ny, int32_t nz);
int32_t jpm_cic_paint_dx_f64(void* stream, double* mesh, const double* disp, const double* weight,
                             double weight_scalar, int32_t nx, int32_t ny, int32_t nz, int32_t hx, int32_t hy);
int32_t jpm_cic_read_f64(void* stream, double* out, const double* mesh, const double* positions, int64_t np,
                         int32_t nx, int32_t ny, int32_t nz);
int32_t jpm_cic_read_dx_f64(void* stream, double* out, const double* mesh, const double* disp, int32_t nx,
                            int32_t ny, int32_t nz, int32_t hx, int32_t hy);

/* The three reads + jnp.stack of jaxpm/pm.py:54-56 in one pass: out[np][3] =
 * scale * (read(fx), read(fy), read(fz)).  relative != 0 selects the
 * painting_utils rule with halo offsets (then np = nx*ny*nz of the particle grid
 * and the meshes are [nx+2hx][ny+2hy][nz]). */
int32_t jpm_cic_read3_f32(void* stream, float* out, const float* fx, const float* fy,
                          const float* fz, const float* pos_or_disp, float scale, int64_t np,
                          int32_t nx, int32_t ny, int32_t nz, int32_t hx, int32_t hy,
                          int32_t relative);

/* read3 fused with the ODE update (jaxpm/ode.py:100-117 kick, :91-98 drift; the
 * FastPM variants :19-58 only change the two scalars):
 *     F       = (read(fx), read(fy), read(fz)) at pos_in
 *     vel_out = vel_prev + kick_coef  * F          (kick_coef = dt*1.5*Om/(a^2 E) ...)
 *     pos_out = pos_prev + drift_coef * vel_drift  (vel_drift = vel_out if use_new_vel
 *                                                   else vel_in)
 * Kick-drift (symplectic): pos_prev=pos_in, vel_prev=vel_in, use_new_vel=1, all in place.
 * Leapfrog-midpoint (diffrax): pos_prev/vel_prev = state n-1, use_new_vel=0, vel_in = v_n.
 * Output pointers may alias the *_prev pointers.  forces_out (nullable) receives F. */
int32_t jpm_cic_read3_kick_drift_f32(void* stream, float* pos_out, float* vel_out,
                                     float* forces_out, const float* fx, const float* fy,
                                     const float* fz, const float* pos_in, const float* vel_in,
                                     const float* pos_prev, const float* vel_prev,
                                     float kick_coef, float drift_coef, int32_t use_new_vel,
                                     int64_t np, int32_t nx, int32_t ny, int32_t nz, int32_t hx,
                                     int32_t hy, int32_t relative);

/* ------------------------------------------------------------------------
 * K6  adjoints (what jax.grad of paint/read transposes to)
 * ---------------------------------------------------------------------- */

/* value[np] = read(mesh, pos) and grad[np][3] = s_p * d value / d pos with
 * d(1-|t|)/dt = -sign(t), sign(0)=0 (JAX's abs' convention).  Either output may be NULL.
 * s_p = grad_scale_scalar * (grad_scale ? grad_scale[p] : 1) folds the cotangent / weight in.
 * Used for: read VJP wrt positions (cot * grad), paint VJP wrt positions
 * (weight * grad of read(cotangent mesh)), paint VJP wrt weights (value).
 * relative as in jpm_cic_read3_f32. */
int32_t jpm_cic_readgrad_f32(void* stream, float* value, float* grad, const float* mesh,
                             const float* pos_or_disp, const float* grad_scale,
                             float grad_scale_scalar, int64_t np, int32_t nx, int32_t ny,
                             int32_t nz, int32_t hx, int32_t hy, int32_t relative);

/* Forward mode (what jax.jvp / jacfwd of the paint produces, tests/test_distributed_pm.py:313-320): the tangent of
 * the painted mesh for a position tangent `tangent[np][3]`,
 *     mesh[c] += w_p * sum_d tangent[p][d] * d K(x_p - c) / d x_d      (transpose of jpm_cic_readgrad_f32).
 * ACCUMULATES into `mesh`.  The JVP of a read is jpm_cic_readgrad_f32's gradient dotted with the tangent plus a plain
 * read of the mesh tangent, so paint / read / pm_forces have both modes from the same four kernels. */
int32_t jpm_cic_paintgrad_f32(void* stream, float* mesh, const float* pos_or_disp, const float* tangent,
                              const float* weight, float weight_scalar, int64_t np, int32_t nx, int32_t ny,
                              int32_t nz, int32_t hx, int32_t hy, int32_t relative);

/* The reverse-mode building blocks of pm_forces (jaxpm/pm.py:12-58) fused per pass over the particles - what
 * jax.grad of the force loop transposes to (tests/test_gradients.py:30-80), three launches + three axpys each before:
 *   jpm_cic_readgrad3_f32: grad[p] (+)= scale * sum_d u[p][d] * d read(m_d)(x_p) / d x_p   (m1 = m2 = NULL, cotangent =
 *                          NULL: one mesh, grad[p] (+)= scale * d read(m0)(x_p) / d x_p - the paint adjoint);
 *   jpm_cic_paint3_f32:    mesh3[d] += paint(x; weight = scale * u[:, d]) for d = 0..2 (mesh3: 3 contiguous meshes);
 *   jpm_fd_divergence3_f32: out = sum_d D_d g3[d], D = the 4th-order central difference the reference's gradient
 *                          kernel is the symbol of, so that sum_d L_d^T G_d = -jpm_density_to_potential_fused(out):
 *                          one transform pair on the fused chain instead of three forward transforms + a k-space pass. */
int32_t jpm_cic_readgrad3_f32(void* stream, float* grad, const float* m0, const float* m1, const float* m2,
                              const float* pos_or_disp, const float* cotangent, float scale, int64_t np, int32_t nx,
                              int32_t ny, int32_t nz, int32_t hx, int32_t hy, int32_t relative, int32_t accumulate);
int32_t jpm_cic_paint3_f32(void* stream, float* mesh3, const float* pos_or_disp, const float* weights3, float scale,
                           int64_t np, int32_t nx, int32_t ny, int32_t nz, int32_t hx, int32_t hy, int32_t relative);
int32_t jpm_fd_divergence3_f32(void* stream, float* out, const float* g3, int32_t nx, int32_t ny, int32_t nz);

/* 2-D CIC paint of projected particles (light-cone density planes): mesh[nx][ny] += paint(pos2[np][2]) * weight[np]
 * (weight may be NULL = 1).  Replaces jaxpm/painting.py:131-158 (cic_paint_2d), same index / weight rule. */
int32_t jpm_cic_paint_2d_f32(void* stream, float* mesh, const float* pos2, const float* weight, int64_t np,
                             int32_t nx, int32_t ny);

/* Un-normalised density plane of a light cone in one pass over pos3[np][3] (cell units): particles with
 * center - width/2 < z <= center + width/2 are painted (2-D CIC) at mod(xy, box_nx) / box_nx * plane_resolution onto
 * plane[res][res] (accumulated into).  Replaces the particle part of jaxpm/lensing.py:11-44 (density_plane :20-35). */
int32_t jpm_density_plane_f32(void* stream, float* plane, const float* pos3, int64_t np, float box_nx, double center,
                              double width, int32_t plane_resolution);

/* ------------------------------------------------------------------------
 * K2/K3/K4  FFT plan and the fused k-space pass
 * ---------------------------------------------------------------------- */

typedef struct jpm_plan jpm_plan; /* opaque; one stream at a time per plan */

/* Host call; allocates cuFFT plans + work areas + the per-axis tables of
 * jaxpm/kernels.py:10-23 `fftk` (w_d = 2*pi*fftfreq(N_d)), :62-66
 * `gradient_kernel` order 1 (a_d = (8 sin w - sin 2w)/6) built in float64 on the host. */
int32_t jpm_plan_create(jpm_plan** plan, int32_t nx, int32_t ny, int32_t nz);
int32_t jpm_plan_destroy(jpm_plan* plan);

/* jaxpm/distributed.py:37-38 `fft3d` on a real field: unnormalised forward R2C.
 * in: real [nx][ny][nz]; out: complex64 [nx][ny][nz/2+1]. */
int32_t jpm_fft3d_r2c(jpm_plan* plan, void* stream, const float* in, void* out);
/* jaxpm/distributed.py:41-42 `ifft3d` (`.real` of the inverse) for `batch` Hermitian
 * half-spectra stored contiguously; UNNORMALISED (callers fold 1/Nc into the
 * k-space pass). The input spectra are destroyed (cuFFT C2R). batch in {1,3}. */
int32_t jpm_ifft3d_c2r(jpm_plan* plan, void* stream, void* in, float* out, int32_t batch);

/* Fused k-space pass of jaxpm/pm.py:49-56: for d in 0..2
 *     out_d = -(i a_d) * ( -1/k^2 [k=0 -> 0] ) * exp(-k^2 r_split^2) * filter(|k|) * delta_k * norm
 * i.e. jaxpm/kernels.py:69-92 `invlaplace_kernel`, :95-115 `longrange_kernel`, :41-66
 * `gradient_kernel`, and the optional radial filter slot of jaxpm/ode.py:194-196 /
 * jaxpm/kernels.py:139-165 (table `filter_tab[n_tab]`, linear in |k| on [0, filter_kmax];
 * NULL = none).  out: 3 contiguous half-spectra.  norm is typically 1/(nx*ny*nz). */
int32_t jpm_greens_grad_c64(jpm_plan* plan, void* stream, const void* delta_k, void* out3,
                            float norm, float r_split, const float* filter_tab, int32_t n_tab,
                            float filter_kmax);

/* Transpose (VJP) of jpm_greens_grad_c64: out = sum_d (-i a_d)(1/k^2) G filter * in_d * norm, the
 * k-space half of the adjoint of jaxpm/pm.py:49-56 (in3: 3 contiguous half-spectra). */
int32_t jpm_greens_div_c64(jpm_plan* plan, void* stream, const void* in3, void* out, float norm,
                           float r_split, const float* filter_tab, int32_t n_tab,
                           float filter_kmax);

/* 2LPT shear spectra of jaxpm/pm.py:95-109: out6 = (-a_i a_j)(-1/k^2) delta_k * norm for
 * (i,j) in (00,11,22,01,02,12) — `gradient_kernel(i)*gradient_kernel(j)*pot_k`. */
int32_t jpm_lpt2_shear_c64(jpm_plan* plan, void* stream, const void* delta_k, void* out6,
                           float norm);
/* delta2 = s00*s11 + s22*(s00+s11) - s01^2 - s02^2 - s12^2 (pm.py:92-109 accumulated form). */
int32_t jpm_lpt2_source_f32(void* stream, float* delta2, const float* shear6, int64_t ncell);
/* Reverse mode of the 2LPT source (what jax.grad of jaxpm/pm.py:88-111 transposes to):
 * t6[q] = cot * d delta2 / d s_q for the 6 shear meshes (00,11,22,01,02,12) ... */
int32_t jpm_lpt2_source_adj_f32(void* stream, float* t6, const float* shear6, const float* cot, int64_t ncell);
/* ... and the transpose of jpm_lpt2_shear_c64 (its multipliers a_i a_j / k^2 are real and even, so the operator is
 * self-adjoint): out = sum_q (a_i a_j)(1/k^2) in6[q] * norm, in6 = the R2C spectra of t6. */
int32_t jpm_lpt2_shear_adj_c64(jpm_plan* plan, void* stream, const void* in6, void* out, float norm);

/* Generic k-space multiply used by linear_field (pm.py:134-143): out = in * tab(|k| scaled)
 * where kphys^2 = sum_d (w_d * kscale_d)^2 and the table is linear in log10(kphys) on
 * [log10_kmin, log10_kmax] (values are sqrt(P(k) * Nc / V)); k=0 -> tab value at kmin. */
int32_t jpm_kfilter_logtab_c64(jpm_plan* plan, void* stream, const void* in, void* out,
                               const float* tab, int32_t n_tab, float log10_kmin,
                               float log10_kmax, float kscale_x, float kscale_y, float kscale_z,
                               float norm);

/* ------------------------------------------------------------------------
 * composed hot path
 * ---------------------------------------------------------------------- */

/* jaxpm/pm.py:12-58 `pm_forces` given a painted density: real `density` [nx][ny][nz] ->
 * three force meshes (fx,fy,fz contiguous, each [nx][ny][nz]).  Uses plan scratch. */
int32_t jpm_density_to_force_meshes(jpm_plan* plan, void* stream, const float* density,
                                    float* force3, float r_split, const float* filter_tab,
                                    int32_t n_tab, float filter_kmax);

/* Same result as jpm_density_to_force_meshes, computed on the plan's ghost-zone meshes by the fused
 * five-pass FFT chain of csrc/pmfft.cu (power-of-two shapes: R2C along z, FFT along y, FFT along x with
 * the Green's function x gradient and the three inverse x transforms in the same kernel, inverse y, C2R
 * along z) or, for other shapes, by cuFFT on the padded arrays.  Replaces jaxpm/pm.py:41-56. */
int32_t jpm_density_to_force_meshes_fused(jpm_plan* plan, void* stream, const float* density,
                                          float* force3, float r_split, const float* filter_tab,
                                          int32_t n_tab, float filter_kmax);

/* The POTENTIAL chain (power-of-two meshes): real `density` -> psi = IFFT(G(k) delta_k / k^2) = -phi, ONE inverse
 * transform.  The reference's force spectra -gradient_kernel(k, d) * pot_k (jaxpm/pm.py:51-56 with
 * jaxpm/kernels.py:62-66) are EXACTLY the 4th-order central differences of this mesh,
 *     F_d(c) = [8 (psi(c + e_d) - psi(c - e_d)) - (psi(c + 2 e_d) - psi(c - 2 e_d))] / 12,
 * which the resident step forms in shared memory while it stages a tile (jpm_sim_set_force_mode). */
int32_t jpm_density_to_potential_fused(jpm_plan* plan, void* stream, const float* density, float* psi,
                                       float r_split, const float* filter_tab, int32_t n_tab,
                                       float filter_kmax);

/* One full PM step on resident particles (single GPU, absolute or relative positions):
 * memset mesh -> paint -> R2C -> greens-grad -> 3x C2R -> read3+kick+drift (kick-drift
 * form, in place on pos/vel).  jaxpm/ode.py:100-117 + :91-98 around jaxpm/pm.py:12-58. */
int32_t jpm_pm_step_f32(jpm_plan* plan, void* stream, float* pos, float* vel, float kick_coef,
                        float drift_coef, int32_t relative);

/* Same step through HOST buffers (pinned or pageable): H2D pos/vel, step, D2H pos/vel.
 * Synchronises `stream`.  This is the end-to-end entry a host-side caller binds. */
int32_t jpm_pm_step_host_f32(jpm_plan* plan, void* stream, float* pos_host, float* vel_host,
                             float* pos_dev, float* vel_dev, float kick_coef, float drift_coef,
                             int32_t relative);

/* ------------------------------------------------------------------------
 * power-spectrum estimator on the R2C half-spectrum and its adjoint
 *   replaces jaxpm/utils.py:14-73 (_initialize_pk) + :76-128 (power_spectrum)
 * ---------------------------------------------------------------------- */
/* spec_a (and spec_b for a cross spectrum, else NULL): complex64 half-spectra [nx][ny][nz/2+1] of UNNORMALISED
 * forward transforms; norm = 1/(nx ny nz) gives the reference's 'ortho' convention.  kx[nx], ky[ny], kz[nz/2+1]:
 * DEVICE float64 physical wavenumbers per axis ((2 pi m / l) fftfreq(m), utils.py:51-53); kedges[n_edges]: DEVICE
 * float64 bin edges.  ells[n_ell] (HOST, each 0, 2 or 4) and los3 (HOST unit vector; may be NULL when all ells
 * are 0).  out (DEVICE float64, zeroed here): [n_ell][nb] real sums | [n_ell][nb] imaginary sums |
 * [nb] mode counts | [nb] sums of |k| (the last two only if want_counts), nb = n_edges + 1 bins of np.digitize. */
int32_t jpm_pk_bin_c64(void* stream, const void* spec_a, const void* spec_b, int32_t nx, int32_t ny, int32_t nz,
                       const double* kx, const double* ky, const double* kz, const double* kedges, int32_t n_edges,
                       const int32_t* ells, int32_t n_ell, const float* los3, float norm, int32_t want_counts,
                       double* out);
/* Adjoint of the auto spectrum: out_k = spec_a_k * 2 norm sum_l wbin[l][bin(k)] (2l+1) L_l(mu_k); the unnormalised
 * C2R of `out` is d(sum_lb g_lb pk_l[b]) / d mesh for wbin[l][b] = g_lb * cell volume / kcount[b] (DEVICE float64). */
int32_t jpm_pk_weight_c64(void* stream, const void* spec_a, void* out, int32_t nx, int32_t ny, int32_t nz,
                          const double* kx, const double* ky, const double* kz, const double* kedges, int32_t n_edges,
                          const int32_t* ells, int32_t n_ell, const float* los3, const double* wbin, float norm);

/* Half-spectrum times a separable real filter: out_k = in_k * norm * tx[ix] ty[iy] tz[iz] (tables are DEVICE
 * float32 of lengths nx, ny, nz/2+1; in place allowed).  compensate_cic of jaxpm/painting.py:263-275 is
 * R2C -> this with t_d = sinc(k_d / 2 pi)^-2 (kernels.py:118-136) -> C2R. */
int32_t jpm_kseparable_c64(void* stream, const void* in, void* out, const float* tx, const float* ty, const float* tz,
                           int32_t nx, int32_t ny, int32_t nz, float norm);

/* Test / debug access to the ghost-zone meshes the resident step (jpm_sim_step) works on: which = 0 the
 * painted density (ghosts not folded), 1..3 a force component (ghosts filled).  dims3 (nullable) receives the
 * padded extents; dst (nullable) the whole padded array [dims3[0]][dims3[1]][dims3[2]]. */
int32_t jpm_plan_padded_get_f32(jpm_plan* plan, void* stream, int32_t which, float* dst, int32_t* dims3);

/* ------------------------------------------------------------------------
 * multi-GPU x-slab plan: the fused FFT chain + halo protocol over NVLink peer memory
 *   replaces, for pdims = (P, 1): [ext] jaxdecomp.pfft3d / pifft3d (jaxpm/distributed.py:37-42) and
 *   halo_exchange + slice_unpad (jaxpm/distributed.py:45-85, painting.py:192-215, :239-260)
 * ---------------------------------------------------------------------- */
/* Host call, one per rank (= GPU of the box, <= 8).  Global mesh [nx][ny][nz] (powers of two); rank r owns
 * x planes [r nx/P, (r+1) nx/P) plus gx ghost planes per side (the reference's halo; gx <= nx/P).
 * Allocates ONE device block holding the rank's density / force / spectrum buffers and barrier flags. */
int32_t jpm_slab_create(jpm_plan** plan, int32_t nx, int32_t ny, int32_t nz, int32_t nranks, int32_t rank,
                        int32_t gx);
/* Pencil process grids (px, py), px py <= 8 ranks, rank = a py + b as in jax.make_mesh(pdims) (jaxpm/distributed.py:
 * 116-129, 168-190; tests/test_distributed_pm.py:28 pdims (4,2), (2,4), (1,8)): rank (a, b) owns the particles and the
 * real meshes of x in [a nx/px, (a+1) nx/px), y in [b ny/py, (b+1) ny/py) plus gx ghost planes / gy ghost rows per
 * side (the reference's halo per sharded axis).  The FFT chain keeps the x slabs of the px py ranks: the z passes do
 * the row-group transpose while they load / store (pull the density rows of the slab from the pencils of the row
 * group with their x / y / corner ghosts folded in, push the force rows back with the ghost images), so the exchange
 * volume equals that of a two-transpose pencil FFT and no NCCL call is on the data path.  py == 1: jpm_slab_create.
 * Needs ny / py % 16 == 0.  Every jpm_slab_* entry below serves both; "interior" is the rank's [nx/px][ny/py][nz]. */
int32_t jpm_slab_create_ex(jpm_plan** plan, int32_t nx, int32_t ny, int32_t nz, int32_t px, int32_t py, int32_t rank,
                           int32_t gx, int32_t gy);
/* cudaIpcMemHandle_t (64 bytes) of the rank's block: gather them over the ranks (any transport), then */
int32_t jpm_slab_ipc_handle(jpm_plan* plan, void* handle_out, int32_t handle_bytes);
/* map every peer's block: handles[nranks][64] in rank order (other processes, cudaIpcOpenMemHandle) ... */
int32_t jpm_slab_attach_ipc(jpm_plan* plan, const void* handles, int32_t nranks);
/* ... or bases[nranks] = device pointers valid in THIS process (ranks sharing a process / a device). */
int32_t jpm_slab_attach_ptrs(jpm_plan* plan, void* const* bases, int32_t nranks);
int32_t jpm_slab_base(jpm_plan* plan, void** base_out, int64_t* bytes_out);
/* interior [nx/P][ny][nz] of this rank: which = 0 density, 1..3 force component -> dst (compact) */
int32_t jpm_slab_get_interior_f32(jpm_plan* plan, void* stream, int32_t which, float* dst);
/* compact local density block -> the rank's density mesh (ghosts zero) */
int32_t jpm_slab_set_density_f32(jpm_plan* plan, void* stream, const float* src);
/* COLLECTIVE (every rank must call, each on its own stream): density with unfolded ghosts -> the three
 * force meshes with ghosts filled (jaxpm/pm.py:41-56).  Kernels only enqueue; ranks meet in four
 * in-stream flag barriers over peer memory. */
int32_t jpm_slab_forces(jpm_plan* plan, void* stream, float r_split);
/* Synchronises `stream`; error if one of this rank's flag barriers timed out (peer lost). */
int32_t jpm_slab_check(jpm_plan* plan, void* stream);
/* Ghost planes per side the last force evaluation exchanged: gx, or - when the density was painted by a
 * jpm_sim bound to this plan - the maximum over the ranks of what their particles actually reach. */
int32_t jpm_slab_ghost_width(jpm_plan* plan, void* stream, int32_t* out);
/* *out = 1 if some step since creation saw particles of this rank on its outermost ghost plane: gx (the
 * reference's halo_size) is too small for the displacement field; like in the reference those particles are then
 * painted / read at wrapped positions, but here the condition is detectable. */
int32_t jpm_slab_halo_exceeded(jpm_plan* plan, void* stream, int32_t* out);

/* ------------------------------------------------------------------------
 * tile-sorted resident particle state (the fast path for many steps)
 * ---------------------------------------------------------------------- */
typedef struct jpm_sim jpm_sim; /* opaque; owns a tile-sorted copy of (pos, vel) */

/* Host call.  Mesh [nx][ny][nz]; particles form a [pnx][pny][pnz] grid (ids in that order).
 * relative != 0: state holds displacements and nx = pnx+2hx, ny = pny+2hy, nz = pnz
 * (jaxpm/painting.py:161-189); else absolute positions (jaxpm/painting.py:15-45).
 * tile in {8,16,32} cells, margin = cells a particle may move per step before the slow
 * (global-memory) fallback is used.  `plan` (nullable) enables jpm_sim_step. */
int32_t jpm_sim_create(jpm_sim** sim, jpm_plan* plan, int32_t nx, int32_t ny, int32_t nz,
                       int32_t pnx, int32_t pny, int32_t pnz, int32_t hx, int32_t hy,
                       int32_t relative, int32_t tile, int32_t margin);
/* Same, with flags: JPM_SIM_POSITIONS_ONLY keeps no velocities and no second ordering (16 instead of 56 bytes per
 * particle): the state can be loaded (vel = NULL), painted and asked for forces, not stepped. */
#define JPM_SIM_POSITIONS_ONLY 1
int32_t jpm_sim_create_ex(jpm_sim** sim, jpm_plan* plan, int32_t nx, int32_t ny, int32_t nz,
                          int32_t pnx, int32_t pny, int32_t pnz, int32_t hx, int32_t hy,
                          int32_t relative, int32_t tile, int32_t margin, int32_t flags);
int32_t jpm_sim_destroy(jpm_sim* sim);
/* Build the sorted state from user-order arrays pos[np][3], vel[np][3] / write it back. */
int32_t jpm_sim_load(jpm_sim* sim, void* stream, const float* pos, const float* vel);
int32_t jpm_sim_store(jpm_sim* sim, void* stream, float* pos, float* vel);
/* mesh += paint(state) (same arithmetic as jpm_cic_paint[_dx]_f32, weight 1); also records the
 * tile occupancy that the following jpm_sim_read_kick_drift uses to re-sort. */
int32_t jpm_sim_paint(jpm_sim* sim, void* stream, float* mesh);
/* vel += kick*F(pos); pos += drift*vel for every particle (jaxpm/ode.py:91-117), F gathered from
 * the three force meshes through shared-memory boxes; writes the next tile ordering. */
int32_t jpm_sim_read_kick_drift(jpm_sim* sim, void* stream, const float* fx, const float* fy,
                                const float* fz, float kick_coef, float drift_coef);
/* jaxpm/pm.py:12-58 `pm_forces` on the tile kernels: paint the loaded state (shared-memory boxes, TMA reduce-add),
 * fused FFT chain with the k-space kernels of pm.py:49-56 (r_split, optional radial filter table as in
 * jpm_greens_grad_c64), then the three reads + stack of pm.py:54-56: out[np][3] = scale * F, in the CALLER's
 * particle order (the order of the arrays given to jpm_sim_load). */
int32_t jpm_sim_forces(jpm_sim* sim, void* stream, float* out, float scale, float r_split,
                       const float* filter_tab, int32_t n_tab, float filter_kmax);
/* Batched form (what jax.vmap over a leading axis of the positions lowers to, tests/test_distributed_pm.py:335-409):
 * positions[nbatch][np][3] -> out[nbatch][np][3], element b == jpm_sim_load(pos_b) + jpm_sim_forces.  The batch
 * shares the sim's meshes and plans (sequential on the stream; a batch element is a full-GPU workload). */
int32_t jpm_sim_forces_batched(jpm_sim* sim, void* stream, const float* positions, float* out, int32_t nbatch,
                               float scale, float r_split, const float* filter_tab, int32_t n_tab,
                               float filter_kmax);
/* One PM step on the resident state: memset, paint, R2C, greens-grad, 3x C2R, read+kick+drift. */
int32_t jpm_sim_step(jpm_sim* sim, void* stream, float kick_coef, float drift_coef);
/* One step end to end through HOST buffers (pinned or pageable) on the tile kernels: H2D pos / vel [np][3], tile sort,
 * jpm_sim_step, un-sort, D2H.  pos_dev / vel_dev: device staging [np][3].  Synchronises `stream`. */
int32_t jpm_sim_step_host_f32(jpm_sim* sim, void* stream, float* pos_host, float* vel_host, float* pos_dev,
                              float* vel_dev, float kick_coef, float drift_coef);
/* A batch of independent particle states, each taken through the sequence of jpm_sim_step_host_f32 (what jax.vmap of a
 * one-step function over host-resident states, or a caller streaming states from host memory, issues).  The legs of
 * consecutive elements overlap: element b + 1 uploads on a copy stream while element b computes on `stream` and
 * element b - 1 downloads on a second copy stream, through double-buffered device staging owned by the sim (4 x
 * np x 24 bytes, allocated at the first call).  pos_hosts / vel_hosts: nbatch host pointers ([np][3] each, updated in
 * place; pinned memory for overlap; a pointer may repeat - an upload waits for the last download into the same
 * buffer).  kick_coefs / drift_coefs: nbatch HOST floats.  Synchronises `stream` and both copy streams. */
int32_t jpm_sim_steps_host_f32(jpm_sim* sim, void* stream, int32_t nbatch, float* const* pos_hosts,
                               float* const* vel_hosts, const float* kick_coefs, const float* drift_coefs);
/* One jpm_sim_step with a CUDA event recorded on `stream` at every stage boundary (memset, paint, each
 * FFT pass, read).  Synchronises the stream; fills names_out[i] (static strings) and ms_out[i] for the
 * *n_out <= cap stages.  This is how bench.py measures the per-kernel roofline live. */
int32_t jpm_sim_step_profile(jpm_sim* sim, void* stream, float kick_coef, float drift_coef,
                             const char** names_out, float* ms_out, int32_t cap, int32_t* n_out);

/* Force path of jpm_sim_step (resident state on a power-of-two mesh, margin 1):
 *   JPM_FORCE_SPECTRAL  three inverse transforms of i a_d(k) delta_k / k^2 (jaxpm/pm.py:54-56 literally);
 *   JPM_FORCE_POTENTIAL one inverse transform of delta_k / k^2, forces by the 4th-order difference stencil
 *                       the reference's gradient kernel is the symbol of (same operator; fp32 differencing
 *                       adds an absolute error ~ 5e-7 max|psi|, negligible once the field is clustered);
 *   JPM_FORCE_AUTO      per step, from the device-measured bound 2.7e-6 rms(psi) / max|F|: potential when it is
 *                       below 4e-6, spectral when above 6e-6 - every step stays within 1e-5 of the reference. */
#define JPM_FORCE_SPECTRAL 0
#define JPM_FORCE_POTENTIAL 1
#define JPM_FORCE_AUTO 2
int32_t jpm_sim_set_force_mode(jpm_sim* sim, int32_t mode);
/* out6_host = {mode, mode of the next step, last evaluated error bound (< 0: none), steps run spectral, steps run
 * potential, 1 if the potential path is available}.  Synchronises the stream. */
int32_t jpm_sim_force_info(jpm_sim* sim, void* stream, double* out6_host);

/* out4_host[0..1] = particles that took the global-memory fallback (drifted beyond the margin) in
 * paint / read so far; [2..3] = particles that took the generic (periodic-wrap) stencil inside the
 * shared-memory box in paint / read.  Synchronises the stream. */
int32_t jpm_sim_stats_host(jpm_sim* sim, void* stream, int64_t* out4_host);

/* Number of kernels of THIS library launched since process start (for bench accounting). */
int64_t jpm_kernel_launch_count(void);

/* ------------------------------------------------------------------------
 * Gaussian initial conditions on the device
 *   replaces jaxpm/distributed.py:193-223 `normal_field` and jaxpm/pm.py:129-144 `linear_field`
 * ---------------------------------------------------------------------- */
/* N(0,1) white noise, local block [lx][ly][nz] at offset (ox, oy) of a global mesh with global_ny rows:
 * Philox4x32-10 keyed by `seed`, counter = global index of the 4-cell group, Box-Muller.  The value of a cell
 * depends only on (seed, stream_id, global cell index): a sharded call draws the single-device field (the reference
 * draws one independent stream per device).  Not JAX's threefry stream. */
int32_t jpm_normal_field_f32(void* stream, float* out, int32_t lx, int32_t ly, int32_t nz, int32_t ox, int32_t oy,
                             int32_t global_ny, uint64_t seed, uint32_t stream_id);
/* pm.py:134-143 on the fused FFT chain (power-of-two meshes): out = IFFT( FFT(white) * amp(|k_phys|) ) / Nc with
 * k_phys^2 = sum_d (w_d kscale_d)^2 and amp tabulated linearly in log10 k on [log10_kmin, log10_kmax]
 * (values sqrt(P(k) Nc / V)); the k = 0 mode is multiplied by dc_amp (sqrt(P(0) Nc / V), 0 for a power law).
 * Three forward passes, the table multiply inside the x pass, three inverse passes: 40 B/cell of HBM traffic. */
int32_t jpm_linear_field_f32(jpm_plan* plan, void* stream, const float* white, float* out, const float* tab,
                             int32_t n_tab, float log10_kmin, float log10_kmax, float kscale_x, float kscale_y,
                             float kscale_z, float dc_amp);

/* ------------------------------------------------------------------------
 * small elementwise helpers (keep host-side glue off third-party ops)
 * ---------------------------------------------------------------------- */
/* out = a*x + b*y (y may be NULL -> a*x); out may alias x or y. */
int32_t jpm_axpby_f32(void* stream, float* out, float a, const float* x, float b, const float* y,
                      int64_t n);
/* out[i][3] = (float)grid(i)[d] + disp[i][d]: absolute positions from a relative state. */
int32_t jpm_grid_plus_disp_f32(void* stream, float* out, const float* disp, int32_t nx,
                               int32_t ny, int32_t nz, int32_t ox, int32_t oy);

/* ------------------------------------------------------------------------
 * multi-GPU building blocks (jaxpm/distributed.py:45-113; one process per GPU)
 * ---------------------------------------------------------------------- */
/* 1-D batched FFT plans for the slab/pencil transform ([ext] jaxdecomp.pfft3d, called from
 * jaxpm/distributed.py:37-42): `batch` contiguous transforms of length n.
 * kind 0 = R2C (n reals -> n/2+1 complex), 1 = C2R, 2 = C2C (direction chosen at exec). Host calls. */
typedef struct jpm_fft1d jpm_fft1d;
int32_t jpm_fft1d_create(jpm_fft1d** plan, int32_t n, int64_t batch, int32_t kind);
int32_t jpm_fft1d_destroy(jpm_fft1d* plan);
int32_t jpm_fft1d_exec(jpm_fft1d* plan, void* stream, void* in, void* out, int32_t inverse);
/* Batched tiled transpose of complex64: dst[b*dsb + j*dsj + i] = src[b*ssb + i*ssi + j]
 * (pack/unpack around the all-to-all transposes). Strides in complex elements. */
int32_t jpm_transpose_c64(void* stream, void* dst, const void* src, int32_t ni, int32_t nj,
                          int64_t nb, int64_t src_stride_i, int64_t src_stride_b,
                          int64_t dst_stride_j, int64_t dst_stride_b);
/* Strided row copy of complex64: dst[r*drs + c] = src[r*srs + c], c < ncols. */
int32_t jpm_copy2d_c64(void* stream, void* dst, const void* src, int64_t nrows, int32_t ncols,
                       int64_t src_row_stride, int64_t dst_row_stride);
/* The fused k-space passes (jpm_greens_grad_c64 = kind 0, jpm_lpt2_shear_c64 = kind 1,
 * jpm_greens_div_c64 = kind 2) on a LOCAL block [n0][n1][n2] of a distributed spectrum whose array
 * axes are a permutation of (x,y,z): w0..2 and a0..2 are the per-array-axis slices of the tables of
 * jaxpm/kernels.py:10-23 and :62-66; axis_of_{x,y,z} map physical directions to array axes. */
int32_t jpm_kspace_local_c64(void* stream, int32_t kind, const void* in, void* out, const float* w0,
                             const float* w1, const float* w2, const float* a0, const float* a1,
                             const float* a2, int32_t n0, int32_t n1, int32_t n2, int32_t axis_of_x,
                             int32_t axis_of_y, int32_t axis_of_z, float norm, float r_split,
                             const float* filter_tab, int32_t n_tab, float filter_kmax);

/* Copy / add a [x0:x1) x [y0:y1) x nz sub-box between a strided mesh and a packed buffer. */
int32_t jpm_pack_box_f32(void* stream, float* packed, const float* mesh, int32_t ny, int32_t nz,
                         int32_t x0, int32_t x1, int32_t y0, int32_t y1);
int32_t jpm_unpack_box_f32(void* stream, float* mesh, const float* packed, int32_t ny, int32_t nz,
                           int32_t x0, int32_t x1, int32_t y0, int32_t y1, int32_t accumulate);

#ifdef __cplusplus
}
#endif
#endif /* JAXPM_B200_H_ */
