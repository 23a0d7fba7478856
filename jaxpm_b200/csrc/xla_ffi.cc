// Thin XLA FFI handlers over the C ABI (include/jaxpm_b200.h).  Compiled only where jaxlib's FFI
// headers are available (`-I$(python -c "import jax.ffi; print(jax.ffi.include_dir())")`); JAX is not
// installable in this image, so this file is NOT part of the default build.
// Each handler forwards XLA-owned device buffers and the compute stream; nothing allocates or blocks.
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define JPM_HAVE_XLA_FFI 1
#endif
#endif

#ifdef JPM_HAVE_XLA_FFI
#include <cuda_runtime.h>

#include "../../include/jaxpm_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

static ffi::Error status(int32_t rc) {
  return rc ? ffi::Error::Internal(jpm_last_error_string()) : ffi::Error::Success();
}

// jaxpm/painting.py:15-45; mesh_out aliases mesh_in (input_output_aliases={0: 0}).
static ffi::Error PaintImpl(cudaStream_t s, ffi::Buffer<ffi::F32> mesh_in, ffi::Buffer<ffi::F32> pos,
                            ffi::ResultBuffer<ffi::F32> mesh_out, float weight) {
  auto d = mesh_in.dimensions();
  if (mesh_out->typed_data() != mesh_in.typed_data())
    cudaMemcpyAsync(mesh_out->typed_data(), mesh_in.typed_data(), mesh_in.size_bytes(), cudaMemcpyDeviceToDevice, s);
  return status(jpm_cic_paint_f32(s, mesh_out->typed_data(), pos.typed_data(), nullptr, weight,
                                  (int64_t)pos.element_count() / 3, d[0], d[1], d[2], d[0], d[1], d[2]));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(JpmCicPaint, PaintImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Attr<float>("weight"));

// jaxpm/painting.py:78-106
static ffi::Error ReadImpl(cudaStream_t s, ffi::Buffer<ffi::F32> mesh, ffi::Buffer<ffi::F32> pos,
                           ffi::ResultBuffer<ffi::F32> out) {
  auto d = mesh.dimensions();
  return status(jpm_cic_read_f32(s, out->typed_data(), mesh.typed_data(), pos.typed_data(),
                                 (int64_t)pos.element_count() / 3, d[0], d[1], d[2]));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(JpmCicRead, ReadImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>());

// jaxpm/painting.py:161-189 (per shard; hx, hy = halo offsets)
static ffi::Error PaintDxImpl(cudaStream_t s, ffi::Buffer<ffi::F32> disp, ffi::ResultBuffer<ffi::F32> mesh,
                              float weight, int32_t hx, int32_t hy) {
  auto d = disp.dimensions();
  cudaMemsetAsync(mesh->typed_data(), 0, mesh->size_bytes(), s);
  return status(jpm_cic_paint_dx_f32(s, mesh->typed_data(), disp.typed_data(), nullptr, weight, d[0], d[1], d[2],
                                     hx, hy));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(JpmCicPaintDx, PaintDxImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Attr<float>("weight")
                                  .Attr<int32_t>("hx")
                                  .Attr<int32_t>("hy"));

// the three reads + stack of jaxpm/pm.py:54-56 (force3 = [3, mx, my, mz])
static ffi::Error Read3Impl(cudaStream_t s, ffi::Buffer<ffi::F32> force3, ffi::Buffer<ffi::F32> pos,
                            ffi::ResultBuffer<ffi::F32> out, int32_t hx, int32_t hy, int32_t relative) {
  auto d = force3.dimensions();
  const int64_t nc = (int64_t)d[1] * d[2] * d[3];
  const float* f = force3.typed_data();
  return status(jpm_cic_read3_f32(s, out->typed_data(), f, f + nc, f + 2 * nc, pos.typed_data(), 1.0f,
                                  (int64_t)pos.element_count() / 3, d[1], d[2], d[3], hx, hy, relative));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(JpmCicRead3, Read3Impl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Attr<int32_t>("hx")
                                  .Attr<int32_t>("hy")
                                  .Attr<int32_t>("relative"));
// ---- the fused hot path -----------------------------------------------------------------------------------
// Opaque handles (jpm_plan / jpm_sim) are created once outside jit and travel as int64 attributes, the way
// jaxdecomp keeps its cuFFT plans.

// jaxpm/ode.py:91-117 around the three reads of pm.py:54-56: (pos, vel) -> (pos', vel'), outputs alias the inputs
static ffi::Error Read3KickDriftImpl(cudaStream_t s, ffi::Buffer<ffi::F32> force3, ffi::Buffer<ffi::F32> pos,
                                     ffi::Buffer<ffi::F32> vel, ffi::ResultBuffer<ffi::F32> pos_out,
                                     ffi::ResultBuffer<ffi::F32> vel_out, float kick, float drift, int32_t hx,
                                     int32_t hy, int32_t relative) {
  auto d = force3.dimensions();
  const int64_t nc = (int64_t)d[1] * d[2] * d[3];
  const float* f = force3.typed_data();
  return status(jpm_cic_read3_kick_drift_f32(s, pos_out->typed_data(), vel_out->typed_data(), nullptr, f, f + nc,
                                             f + 2 * nc, pos.typed_data(), vel.typed_data(), pos.typed_data(),
                                             vel.typed_data(), kick, drift, 1, (int64_t)pos.element_count() / 3,
                                             d[1], d[2], d[3], hx, hy, relative));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(JpmRead3KickDrift, Read3KickDriftImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Attr<float>("kick")
                                  .Attr<float>("drift")
                                  .Attr<int32_t>("hx")
                                  .Attr<int32_t>("hy")
                                  .Attr<int32_t>("relative"));

// jaxpm/pm.py:41-56 on the fused FFT chain: density [nx, ny, nz] -> force3 [3, nx, ny, nz]
static ffi::Error ForceMeshesImpl(cudaStream_t s, ffi::Buffer<ffi::F32> density, ffi::ResultBuffer<ffi::F32> force3,
                                  int64_t plan, float r_split) {
  return status(jpm_density_to_force_meshes_fused(reinterpret_cast<jpm_plan*>(plan), s, density.typed_data(),
                                                  force3->typed_data(), r_split, nullptr, 0, 0.f));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(JpmDensityToForceMeshes, ForceMeshesImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Attr<int64_t>("plan")
                                  .Attr<float>("r_split"));

// pm_forces(positions) on the tile kernels: positions [np, 3] -> forces [np, 3], the caller's particle order
static ffi::Error SimForcesImpl(cudaStream_t s, ffi::Buffer<ffi::F32> pos, ffi::ResultBuffer<ffi::F32> forces,
                                int64_t sim, float r_split) {
  jpm_sim* h = reinterpret_cast<jpm_sim*>(sim);
  int32_t rc = jpm_sim_load(h, s, pos.typed_data(), nullptr);
  if (rc) return status(rc);
  return status(jpm_sim_forces(h, s, forces->typed_data(), 1.0f, r_split, nullptr, 0, 0.f));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(JpmSimForces, SimForcesImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Attr<int64_t>("sim")
                                  .Attr<float>("r_split"));

// K resident steps: (pos, vel) in the caller's order -> tile sort, K x jpm_sim_step, un-sort.  kick / drift: [K] on
// the HOST side are attributes of a scan body in practice; here a single step per call (lax.scan / fori_loop body).
static ffi::Error SimStepImpl(cudaStream_t s, ffi::Buffer<ffi::F32> pos, ffi::Buffer<ffi::F32> vel,
                              ffi::ResultBuffer<ffi::F32> pos_out, ffi::ResultBuffer<ffi::F32> vel_out, int64_t sim,
                              float kick, float drift, int32_t load, int32_t store) {
  jpm_sim* h = reinterpret_cast<jpm_sim*>(sim);
  int32_t rc = 0;
  if (load && (rc = jpm_sim_load(h, s, pos.typed_data(), vel.typed_data()))) return status(rc);
  if ((rc = jpm_sim_step(h, s, kick, drift))) return status(rc);
  if (store) rc = jpm_sim_store(h, s, pos_out->typed_data(), vel_out->typed_data());
  return status(rc);
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(JpmSimStep, SimStepImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Attr<int64_t>("sim")
                                  .Attr<float>("kick")
                                  .Attr<float>("drift")
                                  .Attr<int32_t>("load")
                                  .Attr<int32_t>("store"));

// ---- adjoint / tangent building blocks (custom_jvp + linear_call rules) -----------
// value [np] and grad [np, 3] of a read (read VJP / JVP wrt positions, paint VJP wrt positions and weights)
static ffi::Error ReadGradImpl(cudaStream_t s, ffi::Buffer<ffi::F32> mesh, ffi::Buffer<ffi::F32> pos,
                               ffi::Buffer<ffi::F32> scale, ffi::ResultBuffer<ffi::F32> value,
                               ffi::ResultBuffer<ffi::F32> grad, int32_t hx, int32_t hy, int32_t relative) {
  auto d = mesh.dimensions();
  return status(jpm_cic_readgrad_f32(s, value->typed_data(), grad->typed_data(), mesh.typed_data(), pos.typed_data(),
                                     scale.typed_data(), 1.0f, (int64_t)pos.element_count() / 3, d[0], d[1], d[2],
                                     hx, hy, relative));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(JpmCicReadGrad, ReadGradImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Attr<int32_t>("hx")
                                  .Attr<int32_t>("hy")
                                  .Attr<int32_t>("relative"));

// tangent of a paint for a position tangent (forward mode; transpose of ReadGrad)
static ffi::Error PaintGradImpl(cudaStream_t s, ffi::Buffer<ffi::F32> pos, ffi::Buffer<ffi::F32> tangent,
                                ffi::Buffer<ffi::F32> weight, ffi::ResultBuffer<ffi::F32> mesh, int32_t hx, int32_t hy,
                                int32_t relative) {
  auto d = mesh->dimensions();
  cudaMemsetAsync(mesh->typed_data(), 0, mesh->size_bytes(), s);
  return status(jpm_cic_paintgrad_f32(s, mesh->typed_data(), pos.typed_data(), tangent.typed_data(),
                                      weight.typed_data(), 1.0f, (int64_t)pos.element_count() / 3, d[0], d[1], d[2],
                                      hx, hy, relative));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(JpmCicPaintGrad, PaintGradImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Attr<int32_t>("hx")
                                  .Attr<int32_t>("hy")
                                  .Attr<int32_t>("relative"));

// transpose of the k-space pass of pm.py:49-56: 3 half-spectra -> 1 (reverse mode of pm_forces)
static ffi::Error GreensDivImpl(cudaStream_t s, ffi::Buffer<ffi::C64> in3, ffi::ResultBuffer<ffi::C64> out, int64_t plan,
                                float norm, float r_split) {
  return status(jpm_greens_div_c64(reinterpret_cast<jpm_plan*>(plan), s, in3.typed_data(), out->typed_data(), norm,
                                   r_split, nullptr, 0, 0.f));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(JpmGreensDiv, GreensDivImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::C64>>()
                                  .Ret<ffi::Buffer<ffi::C64>>()
                                  .Attr<int64_t>("plan")
                                  .Attr<float>("norm")
                                  .Attr<float>("r_split"));

// the whole vector-Jacobian product of pm_forces with respect to the positions on the fused passes (what the
// custom_vjp rule of pm_forces binds: jaxpm_b200/pm.py:_pm_forces_vjp_fused).  f3 = the force meshes recomputed by
// JpmDensityToForceMeshes; workspace: three meshes G3 + one mesh S + one mesh psi, owned by XLA (Ret buffers).
static ffi::Error PmForcesVjpImpl(cudaStream_t s, ffi::Buffer<ffi::F32> f3, ffi::Buffer<ffi::F32> pos,
                                  ffi::Buffer<ffi::F32> cotangent, ffi::ResultBuffer<ffi::F32> grad,
                                  ffi::ResultBuffer<ffi::F32> g3, ffi::ResultBuffer<ffi::F32> div,
                                  ffi::ResultBuffer<ffi::F32> psi, int64_t plan, float r_split, int32_t relative) {
  auto d = f3.dimensions();   // [3, nx, ny, nz]
  const int64_t nc = (int64_t)d[1] * d[2] * d[3], np = (int64_t)pos.element_count() / 3;
  int32_t rc = jpm_cic_readgrad3_f32(s, grad->typed_data(), f3.typed_data(), f3.typed_data() + nc,
                                     f3.typed_data() + 2 * nc, pos.typed_data(), cotangent.typed_data(), 1.0f, np, d[1],
                                     d[2], d[3], 0, 0, relative, 0);
  if (rc) return status(rc);
  cudaMemsetAsync(g3->typed_data(), 0, g3->size_bytes(), s);
  if ((rc = jpm_cic_paint3_f32(s, g3->typed_data(), pos.typed_data(), cotangent.typed_data(), 1.0f, np, d[1], d[2], d[3],
                               0, 0, relative)))
    return status(rc);
  if ((rc = jpm_fd_divergence3_f32(s, div->typed_data(), g3->typed_data(), d[1], d[2], d[3]))) return status(rc);
  if ((rc = jpm_density_to_potential_fused(reinterpret_cast<jpm_plan*>(plan), s, div->typed_data(), psi->typed_data(),
                                           r_split, nullptr, 0, 0.f)))
    return status(rc);
  return status(jpm_cic_readgrad3_f32(s, grad->typed_data(), psi->typed_data(), nullptr, nullptr, pos.typed_data(),
                                      nullptr, -1.0f, np, d[1], d[2], d[3], 0, 0, relative, 1));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(JpmPmForcesVjp, PmForcesVjpImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Attr<int64_t>("plan")
                                  .Attr<float>("r_split")
                                  .Attr<int32_t>("relative"));

// ---- multi-GPU slab plan under shard_map (one call per shard, collective: every shard must be launched) --------
static ffi::Error SlabForcesImpl(cudaStream_t s, ffi::Buffer<ffi::F32> density_local, ffi::ResultBuffer<ffi::F32> force3,
                                 int64_t plan, float r_split) {
  jpm_plan* p = reinterpret_cast<jpm_plan*>(plan);
  int32_t rc = jpm_slab_set_density_f32(p, s, density_local.typed_data());
  if (rc) return status(rc);
  if ((rc = jpm_slab_forces(p, s, r_split))) return status(rc);
  const int64_t nl = (int64_t)density_local.element_count();
  for (int c = 0; c < 3 && !rc; ++c) rc = jpm_slab_get_interior_f32(p, s, 1 + c, force3->typed_data() + c * nl);
  return status(rc);
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(JpmSlabForces, SlabForcesImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Attr<int64_t>("plan")
                                  .Attr<float>("r_split"));

// ---- initial conditions -------------------------------------------------------------------------------------
static ffi::Error NormalFieldImpl(cudaStream_t s, ffi::ResultBuffer<ffi::F32> out, int64_t seed, int32_t ox, int32_t oy,
                                  int32_t global_ny) {
  auto d = out->dimensions();
  return status(jpm_normal_field_f32(s, out->typed_data(), d[0], d[1], d[2], ox, oy, global_ny, (uint64_t)seed, 0u));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(JpmNormalField, NormalFieldImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Attr<int64_t>("seed")
                                  .Attr<int32_t>("ox")
                                  .Attr<int32_t>("oy")
                                  .Attr<int32_t>("global_ny"));
#endif  // JPM_HAVE_XLA_FFI
