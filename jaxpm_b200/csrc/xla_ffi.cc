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
#endif  // JPM_HAVE_XLA_FFI
