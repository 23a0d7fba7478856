// MINIMAL STAND-IN for jaxlib's xla/ffi/api/ffi.h, only to type-check jaxpm_b200/csrc/xla_ffi.cc in an image without
// JAX (tests/test_abi.py::test_xla_ffi_handlers_type_check, `make -C jaxpm_b200/csrc ffi-check`).  It declares the
// handful of names the handlers use with the shapes the real header gives them (Buffer<T>::typed_data / dimensions /
// element_count / size_bytes, ResultBuffer<T> as a pointer-like Result, Error::Success / Internal, the Bind() builder,
// XLA_FFI_DEFINE_HANDLER_SYMBOL).  It checks the handler BODIES against the C ABI (argument order and types of every
// jpm_* call); it does not check the binding against the handler signature - the real header does that.
#pragma once
#include <complex>
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace xla {
namespace ffi {
enum DataType { F32, C64 };
template <DataType> struct NativeType;
template <> struct NativeType<F32> { using type = float; };
template <> struct NativeType<C64> { using type = std::complex<float>; };

struct Span {
  std::vector<int64_t> v;
  int64_t operator[](size_t i) const { return v[i]; }
  size_t size() const { return v.size(); }
};

template <DataType T>
class Buffer {
 public:
  using Native = typename NativeType<T>::type;
  Native* typed_data() const { return data_; }
  Span dimensions() const { return dims_; }
  size_t element_count() const { return count_; }
  size_t size_bytes() const { return count_ * sizeof(Native); }

 private:
  Native* data_ = nullptr;
  Span dims_;
  size_t count_ = 0;
};

template <typename T>
class Result {
 public:
  T* operator->() { return &value_; }
  T& operator*() { return value_; }

 private:
  T value_;
};
template <DataType T> using ResultBuffer = Result<Buffer<T>>;

class Error {
 public:
  static Error Success() { return Error(); }
  static Error Internal(std::string m) { Error e; e.msg_ = std::move(m); return e; }

 private:
  std::string msg_;
};

template <typename T> struct PlatformStream {};

struct Binding {
  template <typename T> Binding Ctx() const { return *this; }
  template <typename T> Binding Arg() const { return *this; }
  template <typename T> Binding Ret() const { return *this; }
  template <typename T> Binding Attr(const char*) const { return *this; }
};
struct Ffi {
  static Binding Bind() { return Binding(); }
};
}  // namespace ffi
}  // namespace xla

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, impl, binding)          \
  extern "C" const void* name() {                                   \
    (void)(binding);                                                \
    return reinterpret_cast<const void*>(&impl);                    \
  }
