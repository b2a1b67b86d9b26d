// xla_ffi_stub.h -- TEST INFRASTRUCTURE: the handful of declarations of xla/ffi/api/ffi.h that ffi/jr_ffi.cc uses,
// so that `g++ -fsyntax-only -DJR_FFI_STUB` can check the file in an image without jaxlib / CUDA headers.  Signatures
// follow jaxlib 0.4.31+'s xla/ffi/api/ffi.h; nothing here is linked or shipped.
#pragma once
#include <cstddef>
#include <cstdint>
#include <optional>
#include <string>
#include <vector>

typedef struct CUstream_st* cudaStream_t;
enum cudaMemcpyKind { cudaMemcpyDeviceToDevice = 3 };
inline int cudaMemcpyAsync(void*, const void*, size_t, cudaMemcpyKind, cudaStream_t) { return 0; }
inline int cudaMemsetAsync(void*, int, size_t, cudaStream_t) { return 0; }

namespace xla { namespace ffi {
template <typename T> struct Span { const T* p; size_t n; size_t size() const { return n; } const T& operator[](size_t i) const { return p[i]; } };
class Error {
 public:
  static Error Success() { return Error(); }
  static Error Internal(std::string) { return Error(true); }
  static Error InvalidArgument(std::string) { return Error(true); }
  bool failure() const { return bad_; }
 private:
  explicit Error(bool b = false) : bad_(b) {}
  bool bad_;
};
class AnyBuffer {
 public:
  Span<int64_t> dimensions() const { return {nullptr, 0}; }
  void* untyped_data() const { return nullptr; }
  size_t size_bytes() const { return 0; }
};
template <typename T> class Result {
 public:
  T* operator->() { return &v_; }
 private:
  T v_;
};
class RemainingArgs {
 public:
  size_t size() const { return 0; }
  template <typename T> std::optional<T> get(size_t) const { return T(); }
};
class RemainingRets {
 public:
  size_t size() const { return 0; }
  template <typename T> std::optional<Result<T>> get(size_t) const { return Result<T>(); }
};
template <typename T> struct PlatformStream {};
struct Binding {
  template <typename T> Binding Ctx() { return *this; }
  template <typename T> Binding Attr(const char*) { return *this; }
  Binding RemainingArgs() { return *this; }
  Binding RemainingRets() { return *this; }
};
struct Ffi { static Binding Bind() { return Binding(); } };
}}  // namespace xla::ffi
#define XLA_FFI_DEFINE_HANDLER_SYMBOL(sym, impl, binding) \
  extern "C" void* sym() { auto b = binding; (void)b; return reinterpret_cast<void*>(&impl); }
