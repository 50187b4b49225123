// Minimal STAND-IN for XLA's "xla/ffi/api/ffi.h" (jaxlib is not installed in this image): just the
// declarations somax_b200/csrc/jax_ffi_shim.cc uses, with the real header's names and call shapes,
// so that tests/test_abi.py can compile-check the shim (g++ -fsyntax-only) and it cannot rot.
// Not a replacement: build the shim against `jax.ffi.include_dir()`.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>

namespace xla {
namespace ffi {

enum class DataType { F32, F64 };
enum class ErrorCode { kInternal, kInvalidArgument, kUnimplemented };

class Error {
 public:
  Error(ErrorCode code, std::string message) : code_(code), message_(std::move(message)) {}
  static Error Success() { return Error(); }
 private:
  Error() : code_(ErrorCode::kInternal), ok_(true) {}
  ErrorCode code_;
  std::string message_;
  bool ok_ = false;
};

template <typename T>
class Span {
 public:
  const T* begin() const { return data_; }
  const T* end() const { return data_ + size_; }
  size_t size() const { return size_; }
  const T& operator[](size_t i) const { return data_[i]; }
 private:
  const T* data_ = nullptr;
  size_t size_ = 0;
};

class AnyBuffer {
 public:
  Span<const int64_t> dimensions() const { return {}; }
  DataType element_type() const { return DataType::F32; }
  void* untyped_data() const { return nullptr; }
  size_t size_bytes() const { return 0; }
  size_t element_count() const { return 0; }
};

template <typename T>
class Result {
 public:
  T* operator->() { return &value_; }
  T& operator*() { return value_; }
 private:
  T value_;
};

template <typename T> struct PlatformStream {};

struct Binding {
  template <typename T> Binding& Ctx() { return *this; }
  template <typename T> Binding& Arg() { return *this; }
  template <typename T> Binding& Ret() { return *this; }
  template <typename T> Binding& Attr(const char*) { return *this; }
};

struct Ffi {
  static Binding Bind() { return Binding(); }
};

}  // namespace ffi
}  // namespace xla

// The real macro defines an exported XLA_FFI_Handler symbol; here it only forces the handler and its
// binding expression through the compiler.
#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, impl, binding)           \
  extern "C" const void* name() {                                    \
    (void)(binding);                                                 \
    return reinterpret_cast<const void*>(&impl);                     \
  }
