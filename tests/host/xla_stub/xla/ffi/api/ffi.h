// Stand-in for XLA's typed-FFI header (xla/ffi/api/ffi.h), just large enough to COMPILE AND RUN
// the handler bodies of jax-cfd_b200/csrc/xla_ffi_shim.cc in an image without XLA / JAX.  TEST
// INFRASTRUCTURE: it mirrors the names and signatures the shim uses (Buffer, Result, Span, Error,
// Ffi::Bind().Ctx/Arg/Ret/Attr, XLA_FFI_DEFINE_HANDLER_SYMBOL); the binding itself is a no-op, the
// test harness (tests/host/xla_shim_harness.cc) calls the *Impl functions directly.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

namespace xla {
namespace ffi {

enum DataType { F32 = 11 };
enum class ErrorCode { kOk = 0, kInvalidArgument = 3, kInternal = 13 };

class Error {
 public:
  Error() = default;
  Error(ErrorCode code, std::string message) : code_(code), message_(std::move(message)) {}
  static Error Success() { return Error(); }
  bool success() const { return code_ == ErrorCode::kOk; }
  bool failure() const { return !success(); }
  ErrorCode errc() const { return code_; }
  const std::string& message() const { return message_; }

 private:
  ErrorCode code_ = ErrorCode::kOk;
  std::string message_;
};

template <typename T>
class Span {
 public:
  Span() = default;
  Span(const T* data, size_t size) : data_(data), size_(size) {}
  Span(const std::vector<typename std::remove_const<T>::type>& v) : data_(v.data()), size_(v.size()) {}
  size_t size() const { return size_; }
  const T& operator[](size_t i) const { return data_[i]; }
  const T* begin() const { return data_; }
  const T* end() const { return data_ + size_; }

 private:
  const T* data_ = nullptr;
  size_t size_ = 0;
};

template <DataType dtype>
class Buffer {
 public:
  Buffer() = default;
  Buffer(float* data, std::vector<int64_t> dims) : data_(data), dims_(std::move(dims)) {}
  float* typed_data() const { return data_; }
  Span<const int64_t> dimensions() const { return Span<const int64_t>(dims_.data(), dims_.size()); }
  size_t element_count() const {
    size_t n = 1;
    for (auto d : dims_) n *= (size_t)d;
    return n;
  }

 private:
  float* data_ = nullptr;
  std::vector<int64_t> dims_;
};

template <typename T>
class Result {
 public:
  Result() = default;
  explicit Result(T value) : value_(std::move(value)) {}
  T* operator->() { return &value_; }
  const T* operator->() const { return &value_; }
  T& operator*() { return value_; }

 private:
  T value_;
};
template <DataType dtype>
using ResultBuffer = Result<Buffer<dtype>>;

template <typename T>
struct PlatformStream {};

enum class Traits { kCmdBufferCompatible = 1 };

struct Binding {
  template <typename T>
  Binding Ctx() const { return *this; }
  template <typename T>
  Binding Arg() const { return *this; }
  template <typename T>
  Binding Ret() const { return *this; }
  template <typename T>
  Binding Attr(const char*) const { return *this; }
};
struct Ffi {
  static Binding Bind() { return Binding(); }
};

}  // namespace ffi
}  // namespace xla

// the real macro defines `extern "C" XLA_FFI_Error* symbol(XLA_FFI_CallFrame*)`; here it only keeps
// the implementation and the binding expression alive
#define XLA_FFI_DEFINE_HANDLER_SYMBOL(symbol, impl, ...) \
  extern "C" void* symbol##_stub() {                     \
    (void)(__VA_ARGS__);                                 \
    return reinterpret_cast<void*>(&impl);               \
  }
