// DECLARATION-ONLY STAND-IN of the header-only XLA FFI C++ API (xla/ffi/api/ffi.h, shipped by jaxlib under
// jax.ffi.include_dir()), restricted to what ffi/xla_ffi_shim.cc uses.  Purpose: let `make -C ffi check` type-check
// the shim in a container without jaxlib.  It models the binding's type-level contract -- the handler implementation
// must be invocable with the decoded types of the Ctx / Attr / Arg / RemainingArgs / Ret chain, in order -- and nothing
// else: the handler symbols it defines do nothing.  NEVER ship a library built against this file.
#pragma once
#include <cstddef>
#include <cstdint>
#include <optional>
#include <string>
#include <type_traits>
#include <utility>

struct XLA_FFI_Error;
struct XLA_FFI_CallFrame;

namespace xla::ffi {

enum DataType { PRED, S8, S16, S32, S64, U8, U16, U32, U64, F16, F32, F64, BF16, C64, C128 };
enum class ErrorCode { kOk, kCancelled, kUnknown, kInvalidArgument, kNotFound, kResourceExhausted, kUnimplemented, kInternal };

class Error {
 public:
  Error() = default;
  Error(ErrorCode code, std::string message) : code_(code), message_(std::move(message)) {}
  static Error Success() { return Error(); }
  static Error InvalidArgument(std::string m) { return Error(ErrorCode::kInvalidArgument, std::move(m)); }
  static Error Internal(std::string m) { return Error(ErrorCode::kInternal, std::move(m)); }
  bool success() const { return code_ == ErrorCode::kOk; }
  bool failure() const { return !success(); }
  const std::string& message() const { return message_; }

 private:
  ErrorCode code_ = ErrorCode::kOk;
  std::string message_;
};

template <typename T>
class ErrorOr {
 public:
  ErrorOr(T v) : value_(std::move(v)) {}
  ErrorOr(Error e) : error_(std::move(e)) {}
  bool has_value() const { return value_.has_value(); }
  T& value() { return *value_; }
  T* operator->() { return &*value_; }
  T& operator*() { return *value_; }
  const Error& error() const { return error_; }

 private:
  std::optional<T> value_;
  Error error_;
};

template <typename T>
class Span {
 public:
  Span() = default;
  Span(const T* d, size_t n) : data_(d), size_(n) {}
  const T* begin() const { return data_; }
  const T* end() const { return data_ + size_; }
  size_t size() const { return size_; }
  const T& operator[](size_t i) const { return data_[i]; }

 private:
  const T* data_ = nullptr;
  size_t size_ = 0;
};

template <DataType dtype>
struct NativeTypeOf { using type = void; };
template <> struct NativeTypeOf<F32> { using type = float; };
template <> struct NativeTypeOf<U8> { using type = uint8_t; };
template <> struct NativeTypeOf<S32> { using type = int32_t; };

class AnyBuffer {
 public:
  DataType element_type() const { return F32; }
  void* untyped_data() const { return nullptr; }
  Span<int64_t> dimensions() const { return {}; }
  size_t element_count() const { return 0; }
  size_t size_bytes() const { return 0; }
};

template <DataType dtype>
class Buffer {
 public:
  using T = typename NativeTypeOf<dtype>::type;
  T* typed_data() const { return nullptr; }
  void* untyped_data() const { return nullptr; }
  Span<int64_t> dimensions() const { return {}; }
  size_t element_count() const { return 0; }
  size_t size_bytes() const { return 0; }
};

template <typename T>
class Result {
 public:
  T* operator->() { return &value_; }
  T& operator*() { return value_; }

 private:
  T value_;
};
template <DataType dtype>
using ResultBuffer = Result<Buffer<dtype>>;

class RemainingArgs {
 public:
  size_t size() const { return 0; }
  template <typename T>
  ErrorOr<T> get(size_t) const { return ErrorOr<T>(Error::InvalidArgument("stub")); }
};

class ScratchAllocator {
 public:
  std::optional<void*> Allocate(size_t, size_t = 1) { return std::nullopt; }
};

template <typename T>
struct PlatformStream {};

namespace internal {
template <typename T> struct CtxDecoded { using type = T; };
template <typename T> struct CtxDecoded<PlatformStream<T>> { using type = T; };
}  // namespace internal

template <typename... Ts>
struct Binding {
  template <typename T> Binding<Ts..., typename internal::CtxDecoded<T>::type> Ctx() const { return {}; }
  template <typename T> Binding<Ts..., T> Arg() const { return {}; }
  template <typename T> Binding<Ts..., Result<T>> Ret() const { return {}; }
  template <typename T> Binding<Ts..., T> Attr(const char*) const { return {}; }
  Binding<Ts..., ::xla::ffi::RemainingArgs> RemainingArgs() const { return {}; }
  template <typename Fn>
  static constexpr bool Accepts() { return std::is_invocable_r_v<Error, Fn, Ts...>; }
};

struct Ffi {
  static Binding<> Bind() { return {}; }
};

}  // namespace xla::ffi

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(symbol, impl, binding)                                                      \
  extern "C" XLA_FFI_Error* symbol(XLA_FFI_CallFrame*) {                                                          \
    using B_ = decltype(binding);                                                                                 \
    static_assert(B_::template Accepts<decltype(&impl)>(), #impl " does not match the binding of " #symbol);     \
    return nullptr;                                                                                               \
  }
