// TEST INFRASTRUCTURE: an API-shaped stand-in for jaxlib's `xla/ffi/api/ffi.h` (which is not installable in this
// image), just large enough to TYPE-CHECK deluca_b200/csrc/xla_ffi/deluca_xla_ffi.cc: buffer accessors,
// Error, Span, and a Bind() chain whose To() verifies that the handler is invocable with exactly the bound context /
// argument / attribute / result types, in order -- the property the real binding enforces.  It executes nothing.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <tuple>
#include <type_traits>
#include <utility>

namespace xla::ffi {

enum DataType { F32, F64, S32, U8 };
template <DataType> struct NativeOf;
template <> struct NativeOf<F32> { using type = float; };
template <> struct NativeOf<F64> { using type = double; };
template <> struct NativeOf<S32> { using type = int32_t; };
template <> struct NativeOf<U8> { using type = uint8_t; };

template <typename T>
class Span {
 public:
  Span() = default;
  Span(const T* d, size_t n) : d_(d), n_(n) {}
  const T* begin() const { return d_; }
  const T* end() const { return d_ + n_; }
  size_t size() const { return n_; }
  const T& operator[](size_t i) const { return d_[i]; }
  const T& back() const { return d_[n_ - 1]; }
 private:
  const T* d_ = nullptr;
  size_t n_ = 0;
};

template <DataType dtype>
class Buffer {
 public:
  using Native = typename NativeOf<dtype>::type;
  Native* typed_data() const { return static_cast<Native*>(data_); }
  void* untyped_data() const { return data_; }
  Span<const int64_t> dimensions() const { return Span<const int64_t>(dims_, rank_); }
  size_t element_count() const { size_t n = 1; for (size_t i = 0; i < rank_; ++i) n *= dims_[i]; return n; }
  size_t size_bytes() const { return element_count() * sizeof(Native); }
 private:
  void* data_ = nullptr;
  const int64_t* dims_ = nullptr;
  size_t rank_ = 0;
};

template <typename T>
class Result {
 public:
  T* operator->() { return &v_; }
  T& operator*() { return v_; }
 private:
  T v_;
};
template <DataType dtype> using ResultBuffer = Result<Buffer<dtype>>;

enum class ErrorCode { kOk, kInternal, kInvalidArgument, kUnimplemented };
class Error {
 public:
  Error() = default;
  Error(ErrorCode c, std::string m) : code_(c), msg_(std::move(m)) {}
  static Error Success() { return Error(); }
  bool success() const { return code_ == ErrorCode::kOk; }
 private:
  ErrorCode code_ = ErrorCode::kOk;
  std::string msg_;
};

template <typename T> struct PlatformStream { using type = T; };

namespace internal {
template <typename T> struct CtxArg;
template <typename T> struct CtxArg<PlatformStream<T>> { using type = T; };
template <typename... Ts>
struct Binding {
  template <typename C> Binding<Ts..., typename CtxArg<C>::type> Ctx() const { return {}; }
  template <typename A> Binding<Ts..., A> Arg() const { return {}; }
  template <typename A> Binding<Ts..., A> Attr(const char*) const { return {}; }
  template <typename R> Binding<Ts..., Result<R>> Ret() const { return {}; }
  template <typename Fn>
  int To(Fn fn) const {
    static_assert(std::is_invocable_r<Error, Fn, Ts...>::value,
                  "handler signature does not match the bound context / argument / attribute / result types");
    (void)fn;
    return 0;
  }
};
}  // namespace internal

struct Ffi {
  static internal::Binding<> Bind() { return {}; }
};

}  // namespace xla::ffi

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, impl, binding) \
  extern "C" int name##_typechecked(void) { return (binding).To(impl); }
