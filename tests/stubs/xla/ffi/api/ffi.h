// Minimal STAND-IN for jaxlib's xla/ffi/api/ffi.h (not installed in this image), just enough surface to type-check
// bindings/xla_ffi_shim.cc: the binder records the C++ type each Ctx / Arg / Ret / Attr contributes, and
// XLA_FFI_DEFINE_HANDLER_SYMBOL static_asserts that the implementation function is invocable with exactly those types and
// returns ffi::Error -- the same contract the real header enforces.  Test infrastructure only (tests/test_ffi_shim.py).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

namespace xla {
namespace ffi {

enum class ErrorCode { kOk, kInvalidArgument, kInternal };

class Error {
 public:
  Error() = default;
  Error(ErrorCode code, std::string msg) : code_(code), msg_(std::move(msg)) {}
  static Error Success() { return Error(); }
  bool failure() const { return code_ != ErrorCode::kOk; }

 private:
  ErrorCode code_ = ErrorCode::kOk;
  std::string msg_;
};

template <typename T>
struct Span {
  const T* p = nullptr;
  size_t n = 0;
  size_t size() const { return n; }
  const T& operator[](size_t i) const { return p[i]; }
};

class AnyBuffer {
 public:
  void* untyped_data() const { return nullptr; }
  Span<int64_t> dimensions() const { return {}; }
  size_t size_bytes() const { return 0; }
  size_t element_count() const { return 0; }
};

template <typename T>
class Result {
 public:
  T* operator->() { return &v_; }
  T& operator*() { return v_; }

 private:
  T v_;
};

template <typename T>
struct PlatformStream {};

template <typename... Ts>
struct TypeList {};

template <typename... Ts>
struct Binding {
  template <typename T>
  Binding<Ts..., T> Ctx_() const { return {}; }
  // PlatformStream<S> contributes S
  template <typename C>
  auto Ctx() const { return CtxImpl(static_cast<C*>(nullptr)); }
  template <typename S>
  Binding<Ts..., S> CtxImpl(PlatformStream<S>*) const { return {}; }
  template <typename T>
  Binding<Ts..., T> Arg() const { return {}; }
  template <typename T>
  Binding<Ts..., Result<T>> Ret() const { return {}; }
  template <typename T>
  Binding<Ts..., T> Attr(const char*) const { return {}; }
  template <typename Fn>
  static constexpr bool Matches() {
    return std::is_invocable_r_v<Error, Fn, Ts...>;
  }
};

struct Ffi {
  static Binding<> Bind() { return {}; }
};

}  // namespace ffi
}  // namespace xla

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, impl, binding)                                                     \
  static_assert(decltype(binding)::template Matches<decltype(&impl)>(),                                        \
                #impl " does not match its xla::ffi binding (Ctx/Arg/Ret/Attr order and types)");            \
  extern "C" void* name() { return reinterpret_cast<void*>(&impl); }
