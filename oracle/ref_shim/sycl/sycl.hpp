// sycl/sycl.hpp — single-threaded HOST shim of the small part of SYCL 2020 that srtb's operator
// headers use, so that the reference's own headers (compiled from /root/reference, never copied)
// can run on a CPU that has no SYCL compiler. TEST INFRASTRUCTURE (oracle/_ref). Kernels run as
// plain loops; nd_range kernels run their work-items as ucontext fibers so barrier() works.
#pragma once
#include <ucontext.h>

#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <string>
#include <type_traits>
#include <vector>

#define SYCL_LANGUAGE_VERSION 202001

namespace sycl {

enum class backend { host, cuda, hip, omp };

template <int D = 1>
struct range {
  size_t v;
  range(size_t a = 0) : v{a} {}
  size_t operator[](int) const { return v; }
  size_t get(int) const { return v; }
  size_t size() const { return v; }
  friend range operator*(range a, range b) { return range{a.v * b.v}; }
};
template <class T> range(T) -> range<1>;

template <int D = 1>
struct id {
  size_t v[3] = {0, 0, 0};
  id() = default;
  id(size_t a) { v[0] = a; }
  id(size_t a, size_t b, size_t c) { v[0] = a; v[1] = b; v[2] = c; }
  operator size_t() const { return v[0]; }
  size_t operator[](int i) const { return v[i]; }
  size_t get(int i) const { return v[i]; }
};

template <int D = 1>
struct item {
  size_t i;
  size_t get_id(int = 0) const { return i; }
  size_t get_linear_id() const { return i; }
  operator size_t() const { return i; }
};

template <int D = 1>
struct nd_range {
  range<D> global, local;
  nd_range(range<D> g, range<D> l) : global{g}, local{l} {}
};

namespace access {
enum class fence_space { local_space, global_space, global_and_local };
enum class mode { read, write, read_write };
enum class target { device, local };
}  // namespace access

namespace detail {
struct fiber_state {
  ucontext_t sched{};
  ucontext_t* current = nullptr;
};
inline thread_local fiber_state fibers;
inline thread_local std::function<void()>* fiber_entry = nullptr;
inline void fiber_trampoline() { (*fiber_entry)(); }
}  // namespace detail

template <int D = 1>
struct nd_item {
  size_t group, local, local_range, global_range;
  size_t get_group(int = 0) const { return group; }
  size_t get_local_id(int = 0) const { return local; }
  size_t get_global_id(int = 0) const { return group * local_range + local; }
  size_t get_local_range(int = 0) const { return local_range; }
  size_t get_global_range(int = 0) const { return global_range; }
  void barrier(access::fence_space = access::fence_space::global_and_local) const {
    swapcontext(detail::fibers.current, &detail::fibers.sched);  // yield to the group scheduler
  }
};

struct event {
  void wait() {}
  void wait_and_throw() {}
};

namespace info {
namespace device {
struct max_work_group_size { using return_type = size_t; };
template <int D = 3> struct max_work_item_sizes { using return_type = id<3>; };
struct local_mem_size { using return_type = size_t; };
struct max_compute_units { using return_type = size_t; };
struct name { using return_type = std::string; };
}  // namespace device
}  // namespace info

class device {
 public:
  template <class Info>
  typename Info::return_type get_info() const {
    using R = typename Info::return_type;
    if constexpr (std::is_same_v<Info, info::device::max_work_group_size>) return R{32};
    else if constexpr (std::is_same_v<Info, info::device::local_mem_size>) return R{64 * 1024};
    else if constexpr (std::is_same_v<Info, info::device::max_compute_units>) return R{8};
    else if constexpr (std::is_same_v<Info, info::device::name>) return R{"srtb host shim"};
    else return R{32, 32, 32};
  }
  bool is_cpu() const { return true; }
  bool is_gpu() const { return false; }
  static std::vector<device> get_devices() { return {device{}}; }
};

class handler;

template <class T, int D = 1>
struct local_accessor {
  std::shared_ptr<std::vector<T>> store;
  local_accessor(range<D> r, handler&) : store{std::make_shared<std::vector<T>>(r.size())} {}
  T& operator[](size_t i) const { return (*store)[i]; }
};

class handler {
 public:
  template <class K>
  void parallel_for(range<1> r, K k) {
    for (size_t i = 0; i < r.size(); i++) {
      if constexpr (std::is_invocable_v<K, item<1>>) k(item<1>{i});
      else k(id<1>{i});
    }
  }
  template <class K>
  void parallel_for(nd_range<1> r, K k) {
    const size_t nl = r.local.size(), ng = r.global.size() / nl;
    constexpr size_t stack_bytes = 256 * 1024;
    std::vector<std::unique_ptr<char[]>> stacks(nl);
    for (auto& s : stacks) s.reset(new char[stack_bytes]);
    for (size_t g = 0; g < ng; g++) {
      std::vector<ucontext_t> ctx(nl);
      std::vector<char> done(nl, 0);
      std::vector<std::function<void()>> entry(nl);
      for (size_t l = 0; l < nl; l++) {
        entry[l] = [&, l]() {
          k(nd_item<1>{g, l, nl, r.global.size()});
          done[l] = 1;
          swapcontext(&ctx[l], &detail::fibers.sched);
        };
        getcontext(&ctx[l]);
        ctx[l].uc_stack.ss_sp = stacks[l].get();
        ctx[l].uc_stack.ss_size = stack_bytes;
        ctx[l].uc_link = nullptr;
        makecontext(&ctx[l], &detail::fiber_trampoline, 0);
      }
      std::vector<char> started(nl, 0);
      size_t remaining = nl;
      while (remaining) {
        for (size_t l = 0; l < nl; l++) {
          if (done[l]) continue;
          detail::fibers.current = &ctx[l];
          if (!started[l]) {
            started[l] = 1;
            detail::fiber_entry = &entry[l];
          }
          swapcontext(&detail::fibers.sched, &ctx[l]);
          if (done[l]) remaining--;
        }
      }
    }
  }
  template <class R, class K>
  void parallel_for(range<1>, R, K) {}  // reduction overload: never instantiated
  template <class K>
  void single_task(K k) { k(); }
};

class queue {
 public:
  queue() = default;
  template <class K>
  event parallel_for(range<1> r, K k) {
    handler h;
    h.parallel_for(r, k);
    return {};
  }
  template <class K>
  event parallel_for(nd_range<1> r, K k) {
    handler h;
    h.parallel_for(r, k);
    return {};
  }
  template <class K>
  event single_task(K k) {
    k();
    return {};
  }
  template <class T>
  event copy(const T* src, T* dst, size_t n) {
    std::memcpy(dst, src, n * sizeof(T));
    return {};
  }
  template <class T>
  event fill(T* p, const T& v, size_t n) {
    for (size_t i = 0; i < n; i++) p[i] = v;
    return {};
  }
  template <class F>
  event submit(F f) {
    handler h;
    f(h);
    return {};
  }
  device get_device() const { return {}; }
  void wait() {}
  void wait_and_throw() {}
};

struct interop_handle {};
// declared only: srtb compiles the sycl::reduction path out (config.hpp:50 use_sycl_reduction = false)
template <class... A>
inline int reduction(A&&...) { return 0; }
template <backend B, class Q>
inline void* get_native(Q&) { return nullptr; }

// ---- math (fp32 overloads call the f-suffixed libm functions, as a SYCL CPU backend does)
inline float cos(float x) { return ::cosf(x); }
inline double cos(double x) { return ::cos(x); }
inline float sin(float x) { return ::sinf(x); }
inline double sin(double x) { return ::sin(x); }
inline float sqrt(float x) { return ::sqrtf(x); }
inline double sqrt(double x) { return ::sqrt(x); }
inline float fabs(float x) { return ::fabsf(x); }
inline double fabs(double x) { return ::fabs(x); }
inline float hypot(float a, float b) { return ::hypotf(a, b); }
inline double hypot(double a, double b) { return ::hypot(a, b); }
template <class T, std::enable_if_t<std::is_integral_v<T>, int> = 0>
inline T abs(T x) { return x < 0 ? -x : x; }

template <class T>
struct decorated_private_ptr {
  T* p;
  explicit decorated_private_ptr(T* p_) : p{p_} {}
  T* get() const { return p; }
};
template <class T>
using private_ptr = decorated_private_ptr<T>;

inline float sincos(float x, decorated_private_ptr<float> c) {
  *c.p = ::cosf(x);
  return ::sinf(x);
}
inline double sincos(double x, decorated_private_ptr<double> c) {
  *c.p = ::cos(x);
  return ::sin(x);
}
inline float modf(float x, decorated_private_ptr<float> ip) { return ::modff(x, ip.p); }
inline double modf(double x, decorated_private_ptr<double> ip) { return ::modf(x, ip.p); }

template <class T = void>
struct plus {
  T operator()(const T& a, const T& b) const { return a + b; }
};
template <>
struct plus<void> {
  template <class A, class B>
  auto operator()(const A& a, const B& b) const { return a + b; }
};

template <class T>
struct is_device_copyable : std::true_type {};

}  // namespace sycl
