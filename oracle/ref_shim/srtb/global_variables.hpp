// srtb/global_variables.hpp (shim) — the globals every operator header expects:
// srtb::config and host/device allocators with allocate_shared / allocate_unique
// (reference: global_variables.hpp:42-61 over memory/cached_allocator.hpp:75-154; here plain new[]).
#pragma once
#include <map>
#include <memory>
#include <string>

#include "srtb/config.hpp"
#include "srtb/sycl.hpp"
#include "srtb/work.hpp"

namespace srtb {

inline srtb::configs config;
inline std::map<std::string, std::string> changed_configs;

namespace memory {
struct plain_allocator {
  template <class T>
  std::shared_ptr<T> allocate_shared(size_t n) {
    return std::shared_ptr<T>(new T[n ? n : 1](), [](T* p) { delete[] p; });
  }
  template <class T>
  auto allocate_unique(size_t n) {
    auto del = [](T* p) { delete[] p; };
    return std::unique_ptr<T, decltype(del)>(new T[n ? n : 1](), del);
  }
  void deallocate_all_free_ptrs() {}
};
}  // namespace memory

inline memory::plain_allocator host_allocator;
inline memory::plain_allocator device_allocator;

}  // namespace srtb
