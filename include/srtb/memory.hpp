// srtb/memory.hpp — cached device / pinned-host allocators with the reference's surface
// (reference: memory/cached_allocator.hpp:75-154 allocate_shared / allocate_unique whose deleters
// return the block to a size-keyed cache; global_variables.hpp:47-61 host_allocator /
// device_allocator). Re-hosted on cudaMalloc / cudaMallocHost; blocks are cached per (device, size).
#pragma once
#include <cuda_runtime_api.h>

#include <cstddef>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

namespace srtb {
namespace memory {

enum class space { device, host };

template <space Space>
class cached_allocator {
  struct pool {
    std::mutex m;
    std::multimap<std::pair<int, size_t>, void*> free_blocks;  // (device, bytes) -> ptr
    size_t cached_bytes = 0;
  };
  std::shared_ptr<pool> p_ = std::make_shared<pool>();

  static void* raw_alloc(size_t bytes) {
    void* ptr = nullptr;
    const cudaError_t e = (Space == space::device) ? cudaMalloc(&ptr, bytes) : cudaMallocHost(&ptr, bytes);
    if (e != cudaSuccess) throw std::bad_alloc{};
    return ptr;
  }
  static void raw_free(void* ptr) {
    if (Space == space::device) cudaFree(ptr);
    else cudaFreeHost(ptr);
  }

 public:
  /** bytes are rounded up to 256 so near-equal requests share a bucket */
  void* allocate_bytes(size_t bytes, int& device_out) {
    bytes = (bytes + 255) / 256 * 256;
    int device = 0;
    cudaGetDevice(&device);
    device_out = device;
    {
      std::lock_guard<std::mutex> g{p_->m};
      auto it = p_->free_blocks.find({device, bytes});
      if (it != p_->free_blocks.end()) {
        void* ptr = it->second;
        p_->free_blocks.erase(it);
        p_->cached_bytes -= bytes;
        return ptr;
      }
    }
    return raw_alloc(bytes);
  }

  template <typename T>
  std::shared_ptr<T> allocate_shared(size_t count) {
    const size_t bytes = (count * sizeof(T) + 255) / 256 * 256;
    int device = 0;
    void* ptr = allocate_bytes(bytes, device);
    std::weak_ptr<pool> wp = p_;
    return std::shared_ptr<T>(static_cast<T*>(ptr), [wp, bytes, device](T* q) {
      if (auto p = wp.lock()) {
        std::lock_guard<std::mutex> g{p->m};
        p->free_blocks.emplace(std::make_pair(device, bytes), static_cast<void*>(q));
        p->cached_bytes += bytes;
      } else {
        raw_free(q);
      }
    });
  }

  template <typename T>
  auto allocate_unique(size_t count) {
    auto sp = allocate_shared<T>(count);
    auto del = [sp](T*) mutable { sp.reset(); };
    return std::unique_ptr<T, decltype(del)>(sp.get(), del);
  }

  /** free every cached block (exit_handler.hpp:28-39 calls deallocate_all_free_ptrs) */
  void deallocate_all_free_ptrs() {
    std::lock_guard<std::mutex> g{p_->m};
    for (auto& kv : p_->free_blocks) raw_free(kv.second);
    p_->free_blocks.clear();
    p_->cached_bytes = 0;
  }
  size_t cached_bytes() const { return p_->cached_bytes; }
};

}  // namespace memory

inline memory::cached_allocator<memory::space::host> host_allocator;
inline memory::cached_allocator<memory::space::device> device_allocator;

}  // namespace srtb
