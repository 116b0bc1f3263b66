// boost/lockfree/spsc_queue.hpp — declaration-only stand-in (srtb/work.hpp derives work_queue from
// it; the oracle/_ref build never instantiates a queue)
#pragma once
#include <cstddef>
namespace boost {
namespace lockfree {
template <size_t N>
struct capacity {};
template <class T, class... Options>
class spsc_queue {
 public:
  spsc_queue() = default;
  explicit spsc_queue(size_t) {}
  bool push(const T&) { return false; }
  bool pop(T&) { return false; }
  size_t read_available() const { return 0; }
  bool empty() const { return true; }
};
}  // namespace lockfree
}  // namespace boost
