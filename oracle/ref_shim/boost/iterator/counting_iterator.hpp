#pragma once
#include <cstddef>
#include <iterator>
namespace boost {
namespace iterators {
template <class T>
struct counting_iterator {
  using value_type = T;
  using difference_type = std::ptrdiff_t;
  using reference = T;
  using pointer = const T*;
  using iterator_category = std::random_access_iterator_tag;
  T v;
  T operator*() const { return v; }
  T operator[](difference_type i) const { return v + static_cast<T>(i); }
  counting_iterator operator+(difference_type i) const { return {v + static_cast<T>(i)}; }
};
}  // namespace iterators
using iterators::counting_iterator;
}  // namespace boost
