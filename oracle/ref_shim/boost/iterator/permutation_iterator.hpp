#pragma once
#include <cstddef>
#include <iterator>
namespace boost {
namespace iterators {
template <class ElementIterator, class IndexIterator>
struct permutation_iterator {
  ElementIterator e;
  IndexIterator idx;
  permutation_iterator(ElementIterator e_, IndexIterator i_) : e{e_}, idx{i_} {}
  using value_type = typename std::iterator_traits<ElementIterator>::value_type;
  using difference_type = std::ptrdiff_t;
  using reference = value_type&;
  using pointer = value_type*;
  using iterator_category = std::random_access_iterator_tag;
  decltype(auto) operator*() const { return e[*idx]; }
  decltype(auto) operator[](difference_type i) const { return e[idx[i]]; }
};
}  // namespace iterators
using iterators::permutation_iterator;
}  // namespace boost
