#pragma once
#include <cstddef>
#include <iterator>
#include <type_traits>
namespace boost {
namespace iterators {
template <class F, class It>
struct transform_iterator {
  It it;
  F f;
  transform_iterator(It i, F fn) : it{i}, f{fn} {}
  using value_type = std::remove_cv_t<std::remove_reference_t<decltype(std::declval<F>()(*std::declval<It>()))>>;
  using difference_type = std::ptrdiff_t;
  using reference = value_type;
  using pointer = const value_type*;
  using iterator_category = std::random_access_iterator_tag;
  value_type operator*() const { return f(*it); }
  value_type operator[](difference_type i) const { return f(it[i]); }
};
template <class It, class F>
transform_iterator(It, F) -> transform_iterator<F, It>;
}  // namespace iterators
using iterators::transform_iterator;
}  // namespace boost
