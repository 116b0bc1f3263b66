// boost/algorithm/string.hpp — boost::split / is_any_of / token_compress_on as srtb uses them
// (spectrum/rfi_mitigation.hpp:69-74). Semantics of boost::split with token_compress_on:
// adjacent separators are merged; leading/trailing separators still delimit an empty token.
#pragma once
#include <string>
#include <vector>

namespace boost {
enum token_compress_mode_type { token_compress_on, token_compress_off };
struct is_any_of_pred {
  std::string set;
  bool operator()(char c) const { return set.find(c) != std::string::npos; }
};
inline is_any_of_pred is_any_of(const std::string& s) { return {s}; }

template <class Container, class Pred>
inline Container& split(Container& out, const std::string& in, Pred pred,
                        token_compress_mode_type mode = token_compress_off) {
  out.clear();
  std::string cur;
  bool prev_sep = false;
  for (char c : in) {
    if (pred(c)) {
      if (!(mode == token_compress_on && prev_sep)) {
        out.push_back(cur);
        cur.clear();
      }
      prev_sep = true;
    } else {
      cur.push_back(c);
      prev_sep = false;
    }
  }
  out.push_back(cur);
  return out;
}
}  // namespace boost
