// srtb/pipeline/framework/dummy_pipe.hpp — a pipe that swallows its input
// (reference: userspace/include/srtb/pipeline/framework/dummy_pipe.hpp)
#pragma once
#include <optional>
#include <stop_token>

#include "srtb/work.hpp"

namespace srtb {
namespace pipeline {

template <typename InWork = srtb::work::dummy_work>
class dummy_pipe {
 public:
  dummy_pipe() = default;
  template <typename... Args>
  explicit dummy_pipe(Args...) {}
  std::optional<srtb::work::dummy_work> operator()(std::stop_token, InWork) {
    return srtb::work::dummy_work{};
  }
};

}  // namespace pipeline
}  // namespace srtb
