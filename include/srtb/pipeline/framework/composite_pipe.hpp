// srtb/pipeline/framework/composite_pipe.hpp — run several pipes back to back on one thread
// (reference: userspace/include/srtb/pipeline/framework/composite_pipe.hpp:29-51; its one in-tree
// user is baseband_receiver.cpp:73-76). Here it is also how stream-ordered fusion is expressed:
// pipes sharing a cuda_queue in fast mode skip the per-stage host wait.
#pragma once
#include <optional>
#include <stop_token>
#include <tuple>
#include <utility>

namespace srtb {
namespace pipeline {

template <typename Pipe1, typename... Pipes>
class composite_pipe {
 public:
  Pipe1 pipe_1;
  composite_pipe<Pipes...> pipes;

  composite_pipe() = default;
  /** every member pipe is constructed from the same argument(s), e.g. the cuda_queue */
  template <typename... Args>
  explicit composite_pipe(Args... args) : pipe_1{args...}, pipes{args...} {}

  template <typename Work>
  auto operator()(std::stop_token st, Work in_work) {
    auto opt = pipe_1(st, std::move(in_work));
    using out_type = decltype(pipes(st, std::move(opt.value())));
    if (!opt) return out_type{};
    return pipes(st, std::move(opt.value()));
  }
};

template <typename Pipe>
class composite_pipe<Pipe> : public Pipe {
 public:
  using Pipe::Pipe;
  composite_pipe() = default;
  template <typename... Args>
  explicit composite_pipe(Args... args) : Pipe{args...} {}
};

}  // namespace pipeline
}  // namespace srtb
