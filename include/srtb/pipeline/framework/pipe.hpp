// srtb/pipeline/framework/pipe.hpp — thread-per-pipe runner with the reference's contract
// (reference: userspace/include/srtb/pipeline/framework/pipe.hpp:108-142 pipe::run,
//  :148-175 start_pipe): the functor is constructed ON its own thread from `args...`, the loop is
// in_functor -> pipe_functor -> out_functor, an empty optional or a stop request ends the thread,
// and start_pipe returns once the functor exists.
#pragma once
#include <atomic>
#include <memory>
#include <optional>
#include <stop_token>
#include <string>
#include <thread>
#include <typeinfo>

#if __has_include(<pthread.h>)
#include <pthread.h>
#endif
#if __has_include(<cxxabi.h>)
#include <cxxabi.h>
#endif

#include "srtb/log.hpp"

namespace srtb {
namespace pipeline {

inline namespace detail {

/** unqualified class name without template arguments, e.g. "unpack_pipe" */
template <typename Type>
inline std::string class_name() {
  std::string full = typeid(Type).name();
#if __has_include(<cxxabi.h>)
  int status = 0;
  if (char* dem = abi::__cxa_demangle(full.c_str(), nullptr, nullptr, &status)) {
    if (status == 0) full = dem;
    std::free(dem);
  }
#endif
  const size_t lt = full.find('<');
  const std::string head = full.substr(0, lt);
  const size_t ns = head.rfind("::");
  return (ns == std::string::npos) ? head : head.substr(ns + 2);
}

/** pthread names are limited to 15 characters */
template <typename Type>
inline std::string generate_thread_name() {
  return class_name<Type>().substr(0, 15);
}

inline bool is_running(const std::stop_token& st) { return !st.stop_possible() || !st.stop_requested(); }

}  // namespace detail

template <typename PipeFunctor, typename InFunctor, typename OutFunctor>
class pipe {
 public:
  PipeFunctor pipe_functor;
  InFunctor in_functor;
  OutFunctor out_functor;

  void run(std::stop_token stop_token) {
    const std::string tag = " [" + class_name<PipeFunctor>() + "] ";
    SRTB_LOGD << tag << "starting";
    while (is_running(stop_token)) {
      auto opt_in = in_functor(stop_token);
      if (!is_running(stop_token) || !opt_in) break;
      SRTB_LOGD << tag << "got work";
      auto opt_out = pipe_functor(stop_token, std::move(opt_in.value()));
      if (!is_running(stop_token) || !opt_out) break;
      out_functor(stop_token, std::move(opt_out.value()));
      SRTB_LOGD << tag << "work finished";
    }
    SRTB_LOGD << tag << "stopped";
  }
};

template <typename PipeFunctor, typename InFunctor, typename OutFunctor, typename... Args>
static std::jthread start_pipe(InFunctor in_functor, OutFunctor out_functor, Args... args) {
  auto ready = std::make_shared<std::atomic<int>>(0);  // 0 = constructing, 1 = ok, -1 = failed
  // the arguments are moved into the thread and from there into the functor, which is constructed ON the pipe's own
  // thread (pipe.hpp:148-161 of the reference): move-only arguments (a socket-owning packet provider) work
  std::jthread thread{[ready, in_functor, out_functor](std::stop_token st, Args... a) mutable {
                        try {
                          pipe<PipeFunctor, InFunctor, OutFunctor> p{PipeFunctor{std::move(a)...}, in_functor, out_functor};
                          ready->store(1);
                          p.run(st);
                        } catch (const std::exception& e) {
                          SRTB_LOGE << " [" << class_name<PipeFunctor>() << "] " << "exception: " << e.what();
                          ready->store(-1);
                          throw;  // uncaught in a pipe thread terminates, as in the reference
                        }
                      },
                      std::move(args)...};
#if __has_include(<pthread.h>)
  pthread_setname_np(thread.native_handle(), generate_thread_name<PipeFunctor>().c_str());
#endif
  while (ready->load() == 0) std::this_thread::yield();
  return thread;
}

}  // namespace pipeline
}  // namespace srtb
