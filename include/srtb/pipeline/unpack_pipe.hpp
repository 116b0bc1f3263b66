// srtb/pipeline/unpack_pipe.hpp — the unpack pipes, same class names and work semantics as the
// reference (userspace/include/srtb/pipeline/unpack_pipe.hpp:33-136 unpack_pipe, :145-259
// unpack_interleaved_samples_2_pipe, :262-325 unpack_gznupsr_a1_pipe, :327-390
// unpack_gznupsr_a1_v2_1_pipe, :392-413 start_unpack_pipe). unpack_work.count is in BYTES;
// every output buffer is over-allocated by 2 floats for the in-place R2C (:65-67); fan-out pipes
// return std::array<fft_1d_r2c_work, S> with data_stream_id = S * id + s.
#pragma once
#include <array>
#include <cstdlib>
#include <optional>
#include <stdexcept>
#include <stop_token>
#include <string>
#include <string_view>

#include "srtb/config.hpp"
#include "srtb/cuda_queue.hpp"
#include "srtb/memory.hpp"
#include "srtb/pipeline/framework/pipe.hpp"
#include "srtb/pipeline/framework/pipe_io.hpp"
#include "srtb/pipeline/mode.hpp"
#include "srtb/work.hpp"

namespace srtb {
namespace pipeline {

inline namespace detail {

/** S output streams through one srtb_b200_unpack call */
template <size_t S>
inline auto unpack_to_streams(const srtb::cuda_queue& q, srtb::work::unpack_work& in_work, int bits,
                              int format, size_t out_count) {
  std::array<std::shared_ptr<srtb::real>, S> outs;
  float* raw[4] = {nullptr, nullptr, nullptr, nullptr};
  cuda_check(cudaSetDevice(q.device()), "cudaSetDevice");
  for (size_t i = 0; i < S; i++) {
    outs[i] = srtb::device_allocator.allocate_shared<srtb::real>(out_count + 2);  // +2: in-place R2C
    raw[i] = outs[i].get();
  }
  q.check(srtb_b200_unpack(q.ctx(), in_work.ptr.get(), in_work.count, bits, format,
                           SRTB_B200_WINDOW_RECTANGLE /* default_window, fft_window.hpp:83 */, raw, out_count));
  end_of_pipe(q);
  in_work.ptr.reset();
  std::array<srtb::work::fft_1d_r2c_work, S> works;
  for (size_t i = 0; i < S; i++) {
    works[i].copy_parameter_from(in_work);
    works[i].data_stream_id = static_cast<uint32_t>(S * in_work.data_stream_id + i);
    works[i].ptr = outs[i];
    works[i].count = out_count;
    works[i].batch_size = 1;
  }
  return works;
}

}  // namespace detail

/** baseband_format_type = "simple" / "fastmb_roach2": one stream in, one stream out */
class unpack_pipe {
 protected:
  srtb::cuda_queue q;

 public:
  explicit unpack_pipe(srtb::cuda_queue q_) : q{q_} {}

  auto operator()(std::stop_token, srtb::work::unpack_work unpack_work) {
    const int bits = srtb::config.baseband_input_bits;
    if (bits == 0) throw std::runtime_error("[unpack pipe] unsupported baseband_input_bits = 0");
    const size_t out_count = unpack_work.count * srtb::BITS_PER_BYTE / static_cast<size_t>(std::abs(bits));
    auto works = unpack_to_streams<1>(q, unpack_work, bits, SRTB_B200_FORMAT_SIMPLE, out_count);
    works[0].data_stream_id = unpack_work.data_stream_id;  // 1 -> 1: id unchanged (:131-135)
    return std::optional{works[0]};
  }
};

/** "interleaved_samples_2" and "naocpsr_snap1": 2 polarisations in one block */
class unpack_interleaved_samples_2_pipe {
 public:
  using in_work_type = srtb::work::unpack_work;
  using out_work_type = srtb::work::fft_1d_r2c_work;

 protected:
  srtb::cuda_queue q;

 public:
  explicit unpack_interleaved_samples_2_pipe(srtb::cuda_queue q_) : q{q_} {}

  auto operator()(std::stop_token, srtb::work::unpack_work unpack_work) {
    const int bits = srtb::config.baseband_input_bits;
    if (bits == 0) throw std::runtime_error("[unpack_2pol_interleave_pipe] unsupported baseband_input_bits = 0");
    const size_t out_count =
        unpack_work.count * srtb::BITS_PER_BYTE / static_cast<size_t>(std::abs(bits)) / 2;
    const bool snap1 = srtb::config.baseband_format_type.find("naocpsr_snap1") != std::string::npos;
    const int format = (snap1 && bits == -8) ? SRTB_B200_FORMAT_NAOCPSR_SNAP1 : SRTB_B200_FORMAT_INTERLEAVED_2;
    return std::optional{unpack_to_streams<2>(q, unpack_work, bits, format, out_count)};
  }
};

/** gznupsr_a1, 4 output streams (defined in the reference but not started) */
class unpack_gznupsr_a1_pipe {
 public:
  using in_work_type = srtb::work::unpack_work;
  using out_work_type = srtb::work::fft_1d_r2c_work;

 protected:
  srtb::cuda_queue q;

 public:
  explicit unpack_gznupsr_a1_pipe(srtb::cuda_queue q_) : q{q_} {}
  auto operator()(std::stop_token, srtb::work::unpack_work unpack_work) {
    const size_t out_count = unpack_work.count / 4;  // 8-bit, 4 streams
    return std::optional{unpack_to_streams<4>(q, unpack_work, 8, SRTB_B200_FORMAT_GZNUPSR_A1_4, out_count)};
  }
};

/** gznupsr_a1 v2.1, 2 output streams — the one start_unpack_pipe selects */
class unpack_gznupsr_a1_v2_1_pipe {
 public:
  using in_work_type = srtb::work::unpack_work;
  using out_work_type = srtb::work::fft_1d_r2c_work;

 protected:
  srtb::cuda_queue q;

 public:
  explicit unpack_gznupsr_a1_v2_1_pipe(srtb::cuda_queue q_) : q{q_} {}
  auto operator()(std::stop_token, srtb::work::unpack_work unpack_work) {
    const size_t out_count = unpack_work.count / 2;  // 8-bit, 2 streams
    return std::optional{unpack_to_streams<2>(q, unpack_work, 8, SRTB_B200_FORMAT_GZNUPSR_A1_2, out_count)};
  }
};

/** format names and aliases of io/backend_registry.hpp:36-181 */
inline std::string_view resolve_format_alias(std::string_view name) {
  if (name == "interleaved_samples_2") return "naocpsr_snap1";  // same unpack pipe, layout picked from bits/name
  if (name == "naocpsr_roach2") return "fastmb_roach2";
  return name;
}

template <typename InFunctor, typename OutFunctor, typename... Args>
inline auto start_unpack_pipe(std::string_view format_name, InFunctor in_functor, OutFunctor out_functor,
                              Args... args) {
  format_name = resolve_format_alias(format_name);
  if (format_name == "simple" || format_name == "fastmb_roach2")
    return start_pipe<unpack_pipe>(in_functor, out_functor, args...);
  if (format_name == "naocpsr_snap1")
    return start_pipe<unpack_interleaved_samples_2_pipe>(in_functor, multiple_works_out_functor{out_functor}, args...);
  if (format_name == "gznupsr_a1")
    return start_pipe<unpack_gznupsr_a1_v2_1_pipe>(in_functor, multiple_works_out_functor{out_functor}, args...);
  throw std::invalid_argument{"[start_unpack_pipe] Unknown format name: " + std::string{format_name}};
}

}  // namespace pipeline
}  // namespace srtb
