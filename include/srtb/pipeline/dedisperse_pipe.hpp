// srtb/pipeline/dedisperse_pipe.hpp — coherent dedispersion pipe
// (reference: userspace/include/srtb/pipeline/dedisperse_pipe.hpp:27-48): df = bandwidth / N,
// f_min = freq_low, f_c = freq_low + bandwidth, all float, read from config per call.
#pragma once
#include <optional>
#include <stop_token>

#include "srtb/config.hpp"
#include "srtb/cuda_queue.hpp"
#include "srtb/pipeline/mode.hpp"
#include "srtb/work.hpp"

namespace srtb {
namespace coherent_dedispersion {
/** reference: coherent_dedispersion.hpp:103-128 (reads the global config; may clear reserve_sample) */
inline size_t nsamps_reserved() {
  const size_t r = srtb_b200_nsamps_reserved(
      srtb::config.baseband_input_count, srtb::config.spectrum_channel_count, srtb::config.baseband_freq_low,
      srtb::config.baseband_bandwidth, srtb::config.baseband_sample_rate, srtb::config.dm,
      srtb::config.baseband_reserve_sample ? 1 : 0);
  if (r == 0 && srtb::config.baseband_reserve_sample) srtb::config.baseband_reserve_sample = false;
  return r;
}
}  // namespace coherent_dedispersion
namespace codd = coherent_dedispersion;

namespace pipeline {

class dedisperse_pipe {
 public:
  srtb::cuda_queue q;
  explicit dedisperse_pipe(srtb::cuda_queue q_) : q{q_} {}

  auto operator()(std::stop_token, srtb::work::dedisperse_work in_work) {
    const size_t N = in_work.count;
    const srtb::real df = srtb::config.baseband_bandwidth / N;
    const srtb::real f_min = srtb::config.baseband_freq_low, f_c = f_min + srtb::config.baseband_bandwidth;
    const srtb::real dm = srtb::config.dm;
    cuda_check(cudaSetDevice(q.device()), "cudaSetDevice");
    q.check(srtb_b200_dedisperse(q.ctx(), in_work.ptr.get(), N, f_min, f_c, df, dm));
    end_of_pipe(q);
    srtb::work::watfft_1d_c2c_work out;
    auto ptr = in_work.ptr;
    out.move_parameter_from(std::move(in_work));
    out.ptr = ptr;
    out.count = N;
    out.batch_size = 1;
    return std::optional{out};
  }
};

}  // namespace pipeline
}  // namespace srtb
