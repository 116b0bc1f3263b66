// srtb/pipeline/rfi_mitigation_pipe.hpp — rfi_mitigation_s1_pipe / rfi_mitigation_s2_pipe
// (reference: userspace/include/srtb/pipeline/rfi_mitigation_pipe.hpp:32-101 and :109-130).
// s1: zap bins above threshold * mean power, normalise the rest by (Nc^2/C)^-1/2, then zero the
// manually listed MHz ranges (list re-parsed only when the config string changes, :84-88).
// s2: spectral kurtosis per channel row.
#pragma once
#include <optional>
#include <stop_token>
#include <string>
#include <utility>
#include <vector>

#include "srtb/config.hpp"
#include "srtb/cuda_queue.hpp"
#include "srtb/log.hpp"
#include "srtb/pipeline/mode.hpp"
#include "srtb/work.hpp"

namespace srtb {
namespace spectrum {
using rfi_range_type = std::pair<srtb::real, srtb::real>;

/** "a-b, c-d" -> MHz pairs (reference: spectrum/rfi_mitigation.hpp:64-88) */
inline std::vector<rfi_range_type> eval_rfi_ranges(const std::string& list) {
  std::vector<float> flat(2 * 64);
  const size_t n = srtb_b200_eval_rfi_ranges(list.c_str(), flat.data(), 64);
  std::vector<rfi_range_type> out;
  for (size_t i = 0; i < n && i < 64; i++) out.emplace_back(flat[2 * i], flat[2 * i + 1]);
  return out;
}
}  // namespace spectrum

namespace pipeline {

class rfi_mitigation_s1_pipe {
 protected:
  srtb::cuda_queue q;
  std::string mitigate_rfi_freq_list;
  std::vector<srtb::spectrum::rfi_range_type> rfi_ranges;

 public:
  explicit rfi_mitigation_s1_pipe(srtb::cuda_queue q_) : q{q_} {}

  auto operator()(std::stop_token, srtb::work::rfi_mitigation_s1_work in_work) {
    const size_t in_count = in_work.count;
    if (srtb::config.mitigate_rfi_freq_list != mitigate_rfi_freq_list) {
      mitigate_rfi_freq_list = srtb::config.mitigate_rfi_freq_list;
      rfi_ranges = srtb::spectrum::eval_rfi_ranges(mitigate_rfi_freq_list);
    }
    std::vector<size_t> bins;
    for (auto [f1, f2] : rfi_ranges) {
      size_t lo = 0, hi = 0;
      if (srtb_b200_rfi_range_to_bins(f1, f2, srtb::config.baseband_freq_low, srtb::config.baseband_bandwidth,
                                      in_count, &lo, &hi)) {
        bins.push_back(lo);
        bins.push_back(hi);
      } else {
        SRTB_LOGW << " [mitigate_rfi_manual] " << "RFI frequency range is out of bounds: " << f1 << " - " << f2
                  << " MHz";
      }
    }
    const float coef = srtb_b200_norm_coefficient(in_count, srtb::config.spectrum_channel_count);
    cuda_check(cudaSetDevice(q.device()), "cudaSetDevice");
    q.check(srtb_b200_rfi_s1(q.ctx(), in_work.ptr.get(), in_count, srtb::config.mitigate_rfi_average_method_threshold,
                             coef, bins.empty() ? nullptr : bins.data(), bins.size() / 2, nullptr));
    end_of_pipe(q);
    srtb::work::dedisperse_work out;
    auto ptr = in_work.ptr;
    out.move_parameter_from(std::move(in_work));
    out.ptr = ptr;
    out.count = in_count;
    out.batch_size = 1;
    return std::optional{out};
  }
};

class rfi_mitigation_s2_pipe {
 public:
  srtb::cuda_queue q;
  explicit rfi_mitigation_s2_pipe(srtb::cuda_queue q_) : q{q_} {}

  auto operator()(std::stop_token, srtb::work::rfi_mitigation_s2_work in_work) {
    const size_t time_sample_count = in_work.count, frequency_bin_count = in_work.batch_size;
    cuda_check(cudaSetDevice(q.device()), "cudaSetDevice");
    q.check(srtb_b200_rfi_s2_sk(q.ctx(), in_work.ptr.get(), time_sample_count, frequency_bin_count,
                                srtb::config.mitigate_rfi_spectral_kurtosis_threshold, nullptr));
    end_of_pipe(q);
    srtb::work::signal_detect_work out;
    auto ptr = in_work.ptr;
    out.move_parameter_from(std::move(in_work));
    out.ptr = ptr;
    out.count = time_sample_count;
    out.batch_size = frequency_bin_count;
    return std::optional{out};
  }
};

}  // namespace pipeline
}  // namespace srtb
