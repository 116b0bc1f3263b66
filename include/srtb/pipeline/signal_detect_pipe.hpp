// srtb/pipeline/signal_detect_pipe.hpp — signal_detect_pipe_2 (and the v1 alternate, signal_detect_pipe, at the end)
// (reference: userspace/include/srtb/pipeline/signal_detect_pipe.hpp:244-443): zero-channel count,
// time series, baseline removal, count_signal on the series and on boxcars 2,4,..; every series with
// at least one sample over threshold is handed on as a host time_series_holder. The holders carry
// the INTENDED series for their boxcar, already copied when the pipe returns.
#pragma once
#include <cstring>
#include <optional>
#include <stop_token>
#include <vector>

#include "srtb/config.hpp"
#include "srtb/cuda_queue.hpp"
#include "srtb/log.hpp"
#include "srtb/memory.hpp"
#include "srtb/pipeline/dedisperse_pipe.hpp"
#include "srtb/work.hpp"

namespace srtb {
namespace pipeline {

class signal_detect_pipe_2 {
 protected:
  srtb::cuda_queue q;

 public:
  explicit signal_detect_pipe_2(srtb::cuda_queue q_) : q{q_} {}

  auto operator()(std::stop_token, srtb::work::signal_detect_work in_work) {
    const size_t time_sample_count = in_work.count, frequency_bin_count = in_work.batch_size;
    const size_t time_reserved_count = srtb::codd::nsamps_reserved() / frequency_bin_count;
    if (time_sample_count <= time_reserved_count)
      SRTB_LOGW << " [signal_detect_pipe_2] " << "time_sample_count = " << time_sample_count
                << " <= time_reserved_count = " << time_reserved_count;
    cuda_check(cudaSetDevice(q.device()), "cudaSetDevice");
    auto h_all = srtb::host_allocator.allocate_shared<srtb::real>(SRTB_B200_MAX_BOXCARS * time_sample_count);
    srtb_b200_detect_result res;
    q.check(srtb_b200_signal_detect(q.ctx(), in_work.ptr.get(), time_sample_count, frequency_bin_count,
                                    time_reserved_count, srtb::config.signal_detect_signal_noise_threshold,
                                    srtb::config.signal_detect_channel_threshold,
                                    srtb::config.signal_detect_max_boxcar_length, &res, h_all.get(), 0));
    SRTB_LOGD << " [signal_detect_pipe_2] " << "zero_count = " << res.zero_count;

    srtb::work::write_signal_work out;
    auto ptr = in_work.ptr;
    out.move_parameter_from(std::move(in_work));
    out.ptr = ptr;
    out.count = time_sample_count;
    out.batch_size = frequency_bin_count;
    out.zero_count = res.zero_count;
    for (int b = 0; b < res.n_boxcars; b++) {
      if (res.signal_count[b] == 0) continue;
      srtb::work::time_series_holder holder;
      holder.time_series_length = res.series_length[b];
      holder.boxcar_length = res.boxcar_length[b];
      holder.signal_count = res.signal_count[b];
      // alias into the pinned block: one allocation per work, rows share ownership
      holder.h_time_series = std::shared_ptr<srtb::real>(h_all, h_all.get() + (size_t)b * time_sample_count);
      out.time_series.push_back(holder);
    }
    if (!out.time_series.empty())
      SRTB_LOGI << " [signal_detect_pipe_2] " << " signal detected in " << out.time_series.size() << " time series";
    else
      SRTB_LOGD << " [signal_detect_pipe_2] " << "no signal detected";
    return std::optional{out};
  }
};

/** signal_detect_pipe (v1; reference: pipeline/signal_detect_pipe.hpp:51-230, defined but not wired in main.cpp):
 *  takes the refft path's spectra [batch_size = time][count = frequency], applies spectral kurtosis v1 in place
 *  (spectrum/rfi_mitigation.hpp:181-275), counts masked channels over the first spectrum, sums every spectrum to one
 *  time-series value, removes the baseline and runs count_signal on the series and its boxcars. */
class signal_detect_pipe {
 protected:
  srtb::cuda_queue q;

 public:
  explicit signal_detect_pipe(srtb::cuda_queue q_) : q{q_} {}

  auto operator()(std::stop_token, srtb::work::signal_detect_work in_work) {
    const size_t count_per_batch = in_work.count, batch_size = in_work.batch_size;
    cuda_check(cudaSetDevice(q.device()), "cudaSetDevice");
    auto h_all = srtb::host_allocator.allocate_shared<srtb::real>(SRTB_B200_MAX_BOXCARS * batch_size);
    srtb_b200_detect_result res;
    q.check(srtb_b200_signal_detect_v1(q.ctx(), in_work.ptr.get(), count_per_batch, batch_size,
                                       srtb::config.mitigate_rfi_spectral_kurtosis_threshold,
                                       srtb::config.signal_detect_signal_noise_threshold,
                                       srtb::config.signal_detect_channel_threshold,
                                       srtb::config.signal_detect_max_boxcar_length, &res, h_all.get(), 0));
    srtb::work::write_signal_work out;
    auto ptr = in_work.ptr;
    out.move_parameter_from(std::move(in_work));
    out.ptr = ptr;
    out.count = count_per_batch;
    out.batch_size = batch_size;
    out.zero_count = res.zero_count;
    for (int b = 0; b < res.n_boxcars; b++) {
      if (res.signal_count[b] == 0) continue;
      srtb::work::time_series_holder holder;
      holder.time_series_length = res.series_length[b];
      holder.boxcar_length = res.boxcar_length[b];
      holder.signal_count = res.signal_count[b];
      holder.h_time_series = std::shared_ptr<srtb::real>(h_all, h_all.get() + (size_t)b * batch_size);
      out.time_series.push_back(holder);
    }
    if (!out.time_series.empty())
      SRTB_LOGI << " [signal_detect_pipe] " << " signal detected in " << out.time_series.size() << " time series";
    else
      SRTB_LOGD << " [signal_detect_pipe] " << "no signal detected";
    return std::optional{out};
  }
};

}  // namespace pipeline
}  // namespace srtb
