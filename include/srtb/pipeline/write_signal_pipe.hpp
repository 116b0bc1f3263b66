// srtb/pipeline/write_signal_pipe.hpp — candidate sink
// (reference: userspace/include/srtb/pipeline/write_signal_pipe.hpp:77-284): a work is kept iff it
// carries at least one time series (file mode), or — in real-time mode (no input_file_path) — also if
// its timestamp lies within 0.45 block lengths of a recent positive (:96-140, so the other
// polarisation / neighbouring blocks of a detection are kept). For a kept work it writes
//   <prefix><counter>.bin            raw baseband block                                (:159-206)
//   <prefix><counter>.<i>.npy        complex64 dynamic spectrum, shape {batch_size, count} (:209-246)
//   <prefix><counter>.<boxcar>.tim   float32 time series per detected boxcar           (:249-280)
// counter = udp_packet_counter, or the timestamp when there is none (:145-148).
// Differences: files are written synchronously on the pipe's thread (the reference posts to thread
// pools), and the .npy index is incremented until the name is free (the reference's loop never
// increments `i`, :232-236).
#pragma once
#include <complex>
#include <cstdint>
#include <deque>
#include <filesystem>
#include <fstream>
#include <optional>
#include <stop_token>
#include <string>
#include <vector>

#include "srtb/config.hpp"
#include "srtb/cuda_queue.hpp"
#include "srtb/io/npy.hpp"
#include "srtb/log.hpp"
#include "srtb/work.hpp"

namespace srtb {
namespace pipeline {

class write_signal_pipe {
 protected:
  srtb::cuda_queue q;
  std::deque<uint64_t> recent_positive_timestamps;
  std::deque<srtb::work::write_signal_work> recent_negative_works;
  size_t written_ = 0;

 public:
  explicit write_signal_pipe(srtb::cuda_queue q_) : q{q_} {
    const std::string check = srtb::config.baseband_output_file_prefix + ".check";
    std::ofstream f(check, std::ios::binary);
    if (!f) {
      SRTB_LOGE << " [write_signal_pipe] " << "cannot open file " << check;
      throw std::runtime_error("[write_signal_pipe] cannot open file " + check);
    }
  }
  size_t written() const { return written_; }

  /** the keep/drop decision alone (host logic; testable without a GPU) */
  std::optional<srtb::work::write_signal_work> select(srtb::work::write_signal_work work) {
    std::optional<srtb::work::write_signal_work> keep;
    const bool has_signal = !work.time_series.empty();
    const bool real_time = srtb::config.input_file_path.empty();
    const double window = 0.45 * 1e9 * (double)srtb::config.baseband_input_count / srtb::config.baseband_sample_rate;
    auto near_positive = [&](uint64_t ts) {
      for (auto t : recent_positive_timestamps)
        if (std::abs(static_cast<double>(static_cast<int64_t>(ts - t))) < window) return true;
      return false;
    };
    while (real_time && !recent_positive_timestamps.empty() &&
           static_cast<int64_t>(work.timestamp - recent_positive_timestamps.front()) > 5 * window)
      recent_positive_timestamps.pop_front();
    if (has_signal) {
      recent_positive_timestamps.push_back(work.timestamp);
      keep = std::move(work);
    } else if (real_time) {
      if (near_positive(work.timestamp)) keep = std::move(work);
      else recent_negative_works.push_back(std::move(work));
    }
    if (real_time && !keep && !recent_negative_works.empty()) {
      auto w2 = std::move(recent_negative_works.front());
      recent_negative_works.pop_front();
      if (near_positive(w2.timestamp)) keep = std::move(w2);
    }
    return keep;
  }

  auto operator()(std::stop_token, srtb::work::write_signal_work work) {
    if (auto keep = select(std::move(work))) write(*keep);
    return std::optional{srtb::work::dummy_work{}};
  }

  void write(const srtb::work::write_signal_work& w) {
    uint64_t counter = w.udp_packet_counter;
    if (counter == w.no_udp_packet_counter) counter = w.timestamp;
    const std::string stem = srtb::config.baseband_output_file_prefix + std::to_string(counter);
    SRTB_LOGI << " [write_signal_pipe] " << "Begin writing baseband data, file_counter = " << counter;
    if (w.baseband_data.baseband_ptr) {
      std::ofstream f(stem + ".bin", std::ios::binary | std::ios::trunc);
      f.write(reinterpret_cast<const char*>(w.baseband_data.baseband_ptr.get()),
              (std::streamsize)w.baseband_data.baseband_input_bytes);
    } else {
      SRTB_LOGE << " [write_signal_pipe] " << "baseband pointer not valid!";
    }
    if (w.ptr) {
      const size_t total = w.count * w.batch_size;
      std::vector<std::complex<srtb::real>> h(total);
      cuda_check(cudaSetDevice(q.device()), "cudaSetDevice");
      cuda_check(cudaMemcpyAsync(h.data(), w.ptr.get(), total * sizeof(h[0]), cudaMemcpyDeviceToHost, q.stream()),
                 "D2H spectrum");
      q.wait();
      size_t i = 0;
      std::string path;
      do {
        path = stem + "." + std::to_string(i++) + ".npy";
      } while (std::filesystem::exists(path));
      srtb::io::npy_save(path, h.data(), std::vector<size_t>{w.batch_size, w.count});
    }
    for (const auto& ts : w.time_series) {
      std::ofstream f(stem + "." + std::to_string(ts.boxcar_length) + ".tim", std::ios::binary | std::ios::trunc);
      f.write(reinterpret_cast<const char*>(ts.h_time_series.get()), (std::streamsize)(ts.time_series_length * sizeof(srtb::real)));
    }
    written_++;
  }
};

}  // namespace pipeline
}  // namespace srtb
