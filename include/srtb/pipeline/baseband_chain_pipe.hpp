// srtb/pipeline/baseband_chain_pipe.hpp — the whole device chain as ONE pipe (the fused
// composites behind the pipe API). Stands where main.cpp:170-204 chains copy_to_device -> unpack -> fft_1d_r2c ->
// rfi_mitigation_s1 -> dedisperse -> watfft_1d_c2c -> rfi_mitigation_s2 -> signal_detect_pipe_2: takes the
// copy_to_device_work a source pipe produced (pinned host block) and returns one write_signal_work per data stream,
// the same objects signal_detect_pipe_2 hands to write_signal_pipe. Inside it is one
// srtb_b200_submit_block_ex / collect_block_ex pair per block, i.e. the fused kernels (unpack in the first FFT
// sweep, R2C split + power sum in the last, s1 + chirp + waterfall FFT + SK + column sums in one kernel), one H2D,
// one 1 KiB D2H per stream, and the positive boxcar series written to pinned host memory by the detector kernel.
// Throughput across blocks: start several of these pipes on their own cuda_queue (one CUDA stream + context each)
// popping from one MPMC work queue — blocks then alternate over contexts and overlap on the GPU.
#pragma once
#include <cstdlib>
#include <cstring>
#include <deque>
#include <optional>
#include <stop_token>
#include <string>
#include <vector>

#include "srtb/config.hpp"
#include "srtb/cuda_queue.hpp"
#include "srtb/log.hpp"
#include "srtb/memory.hpp"
#include "srtb/pipeline/rfi_mitigation_pipe.hpp"  // srtb::spectrum::eval_rfi_ranges
#include "srtb/pipeline/unpack_pipe.hpp"          // resolve_format_alias
#include "srtb/work.hpp"

namespace srtb {
namespace pipeline {

/** the subset of srtb::config the device path reads, as the C ABI wants it (config.hpp:80-249) */
struct block_config_holder {
  srtb_b200_block_config cfg{};
  std::vector<float> rfi_pairs;

  static int format_of(std::string_view name, int bits) {
    name = resolve_format_alias(name);
    if (name == "simple" || name == "fastmb_roach2") return SRTB_B200_FORMAT_SIMPLE;  // one stream (unpack_pipe.hpp:396)
    if (name == "interleaved_samples_2") return SRTB_B200_FORMAT_INTERLEAVED_2;
    if (name == "naocpsr_snap1") return bits == -8 ? SRTB_B200_FORMAT_NAOCPSR_SNAP1 : SRTB_B200_FORMAT_INTERLEAVED_2;
    if (name == "gznupsr_a1") return SRTB_B200_FORMAT_GZNUPSR_A1_2;
    if (name == "gznupsr_a1_4") return SRTB_B200_FORMAT_GZNUPSR_A1_4;
    throw std::invalid_argument("[start_unpack_pipe] Unknown format name: " + std::string{name});
  }
  static int stream_count(int format) {
    return format == SRTB_B200_FORMAT_SIMPLE ? 1 : (format == SRTB_B200_FORMAT_GZNUPSR_A1_4 ? 4 : 2);
  }

  /** re-read srtb::config (the reference's pipes read it on every call) */
  void refresh() {
    const auto& c = srtb::config;
    cfg.baseband_input_count = c.baseband_input_count;
    cfg.baseband_input_bits = c.baseband_input_bits;
    cfg.baseband_format = format_of(c.baseband_format_type, c.baseband_input_bits);
    cfg.window = SRTB_B200_WINDOW_RECTANGLE;  // default_window, fft_window.hpp:83
    cfg.baseband_reserve_sample = c.baseband_reserve_sample ? 1 : 0;
    cfg.baseband_freq_low = static_cast<float>(c.baseband_freq_low);
    cfg.baseband_bandwidth = static_cast<float>(c.baseband_bandwidth);
    cfg.baseband_sample_rate = static_cast<float>(c.baseband_sample_rate);
    cfg.dm = static_cast<float>(c.dm);
    cfg.mitigate_rfi_average_method_threshold = static_cast<float>(c.mitigate_rfi_average_method_threshold);
    cfg.mitigate_rfi_spectral_kurtosis_threshold = static_cast<float>(c.mitigate_rfi_spectral_kurtosis_threshold);
    cfg.spectrum_channel_count = c.spectrum_channel_count;
    cfg.signal_detect_signal_noise_threshold = static_cast<float>(c.signal_detect_signal_noise_threshold);
    cfg.signal_detect_channel_threshold = static_cast<float>(c.signal_detect_channel_threshold);
    cfg.signal_detect_max_boxcar_length = c.signal_detect_max_boxcar_length;
    rfi_pairs.clear();
    for (const auto& r : srtb::spectrum::eval_rfi_ranges(c.mitigate_rfi_freq_list)) {
      rfi_pairs.push_back(r.first);
      rfi_pairs.push_back(r.second);
    }
    cfg.rfi_freq_pairs = rfi_pairs.data();
    cfg.n_rfi_freq_pairs = rfi_pairs.size() / 2;
  }
};

class baseband_chain_pipe {
 protected:
  srtb::cuda_queue q;
  block_config_holder holder;
  struct pending_block {
    int ticket;
    int streams;
    size_t C, L;
    srtb::work::copy_to_device_work work;
    // what the block leaves behind, owned by the work items it turns into (zero copy, like the reference's d_in):
    std::vector<std::shared_ptr<srtb::complex<srtb::real>>> d_spectrum;  // per stream: N + 2 floats, [C][L] when done
    std::shared_ptr<srtb::real> h_series;                                 // pinned [streams][MAX_BOXCARS][L]
  };
  std::deque<pending_block> pending;  // blocks in the pinned-host ring of this queue's context

 public:
  /** ring_depth blocks are kept in flight through srtb_b200_submit_block_ex / collect_block_ex (block k's H2D
   *  overlaps block k-1's compute; results lag the input by ring_depth - 1 works and an idle tick from
   *  idle_queue_in_functor flushes). ring_depth <= 1: every work is collected before operator() returns.
   *  Every write_signal_work carries its dynamic spectrum (work.ptr) and the host series of its positive boxcars,
   *  exactly what signal_detect_pipe_2 forwards (signal_detect_pipe.hpp:347-366,405-441): write_signal_pipe's
   *  "neighbour of a positive" rule (write_signal_pipe.hpp:102-115) therefore sees a spectrum on negatives too.
   *  keep_every_spectrum is kept for source compatibility; spectra are always kept now. */
  explicit baseband_chain_pipe(srtb::cuda_queue q_, bool keep_every_spectrum_ = false, int ring_depth_ = 1)
      : q{q_}, keep_every_spectrum{keep_every_spectrum_},
        ring_depth{std::max(1, std::min(ring_depth_, (int)SRTB_B200_RING_SLOTS))} {}

  using out_type = std::vector<srtb::work::write_signal_work>;

  std::optional<out_type> operator()(std::stop_token, srtb::work::copy_to_device_work in_work) {
    cuda_check(cudaSetDevice(q.device()), "cudaSetDevice");
    const bool idle_tick = (in_work.count == 0 && !in_work.baseband_data.baseband_ptr);
    out_type out;
    if (!idle_tick) submit(std::move(in_work));
    while (!pending.empty() && (idle_tick || (int)pending.size() >= ring_depth)) collect_front(out);
    return std::optional{std::move(out)};
  }

  bool keep_every_spectrum = false;
  int ring_depth = 1;

 protected:
  void submit(srtb::work::copy_to_device_work in_work) {
    holder.refresh();
    const auto& cfg = holder.cfg;
    pending_block p;
    p.streams = block_config_holder::stream_count(cfg.baseband_format);
    const size_t N = cfg.baseband_input_count, Nc = N / 2;
    p.C = std::min<size_t>(cfg.spectrum_channel_count, Nc);
    p.L = p.C ? Nc / p.C : 0;
    srtb_b200_block_outputs outputs{};
    for (int s = 0; s < p.streams; s++) {
      p.d_spectrum.push_back(srtb::device_allocator.allocate_shared<srtb::complex<srtb::real>>(Nc + 1));
      outputs.d_spectrum[s] = reinterpret_cast<float*>(p.d_spectrum.back().get());
    }
    p.h_series = srtb::host_allocator.allocate_shared<srtb::real>((size_t)p.streams * SRTB_B200_MAX_BOXCARS * p.L);
    outputs.h_series = p.h_series.get();
    p.ticket = srtb_b200_submit_block_ex(q.ctx(), &cfg, in_work.baseband_data.baseband_ptr.get(),
                                         in_work.baseband_data.baseband_input_bytes, /*on_device=*/0, &outputs);
    q.check(p.ticket);
    p.work = std::move(in_work);
    pending.push_back(std::move(p));
  }

  void collect_front(out_type& out) {
    pending_block p = std::move(pending.front());
    pending.pop_front();
    srtb_b200_detect_result res[4];
    const int n = srtb_b200_collect_block_ex(q.ctx(), p.ticket, res, nullptr, nullptr);
    q.check(n);
    for (int s = 0; s < n; s++) {
      srtb::work::write_signal_work w;
      w.copy_parameter_from(p.work);
      w.data_stream_id = p.work.data_stream_id * static_cast<uint32_t>(p.streams) + static_cast<uint32_t>(s);
      w.count = p.L;
      w.batch_size = p.C;
      w.zero_count = res[s].zero_count;
      w.ptr = p.d_spectrum[s];
      srtb::real* base = p.h_series.get() + (size_t)s * SRTB_B200_MAX_BOXCARS * p.L;
      for (int b = 0; b < res[s].n_boxcars; b++) {
        if (res[s].signal_count[b] == 0) continue;
        srtb::work::time_series_holder h;
        h.time_series_length = res[s].series_length[b];
        h.boxcar_length = res[s].boxcar_length[b];
        h.signal_count = res[s].signal_count[b];
        h.h_time_series = std::shared_ptr<srtb::real>(p.h_series, base + (size_t)b * p.L);
        w.time_series.push_back(h);
      }
      out.push_back(std::move(w));
    }
  }
};

}  // namespace pipeline
}  // namespace srtb
