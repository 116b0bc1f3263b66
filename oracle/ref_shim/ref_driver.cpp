// ref_driver.cpp — C entry points over the REFERENCE'S OWN operator and pipe headers, compiled from
// /root/reference/userspace/include through the host shims in this directory
// (SYCL runtime, Boost, SyclCPLX, SyclParallelSTL are third-party and shimmed; every srtb header
// named below is the reference's file, included where it lies — nothing is copied).
// Output: oracle/_ref/libsrtb_ref.so. TEST INFRASTRUCTURE: used to pin oracle/srtb_oracle.cpp and to
// generate tests/golden/*.npz (tests/golden/make_golden.py).
#include <cstring>
#include <stop_token>
#include <vector>

#include "srtb/commons.hpp"
// reference operator headers (verbatim)
#include "srtb/coherent_dedispersion.hpp"
#include "srtb/fft/fft_window.hpp"
#include "srtb/fft/naive_fft.hpp"
#include "srtb/signal_detect.hpp"
#include "srtb/spectrum/rfi_mitigation.hpp"
#include "srtb/unpack.hpp"
// reference pipe headers (verbatim)
#include "srtb/pipeline/dedisperse_pipe.hpp"
#include "srtb/pipeline/rfi_mitigation_pipe.hpp"
#include "srtb/pipeline/signal_detect_pipe.hpp"

namespace {

using C = srtb::complex<srtb::real>;
sycl::queue& queue() {
  static sycl::queue q;
  return q;
}

template <typename Window>
auto window_functor(size_t n) {
  return srtb::fft::fft_window_functor_manager<srtb::real, Window, false>{Window{}, n, queue()}.functor;
}

template <typename F>
int with_window(int window, size_t n, F f) {
  if (window == 0) return f(window_functor<srtb::fft::window::rectangle<>>(n));
  if (window == 1) return f(window_functor<srtb::fft::window::hann<>>(n));
  if (window == 2) return f(window_functor<srtb::fft::window::hamming<>>(n));
  return -1;
}

template <typename T>
std::shared_ptr<C> to_device(const float* x, size_t count) {
  auto p = srtb::device_allocator.allocate_shared<C>(count);
  std::memcpy(p.get(), x, count * sizeof(C));
  return p;
}

}  // namespace

extern "C" {

// ---- srtb::unpack::unpack<BITS> with the bits dispatch of pipeline/unpack_pipe.hpp:72-127
int srtb_ref_unpack(const void* in, size_t out_count, int bits, int window, float* out) {
  auto& q = queue();
  return with_window(window, out_count, [&](auto functor) -> int {
    auto d_in = reinterpret_cast<std::byte*>(const_cast<void*>(in));
    switch (bits) {
      case 1: srtb::unpack::unpack<1>(d_in, out, out_count, functor, q); return 0;
      case 2: srtb::unpack::unpack<2>(d_in, out, out_count, functor, q); return 0;
      case 4: srtb::unpack::unpack<4>(d_in, out, out_count, functor, q); return 0;
      case 8: srtb::unpack::unpack<8>(reinterpret_cast<uint8_t*>(d_in), out, out_count, functor, q); return 0;
      case -8: srtb::unpack::unpack<8>(reinterpret_cast<int8_t*>(d_in), out, out_count, functor, q); return 0;
      case 16: srtb::unpack::unpack<16>(reinterpret_cast<uint16_t*>(d_in), out, out_count, functor, q); return 0;
      case -16: srtb::unpack::unpack<16>(reinterpret_cast<int16_t*>(d_in), out, out_count, functor, q); return 0;
      case 32: srtb::unpack::unpack<32>(reinterpret_cast<float*>(d_in), out, out_count, functor, q); return 0;
      case 64: srtb::unpack::unpack<64>(reinterpret_cast<double*>(d_in), out, out_count, functor, q); return 0;
      default: return -1;
    }
  });
}

// the handwritten 1/2/4-bit item functions (unpack.hpp:77-140), which test-unpack.cpp compares
int srtb_ref_unpack_handwritten(const void* in, size_t out_count, int bits, float* out) {
  auto& q = queue();
  auto d_in = reinterpret_cast<std::byte*>(const_cast<void*>(in));
  auto functor = window_functor<srtb::fft::window::rectangle<>>(out_count);
  switch (bits) {
    case 1: srtb::unpack::unpack<1, true>(d_in, out, out_count, functor, q); return 0;
    case 2: srtb::unpack::unpack<2, true>(d_in, out, out_count, functor, q); return 0;
    case 4: srtb::unpack::unpack<4, true>(d_in, out, out_count, functor, q); return 0;
    default: return -1;
  }
}

int srtb_ref_unpack_interleaved_2(const void* in, size_t out_count, int bits, int window, float* o1, float* o2) {
  auto& q = queue();
  return with_window(window, out_count, [&](auto functor) -> int {
    void* p = const_cast<void*>(in);
    switch (bits) {
      case 8: srtb::unpack::unpack<8>(static_cast<uint8_t*>(p), o1, o2, out_count, functor, q); return 0;
      case -8: srtb::unpack::unpack<8>(static_cast<int8_t*>(p), o1, o2, out_count, functor, q); return 0;
      case 16: srtb::unpack::unpack<16>(static_cast<uint16_t*>(p), o1, o2, out_count, functor, q); return 0;
      case -16: srtb::unpack::unpack<16>(static_cast<int16_t*>(p), o1, o2, out_count, functor, q); return 0;
      case 32: srtb::unpack::unpack<32>(static_cast<float*>(p), o1, o2, out_count, functor, q); return 0;
      case 64: srtb::unpack::unpack<64>(static_cast<double*>(p), o1, o2, out_count, functor, q); return 0;
      default: return -1;
    }
  });
}

int srtb_ref_unpack_snap1(const void* in, size_t out_count, int window, float* o1, float* o2) {
  return with_window(window, out_count, [&](auto functor) -> int {
    srtb::unpack::unpack_naocpsr_snap1(static_cast<int8_t*>(const_cast<void*>(in)), o1, o2, out_count, functor,
                                       queue());
    return 0;
  });
}

int srtb_ref_unpack_gznupsr_a1(const void* in, size_t out_count, int streams, int window, float* const* out) {
  return with_window(window, out_count, [&](auto functor) -> int {
    auto p = static_cast<int8_t*>(const_cast<void*>(in));
    if (streams == 4)
      srtb::unpack::unpack_gznupsr_a1(p, out[0], out[1], out[2], out[3], out_count, functor, queue());
    else if (streams == 2)
      srtb::unpack::unpack_gznupsr_a1(p, out[0], out[1], out_count, functor, queue());
    else
      return -1;
    return 0;
  });
}

float srtb_ref_window(int window, size_t i, size_t n) {
  float r = 0;
  with_window(window, n, [&](auto functor) -> int {
    r = functor(i, 1.0f);
    return 0;
  });
  return r;
}

// ---- naive_fft (fft/naive_fft.hpp:155-176, 221-261)
void srtb_ref_fft_c2c(float* x, size_t n, int direction) {
  size_t k = 0;
  while ((size_t{1} << k) < n) k++;
  auto p = reinterpret_cast<C*>(x);
  naive_fft::fft_1d_c2c<srtb::real, C>(k, p, p, direction, queue());
}
void srtb_ref_fft_r2c(float* inout, size_t n_real) {
  size_t k = 0;
  while ((size_t{1} << k) < n_real) k++;
  naive_fft::fft_1d_r2c<srtb::real, C>(k, inout, reinterpret_cast<C*>(inout), queue());
}
void srtb_ref_watfft(float* x, size_t length, size_t batch) {
  // batched backward C2C as fft/naive_fft_wrapper.hpp:88-91 loops it
  size_t k = 0;
  while ((size_t{1} << k) < length) k++;
  for (size_t b = 0; b < batch; b++) {
    auto p = reinterpret_cast<C*>(x) + b * length;
    naive_fft::fft_1d_c2c<srtb::real, C>(k, p, p, -1, queue());
  }
}

// ---- rfi_mitigation_s1_pipe (pipeline/rfi_mitigation_pipe.hpp:43-101), in place on x
void srtb_ref_rfi_s1_pipe(float* x, size_t count, float threshold, size_t channel_count, float freq_low,
                          float bandwidth, const char* freq_list) {
  srtb::config.mitigate_rfi_average_method_threshold = threshold;
  srtb::config.spectrum_channel_count = channel_count;
  srtb::config.baseband_freq_low = freq_low;
  srtb::config.baseband_bandwidth = bandwidth;
  srtb::config.mitigate_rfi_freq_list = freq_list ? freq_list : "";
  srtb::pipeline::rfi_mitigation_s1_pipe pipe{queue()};
  srtb::work::rfi_mitigation_s1_work w;
  w.ptr = to_device<C>(x, count);
  w.count = count;
  w.batch_size = 1;
  auto out = pipe(std::stop_token{}, w);
  std::memcpy(x, out.value().ptr.get(), count * sizeof(C));
}

size_t srtb_ref_eval_rfi_ranges(const char* list, float* pairs, size_t max_pairs) {
  auto r = srtb::spectrum::eval_rfi_ranges(std::string{list});
  for (size_t i = 0; i < r.size() && i < max_pairs; i++) {
    pairs[2 * i] = r[i].first;
    pairs[2 * i + 1] = r[i].second;
  }
  return r.size();
}

void srtb_ref_rfi_manual(float* x, size_t count, float freq_low, float bandwidth, const float* pairs, size_t n) {
  std::vector<srtb::spectrum::rfi_range_type> ranges;
  for (size_t i = 0; i < n; i++) ranges.emplace_back(pairs[2 * i], pairs[2 * i + 1]);
  srtb::spectrum::mitigate_rfi_manual(reinterpret_cast<C*>(x), count, freq_low, bandwidth, ranges, queue());
}

// ---- dedisperse_pipe (pipeline/dedisperse_pipe.hpp:31-48), in place on x
void srtb_ref_dedisperse_pipe(float* x, size_t count, float freq_low, float bandwidth, float dm) {
  srtb::config.baseband_freq_low = freq_low;
  srtb::config.baseband_bandwidth = bandwidth;
  srtb::config.dm = dm;
  srtb::pipeline::dedisperse_pipe pipe{queue()};
  srtb::work::dedisperse_work w;
  w.ptr = to_device<C>(x, count);
  w.count = count;
  w.batch_size = 1;
  auto out = pipe(std::stop_token{}, w);
  std::memcpy(x, out.value().ptr.get(), count * sizeof(C));
}
// the operator itself with explicit scalars (coherent_dedispersion.hpp:223-237)
void srtb_ref_dedisperse(float* x, size_t count, float f_min, float f_c, float df, float dm) {
  srtb::coherent_dedispersion::coherent_dedispertion(reinterpret_cast<C*>(x), count, f_min, f_c, df, dm, queue());
}

size_t srtb_ref_nsamps_reserved(size_t n, size_t channel_count, float freq_low, float bandwidth, float sample_rate,
                                float dm, int reserve) {
  srtb::config.baseband_input_count = n;
  srtb::config.spectrum_channel_count = channel_count;
  srtb::config.baseband_freq_low = freq_low;
  srtb::config.baseband_bandwidth = bandwidth;
  srtb::config.baseband_sample_rate = sample_rate;
  srtb::config.dm = dm;
  srtb::config.baseband_reserve_sample = reserve != 0;
  return srtb::coherent_dedispersion::nsamps_reserved();
}

// ---- rfi_mitigation_s2_pipe (pipeline/rfi_mitigation_pipe.hpp:113-130), in place on x [C][L]
void srtb_ref_rfi_s2_pipe(float* x, size_t time_count, size_t chan_count, float sk_threshold) {
  srtb::config.mitigate_rfi_spectral_kurtosis_threshold = sk_threshold;
  srtb::pipeline::rfi_mitigation_s2_pipe pipe{queue()};
  srtb::work::rfi_mitigation_s2_work w;
  w.ptr = to_device<C>(x, time_count * chan_count);
  w.count = time_count;
  w.batch_size = chan_count;
  auto out = pipe(std::stop_token{}, w);
  std::memcpy(x, out.value().ptr.get(), time_count * chan_count * sizeof(C));
}

// ---- signal_detect_pipe_2 (pipeline/signal_detect_pipe.hpp:252-442)
// outputs: n_series holders {boxcar_length, series_length, signal_count} + their host series
// (row i of series_out, stride time_count). returns the number of holders.
int srtb_ref_signal_detect_pipe(const float* x, size_t time_count, size_t chan_count, size_t baseband_input_count,
                                int reserve_sample, float freq_low, float bandwidth, float sample_rate, float dm,
                                float snr, float chan_thr, size_t max_boxcar, unsigned long long* boxcar_length,
                                unsigned long long* series_length, unsigned long long* signal_count,
                                float* series_out, int max_series) {
  auto& cfg = srtb::config;
  cfg.baseband_input_count = baseband_input_count;
  cfg.spectrum_channel_count = chan_count;
  cfg.baseband_reserve_sample = reserve_sample != 0;
  cfg.baseband_freq_low = freq_low;
  cfg.baseband_bandwidth = bandwidth;
  cfg.baseband_sample_rate = sample_rate;
  cfg.dm = dm;
  cfg.signal_detect_signal_noise_threshold = snr;
  cfg.signal_detect_channel_threshold = chan_thr;
  cfg.signal_detect_max_boxcar_length = max_boxcar;
  srtb::pipeline::signal_detect_pipe_2 pipe{queue()};
  srtb::work::signal_detect_work w;
  w.ptr = to_device<C>(x, time_count * chan_count);
  w.count = time_count;
  w.batch_size = chan_count;
  auto out = pipe(std::stop_token{}, w).value();
  int n = 0;
  for (auto& h : out.time_series) {
    if (n >= max_series) break;
    boxcar_length[n] = h.boxcar_length;
    series_length[n] = h.time_series_length;
    std::memcpy(series_out + (size_t)n * time_count, h.h_time_series.get(), h.time_series_length * sizeof(float));
    signal_count[n] = srtb::signal_detect::count_signal<srtb::real>(h.h_time_series.get(), h.time_series_length,
                                                                    snr, queue());
    n++;
  }
  return n;
}

// ---- alternates of the refft path ([time][frequency]): SK v1 (spectrum/rfi_mitigation.hpp:181-275) and
// signal_detect_pipe v1 (pipeline/signal_detect_pipe.hpp:51-230), both as the reference defines them
void srtb_ref_sk_v1(float* x, size_t fft_bins, size_t time_counts, float sk_threshold) {
  auto d = to_device<C>(x, fft_bins * time_counts);
  srtb::spectrum::mitigate_rfi_spectral_kurtosis_method(d.get(), fft_bins, time_counts, sk_threshold, queue());
  std::memcpy(x, d.get(), fft_bins * time_counts * sizeof(C));
}

int srtb_ref_signal_detect_pipe_v1(float* x, size_t count_per_batch, size_t batch_size, float sk_threshold, float snr,
                                   float chan_thr, size_t max_boxcar, unsigned long long* boxcar_length,
                                   unsigned long long* series_length, unsigned long long* signal_count,
                                   float* series_out, int max_series) {
  auto& cfg = srtb::config;
  cfg.mitigate_rfi_spectral_kurtosis_threshold = sk_threshold;
  cfg.signal_detect_signal_noise_threshold = snr;
  cfg.signal_detect_channel_threshold = chan_thr;
  cfg.signal_detect_max_boxcar_length = max_boxcar;
  srtb::pipeline::signal_detect_pipe pipe{queue()};
  srtb::work::signal_detect_work w;
  w.ptr = to_device<C>(x, count_per_batch * batch_size);
  w.count = count_per_batch;
  w.batch_size = batch_size;
  auto out = pipe(std::stop_token{}, w).value();
  std::memcpy(x, out.ptr.get(), count_per_batch * batch_size * sizeof(C));  // spectrum after SK v1
  int n = 0;
  for (auto& h : out.time_series) {
    if (n >= max_series) break;
    h.transfer_event.wait();
    boxcar_length[n] = h.boxcar_length;
    series_length[n] = h.time_series_length;
    std::memcpy(series_out + (size_t)n * batch_size, h.h_time_series.get(), h.time_series_length * sizeof(float));
    signal_count[n] = srtb::signal_detect::count_signal<srtb::real>(h.h_time_series.get(), h.time_series_length,
                                                                    snr, queue());
    n++;
  }
  return n;
}

// count_signal alone (signal_detect.hpp:32-72)
unsigned long long srtb_ref_count_signal(const float* v, size_t n, float snr) {
  return srtb::signal_detect::count_signal<srtb::real>(const_cast<float*>(v), n, snr, queue());
}

}  // extern "C"
