// srtb/config.hpp — the srtb::configs surface the hot path reads, field for field
// (reference: userspace/include/srtb/config.hpp:28-58 compile-time knobs, :80-249 struct configs).
// GUI / FFTW-only fields are kept so the shipped .cfg files still map onto the struct.
#pragma once
#include <complex>
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace srtb {

using real = float;                       // math.hpp:44
template <typename T>
using complex = std::complex<T>;          // layout-compatible with float2 on the device

inline constexpr size_t work_queue_capacity = 2;      // config.hpp:40
inline constexpr bool work_queue_fixed_size = true;   // config.hpp:43
inline constexpr bool fft_window_precompute = false;  // config.hpp:46
inline constexpr bool fft_operate_in_place = true;    // config.hpp:47
inline constexpr size_t MEMORY_ALIGNMENT = 64ul;      // config.hpp:56
inline constexpr size_t BITS_PER_BYTE = 8ul;          // config.hpp:58

struct configs {
  std::string config_file_name = "srtb_config.cfg";
  size_t baseband_input_count = size_t{1} << 28;
  int32_t baseband_input_bits = 8;
  std::string baseband_format_type = "simple";
  srtb::real baseband_freq_low = 1000.0;
  srtb::real baseband_bandwidth = 500.0;
  srtb::real baseband_sample_rate = 1000 * 1e6;
  bool baseband_reserve_sample = true;
  srtb::real dm = 0;
  std::vector<std::string> udp_receiver_address = {"10.0.1.2"};
  std::vector<unsigned short> udp_receiver_port = {12004};
  std::vector<uint32_t> udp_receiver_cpu_preferred = {0};
  std::string input_file_path = "";
  size_t input_file_offset_bytes = 0;
  std::string baseband_output_file_prefix = "srtb_baseband_output_";
  bool baseband_write_all = false;
  std::string fft_fftw_wisdom_path = "srtb_fftw_wisdom.txt";
  srtb::real mitigate_rfi_average_method_threshold = 10;
  srtb::real mitigate_rfi_spectral_kurtosis_threshold = 1.1;
  std::string mitigate_rfi_freq_list = "";
  size_t spectrum_sum_count = 1;
  size_t spectrum_channel_count = size_t{1} << 15;
  srtb::real signal_detect_signal_noise_threshold = 6;
  srtb::real signal_detect_channel_threshold = 0.9;
  size_t signal_detect_max_boxcar_length = 1024;
  size_t thread_query_work_wait_time = 1000;
  bool gui_enable = false;
  size_t gui_pixmap_width = 1920;
  size_t gui_pixmap_height = 1080;
};

// the one global every pipe re-reads at call time (global_variables.hpp:42)
inline srtb::configs config;

}  // namespace srtb
