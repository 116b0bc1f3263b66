// srtb/work.hpp — work items and work queues of the pipeline, same fields and names as the
// reference (userspace/include/srtb/work.hpp:30-72 work_queue, :102-157 work<T>, :162-285 typedefs).
// Boost.Lockfree / moodycamel are not used: a bounded ring (SPSC use) and a mutex deque (MPMC use)
// provide the same push / pop / read_available / empty surface.
#pragma once
#include <array>
#include <atomic>
#include <cstddef>
#include <cstdint>
#include <deque>
#include <memory>
#include <mutex>
#include <vector>

#include "srtb/config.hpp"

namespace srtb {

/** bounded single-producer single-consumer ring, capacity fixed at compile time */
template <typename T, size_t capacity>
class spsc_ring {
  std::array<T, capacity + 1> slots_{};
  std::atomic<size_t> head_{0}, tail_{0};  // head: next pop, tail: next push

 public:
  bool push(const T& v) {
    const size_t t = tail_.load(std::memory_order_relaxed);
    const size_t next = (t + 1) % (capacity + 1);
    if (next == head_.load(std::memory_order_acquire)) return false;  // full
    slots_[t] = v;
    tail_.store(next, std::memory_order_release);
    return true;
  }
  bool pop(T& out) {
    const size_t h = head_.load(std::memory_order_relaxed);
    if (h == tail_.load(std::memory_order_acquire)) return false;  // empty
    out = std::move(slots_[h]);
    slots_[h] = T{};
    head_.store((h + 1) % (capacity + 1), std::memory_order_release);
    return true;
  }
  size_t read_available() const {
    const size_t h = head_.load(std::memory_order_acquire), t = tail_.load(std::memory_order_acquire);
    return (t + capacity + 1 - h) % (capacity + 1);
  }
  bool empty() const { return read_available() == 0; }
};

/** unbounded multi-producer multi-consumer queue */
template <typename T>
class mpmc_queue {
  mutable std::mutex m_;
  std::deque<T> q_;

 public:
  bool push(const T& v) {
    std::lock_guard<std::mutex> g{m_};
    q_.push_back(v);
    return true;
  }
  bool pop(T& out) {
    std::lock_guard<std::mutex> g{m_};
    if (q_.empty()) return false;
    out = std::move(q_.front());
    q_.pop_front();
    return true;
  }
  size_t read_available() const {
    std::lock_guard<std::mutex> g{m_};
    return q_.size();
  }
  bool empty() const { return read_available() == 0; }
};

// work_queue<T, spsc, fixed_size, capacity> as in the reference; spsc = false -> MPMC
template <typename T, bool spsc = true, bool fixed_size = srtb::work_queue_fixed_size,
          size_t capacity = srtb::work_queue_capacity>
class work_queue : public spsc_ring<T, capacity> {
 public:
  using work_type = T;
};
template <typename T, bool fixed_size, size_t capacity>
class work_queue<T, false, fixed_size, capacity> : public mpmc_queue<T> {
 public:
  using work_type = T;
};

namespace pipeline {
struct dummy_work {};  // pipeline/framework/dummy_pipe.hpp
}  // namespace pipeline

namespace work {

using dummy_work = srtb::pipeline::dummy_work;

/** original baseband of a block, kept so a positive detection can be written out (work.hpp:88-96) */
struct baseband_data_holder {
  std::shared_ptr<std::byte> baseband_ptr;
  size_t baseband_input_bytes = 0;
};

template <typename T>
struct work {
  T ptr{};
  size_t count = 0;       // shape[0]
  size_t batch_size = 1;  // shape[1]
  uint64_t timestamp = 0;
  uint64_t udp_packet_counter = static_cast<uint64_t>(-1);
  uint32_t data_stream_id = 0;
  static constexpr uint64_t no_udp_packet_counter = static_cast<uint64_t>(-1);
  baseband_data_holder baseband_data;

  template <typename U>
  void move_parameter_from(work<U>&& other) {
    timestamp = other.timestamp;
    udp_packet_counter = other.udp_packet_counter;
    data_stream_id = other.data_stream_id;
    baseband_data = std::move(other.baseband_data);
  }
  template <typename U>
  void copy_parameter_from(const work<U>& other) {
    timestamp = other.timestamp;
    udp_packet_counter = other.udp_packet_counter;
    data_stream_id = other.data_stream_id;
    baseband_data = other.baseband_data;
  }
};

using complex_ptr = std::shared_ptr<srtb::complex<srtb::real>>;
using copy_to_device_work = work<std::shared_ptr<std::byte>>;
using unpack_work = work<std::shared_ptr<std::byte>>;
using fft_1d_r2c_work = work<std::shared_ptr<srtb::real>>;
using fft_1d_c2c_work = work<complex_ptr>;
using rfi_mitigation_s1_work = work<complex_ptr>;
using dedisperse_work = work<complex_ptr>;
using ifft_1d_c2c_work = work<complex_ptr>;
using refft_1d_c2c_work = work<complex_ptr>;
using watfft_1d_c2c_work = work<complex_ptr>;
using rfi_mitigation_s2_work = work<complex_ptr>;
using signal_detect_work = work<complex_ptr>;
using write_file_work = work<complex_ptr>;

/** one detected time series: host copy + boxcar length (work.hpp:240-247; no transfer_event —
 *  the copy has completed when the detect pipe returns) */
struct time_series_holder {
  std::shared_ptr<srtb::real> h_time_series;
  size_t time_series_length = 0;
  size_t boxcar_length = 0;
  size_t signal_count = 0;
};

struct write_signal_work : public write_file_work {
  std::vector<time_series_holder> time_series;  // non-empty <=> has_signal
  size_t zero_count = 0;
};

}  // namespace work
}  // namespace srtb
