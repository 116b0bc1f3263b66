// srtb/pipeline/udp_receiver_pipe.hpp — source pipe: packets -> zero-filled pinned block -> copy_to_device_work
// (reference: userspace/include/srtb/pipeline/udp_receiver_pipe.hpp:39-161). One instance per receiver
// (data_stream_id = receiver id); the block size is baseband_input_count * |bits| / 8 * streams of the
// backend; the work carries the counter of the block's first packet and a wall-clock timestamp.
// The overlap-save tail (nsamps_reserved) is not re-sent for live streams, exactly like the reference's
// `continuous` worker-less path: consecutive blocks are disjoint.
#pragma once
#include <chrono>
#include <optional>
#include <stop_token>

#include "srtb/config.hpp"
#include "srtb/io/udp_block_assembler.hpp"
#include "srtb/log.hpp"
#include "srtb/memory.hpp"
#include "srtb/work.hpp"

namespace srtb {
namespace pipeline {

template <typename PacketProvider, typename Backend>
class udp_receiver_pipe {
  srtb::io::udp::block_assembler<PacketProvider, Backend> assembler_;
  size_t id_;

 public:
  explicit udp_receiver_pipe(PacketProvider provider, size_t id = 0) : assembler_{std::move(provider)}, id_{id} {}

  std::optional<srtb::work::copy_to_device_work> operator()(std::stop_token stop_token, srtb::work::dummy_work) {
    const size_t bytes = srtb::config.baseband_input_count *
                         static_cast<size_t>(std::abs(srtb::config.baseband_input_bits)) / srtb::BITS_PER_BYTE *
                         Backend::data_stream_count;
    auto h_in = srtb::host_allocator.allocate_shared<std::byte>(bytes);
    const auto first = assembler_.receive(std::span<std::byte>(h_in.get(), bytes), stop_token);
    if (!first.has_value()) return std::nullopt;
    SRTB_LOGD << " [udp receiver pipe] " << "id = " << id_ << ": block from packet " << *first << ", lost so far "
              << assembler_.total_lost_packet_count;
    srtb::work::copy_to_device_work w;
    w.ptr = nullptr;
    w.count = bytes;
    w.baseband_data = {h_in, bytes};
    w.timestamp = (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(
                      std::chrono::system_clock::now().time_since_epoch())
                      .count();
    w.udp_packet_counter = *first;
    w.data_stream_id = id_;
    return w;
  }
  size_t lost_packets() const { return assembler_.total_lost_packet_count; }
  size_t received_packets() const { return assembler_.total_received_packet_count; }
};

}  // namespace pipeline
}  // namespace srtb
