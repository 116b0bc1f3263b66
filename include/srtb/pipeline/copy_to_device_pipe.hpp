// srtb/pipeline/copy_to_device_pipe.hpp — host block -> device block
// (reference: userspace/include/srtb/pipeline/copy_to_device_pipe.hpp:30-52): allocates a device
// buffer of baseband_input_bytes, copies from work.baseband_data (pinned host), forwards as unpack_work.
#pragma once
#include <optional>
#include <stop_token>

#include "srtb/cuda_queue.hpp"
#include "srtb/memory.hpp"
#include "srtb/pipeline/mode.hpp"
#include "srtb/work.hpp"

namespace srtb {
namespace pipeline {

class copy_to_device_pipe {
 protected:
  srtb::cuda_queue q;

 public:
  explicit copy_to_device_pipe(srtb::cuda_queue q_) : q{q_} {}

  auto operator()(std::stop_token, srtb::work::copy_to_device_work in_work) {
    cuda_check(cudaSetDevice(q.device()), "cudaSetDevice");
    const size_t bytes = in_work.baseband_data.baseband_input_bytes;
    auto d = srtb::device_allocator.allocate_shared<std::byte>(bytes);
    cuda_check(cudaMemcpyAsync(d.get(), in_work.baseband_data.baseband_ptr.get(), bytes,
                               cudaMemcpyHostToDevice, q.stream()),
               "cudaMemcpyAsync H2D");
    end_of_pipe(q);
    srtb::work::unpack_work out;
    out.move_parameter_from(std::move(in_work));
    out.ptr = d;
    out.count = bytes;
    out.batch_size = 1;
    return std::optional{out};
  }
};

}  // namespace pipeline
}  // namespace srtb
