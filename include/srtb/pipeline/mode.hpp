// srtb/pipeline/mode.hpp — sync contract of the re-hosted pipes.
// The reference's operators all `.wait()` (e.g. unpack.hpp:196, coherent_dedispersion.hpp:236), so
// a work handed downstream is complete on the device. drop_in (default) keeps that contract with a
// stream synchronise at the end of every pipe; stream_ordered skips it when every pipe of a chain
// shares one cuda_queue (composite_pipe), leaving one host sync in signal_detect.
#pragma once
namespace srtb {
namespace pipeline {
enum class sync_mode { drop_in, stream_ordered };
inline sync_mode pipe_sync_mode = sync_mode::drop_in;
template <typename Queue>
inline void end_of_pipe(const Queue& q) {
  if (pipe_sync_mode == sync_mode::drop_in) q.wait();
}
}  // namespace pipeline
}  // namespace srtb
