// srtb/cuda_queue.hpp — the handle every pipe is constructed with, replacing `sycl::queue`
// (reference: pipes take `sycl::queue q`, pipeline/framework/pipe.hpp:148-161; main.cpp:99).
// One cuda_queue = one GPU + one CUDA stream + one srtb_b200_ctx; it is a cheap copyable handle.
#pragma once
#include <cuda_runtime_api.h>

#include <memory>
#include <stdexcept>
#include <string>

#include "srtb_b200.h"

namespace srtb {

inline void cuda_check(cudaError_t e, const char* what) {
  if (e != cudaSuccess) throw std::runtime_error(std::string{what} + ": " + cudaGetErrorString(e));
}

class cuda_queue {
  struct state {
    int device = 0;
    cudaStream_t stream = nullptr;
    srtb_b200_ctx* ctx = nullptr;
    bool own_stream = false;
    ~state() {
      if (ctx) srtb_b200_ctx_destroy(ctx);
      if (own_stream && stream) {
        cudaSetDevice(device);
        cudaStreamDestroy(stream);
      }
    }
  };
  std::shared_ptr<state> s_;

 public:
  /** new non-blocking stream + context on `device`; throws if there is no CUDA device */
  explicit cuda_queue(int device = 0) : s_{std::make_shared<state>()} {
    s_->device = device;
    cuda_check(cudaSetDevice(device), "cudaSetDevice");
    cuda_check(cudaStreamCreateWithFlags(&s_->stream, cudaStreamNonBlocking), "cudaStreamCreate");
    s_->own_stream = true;
    if (srtb_b200_ctx_create(device, s_->stream, &s_->ctx) != 0)
      throw std::runtime_error(std::string{"srtb_b200_ctx_create: "} + srtb_b200_last_error(nullptr));
  }
  int device() const { return s_->device; }
  cudaStream_t stream() const { return s_->stream; }
  srtb_b200_ctx* ctx() const { return s_->ctx; }
  /** the reference's operators all end in `.wait()`; pipes call this in drop-in mode */
  void wait() const { cuda_check(cudaStreamSynchronize(s_->stream), "cudaStreamSynchronize"); }
  /** throw the reference's exception types on a C-ABI error */
  void check(int rc) const {
    if (rc >= 0) return;
    const std::string msg = srtb_b200_last_error(s_->ctx);
    if (rc == SRTB_B200_E_UNSUPPORTED || rc == SRTB_B200_E_INVALID || rc == SRTB_B200_E_SIZE)
      throw std::invalid_argument(msg);
    throw std::runtime_error(msg);
  }
};

}  // namespace srtb
