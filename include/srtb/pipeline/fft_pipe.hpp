// srtb/pipeline/fft_pipe.hpp — fft_1d_r2c_pipe, watfft_1d_c2c_pipe and the alternates ifft_1d_c2c_pipe / refft_1d_c2c_pipe
// (reference: userspace/include/srtb/pipeline/fft_pipe.hpp:32-80 and :285-372). In place; the R2C
// pipe reinterprets the float buffer as complex and drops the Nyquist bin (count = N/2, :75-77);
// watfft: batch = min(spectrum_channel_count, count), length = count / batch (:318-320), output
// work {count = length (time), batch_size = batch (frequency)}. Plans are not objects here: the
// library re-sizes its scratch when a work's size differs (fft_wrapper::set_size semantics).
#pragma once
#include <algorithm>
#include <memory>
#include <optional>
#include <stop_token>

#include "srtb/config.hpp"
#include "srtb/cuda_queue.hpp"
#include "srtb/log.hpp"
#include "srtb/pipeline/dedisperse_pipe.hpp"  // srtb::codd::nsamps_reserved
#include "srtb/pipeline/mode.hpp"
#include "srtb/work.hpp"

namespace srtb {
namespace pipeline {

class fft_1d_r2c_pipe {
 protected:
  srtb::cuda_queue q;

 public:
  explicit fft_1d_r2c_pipe(srtb::cuda_queue q_) : q{q_} {}

  auto operator()(std::stop_token, srtb::work::fft_1d_r2c_work in_work) {
    const size_t in_count = in_work.count;
    const size_t out_count = in_count / 2 + 1;
    cuda_check(cudaSetDevice(q.device()), "cudaSetDevice");
    q.check(srtb_b200_fft_r2c_inplace(q.ctx(), in_work.ptr.get(), in_count));
    end_of_pipe(q);
    auto d_out = std::reinterpret_pointer_cast<srtb::complex<srtb::real>>(in_work.ptr);
    in_work.ptr.reset();
    srtb::work::rfi_mitigation_s1_work out;
    out.move_parameter_from(std::move(in_work));
    out.ptr = d_out;
    out.count = out_count - 1;  // drop the highest frequency point
    out.batch_size = 1;
    return std::optional{out};
  }
};

class watfft_1d_c2c_pipe {
 protected:
  srtb::cuda_queue q;

 public:
  explicit watfft_1d_c2c_pipe(srtb::cuda_queue q_) : q{q_} {}

  auto operator()(std::stop_token, srtb::work::watfft_1d_c2c_work in_work) {
    const size_t input_count = in_work.count;
    const size_t batch = std::min(srtb::config.spectrum_channel_count, input_count);
    const size_t length = input_count / batch;
    cuda_check(cudaSetDevice(q.device()), "cudaSetDevice");
    q.check(srtb_b200_watfft_c2c_backward(q.ctx(), in_work.ptr.get(), length, batch));
    end_of_pipe(q);
    srtb::work::rfi_mitigation_s2_work out;
    auto ptr = in_work.ptr;
    out.move_parameter_from(std::move(in_work));
    out.ptr = ptr;
    out.count = length;
    out.batch_size = batch;
    return std::optional{out};
  }
};

/** Alternative back half of the chain (reference: fft_pipe.hpp:88-185, not wired in main.cpp): the dedispersed
 *  spectrum back to complex time samples — one backward C2C over the whole block, in place, no normalisation
 *  (cufftExecC2C CUFFT_INVERSE semantics); the overlap-save tail (nsamps_reserved() / 2 complex samples) is cut off
 *  the count. The default window is the rectangle, so no de-windowing step exists. */
class ifft_1d_c2c_pipe {
 protected:
  srtb::cuda_queue q;

 public:
  explicit ifft_1d_c2c_pipe(srtb::cuda_queue q_) : q{q_} {}

  auto operator()(std::stop_token, srtb::work::ifft_1d_c2c_work in_work) {
    const size_t input_count = in_work.count;
    cuda_check(cudaSetDevice(q.device()), "cudaSetDevice");
    q.check(srtb_b200_fft_c2c(q.ctx(), in_work.ptr.get(), input_count, 1, -1));
    end_of_pipe(q);
    const size_t reserved_complex = srtb::codd::nsamps_reserved() / 2;
    size_t output_count = input_count;
    if (reserved_complex < input_count) {
      output_count = input_count - reserved_complex;
    } else {
      SRTB_LOGW << " [ifft 1d c2c pipe] " << "nsamps_reserved_complex = " << reserved_complex
                << " >= input_count = " << input_count;
    }
    srtb::work::refft_1d_c2c_work out;
    auto ptr = in_work.ptr;
    out.move_parameter_from(std::move(in_work));
    out.ptr = ptr;
    out.count = output_count;
    return std::optional{out};
  }
};

/** ... and short forward transforms of spectrum_channel_count points over the dedispersed time samples
 *  (fft_pipe.hpp:197-278): high time resolution, layout [time][frequency]; output work
 *  {count = refft_length (frequency), batch_size = number of spectra (time)}. */
class refft_1d_c2c_pipe {
 protected:
  srtb::cuda_queue q;

 public:
  explicit refft_1d_c2c_pipe(srtb::cuda_queue q_) : q{q_} {}

  auto operator()(std::stop_token, srtb::work::refft_1d_c2c_work in_work) {
    const size_t input_count = in_work.count;
    const size_t refft_length = std::min(srtb::config.spectrum_channel_count, input_count);
    const size_t refft_batch_size = input_count / refft_length;
    cuda_check(cudaSetDevice(q.device()), "cudaSetDevice");
    q.check(srtb_b200_fft_c2c(q.ctx(), in_work.ptr.get(), refft_length, refft_batch_size, +1));
    end_of_pipe(q);
    srtb::work::signal_detect_work out;
    auto ptr = in_work.ptr;
    out.move_parameter_from(std::move(in_work));
    out.ptr = ptr;
    out.count = refft_length;
    out.batch_size = refft_batch_size;
    return std::optional{out};
  }
};

}  // namespace pipeline
}  // namespace srtb
