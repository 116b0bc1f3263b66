// pipeline_main.cpp — the reference's main.cpp wiring (userspace/src/main.cpp:125-228) on the
// re-hosted pipes: one thread per pipe, work queues between them, one cuda_queue per GPU:
//   copy_to_device -> unpack -> fft_1d_r2c -> rfi_mitigation_s1 -> dedisperse -> watfft_1d_c2c
//                  -> rfi_mitigation_s2 -> signal_detect_pipe_2 -> (sink: prints one JSON line per work)
// Input: a raw baseband file (--input) cut into blocks of baseband_input_count samples per stream.
// Used by tests/test_gpu_pipeline.py, which compares the printed results with the CPU oracle.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "srtb/config.hpp"
#include "srtb/cuda_queue.hpp"
#include "srtb/memory.hpp"
#include "srtb/pipeline/baseband_chain_pipe.hpp"
#include "srtb/pipeline/copy_to_device_pipe.hpp"
#include "srtb/pipeline/dedisperse_pipe.hpp"
#include "srtb/pipeline/fft_pipe.hpp"
#include "srtb/pipeline/framework/composite_pipe.hpp"
#include "srtb/pipeline/framework/pipe.hpp"
#include "srtb/pipeline/framework/pipe_io.hpp"
#include "srtb/pipeline/read_file_pipe.hpp"
#include "srtb/pipeline/rfi_mitigation_pipe.hpp"
#include "srtb/pipeline/signal_detect_pipe.hpp"
#include "srtb/pipeline/udp_receiver_pipe.hpp"
#include "srtb/pipeline/unpack_pipe.hpp"
#include "srtb/pipeline/write_signal_pipe.hpp"
#include "srtb/program_options.hpp"
#include "srtb/work.hpp"

namespace {

struct sink_out_functor {
  std::shared_ptr<std::atomic<int>> done;
  std::string dump_prefix;
  std::shared_ptr<srtb::pipeline::write_signal_pipe> writer;  // candidate sink (.bin/.npy/.tim), optional
  void operator()(std::stop_token st, srtb::work::write_signal_work w) {
    if (writer) (*writer)(st, w);
    std::string line = "{\"stream\": " + std::to_string(w.data_stream_id) + ", \"block\": " +
                       std::to_string(w.udp_packet_counter) + ", \"count\": " + std::to_string(w.count) +
                       ", \"batch_size\": " + std::to_string(w.batch_size) + ", \"zero_count\": " +
                       std::to_string(w.zero_count) + ", \"series\": [";
    for (size_t i = 0; i < w.time_series.size(); i++) {
      const auto& h = w.time_series[i];
      double peak = 0;
      for (size_t j = 0; j < h.time_series_length; j++) peak = std::max<double>(peak, h.h_time_series.get()[j]);
      line += std::string(i ? ", " : "") + "{\"boxcar\": " + std::to_string(h.boxcar_length) + ", \"length\": " +
              std::to_string(h.time_series_length) + ", \"count\": " + std::to_string(h.signal_count) +
              ", \"peak\": " + std::to_string(peak) + "}";
    }
    line += "]}";
    if (!dump_prefix.empty()) {  // dynamic spectrum [C][L] complex64, like the reference's .npy payload
      std::vector<char> host(w.count * w.batch_size * sizeof(srtb::complex<srtb::real>));
      srtb::cuda_check(cudaMemcpy(host.data(), w.ptr.get(), host.size(), cudaMemcpyDeviceToHost), "D2H spectrum");
      std::ofstream f(dump_prefix + std::to_string(w.udp_packet_counter) + "." + std::to_string(w.data_stream_id) + ".bin",
                      std::ios::binary);
      f.write(host.data(), (std::streamsize)host.size());
    }
    {
      static std::mutex print_mutex;  // several fused chain pipes may drain into this sink at once
      std::lock_guard<std::mutex> lock(print_mutex);
      std::cout << line << std::endl;
    }
    (*done)++;
  }
};

const char* arg(int argc, char** argv, const char* name, const char* def) {
  for (int i = 1; i + 1 < argc; i++)
    if (!std::strcmp(argv[i], name)) return argv[i + 1];
  return def;
}

}  // namespace

int main(int argc, char** argv) {
  using namespace srtb::pipeline;
  auto& cfg = srtb::config;
  cfg.baseband_input_count = size_t{1} << std::atoi(arg(argc, argv, "--log2n", "20"));
  cfg.baseband_input_bits = std::atoi(arg(argc, argv, "--bits", "-8"));
  cfg.baseband_format_type = arg(argc, argv, "--format", "simple");
  cfg.baseband_freq_low = std::atof(arg(argc, argv, "--freq-low", "1000"));
  cfg.baseband_bandwidth = std::atof(arg(argc, argv, "--bandwidth", "500"));
  cfg.baseband_sample_rate = std::atof(arg(argc, argv, "--sample-rate", "1e9"));
  cfg.baseband_reserve_sample = std::atoi(arg(argc, argv, "--reserve", "0")) != 0;
  cfg.dm = std::atof(arg(argc, argv, "--dm", "0"));
  cfg.spectrum_channel_count = std::strtoull(arg(argc, argv, "--channels", "256"), nullptr, 10);
  cfg.mitigate_rfi_average_method_threshold = std::atof(arg(argc, argv, "--avg-thr", "10"));
  cfg.mitigate_rfi_spectral_kurtosis_threshold = std::atof(arg(argc, argv, "--sk-thr", "1.1"));
  cfg.mitigate_rfi_freq_list = arg(argc, argv, "--freq-list", "");
  cfg.signal_detect_signal_noise_threshold = std::atof(arg(argc, argv, "--snr", "6"));
  cfg.signal_detect_max_boxcar_length = std::strtoull(arg(argc, argv, "--max-boxcar", "64"), nullptr, 10);
  // a reference-style config file (expressions and all) overrides the flags above
  const std::string cfg_file = arg(argc, argv, "--config_file_name", "");
  if (!cfg_file.empty()) {
    std::string a0 = argv[0], a1 = "--config_file_name", a2 = cfg_file;
    char* av[] = {a0.data(), a1.data(), a2.data()};
    srtb::program_options::apply_changed_configs(srtb::program_options::parse_arguments(3, av, cfg_file), cfg);
  }
  const std::string input = cfg.input_file_path.empty() ? arg(argc, argv, "--input", "") : cfg.input_file_path;
  cfg.input_file_path = input;
  const std::string dump = arg(argc, argv, "--dump-prefix", "");
  const bool write_candidates = std::atoi(arg(argc, argv, "--write-candidates", "0")) != 0;
  const bool composite = std::atoi(arg(argc, argv, "--composite", "0")) != 0;
  if (input.empty()) {
    std::fprintf(stderr, "usage: pipeline_main --input <file> [--log2n 20 --bits -8 --format simple ...]\n");
    return 2;
  }
  size_t streams = 1;
  if (cfg.baseband_format_type == "naocpsr_snap1" || cfg.baseband_format_type == "interleaved_samples_2" ||
      cfg.baseband_format_type == "gznupsr_a1")
    streams = 2;

  srtb::cuda_queue q{std::atoi(arg(argc, argv, "--device", "0"))};

  using namespace srtb::work;
  auto copy_q = std::make_shared<srtb::work_queue<copy_to_device_work, false>>();
  auto unpack_q = std::make_shared<srtb::work_queue<unpack_work, false>>();
  auto r2c_q = std::make_shared<srtb::work_queue<fft_1d_r2c_work>>();
  auto s1_q = std::make_shared<srtb::work_queue<rfi_mitigation_s1_work>>();
  auto dd_q = std::make_shared<srtb::work_queue<dedisperse_work>>();
  auto wat_q = std::make_shared<srtb::work_queue<watfft_1d_c2c_work>>();
  auto s2_q = std::make_shared<srtb::work_queue<rfi_mitigation_s2_work>>();
  auto det_q = std::make_shared<srtb::work_queue<signal_detect_work>>();
  auto done = std::make_shared<std::atomic<int>>(0);
  sink_out_functor sink{done, dump, nullptr};
  if (write_candidates) sink.writer = std::make_shared<write_signal_pipe>(q);

  std::vector<std::jthread> threads;
  const int fused = std::atoi(arg(argc, argv, "--fused", "0"));
  const int ring_depth = std::atoi(arg(argc, argv, "--ring", "1"));  // > 1: submit/collect ring inside each chain pipe
  if (fused > 0) {
    // the whole device chain as one pipe (fused kernels), `fused` of them on their own queue each, all fed from
    // the same MPMC queue: blocks alternate over the contexts and overlap on the GPU
    for (int i = 0; i < fused; i++) {
      srtb::cuda_queue qi = (i == 0) ? q : srtb::cuda_queue{q.device()};
      threads.push_back(start_pipe<baseband_chain_pipe>(idle_queue_in_functor{copy_q}, multiple_works_out_functor{sink}, qi,
                                                        !dump.empty(), ring_depth));
    }
  } else {
    // the H2D copy gets its own queue (stream): block k+1 crosses PCIe while block k computes; the hand-over is the
    // pipe's own wait() (drop-in contract: the work is complete when the pipe returns)
    // (not in the stream-ordered composite mode, where pipes do not wait and must share one stream)
    srtb::cuda_queue q_copy = composite ? q : srtb::cuda_queue{q.device()};
    threads.push_back(start_pipe<copy_to_device_pipe>(queue_in_functor{copy_q}, queue_out_functor{unpack_q}, q_copy));
    threads.push_back(
        start_unpack_pipe(cfg.baseband_format_type, queue_in_functor{unpack_q}, queue_out_functor{r2c_q}, q));
  }
  if (fused > 0) {
    // nothing else to start: the chain pipes drain straight into the sink
  } else if (!composite) {
    threads.push_back(start_pipe<fft_1d_r2c_pipe>(queue_in_functor{r2c_q}, queue_out_functor{s1_q}, q));
    threads.push_back(start_pipe<rfi_mitigation_s1_pipe>(queue_in_functor{s1_q}, queue_out_functor{dd_q}, q));
    threads.push_back(start_pipe<dedisperse_pipe>(queue_in_functor{dd_q}, queue_out_functor{wat_q}, q));
    threads.push_back(start_pipe<watfft_1d_c2c_pipe>(queue_in_functor{wat_q}, queue_out_functor{s2_q}, q));
    threads.push_back(start_pipe<rfi_mitigation_s2_pipe>(queue_in_functor{s2_q}, queue_out_functor{det_q}, q));
    threads.push_back(start_pipe<signal_detect_pipe_2>(queue_in_functor{det_q}, sink, q));
  } else {
    // stream-ordered fast mode: the six device stages as one composite pipe on one thread, no
    // per-stage host wait (all share q's stream; signal_detect synchronises once)
    pipe_sync_mode = sync_mode::stream_ordered;
    using chain = composite_pipe<fft_1d_r2c_pipe, rfi_mitigation_s1_pipe, dedisperse_pipe, watfft_1d_c2c_pipe,
                                 rfi_mitigation_s2_pipe, signal_detect_pipe_2>;
    threads.push_back(start_pipe<chain>(queue_in_functor{r2c_q}, sink, q));
  }

  int blocks = 0;
  auto t_start = std::chrono::steady_clock::now();
  auto feed = [&](auto& source, bool renumber) {
    while (auto w = source(std::stop_token{}, srtb::work::dummy_work{})) {
      if (renumber) w->udp_packet_counter = (uint64_t)blocks;  // deterministic file names / JSON keys for the test
      while (copy_q->read_available() >= 2) std::this_thread::sleep_for(std::chrono::microseconds(50));
      copy_q->push(*w);
      blocks++;
    }
  };
  const int drop_every = std::atoi(arg(argc, argv, "--udp-drop-every", "0"));
  if (std::atoi(arg(argc, argv, "--udp-shaped", "0")) != 0) {
    // UDP-shaped source (BASELINE config #5): the file is framed into backend packets (8-byte counter +
    // 4096 payload bytes), every `drop_every`-th packet is lost on the way, and udp_receiver_pipe
    // assembles blocks by counter with zero fill. The block key is the counter of its first packet.
    namespace io = srtb::io;
    std::ifstream f(input, std::ios::binary);
    std::vector<char> bytes((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    auto run = [&]<typename B>(B) {
      auto packets = io::udp::frame_stream<B>(
          std::span<const std::byte>(reinterpret_cast<const std::byte*>(bytes.data()), bytes.size()), 0);
      io::udp::memory_packet_provider prov;
      for (size_t i = 0; i < packets.size(); i++)
        if (!(drop_every > 0 && i % (size_t)drop_every == (size_t)drop_every - 1)) prov.push(std::move(packets[i]));
      udp_receiver_pipe<io::udp::memory_packet_provider, B> receiver{std::move(prov)};
      feed(receiver, false);
      std::fprintf(stderr, "[udp-shaped] received %zu packets, lost %zu\n", receiver.received_packets(),
                   receiver.lost_packets());
    };
    if (streams == 2) run(io::backend_registry::naocpsr_snap1{});
    else run(io::backend_registry::fastmb_roach2{});
  } else if (std::atoi(arg(argc, argv, "--preload", "0")) != 0) {
    // device-path throughput without host file I/O in the timed region: read every block into pinned memory
    // first, then replay the same works --repeat times (each replay is still H2D + the whole chain + D2H)
    read_file_pipe reader;
    std::vector<srtb::work::copy_to_device_work> preloaded;
    while (auto w = reader(std::stop_token{}, srtb::work::dummy_work{})) preloaded.push_back(*w);
    const int repeat = std::atoi(arg(argc, argv, "--repeat", "1"));
    t_start = std::chrono::steady_clock::now();
    for (int r = 0; r < repeat; r++)
      for (auto w : preloaded) {
        w.udp_packet_counter = (uint64_t)blocks;
        while (copy_q->read_available() >= 8) std::this_thread::yield();
        copy_q->push(w);
        blocks++;
      }
  } else {
    // source: read_file_pipe (pinned host block, zero padded tail, overlap-save rewind by nsamps_reserved)
    read_file_pipe reader;
    feed(reader, true);
  }
  while (done->load() < blocks * (int)streams) std::this_thread::yield();
  {
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
    std::fprintf(stderr, "[pipeline_main] %d block(s) x %zu stream(s) of 2^%d samples in %.3f s = %.2f Gsamples/s (host wall clock: source + H2D + chain + D2H)\n",
                 blocks, streams, std::atoi(arg(argc, argv, "--log2n", "20")), dt,
                 (double)blocks * (double)streams * (double)cfg.baseband_input_count / dt / 1e9);
  }
  for (auto& t : threads) t.request_stop();
  threads.clear();
  srtb::device_allocator.deallocate_all_free_ptrs();
  srtb::host_allocator.deallocate_all_free_ptrs();
  return 0;
}
