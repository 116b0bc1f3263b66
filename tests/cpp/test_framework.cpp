// CPU-only test of the re-hosted pipe framework (include/srtb/pipeline/framework/*, srtb/work.hpp):
// queues, queue functors, thread-per-pipe start_pipe, fan-out, tee, loose out, composite_pipe,
// stop behaviour. The reference has no test for its framework; the expectations here
// are its documented contract (pipe.hpp:108-175, pipe_io.hpp:28-152, composite_pipe.hpp:29-51).
#include <array>
#include <atomic>
#include <cassert>
#include <chrono>
#include <cstdio>
#include <memory>
#include <numeric>
#include <vector>

#include "srtb/pipeline/framework/composite_pipe.hpp"
#include "srtb/pipeline/framework/dummy_pipe.hpp"
#include "srtb/pipeline/framework/pipe.hpp"
#include "srtb/io/udp_block_assembler.hpp"
#include "srtb/pipeline/framework/pipe_io.hpp"
#include "srtb/work.hpp"

#define CHECK(...)                                                         \
  do {                                                                      \
    if (!(__VA_ARGS__)) {                                                      \
      std::fprintf(stderr, "CHECK failed: %s (%s:%d)\n", #__VA_ARGS__, __FILE__, __LINE__); \
      return 1;                                                             \
    }                                                                       \
  } while (0)

using int_work = srtb::work::work<int>;

struct add_pipe {
  int delta;
  explicit add_pipe(int d) : delta{d} {}
  auto operator()(std::stop_token, int_work w) {
    w.ptr += delta;
    return std::optional{w};
  }
};
struct double_pipe {
  double_pipe() = default;
  explicit double_pipe(int) {}
  auto operator()(std::stop_token, int_work w) {
    w.ptr *= 2;
    return std::optional{w};
  }
};
struct split_pipe {  // 1 in -> 2 out, data_stream_id = 2*id + s (unpack fan-out contract)
  explicit split_pipe(int) {}
  auto operator()(std::stop_token, int_work w) {
    std::array<int_work, 2> out{w, w};
    out[0].data_stream_id = 2 * w.data_stream_id;
    out[1].data_stream_id = 2 * w.data_stream_id + 1;
    return std::optional{out};
  }
};
struct stop_at_pipe {  // empty optional ends the pipe thread (pipe.hpp:132-135)
  int limit;
  explicit stop_at_pipe(int l) : limit{l} {}
  std::optional<int_work> operator()(std::stop_token, int_work w) {
    if (w.ptr >= limit) return std::nullopt;
    return w;
  }
};

int main() {
  using namespace srtb::pipeline;
  srtb::config.thread_query_work_wait_time = 1000;

  // --- queues: capacity 2 SPSC ring and MPMC deque, same push/pop/read_available/empty surface
  {
    srtb::work_queue<int_work> q;
    int_work w;
    CHECK(q.empty() && q.read_available() == 0 && !q.pop(w));
    w.ptr = 1;
    CHECK(q.push(w));
    w.ptr = 2;
    CHECK(q.push(w));
    w.ptr = 3;
    CHECK(!q.push(w));  // work_queue_capacity = 2 (config.hpp:40)
    CHECK(q.read_available() == 2);
    CHECK(q.pop(w) && w.ptr == 1);
    CHECK(q.pop(w) && w.ptr == 2);
    CHECK(q.empty());
    srtb::work_queue<int_work, false> m;
    for (int i = 0; i < 100; i++) {
      w.ptr = i;
      CHECK(m.push(w));
    }
    CHECK(m.read_available() == 100);
    CHECK(m.pop(w) && w.ptr == 0);
  }
  // --- work parameter propagation
  {
    int_work a;
    a.timestamp = 7;
    a.udp_packet_counter = 9;
    a.data_stream_id = 3;
    a.baseband_data.baseband_input_bytes = 11;
    srtb::work::work<float> b;
    b.copy_parameter_from(a);
    CHECK(b.timestamp == 7 && b.udp_packet_counter == 9 && b.data_stream_id == 3 &&
          b.baseband_data.baseband_input_bytes == 11);
    CHECK(int_work::no_udp_packet_counter == static_cast<uint64_t>(-1));
  }
  // --- three pipes on three threads, chained by capacity-2 queues; order preserved
  {
    auto q0 = std::make_shared<srtb::work_queue<int_work, false>>();
    auto q1 = std::make_shared<srtb::work_queue<int_work>>();
    auto q2 = std::make_shared<srtb::work_queue<int_work>>();
    auto q3 = std::make_shared<srtb::work_queue<int_work, false>>();
    std::jthread t1 = start_pipe<add_pipe>(queue_in_functor{q0}, queue_out_functor{q1}, 1);
    std::jthread t2 = start_pipe<double_pipe>(queue_in_functor{q1}, queue_out_functor{q2}, 0);
    std::jthread t3 = start_pipe<add_pipe>(queue_in_functor{q2}, queue_out_functor{q3}, -3);
    const int n = 200;
    for (int i = 0; i < n; i++) {
      int_work w;
      w.ptr = i;
      w.timestamp = i;
      q0->push(w);
    }
    std::vector<int> got;
    while ((int)got.size() < n) {
      int_work w;
      if (q3->pop(w)) {
        CHECK(w.timestamp == (uint64_t)got.size());
        got.push_back(w.ptr);
      } else {
        std::this_thread::yield();
      }
    }
    for (int i = 0; i < n; i++) CHECK(got[i] == (i + 1) * 2 - 3);
    t1.request_stop();
    t2.request_stop();
    t3.request_stop();
  }
  // --- fan-out pipe + multiple_works_out_functor, tee, loose out
  {
    auto qin = std::make_shared<srtb::work_queue<int_work, false>>();
    auto qa = std::make_shared<srtb::work_queue<int_work, false>>();
    auto qb = std::make_shared<srtb::work_queue<int_work>>();  // capacity 2: loose pushes drop
    std::jthread t = start_pipe<split_pipe>(
        queue_in_functor{qin},
        multiple_works_out_functor{multiple_out_functors_functor{queue_out_functor{qa}, loose_queue_out_functor{qb}}},
        0);
    for (int i = 0; i < 10; i++) {
      int_work w;
      w.ptr = i;
      w.data_stream_id = i;
      qin->push(w);
    }
    while (qa->read_available() < 20) std::this_thread::yield();
    for (int i = 0; i < 20; i++) {
      int_work w;
      CHECK(qa->pop(w));
      CHECK(w.ptr == i / 2 && w.data_stream_id == (uint32_t)i);
    }
    CHECK(qb->read_available() == 2);  // the side branch kept only what fitted
    t.request_stop();
  }
  // --- composite_pipe runs members back to back on one thread; empty optional propagates
  {
    composite_pipe<add_pipe, double_pipe, add_pipe> c{5};
    int_work w;
    w.ptr = 1;
    auto out = c(std::stop_token{}, w);
    CHECK(out && out->ptr == (1 + 5) * 2 + 5);
    composite_pipe<stop_at_pipe, double_pipe> s{3};
    w.ptr = 2;
    CHECK(s(std::stop_token{}, w)->ptr == 4);
    w.ptr = 3;
    CHECK(!s(std::stop_token{}, w));
  }
  // --- a pipe returning an empty optional ends its thread; request_stop ends a blocked pipe
  {
    auto qin = std::make_shared<srtb::work_queue<int_work, false>>();
    auto qout = std::make_shared<srtb::work_queue<int_work, false>>();
    std::jthread t = start_pipe<stop_at_pipe>(queue_in_functor{qin}, queue_out_functor{qout}, 3);
    for (int i = 0; i < 6; i++) {
      int_work w;
      w.ptr = i;
      qin->push(w);
    }
    t.join();  // thread exits by itself at ptr == 3
    CHECK(qout->read_available() == 3);
    std::jthread idle = start_pipe<dummy_pipe<int_work>>(queue_in_functor{qout}, dummy_out_functor<srtb::work::dummy_work>{}, 0);
    while (!qout->empty()) std::this_thread::yield();
    idle.request_stop();
    idle.join();
  }
#ifdef SRTB_HAS_SOCKETS
  {
    // live-socket source: a move-only, socket-owning packet provider goes through start_pipe into a
    // pipe functor built on the pipe's thread; datagrams sent over the loop-back interface come out as assembled
    // blocks; an idle socket does not keep the thread from stopping (bounded recvfrom + stop_token)
    namespace io = srtb::io;
    using B = io::backend_registry::fastmb_roach2;
    constexpr size_t d = B::packet_payload_size - B::packet_header_size;
    struct block_work {
      std::vector<std::byte> bytes;
      uint64_t first = 0;
    };
    struct socket_source_pipe {
      io::udp::block_assembler<io::udp::recvfrom_packet_provider, B> assembler;
      explicit socket_source_pipe(io::udp::recvfrom_packet_provider p) : assembler{std::move(p)} {}
      std::optional<block_work> operator()(std::stop_token st, srtb::work::dummy_work) {
        block_work w;
        w.bytes.resize(4 * d);
        const auto first = assembler.receive(w.bytes, st);
        if (!first) return std::nullopt;
        w.first = *first;
        return w;
      }
    };
    io::udp::recvfrom_packet_provider provider{"127.0.0.1", 0};
    const unsigned short port = provider.bound_port();
    CHECK(port != 0);
    auto blocks = std::make_shared<srtb::work_queue<block_work, false>>();
    std::jthread src = start_pipe<socket_source_pipe>(dummy_in_functor<>{}, queue_out_functor{blocks}, std::move(provider));
    const int tx = ::socket(AF_INET, SOCK_DGRAM, 0);
    sockaddr_in to{};
    to.sin_family = AF_INET;
    to.sin_port = htons(port);
    ::inet_pton(AF_INET, "127.0.0.1", &to.sin_addr);
    std::vector<std::byte> stream(8 * d);
    for (size_t i = 0; i < stream.size(); i++) stream[i] = static_cast<std::byte>((i * 7 + i / d) & 0xff);
    auto packets = io::udp::frame_stream<B>(stream, 500);
    for (size_t i = 0; i < packets.size(); i++) {
      if (i == 2) continue;  // one datagram lost on the way
      CHECK(::sendto(tx, packets[i].data(), packets[i].size(), 0, reinterpret_cast<sockaddr*>(&to), sizeof(to)) ==
            (ssize_t)packets[i].size());
      std::this_thread::sleep_for(std::chrono::milliseconds(1));
    }
    block_work b0, b1;
    const auto deadline = std::chrono::steady_clock::now() + std::chrono::seconds(5);
    while (!blocks->pop(b0)) CHECK(std::chrono::steady_clock::now() < deadline);
    while (!blocks->pop(b1)) CHECK(std::chrono::steady_clock::now() < deadline);
    CHECK(b0.first == 500 && b1.first == 504);
    for (size_t pkt = 0; pkt < 4; pkt++)
      for (size_t j = 0; j < d; j += 97) {
        CHECK(b0.bytes[pkt * d + j] == (pkt == 2 ? std::byte{0} : stream[pkt * d + j]));  // the lost packet is zero-filled
        CHECK(b1.bytes[pkt * d + j] == stream[(4 + pkt) * d + j]);
      }
    // the socket is idle now: the receiver must still stop promptly
    const auto t_stop = std::chrono::steady_clock::now();
    src.request_stop();
    src.join();
    CHECK(std::chrono::steady_clock::now() - t_stop < std::chrono::seconds(2));
    ::close(tx);
  }
#endif
  CHECK(class_name<add_pipe>() == "add_pipe");
  CHECK(class_name<composite_pipe<add_pipe, double_pipe>>() == "composite_pipe");
  CHECK(generate_thread_name<composite_pipe<add_pipe, double_pipe>>().size() <= 15);
  std::printf("framework ok\n");
  return 0;
}
