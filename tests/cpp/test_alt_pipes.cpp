// GPU test of the alternate back half of the chain (ifft_1d_c2c_pipe -> refft_1d_c2c_pipe, reference
// fft_pipe.hpp:88-278): a random spectrum of Nc bins goes back to Nc complex time samples (one backward C2C,
// unnormalised), loses the overlap-save tail, and is cut into spectra of spectrum_channel_count points (forward C2C),
// layout [time][frequency]. Checked against a float64 DFT on the host. Run by tests/test_gpu_pipeline.py.
#include <cmath>
#include <complex>
#include <cstdio>
#include <random>
#include <vector>

#include "srtb/config.hpp"
#include "srtb/cuda_queue.hpp"
#include "srtb/memory.hpp"
#include "srtb/pipeline/fft_pipe.hpp"
#include "srtb/pipeline/signal_detect_pipe.hpp"

#define CHECK(...)                                                                     \
  do {                                                                                 \
    if (!(__VA_ARGS__)) {                                                              \
      std::fprintf(stderr, "CHECK failed: %s (%s:%d)\n", #__VA_ARGS__, __FILE__, __LINE__); \
      return 1;                                                                        \
    }                                                                                  \
  } while (0)

int main() {
  using cd = std::complex<double>;
  using cf = srtb::complex<srtb::real>;
  auto& cfg = srtb::config;
  const size_t Nc = 1 << 12, C = 1 << 4;
  cfg.baseband_input_count = 2 * Nc;
  cfg.spectrum_channel_count = C;
  cfg.baseband_reserve_sample = true;      // a non-zero overlap-save tail
  cfg.baseband_freq_low = 1000.0f;
  cfg.baseband_bandwidth = 500.0f;
  cfg.baseband_sample_rate = 1e9f;
  cfg.dm = 2e-4f;  // dispersive delay of ~460 samples across the band: nsamps_reserved() = 928 of 8192
  const size_t reserved_complex = srtb::codd::nsamps_reserved() / 2;
  CHECK(reserved_complex > 0 && reserved_complex < Nc);
  srtb::cuda_queue q{0};
  std::mt19937 rng(7);
  std::uniform_real_distribution<float> u(-1.f, 1.f);
  std::vector<cf> h(Nc);
  for (auto& v : h) v = cf(u(rng), u(rng));
  auto d = srtb::device_allocator.allocate_shared<cf>(Nc);
  srtb::cuda_check(cudaMemcpy(d.get(), h.data(), Nc * sizeof(cf), cudaMemcpyHostToDevice), "H2D");
  srtb::work::ifft_1d_c2c_work w;
  w.ptr = d;
  w.count = Nc;
  w.udp_packet_counter = 42;
  srtb::pipeline::ifft_1d_c2c_pipe ifft{q};
  srtb::pipeline::refft_1d_c2c_pipe refft{q};
  auto t = ifft(std::stop_token{}, w);
  CHECK(t.has_value() && t->count == Nc - reserved_complex && t->udp_packet_counter == 42);
  auto f = refft(std::stop_token{}, *t);
  CHECK(f.has_value() && f->count == C && f->batch_size == (Nc - reserved_complex) / C);
  std::vector<cf> out(Nc);
  srtb::cuda_check(cudaMemcpy(out.data(), f->ptr.get(), Nc * sizeof(cf), cudaMemcpyDeviceToHost), "D2H");
  // float64 truth: x[n] = sum_k X[k] e^{+2 pi i k n / Nc}; spectrum m = forward DFT of x[mC .. mC + C)
  std::vector<cd> x(Nc);
  for (size_t n = 0; n < Nc; n++) {
    cd a = 0;
    for (size_t k = 0; k < Nc; k++) a += cd(h[k]) * std::polar(1.0, 2.0 * M_PI * (double)((k * n) % Nc) / (double)Nc);
    x[n] = a;
  }
  double num = 0, den = 0;
  for (size_t m = 0; m < f->batch_size; m++)
    for (size_t c = 0; c < C; c++) {
      cd a = 0;
      for (size_t j = 0; j < C; j++) a += x[m * C + j] * std::polar(1.0, -2.0 * M_PI * (double)((c * j) % C) / (double)C);
      num += std::norm(cd(out[m * C + c]) - a);
      den += std::norm(a);
    }
  const double rel = std::sqrt(num / den);
  std::printf("alt pipes ok: rel-L2 %.3e over %zu spectra of %zu channels (tail of %zu samples cut)\n", rel,
              (size_t)f->batch_size, C, reserved_complex);
  CHECK(rel < 1e-5);

  // ---- signal_detect_pipe (v1) on spectra [time][frequency]: SK v1, per-spectrum sums, count_signal; float64 truth
  {
    const size_t nt = 1024, nf = 64;
    std::normal_distribution<float> g(0.f, 1.f);
    std::vector<cf> s(nt * nf);
    for (auto& v : s) v = cf(g(rng), g(rng));
    for (size_t i = 0; i < nt; i++) s[i * nf + 5] = cf(3.f, 0.f);                   // steady tone: zapped by SK
    for (size_t i = 400; i < 416; i++)
      for (size_t j = 0; j < nf; j++) s[i * nf + j] *= 1.6f;                         // mild broadband burst
    cfg.mitigate_rfi_spectral_kurtosis_threshold = 1.4f;
    cfg.signal_detect_signal_noise_threshold = 5.0f;
    cfg.signal_detect_channel_threshold = 0.9f;
    cfg.signal_detect_max_boxcar_length = 64;
    auto ds = srtb::device_allocator.allocate_shared<cf>(nt * nf);
    srtb::cuda_check(cudaMemcpy(ds.get(), s.data(), nt * nf * sizeof(cf), cudaMemcpyHostToDevice), "H2D");
    srtb::work::signal_detect_work sw;
    sw.ptr = ds;
    sw.count = nf;
    sw.batch_size = nt;
    srtb::pipeline::signal_detect_pipe detect{q};
    auto r = detect(std::stop_token{}, sw);
    CHECK(r.has_value() && r->count == nf && r->batch_size == nt);
    // truth
    const double M = (double)nt, lo = (2 - 1.4) * ((M - 1) / (M + 1)) + 1, hi = 1.4 * ((M - 1) / (M + 1)) + 1;
    std::vector<char> zap(nf, 0);
    for (size_t j = 0; j < nf; j++) {
      double s2 = 0, s4 = 0;
      for (size_t i = 0; i < nt; i++) {
        const double p = std::norm(cd(s[i * nf + j]));
        s2 += p;
        s4 += p * p;
      }
      const double sk = M * s4 / (s2 * s2);
      zap[j] = (sk > hi || sk < lo);
    }
    CHECK(zap[5]);
    size_t zc = 0;
    for (size_t j = 0; j < nf; j++) zc += zap[j];
    CHECK(r->zero_count == zc);
    std::vector<double> ts(nt, 0.0);
    double mean = 0;
    for (size_t i = 0; i < nt; i++) {
      for (size_t j = 0; j < nf; j++)
        if (!zap[j]) ts[i] += std::norm(cd(s[i * nf + j]));
      mean += ts[i];
    }
    mean /= (double)nt;
    double var = 0;
    for (auto& v : ts) {
      v -= mean;
      var += v * v;
    }
    const double thr = 5.0 * std::sqrt(var / (double)nt);
    size_t cnt = 0;
    for (auto v : ts) cnt += (v > thr);
    CHECK(cnt >= 8);  // the burst
    const srtb::work::time_series_holder* h1 = nullptr;
    for (const auto& hh : r->time_series)
      if (hh.boxcar_length == 1) h1 = &hh;
    CHECK(h1 != nullptr && h1->time_series_length == nt);
    double worst = 0;
    for (size_t i = 0; i < nt; i++) worst = std::max(worst, std::fabs((double)h1->h_time_series.get()[i] - ts[i]));
    std::printf("signal_detect_pipe v1 ok: %zu masked channels, %zu samples over threshold (pipe: %zu), series max-abs diff %.3e\n",
                zc, cnt, (size_t)h1->signal_count, worst);
    CHECK(worst < 1e-3 * std::sqrt(var / (double)nt));
    CHECK(h1->signal_count + 1 >= cnt && h1->signal_count <= cnt + 1);
    // the spectrum handed on has the zapped channels zeroed
    std::vector<cf> back(nt * nf);
    srtb::cuda_check(cudaMemcpy(back.data(), r->ptr.get(), nt * nf * sizeof(cf), cudaMemcpyDeviceToHost), "D2H");
    for (size_t j = 0; j < nf; j++) CHECK((back[7 * nf + j] == cf(0.f, 0.f)) == (bool)zap[j]);
  }
  return 0;
}
