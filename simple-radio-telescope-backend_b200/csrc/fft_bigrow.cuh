// fft_bigrow.cuh — one-kernel waterfall for long channel rows (L = 2^13, 2^14): a whole row lives in ONE
// CTA's shared memory (64 / 128 KiB), so the chain
//   rfi_mitigation_s1 (zap + normalise)  -> dedisperse chirp -> watfft (backward C2C) -> rfi_mitigation_s2 (SK)
//   -> first stage of signal_detect's column sums
// touches HBM exactly twice (read the spectrum row, write the dynamic-spectrum row). Reference operators:
//   userspace/include/srtb/pipeline/rfi_mitigation_pipe.hpp:66-79, coherent_dedispersion.hpp:133-150,223-237,
//   pipeline/fft_pipe.hpp:313-371, spectrum/rfi_mitigation.hpp:292-341, pipeline/signal_detect_pipe.hpp:296-316.
// (The BASELINE "J1644 shape" has L = Nc / C = 2^25 / 2^11 = 2^14.)
//
// Algorithm (index algebra prototyped and checked against numpy in tools/proto_bigrow.py):
//   * in-place decimation-in-frequency, radices 16 x 16 x 16 x R3 (R3 = L / 4096 = 2 or 4); natural order in,
//     digit-reversed order in shared memory, natural order restored by the final store;
//   * a thread owns TWO adjacent butterflies per stage (32 points) and moves them with 16-byte shared-memory
//     accesses (one LDS.128 / STS.128 per two points, addresses base + i * stride: no per-access integer work);
//   * the row is padded so that every access pattern is bank-conflict free: position p = d0*B1 + d1*B2 + r
//     (B1 = L/16, B2 = L/256) lives at d0*S1 + d1*S2 + r with S2 = B2 + R3, S1 = 16*S2 + 2;
//   * the next row arrives by TMA bulk copies (cp.async.bulk, one per B2-element segment so that the padding is
//     produced by the copy engine) signalled on an mbarrier;
//   * twiddles: stage 0 from W_L^j by repeated multiplication, stages 1 and 2 from small shared-memory tables;
//   * the chirp phase is evaluated in fp64 like the reference; 1/f comes from one correctly rounded reciprocal per
//     thread and Newton steps from the neighbouring bin: one step when (B1*df/f)^2 <= 2^-52 (the reciprocal is then
//     good to an ulp, the accuracy of the reference's own fp64 division chain), two when the fourth power keeps the
//     phase error below 1e-9 cycles, otherwise the exact reciprocal of every bin; the fractional
//     part by the round-to-nearest trick (e^{-2 pi i k} only needs k mod 1), sin/cos from the SFU on [-pi, pi].
#pragma once
#include "fft_engine.cuh"

namespace srtb_b200 {

template <int LOGL>
struct bigrow {
  static_assert(LOGL == 13 || LOGL == 14, "whole-row kernel: L = 8192 or 16384");
  static constexpr int L = 1 << LOGL;
  static constexpr int R3 = L / 4096;
  static constexpr int B1 = L / 16, B2 = L / 256;
  static constexpr int S2 = B2 + R3, S1 = 16 * S2 + 2;
  static constexpr int BUF = 16 * S1;       // padded row, elements
  static constexpr int NT = L / 32;         // threads: 32 points each
  static constexpr int NW = NT / 32;
  static constexpr int CTAS = (LOGL == 13) ? 2 : 1;
  static constexpr int TABN = B1 + 15 * B2 + 15 * R3;  // W_L^j | W_B1^{i j} [15][B2] | W_B2^{i j} [15][R3]
  static constexpr size_t off_colacc = (size_t)BUF * sizeof(float2);
  __host__ __device__ static constexpr size_t off_tab(bool sk) { return off_colacc + (sk ? (size_t)BUF * sizeof(float) : 0); }
  __host__ __device__ static constexpr size_t off_mbar(bool sk) { return off_tab(sk) + (size_t)TABN * sizeof(float2); }
  __host__ __device__ static constexpr size_t off_red(bool sk) { return off_mbar(sk) + 128; }  // 16 mbarriers
  __host__ __device__ static constexpr size_t bytes(bool sk) { return off_red(sk) + 4 * 32 * sizeof(float); }
};

// a[i] *= w^i, i = 1..15 (powers at most four multiplications deep)
__device__ __forceinline__ void bigrow_twiddle_powers(float2 (&a)[16], const float2 w1) {
  const float2 w2 = c_sqr(w1), w3 = c_mul(w2, w1), w4 = c_sqr(w2);
  a[1] = c_mul(a[1], w1);
  a[2] = c_mul(a[2], w2);
  a[3] = c_mul(a[3], w3);
  a[4] = c_mul(a[4], w4);
  const float2 w5 = c_mul(w4, w1), w6 = c_sqr(w3), w7 = c_mul(w4, w3), w8 = c_sqr(w4);
  a[5] = c_mul(a[5], w5);
  a[6] = c_mul(a[6], w6);
  a[7] = c_mul(a[7], w7);
  a[8] = c_mul(a[8], w8);
  a[9] = c_mul(a[9], c_mul(w8, w1));
  a[10] = c_mul(a[10], c_sqr(w5));
  a[11] = c_mul(a[11], c_mul(w8, w3));
  a[12] = c_mul(a[12], c_sqr(w6));
  a[13] = c_mul(a[13], c_mul(w8, w5));
  a[14] = c_mul(a[14], c_sqr(w7));
  a[15] = c_mul(a[15], c_mul(w8, w7));
}

// one radix-16 DIF stage on two adjacent butterflies: elements base + i*STRIDE (16-byte pairs), twiddles
// W^{i j} for the pair (j, j + 1) from a table laid out [15][TB] (row i - 1, column j)
template <bool FWD, int STRIDE, int TB>
__device__ __forceinline__ void bigrow_stage_table(float2* __restrict__ p, const float2* __restrict__ tab) {
  float2 a[16], b[16];
#pragma unroll
  for (int i = 0; i < 16; i++) {
    const float4 q = *reinterpret_cast<const float4*>(p + i * STRIDE);
    a[i] = make_float2(q.x, q.y);
    b[i] = make_float2(q.z, q.w);
  }
  dft16<FWD>(a);
  dft16<FWD>(b);
  *reinterpret_cast<float4*>(p) = make_float4(a[0].x, a[0].y, b[0].x, b[0].y);
#pragma unroll
  for (int i = 1; i < 16; i++) {
    const float4 w = *reinterpret_cast<const float4*>(tab + (i - 1) * TB);
    const float2 ya = c_mul(a[i], make_float2(w.x, w.y)), yb = c_mul(b[i], make_float2(w.z, w.w));
    *reinterpret_cast<float4*>(p + i * STRIDE) = make_float4(ya.x, ya.y, yb.x, yb.y);
  }
}

#ifndef SRTB_BIGROW_FAST_SINCOS
#define SRTB_BIGROW_FAST_SINCOS 1
#endif

// s1 + chirp on one spectrum bin: f and 1/f in fp64 (K12: coherent_dedispersion.hpp:133-150)
__device__ __forceinline__ float2 bigrow_chirp_point(float2 v, double f, double r, const row_chirp_params& cp,
                                                     float limit) {
  const double q = (f - cp.f_c) * cp.inv_fc;
  const double k = (cp.ddm * r) * (q * q);
  // k mod 1 in [-0.5, 0.5]: e^{-2 pi i k} is unchanged by the integer that is dropped
  constexpr double MAGIC = 6755399441055744.0;  // 1.5 * 2^52
  const double kr = __dadd_rn(__dadd_rn(k, MAGIC), -MAGIC);
  const float frac = (float)(k - kr);
  float s, c;
#if SRTB_BIGROW_FAST_SINCOS
  __sincosf(-6.283185307179586f * frac, &s, &c);  // SFU, argument in [-pi, pi]: abs error <= 2^-21
#else
  sincospif(-2.0f * frac, &s, &c);
#endif
  const float scale = (v.x * v.x + v.y * v.y > limit) ? 0.f : cp.coef;  // rfi_mitigation_pipe.hpp:66-79
  const float wr = c * scale, wi = s * scale;
  return make_float2(v.x * wr - v.y * wi, v.x * wi + v.y * wr);
}

// CHIRP: 0 = plain transform; s1 + chirp on load with 1 = one-step Newton reciprocals, 3 = two steps, 2 = exact reciprocals
template <int LOGL, bool FWD, bool SK, int CHIRP>
__global__ void __launch_bounds__(bigrow<LOGL>::NT, bigrow<LOGL>::CTAS)
    fft_bigrow_kernel(const float2* __restrict__ in, float2* __restrict__ out, unsigned nrows,
                      const float2* __restrict__ tabs, row_sk_params skp, row_chirp_params cp) {
  using C = bigrow<LOGL>;
  constexpr int L = C::L, R3 = C::R3, B1 = C::B1, B2 = C::B2, S1 = C::S1, S2 = C::S2, NT = C::NT, NW = C::NW;
  extern __shared__ __align__(128) unsigned char smraw[];
  float2* const buf = reinterpret_cast<float2*>(smraw);
  float* const colacc = reinterpret_cast<float*>(smraw + C::off_colacc);  // padded like buf (SK only)
  float2* const T0 = reinterpret_cast<float2*>(smraw + C::off_tab(SK));
  float2* const T1 = T0 + B1;
  float2* const T2 = T1 + 15 * B2;
  uint64_t* const mbar = reinterpret_cast<uint64_t*>(smraw + C::off_mbar(SK));
  float* const red = reinterpret_cast<float*>(smraw + C::off_red(SK));  // [2 slots][s2 | s4][32]
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

  for (int i = tid; i < C::TABN; i += NT) T0[i] = __ldg(&tabs[i]);
  if constexpr (SK)
    for (int i = tid; i < C::BUF; i += NT) colacc[i] = 0.f;
  if (tid < 16) mbar_init(&mbar[tid], 1);
  if (tid == 0) fence_mbar_init();
  __syncthreads();

  // warp 0: fetch one row into the padded layout, one bulk copy per B2-element segment. The row is sixteen chunks of
  // B1 elements (d0 = 0..15), each signalled on its own mbarrier, fetched in order: stage 0 consumes chunk i (its
  // butterfly input i) while the later chunks are still in flight, so most of the load hides behind the chirp.
  auto issue = [&](unsigned row) {
    fence_proxy_async();
    if (lane < 16) mbar_expect_tx(&mbar[lane], (uint32_t)(B1 * sizeof(float2)));
    __syncwarp();
    const float2* src = in + (size_t)row * L;
#pragma unroll
    for (int seg = lane; seg < 256; seg += 32)  // seg >> 4 = chunk: two chunks per round, in chunk order
      bulk_g2s(buf + (seg >> 4) * S1 + (seg & 15) * S2, src + seg * B2, (uint32_t)(B2 * sizeof(float2)), &mbar[seg >> 4]);
  };

  float limit = 0.f;
  if constexpr (CHIRP) limit = cp.threshold * __ldg(cp.mean);

  unsigned row = blockIdx.x;
  if (row < nrows && wid == 0) issue(row);
  for (unsigned it = 0; row < nrows; row += gridDim.x, it++) {
    // ---- stage 0: butterflies over d0 (stride S1), pair j = 2 tid, 2 tid + 1 of [0, B1)
    {
      const int j = 2 * tid;
      float2* const p = buf + (j / B2) * S2 + (j % B2);
      float2 a[16], b[16];
      if constexpr (!CHIRP) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
          mbar_wait(&mbar[i], it & 1);
          const float4 q = *reinterpret_cast<const float4*>(p + i * S1);
          a[i] = make_float2(q.x, q.y);
          b[i] = make_float2(q.z, q.w);
        }
      }
      if constexpr (CHIRP) {
        // bin index of a[i] is row*L + j + i*B1 (b[i]: + 1); f = f_min + df * index in fp64
        double idx = (double)((size_t)row * L + j);
        double fa = fma(cp.df, idx, cp.f_min);
        double ra = __drcp_rn(fa);
        // 1/f of the next bin: Newton steps from the neighbour's reciprocal (each squares the relative error, which
        // starts at bin distance * df / f), or a correctly rounded reciprocal per bin (CHIRP == 2)
        auto refine = [&](double r, double f) {
          if constexpr (CHIRP == 1) {
            return fma(r, fma(-f, r, 1.0), r);  // one Newton step (the host checked that it is good to 1 ulp)
          } else if constexpr (CHIRP == 3) {
            r = fma(r, fma(-f, r, 1.0), r);     // two steps
            return fma(r, fma(-f, r, 1.0), r);
          } else {
            return __drcp_rn(f);                // widely spaced bins (short test blocks): exact reciprocal
          }
        };
#pragma unroll
        for (int i = 0; i < 16; i++) {
          if (i > 0) {
            idx += (double)B1;
            fa = fma(cp.df, idx, cp.f_min);
            ra = refine(ra, fa);
          }
          mbar_wait(&mbar[i], it & 1);  // chunk i has landed (later chunks still arriving)
          {
            const float4 q = *reinterpret_cast<const float4*>(p + i * S1);
            a[i] = make_float2(q.x, q.y);
            b[i] = make_float2(q.z, q.w);
          }
          a[i] = bigrow_chirp_point(a[i], fa, ra, cp, limit);
          const double fb = fma(cp.df, idx + 1.0, cp.f_min);
          b[i] = bigrow_chirp_point(b[i], fb, refine(ra, fb), cp, limit);
        }
      }
      dft16<FWD>(a);
      dft16<FWD>(b);
      const float4 w = *reinterpret_cast<const float4*>(T0 + j);
      bigrow_twiddle_powers(a, make_float2(w.x, w.y));
      bigrow_twiddle_powers(b, make_float2(w.z, w.w));
#pragma unroll
      for (int i = 0; i < 16; i++)
        *reinterpret_cast<float4*>(p + i * S1) = make_float4(a[i].x, a[i].y, b[i].x, b[i].y);
    }
    __syncthreads();
    // ---- stage 1: block d0, butterflies over d1 (stride S2), pair of [0, B2)
    {
      const int d0 = tid / (B2 / 2), jp = tid % (B2 / 2);
      bigrow_stage_table<FWD, S2, B2>(buf + d0 * S1 + 2 * jp, T1 + 2 * jp);
    }
    __syncthreads();
    // ---- stage 2: block (d0, d1), butterflies over d2 (stride R3), pair of [0, R3)
    {
      const int blk = tid / (R3 / 2), jp = tid % (R3 / 2);
      bigrow_stage_table<FWD, R3, R3>(buf + (blk >> 4) * S1 + (blk & 15) * S2 + 2 * jp, T2 + 2 * jp);
    }
    __syncthreads();
    // ---- last stage, pass 1: radix R3 on R3 contiguous elements, in place; row statistics for SK
    float s2 = 0.f, s4 = 0.f;
    if constexpr (R3 == 4) {
#pragma unroll
      for (int g = 0; g < 8; g++) {
        // lanes vary (d0 bit 0, d2 bits 0-1): the eight 16-byte chunks of a quarter warp are distinct mod 8
        const int u = (tid >> 3) + (NT / 8) * g;
        const int d0 = (lane & 1) | ((u & 7) << 1), d2 = ((lane >> 1) & 3) | ((u >> 7) << 2), d1 = (u >> 3) & 15;
        float2* const p = buf + d0 * S1 + d1 * S2 + d2 * 4;
        const float4 q0 = *reinterpret_cast<const float4*>(p), q1 = *reinterpret_cast<const float4*>(p + 2);
        float2 x0 = make_float2(q0.x, q0.y), x1 = make_float2(q0.z, q0.w);
        float2 x2 = make_float2(q1.x, q1.y), x3 = make_float2(q1.z, q1.w);
        dft4<FWD>(x0, x1, x2, x3);
        *reinterpret_cast<float4*>(p) = make_float4(x0.x, x0.y, x1.x, x1.y);
        *reinterpret_cast<float4*>(p + 2) = make_float4(x2.x, x2.y, x3.x, x3.y);
        if constexpr (SK) {
          const float p0 = x0.x * x0.x + x0.y * x0.y, p1 = x1.x * x1.x + x1.y * x1.y;
          const float p2 = x2.x * x2.x + x2.y * x2.y, p3 = x3.x * x3.x + x3.y * x3.y;
          s2 += (p0 + p1) + (p2 + p3);
          s4 += (p0 * p0 + p1 * p1) + (p2 * p2 + p3 * p3);
        }
      }
    } else {
#pragma unroll
      for (int g = 0; g < 16; g++) {
        const int G = tid + NT * g;  // pair index: d2 = G & 15, d1 = (G >> 4) & 15, d0 = G >> 8
        float2* const p = buf + (G >> 8) * S1 + ((G >> 4) & 15) * S2 + (G & 15) * 2;
        const float4 q = *reinterpret_cast<const float4*>(p);
        const float2 y0 = make_float2(q.x + q.z, q.y + q.w), y1 = make_float2(q.x - q.z, q.y - q.w);
        *reinterpret_cast<float4*>(p) = make_float4(y0.x, y0.y, y1.x, y1.y);
        if constexpr (SK) {
          const float p0 = y0.x * y0.x + y0.y * y0.y, p1 = y1.x * y1.x + y1.y * y1.y;
          s2 += p0 + p1;
          s4 += p0 * p0 + p1 * p1;
        }
      }
    }
    bool zap = false;
    if constexpr (SK) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        s4 += __shfl_xor_sync(0xffffffffu, s4, o);
      }
      float* const slot = red + (it & 1) * 64;  // per-warp partials alternate between two slots
      if (lane == 0) {
        slot[wid] = s2;
        slot[32 + wid] = s4;
      }
    }
    __syncthreads();
    if constexpr (SK) {
      // every warp folds the NW per-warp partials with the same fixed shuffle tree
      const float* const slot = red + (it & 1) * 64;
      float a = (lane < NW) ? slot[lane] : 0.f, b = (lane < NW) ? slot[32 + lane] : 0.f;
#pragma unroll
      for (int o = NW / 2; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
      }
      a = __shfl_sync(0xffffffffu, a, 0);
      b = __shfl_sync(0xffffffffu, b, 0);
      const float sk = (float)L * (b / (a * a));
      zap = (sk > skp.thr_hi || sk < skp.thr_lo);  // NaN (all-zero row): untouched (rfi_mitigation.hpp:333-339)
    }
    // ---- pass 2: natural-order store. lane -> (d0, d1 bit 0): 32 consecutive output indices per instruction
    {
      const int pd0 = lane & 15, pd1 = ((wid & 7) << 1) | (lane >> 4), sel = wid >> 3;
      const int base = pd0 * S1 + pd1 * S2 + 2 * sel;
      float2* const o = out + (size_t)row * L + (pd0 + 16 * pd1 + 4096 * (2 * sel));
      if (zap) {
#pragma unroll
        for (int i2 = 0; i2 < 16; i2++) {
          o[256 * i2] = make_float2(0.f, 0.f);
          o[256 * i2 + 4096] = make_float2(0.f, 0.f);
        }
      } else {
#pragma unroll
        for (int i2 = 0; i2 < 16; i2++) {
          const float4 q = *reinterpret_cast<const float4*>(buf + base + i2 * R3);
          o[256 * i2] = make_float2(q.x, q.y);          // k = d0 + 16 d1 + 256 d2 + 4096 d3, d3 = 2 sel
          o[256 * i2 + 4096] = make_float2(q.z, q.w);   // d3 = 2 sel + 1
          if constexpr (SK) {
            float2* const ca = reinterpret_cast<float2*>(colacc + base + i2 * R3);
            float2 acc = *ca;
            acc.x += q.x * q.x + q.y * q.y;
            acc.y += q.z * q.z + q.w * q.w;
            *ca = acc;
          }
        }
      }
    }
    __syncthreads();  // the row buffer is free: fetch the next row
    const unsigned nxt = row + gridDim.x;
    if (nxt < nrows && wid == 0) issue(nxt);
  }
  if constexpr (SK) {
    // column sums of the surviving rows this CTA transformed: one partial row per CTA, natural order
    __syncthreads();
    for (unsigned k = tid; k < (unsigned)L; k += NT) {
      const int idx = (k & 15) * S1 + ((k >> 4) & 15) * S2 + ((k >> 8) & 15) * R3 + (k >> 12);
      if (k < skp.ts_count) skp.partial[(size_t)blockIdx.x * skp.ts_count + k] = colacc[idx];
    }
  }
}

}  // namespace srtb_b200
