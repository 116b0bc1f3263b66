// fft_engine.cuh — hand-written sm_100a FFT passes (no cuFFT on the hot path).
//
// Replaces the vendor-FFT call sites of the reference:
//   R2C  : userspace/include/srtb/fft/cufft_like_wrapper.hpp:183-191 (cufftExecR2C)
//   C2C  : userspace/include/srtb/fft/cufft_like_wrapper.hpp:205-207 (batched, waterfall)
// and the in-tree fallback userspace/include/srtb/fft/naive_fft.hpp:155-176,221-261,
// whose transform definition (unnormalised, forward = e^{-2 pi i nk/N}) is kept.
//
// One kernel template implements one "pass": a Stockham auto-sort FFT of length
// L = 2^LOGL on a tile of T independent sequences held in shared memory, eight points
// per thread per stage in registers (radix 8/4/2 butterflies), the first stage loading
// straight from HBM and the last stage storing straight to HBM.
//   MODE_ROW   : sequences are contiguous rows (lanes run along the FFT index)
//   MODE_COL   : sequences are strided columns, T neighbours contiguous (lanes run along T)
//   MODE_TRANS : loads like ROW, stores like COL (the transposing last pass of a
//                multi-pass transform, which produces natural order)
// A length-n transform with n > 4096 is n = L1*L2(*L3): COL pass(es) with the inter-pass
// twiddle fused into the store, then one TRANS pass (see fft_plan in srtb_b200_fft.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace srtb_b200 {

enum { MODE_ROW = 0, MODE_COL = 1, MODE_TRANS = 2 };

__device__ __forceinline__ float2 c_add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 c_sub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 c_mul(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ float2 c_sqr(float2 a) {
  return make_float2(fmaf(a.x, a.x, -a.y * a.y), 2.0f * a.x * a.y);
}
// multiply by -i (forward) / +i (backward)
template <bool FWD>
__device__ __forceinline__ float2 c_rot(float2 a) {
  return FWD ? make_float2(a.y, -a.x) : make_float2(-a.y, a.x);
}
template <bool FWD>
__device__ __forceinline__ float2 c_dir(float2 w) {  // table holds forward twiddles
  return FWD ? w : make_float2(w.x, -w.y);
}

// ---- in-register DFTs, natural-order in and out -----------------------------------
template <bool FWD>
__device__ __forceinline__ void dft2(float2& a, float2& b) {
  const float2 t = a;
  a = c_add(t, b);
  b = c_sub(t, b);
}
template <bool FWD>
__device__ __forceinline__ void dft4(float2& a0, float2& a1, float2& a2, float2& a3) {
  const float2 t0 = c_add(a0, a2), t1 = c_sub(a0, a2);
  const float2 t2 = c_add(a1, a3), t3 = c_rot<FWD>(c_sub(a1, a3));
  a0 = c_add(t0, t2);
  a1 = c_add(t1, t3);
  a2 = c_sub(t0, t2);
  a3 = c_sub(t1, t3);
}
template <bool FWD>
__device__ __forceinline__ void dft8(float2& v0, float2& v1, float2& v2, float2& v3, float2& v4,
                                     float2& v5, float2& v6, float2& v7) {
  dft4<FWD>(v0, v2, v4, v6);  // E0..E3 in v0,v2,v4,v6
  dft4<FWD>(v1, v3, v5, v7);  // O0..O3 in v1,v3,v5,v7
  constexpr float c = 0.70710678118654752440f;
  // O1 *= W8^1, O2 *= W8^2, O3 *= W8^3
  const float2 o1 = FWD ? make_float2(c * (v3.x + v3.y), c * (v3.y - v3.x))
                        : make_float2(c * (v3.x - v3.y), c * (v3.x + v3.y));
  const float2 o2 = c_rot<FWD>(v5);
  const float2 o3 = FWD ? make_float2(c * (v7.y - v7.x), -c * (v7.x + v7.y))
                        : make_float2(-c * (v7.x + v7.y), c * (v7.x - v7.y));
  const float2 e0 = v0, e1 = v2, e2 = v4, e3 = v6, o0 = v1;
  v0 = c_add(e0, o0);
  v4 = c_sub(e0, o0);
  v1 = c_add(e1, o1);
  v5 = c_sub(e1, o1);
  v2 = c_add(e2, o2);
  v6 = c_sub(e2, o2);
  v3 = c_add(e3, o3);
  v7 = c_sub(e3, o3);
}

// ---- stage schedule: radix-8 stages first, remainder (radix 4 or 2) last ------------
template <int LOGL>
struct sched {
  static constexpr int S = (LOGL + 2) / 3;
  __host__ __device__ static constexpr int logr(int s) { return (s < S - 1) ? 3 : (LOGL - 3 * (S - 1)); }
  __host__ __device__ static constexpr int logns(int s) { return 3 * s; }  // all stages before s are radix 8
};

// One Stockham stage on the eight values a thread holds.
// Slot e of v[] is the point read at index u + e*U (U = L/8). A radix-r stage treats the
// slots as NB = 8/r butterflies: butterfly m uses slots m + i*NB (i < r) and is butterfly
// number u + m*U of the L/r butterflies in the sequence. Twiddle (DIT) W_{Ns*r}^{k*i},
// k = butterfly % Ns; outputs go to (b / Ns)*Ns*r + k + i*Ns.
template <int LOGL, int LOGR, int LOGNS, bool FWD>
__device__ __forceinline__ void stage_compute(float2 (&v)[8], int u, const float2* __restrict__ tw,
                                              int (&oidx)[8]) {
  constexpr int L = 1 << LOGL, U = L / 8, R = 1 << LOGR, NB = 8 / R, NS = 1 << LOGNS;
#pragma unroll
  for (int m = 0; m < NB; m++) {
    const int b = u + m * U;
    const int k = b & (NS - 1);
    if constexpr (LOGNS > 0) {
      // w1 = W_{Ns*R}^k from the length-L forward table; higher powers by multiplication
      const float2 w1 = c_dir<FWD>(__ldg(&tw[k << (LOGL - LOGNS - LOGR)]));
      if (R == 2) {
        v[m + NB] = c_mul(v[m + NB], w1);
      } else if (R == 4) {
        const float2 w2 = c_sqr(w1), w3 = c_mul(w2, w1);
        v[m + NB] = c_mul(v[m + NB], w1);
        v[m + 2 * NB] = c_mul(v[m + 2 * NB], w2);
        v[m + 3 * NB] = c_mul(v[m + 3 * NB], w3);
      } else {
        const float2 w2 = c_sqr(w1), w3 = c_mul(w2, w1), w4 = c_sqr(w2);
        const float2 w5 = c_mul(w4, w1), w6 = c_sqr(w3), w7 = c_mul(w4, w3);
        v[m + 1] = c_mul(v[m + 1], w1);
        v[m + 2] = c_mul(v[m + 2], w2);
        v[m + 3] = c_mul(v[m + 3], w3);
        v[m + 4] = c_mul(v[m + 4], w4);
        v[m + 5] = c_mul(v[m + 5], w5);
        v[m + 6] = c_mul(v[m + 6], w6);
        v[m + 7] = c_mul(v[m + 7], w7);
      }
    }
    if (R == 2) {
      dft2<FWD>(v[m], v[m + NB]);
    } else if (R == 4) {
      dft4<FWD>(v[m], v[m + NB], v[m + 2 * NB], v[m + 3 * NB]);
    } else {
      dft8<FWD>(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
    }
    const int obase = ((b >> LOGNS) << (LOGNS + LOGR)) + k;
#pragma unroll
    for (int i = 0; i < R; i++) oidx[m + i * NB] = obase + i * NS;
  }
}

// shared-memory layout of the tile
template <int LOGL, int T, int MODE>
struct tile_layout {
  static constexpr int L = 1 << LOGL;
  static constexpr int LPAD = L + (L >> 4);
  static constexpr int ELEMS = (MODE == MODE_ROW) ? T * LPAD : T * L;
  __device__ __forceinline__ static int at(int idx, int t) {
    if (MODE == MODE_ROW) return t * LPAD + idx + (idx >> 4);
    if (MODE == MODE_COL) return idx * T + t;
    return idx * T + ((t + (idx >> 3)) & (T - 1));  // TRANS: rotate so both maps are conflict-free
  }
};

template <int LOGL, int T>
struct pass_threads {
  static constexpr int value = ((1 << LOGL) / 8) * T;
};

// IO concept:
//   bool  IO::tile_valid(int t)                     sequence t of this tile exists
//   float2 IO::load(int t, int pos)                  point `pos` of sequence t
//   void  IO::store8(int t, int u, int U, float2 (&v)[8])   outputs k = u + e*U, e = 0..7
template <int LOGL, int T, int MODE, bool FWD, class IO>
__global__ void __launch_bounds__(pass_threads<LOGL, T>::value)
    fft_pass_kernel(IO io, const float2* __restrict__ tw) {
  using SC = sched<LOGL>;
  using LAY = tile_layout<LOGL, T, MODE>;
  constexpr int L = 1 << LOGL, U = L / 8, S = SC::S;
  extern __shared__ float2 sm[];
  io.init(blockIdx.x, sm + LAY::ELEMS);
  const int tid = threadIdx.x;
  int t0, u0, t1, u1;
  if (MODE == MODE_ROW) {
    t0 = t1 = tid / U;
    u0 = u1 = tid % U;
  } else if (MODE == MODE_COL) {
    t0 = t1 = tid % T;
    u0 = u1 = tid / T;
  } else {
    t0 = tid / U;
    u0 = tid % U;
    t1 = tid % T;
    u1 = tid / T;
  }
  float2 v[8];
  int oidx[8];
  const bool valid0 = io.tile_valid(t0);
#pragma unroll
  for (int e = 0; e < 8; e++) v[e] = valid0 ? io.load(t0, u0 + e * U) : make_float2(0.f, 0.f);

  stage_compute<LOGL, SC::logr(0), 0, FWD>(v, u0, tw, oidx);
  if constexpr (S == 1) {
    if (valid0) io.store8(t0, u0, U, v);
    return;
  } else {
#pragma unroll
  for (int e = 0; e < 8; e++) sm[LAY::at(oidx[e], t0)] = v[e];
  __syncthreads();

  if constexpr (S >= 3) {
#pragma unroll
    for (int e = 0; e < 8; e++) v[e] = sm[LAY::at(u1 + e * U, t1)];
    __syncthreads();
    stage_compute<LOGL, SC::logr(1), SC::logns(1), FWD>(v, u1, tw, oidx);
#pragma unroll
    for (int e = 0; e < 8; e++) sm[LAY::at(oidx[e], t1)] = v[e];
    __syncthreads();
  }
  if constexpr (S >= 4) {
#pragma unroll
    for (int e = 0; e < 8; e++) v[e] = sm[LAY::at(u1 + e * U, t1)];
    __syncthreads();
    stage_compute<LOGL, SC::logr(2), SC::logns(2), FWD>(v, u1, tw, oidx);
#pragma unroll
    for (int e = 0; e < 8; e++) sm[LAY::at(oidx[e], t1)] = v[e];
    __syncthreads();
  }
  // last stage
#pragma unroll
  for (int e = 0; e < 8; e++) v[e] = sm[LAY::at(u1 + e * U, t1)];
  stage_compute<LOGL, SC::logr(S - 1), SC::logns(S - 1), FWD>(v, u1, tw, oidx);
  if (io.tile_valid(t1)) io.store8(t1, u1, U, v);
  }
}

// ---------------------------------------------------------------------------------
// IO functors
// ---------------------------------------------------------------------------------

// contiguous rows, in place or out of place; T rows per CTA
template <int LOGL, int T>
struct row_io {
  const float2* in;
  float2* out;
  size_t nrows;
  size_t row0;
  __device__ __forceinline__ void init(unsigned block, float2*) { row0 = (size_t)block * T; }
  __device__ __forceinline__ bool tile_valid(int t) const { return row0 + t < nrows; }
  __device__ __forceinline__ float2 load(int t, int pos) const {
    return in[((row0 + t) << LOGL) + pos];
  }
  __device__ __forceinline__ void store8(int t, int u, int U, float2 (&v)[8]) const {
    float2* o = out + ((row0 + t) << LOGL) + u;
#pragma unroll
    for (int e = 0; e < 8; e++) o[e * U] = v[e];
  }
};

// three-level twiddle table for W_n^idx, idx < n <= 2^30: idx = a*2^(2q) + b*2^q + c
struct big_twiddle {
  const float2* tab;  // [3][1 << q]: W^(c), W^(b << q), W^(a << 2q)
  int q;
};
__device__ __forceinline__ float2 big_tw_lookup(const float2* s, int q, uint32_t idx) {
  const uint32_t mask = (1u << q) - 1u;
  const float2 w0 = s[idx & mask];
  const float2 w1 = s[(1u << q) + ((idx >> q) & mask)];
  const float2 w2 = s[(2u << q) + (idx >> (2 * q))];
  return c_mul(c_mul(w2, w1), w0);
}

// column pass of a multi-pass transform: view [A][L][B], FFT along L for T adjacent b.
// store multiplies by the inter-pass twiddle W_{L*B}^{k*b} (forward table, conj if !FWD).
template <int LOGL, int T, bool FWD>
struct col_io {
  const float2* in;
  float2* out;
  size_t B;        // elements between consecutive FFT points
  uint32_t btiles; // B / T
  big_twiddle btw;
  size_t base;
  uint32_t b0;
  const float2* stw;
  __device__ __forceinline__ void init(unsigned block, float2* smem_extra) {
    const uint32_t a = block / btiles;
    b0 = (block % btiles) * T;
    base = ((size_t)a << LOGL) * B + b0;
    // stage the three small twiddle tables in shared memory
    float2* s = smem_extra;
    const int n = 3 << btw.q;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s[i] = __ldg(&btw.tab[i]);
    stw = s;
    __syncthreads();
  }
  __device__ __forceinline__ bool tile_valid(int) const { return true; }
  __device__ __forceinline__ float2 load(int t, int pos) const { return in[base + (size_t)pos * B + t]; }
  __device__ __forceinline__ void store8(int t, int u, int U, float2 (&v)[8]) const {
    const uint32_t b = b0 + t;
    // twiddles W^{(u + e*U) * b}: base and ratio from the tables, powers by multiplication
    float2 wb = big_tw_lookup(stw, btw.q, (uint32_t)u * b);
    float2 r1 = big_tw_lookup(stw, btw.q, (uint32_t)U * b);
    if (!FWD) {
      wb.y = -wb.y;
      r1.y = -r1.y;
    }
    const float2 r2 = c_sqr(r1), r4 = c_sqr(r2);
    float2 w[8];
    w[0] = wb;
    w[1] = c_mul(wb, r1);
    w[2] = c_mul(wb, r2);
    w[3] = c_mul(w[1], r2);
    w[4] = c_mul(wb, r4);
    w[5] = c_mul(w[1], r4);
    w[6] = c_mul(w[2], r4);
    w[7] = c_mul(w[3], r4);
    float2* o = out + base + (size_t)u * B + t;
#pragma unroll
    for (int e = 0; e < 8; e++) o[(size_t)e * U * B] = c_mul(v[e], w[e]);
  }
};

// transposing last pass: rows [beta][k1][rest][L] -> natural order
//   in  row = beta*A + k1*S + rest            (A = n / L rows per transform, S = A / L1)
//   out idx = beta*n + k1 + L1*rest + A*k     (k = output index of this pass)
template <int LOGL, int T>
struct trans_io {
  const float2* in;
  float2* out;
  uint32_t A, S, L1;   // rows per transform, rest count, first-pass length
  uint32_t k1tiles;    // L1 / T
  size_t in_row0;      // row of t = 0
  size_t out0;         // out index of (t = 0, k = 0)
  __device__ __forceinline__ void init(unsigned block, float2*) {
    const uint32_t k1t = block % k1tiles;
    const uint32_t r = block / k1tiles;
    const uint32_t rest = r % S;
    const uint32_t beta = r / S;
    const uint32_t k1 = k1t * T;
    in_row0 = (size_t)beta * A + (size_t)k1 * S + rest;
    out0 = ((size_t)beta * A << LOGL) + k1 + (size_t)L1 * rest;
  }
  __device__ __forceinline__ bool tile_valid(int) const { return true; }
  __device__ __forceinline__ float2 load(int t, int pos) const {
    return in[((in_row0 + (size_t)t * S) << LOGL) + pos];
  }
  __device__ __forceinline__ void store8(int t, int u, int U, float2 (&v)[8]) const {
    float2* o = out + out0 + t + (size_t)A * u;
#pragma unroll
    for (int e = 0; e < 8; e++) o[(size_t)A * e * U] = v[e];
  }
};

}  // namespace srtb_b200
