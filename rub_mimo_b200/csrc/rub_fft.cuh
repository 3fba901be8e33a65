// rub_fft.cuh — thread-level pieces of the batched OFDM FFT (replaces fftwf_execute at
// mimo/framing.cc:560 and the volk scale at :561).
//
// Algorithm: Stockham autosort, decimation in time, radix plan {16,8} per size (same plan and
// operation order as oracle/rub_oracle.c restates).  One FFT of size M is computed by NT
// cooperating threads, each holding P = M/NT points in registers per stage; stages exchange
// data through shared memory.  A stage is split into load / compute / store so callers decide
// where barriers go (ping-pong buffers in the staged kernel, in-place in the fused kernel).
//
// Shared-memory layout between stages is padded by one element every 16 (pad_idx) so that the
// stride-R stores of the first stage are bank-conflict free for 8-byte elements.
#pragma once
#include "rub_arith.cuh"

namespace rub {

template <int LOG2M> struct FftPlan;
template <> struct FftPlan<6>  { static constexpr int M = 64,   NSTG = 2, R0 = 8,  R1 = 8,  R2 = 1,  NT = 8; };
template <> struct FftPlan<7>  { static constexpr int M = 128,  NSTG = 2, R0 = 16, R1 = 8,  R2 = 1,  NT = 8; };
template <> struct FftPlan<8>  { static constexpr int M = 256,  NSTG = 2, R0 = 16, R1 = 16, R2 = 1,  NT = 16; };
template <> struct FftPlan<9>  { static constexpr int M = 512,  NSTG = 3, R0 = 8,  R1 = 8,  R2 = 8,  NT = 64; };
template <> struct FftPlan<10> { static constexpr int M = 1024, NSTG = 3, R0 = 16, R1 = 8,  R2 = 8,  NT = 64; };
template <> struct FftPlan<11> { static constexpr int M = 2048, NSTG = 3, R0 = 16, R1 = 16, R2 = 8,  NT = 128; };
template <> struct FftPlan<12> { static constexpr int M = 4096, NSTG = 3, R0 = 16, R1 = 16, R2 = 16, NT = 256; };

RUB_HD int pad_idx(int i) { return i + (i >> 4); }
#if defined(__CUDACC__)
__host__ __device__
#endif
constexpr int fft_padded_size(int M) { return M + (M >> 4); }

// number of stage-twiddle entries of a plan and per-stage offsets into the packed table
// stage s (s >= 1) table: tw_s[(t-1)*Ns + k] = master[t*k*(M/(Ns*R))], t = 1..R-1, k < Ns
template <int LOG2M>
struct FftTw {
  using P = FftPlan<LOG2M>;
  static constexpr int NS1 = P::R0;
  static constexpr int NS2 = P::R0 * P::R1;
  static constexpr int OFF1 = 0;
  static constexpr int CNT1 = (P::R1 - 1) * NS1;
  static constexpr int OFF2 = CNT1;
  static constexpr int CNT2 = (P::NSTG > 2) ? (P::R2 - 1) * NS2 : 0;
  static constexpr int TOTAL = CNT1 + CNT2;
};

#if defined(__CUDA_ARCH__)
RUB_HD cf ld_tw(const cf *p) {
  float2 t = __ldg(reinterpret_cast<const float2 *>(p));
  return mk(t.x, t.y);
}
#else
RUB_HD cf ld_tw(const cf *p) { return *p; }
#endif

// One stage for thread `tid` (0..NT-1) of an M-point FFT: radix R, stride NS (product of the
// previous radices).  v holds P = M/NT values: butterfly b uses v[b*R .. b*R+R-1].
template <int M, int R, int NS, int NT>
struct FftStage {
  static constexpr int Q = M / R;   // butterflies per FFT
  static constexpr int P = M / NT;  // points per thread
  static constexpr int B = P / R;   // butterflies per thread
  static_assert(B >= 1 && B * R == P, "plan");

  // Padded index of element idx0 + t*STRIDE: when STRIDE is a multiple of 16, or the whole
  // butterfly lives inside one 16-element pad group (NS == 1), the pad term does not depend
  // on t and the per-element address is base + t*const (an immediate offset in SASS).
  template <bool PAD, int STRIDE>
  RUB_HD static int elem(int base_padded, int idx0, int t) {
    if (!PAD) return idx0 + t * STRIDE;
    if (STRIDE % 16 == 0) return base_padded + t * (STRIDE + STRIDE / 16);
    if (STRIDE == 1 && R <= 16) return base_padded + t;
    return pad_idx(idx0 + t * STRIDE);
  }

  template <bool PAD_IN>
  RUB_HD static void load(int tid, const cf *in, cf *v) {
#pragma unroll
    for (int b = 0; b < B; b++) {
      const int j = tid + b * NT;
      const int bp = PAD_IN ? pad_idx(j) : j;
#pragma unroll
      for (int t = 0; t < R; t++) v[b * R + t] = in[elem<PAD_IN, Q>(bp, j, t)];
    }
  }
  // twiddle (skipped for the first stage, NS == 1) + butterfly.  TW_DIRECT: tws is dereferenced
  // directly (a shared-memory copy of the table) instead of through the read-only global path
  template <bool TW_DIRECT = false>
  RUB_HD static void compute(int tid, cf *v, const cf *tws) {
#pragma unroll
    for (int b = 0; b < B; b++) {
      const int j = tid + b * NT;
      if (NS > 1) {
        const int k = j % NS;
#pragma unroll
        for (int t = 1; t < R; t++)
          v[b * R + t] = cmul(v[b * R + t], TW_DIRECT ? tws[(t - 1) * NS + k] : ld_tw(tws + (t - 1) * NS + k));
      }
      bfly<R>(v + b * R);
    }
  }
  // The stage twiddles of thread `tid` depend on tid only, so a persistent thread can keep them
  // in registers: tw[b*(R-1) + t-1] = tws[(t-1)*NS + (tid + b*NT) % NS]
  static constexpr int NTW = (NS > 1) ? B * (R - 1) : 0;
  RUB_HD static void load_twiddles(int tid, const cf *tws, cf *tw) {
#pragma unroll
    for (int b = 0; b < B; b++) {
      const int k = (tid + b * NT) % NS;
#pragma unroll
      for (int t = 1; t < R; t++) tw[b * (R - 1) + t - 1] = tws[(t - 1) * NS + k];
    }
  }
  // same arithmetic as compute(), twiddles taken from the caller's registers
  RUB_HD static void compute_pre(cf *v, const cf *tw) {
#pragma unroll
    for (int b = 0; b < B; b++) {
      if (NS > 1) {
#pragma unroll
        for (int t = 1; t < R; t++) v[b * R + t] = cmul(v[b * R + t], tw[b * (R - 1) + t - 1]);
      }
      bfly<R>(v + b * R);
    }
  }
  template <bool PAD_OUT, bool SCALE>
  RUB_HD static void store(int tid, const cf *v, cf *out, float scale) {
#pragma unroll
    for (int b = 0; b < B; b++) {
      const int j = tid + b * NT;
      const int k = j % NS;
      const int base = (j / NS) * NS * R + k;
      const int bp = PAD_OUT ? pad_idx(base) : base;
#pragma unroll
      for (int t = 0; t < R; t++) {
        cf val = v[b * R + t];
        if (SCALE) val = cscale(val, scale);
        out[elem<PAD_OUT, NS>(bp, base, t)] = val;
      }
    }
  }
};

template <int LOG2M>
struct Fft {
  using P = FftPlan<LOG2M>;
  using T = FftTw<LOG2M>;
  static constexpr int M = P::M, NT = P::NT, PTS = M / NT;
  using S0 = FftStage<M, P::R0, 1, NT>;
  using S1 = FftStage<M, P::R1, P::R0, NT>;
  using S2 = FftStage<M, (P::NSTG > 2 ? P::R2 : 8), P::R0 * P::R1, NT>;
};

// Host replay of the device stages for one FFT (used by the transmit-side helpers and by the
// arithmetic tests).  in/out natural order, tws = packed stage-twiddle table.
template <int LOG2M>
inline void fft_host(const cf *in, cf *out, const cf *tws, bool do_scale, float scale) {
  using F = Fft<LOG2M>;
  using P = typename F::P;
  using T = typename F::T;
  constexpr int M = F::M, NT = F::NT, PTS = F::PTS;
  cf bufA[fft_padded_size(M)], bufB[fft_padded_size(M)];
  cf v[PTS];
  for (int tid = 0; tid < NT; tid++) {
    F::S0::template load<false>(tid, in, v);
    F::S0::compute(tid, v, nullptr);
    F::S0::template store<true, false>(tid, v, bufA, 1.f);
  }
  if (P::NSTG == 2) {
    for (int tid = 0; tid < NT; tid++) {
      F::S1::template load<true>(tid, bufA, v);
      F::S1::compute(tid, v, tws + T::OFF1);
      if (do_scale) F::S1::template store<false, true>(tid, v, out, scale);
      else F::S1::template store<false, false>(tid, v, out, scale);
    }
  } else {
    for (int tid = 0; tid < NT; tid++) {
      F::S1::template load<true>(tid, bufA, v);
      F::S1::compute(tid, v, tws + T::OFF1);
      F::S1::template store<true, false>(tid, v, bufB, 1.f);
    }
    for (int tid = 0; tid < NT; tid++) {
      F::S2::template load<true>(tid, bufB, v);
      F::S2::compute(tid, v, tws + T::OFF2);
      if (do_scale) F::S2::template store<false, true>(tid, v, out, scale);
      else F::S2::template store<false, false>(tid, v, out, scale);
    }
  }
}

}  // namespace rub
