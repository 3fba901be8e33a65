// rub_kernels_tx.cuh — batched transmit waveform: framegen::write_sync_words
// (access codes, mimo/framing.cc:191-204) and framegen::assemble_mimo_packet (:210-235) for whole
// batches of frames on the GPU: symbol indices -> liquid square-QAM modulate -> carrier mapping ->
// IFFT -> dft_normalizer -> cyclic prefix.  Same stage code and operation order as the host
// framegen (rub_host.cpp), so the waveform is bit-identical.
#pragma once
#include <cuda_runtime.h>

#include "rub_fft.cuh"

namespace rub {

struct TxArgs {
  const unsigned char *tx_data;  // [n_frames][N][D][Mo] symbol indices
  cf *out;                       // sample n of (frame f, stream s) at out[f*frame_stride + s*stream_stride + n]
  long long frame_stride, stream_stride;
  const cf *tw;                  // packed stage twiddles
  const unsigned short *occ;     // occupied index j -> carrier k
  const unsigned char *scnull;   // [M] 1 = null carrier
  const float *sgn;              // [N][nac][M] access-code signs (0 on null carriers)
  const cf *s1;                  // [N][nac][M] time-domain access codes
  int n_frames, N, nac, T, D, M, Mo, cp, L, q, P;
  int comb;                      // RUB_EST_LS_COMB_INTERP training layout
  float alpha, dn, g1, gain;     // level spacing/2, 1/sqrt(Mo), sqrt(1/M), BASEBAND_GAIN
};

// one CTA of NT threads per (frame, OFDM symbol, stream)
template <int LOG2M>
__global__ void __launch_bounds__(FftPlan<LOG2M>::NT) k_framegen(TxArgs a) {
  using FF = Fft<LOG2M>;
  using PL = FftPlan<LOG2M>;
  using TW = FftTw<LOG2M>;
  constexpr int NT = FF::NT, M = FF::M, PAD = fft_padded_size(M);
  // two exchange buffers: X (natural-order input of stage 0) doubles as the stage-1 output B,
  // and the final natural-order result lands in A again
  extern __shared__ __align__(128) unsigned char smem_raw[];
  cf *X = reinterpret_cast<cf *>(smem_raw), *B = X;
  cf *A = X + PAD;
  const int tid = threadIdx.x;
  const int nsym = a.T + a.D;
  const long long fid = blockIdx.x;
  const int s = (int)(fid % a.N);
  const int sym = (int)((fid / a.N) % nsym);
  const long long frame = fid / ((long long)a.N * nsym);
  cf *dst = a.out + frame * a.frame_stride + (long long)s * a.stream_stride + (long long)sym * a.L;
  const bool training = sym < a.T;
  float scale = a.dn;
  if (training && !a.comb) {
    // TDMA access codes, code-major / stream-minor: only stream sym % N is active
    const int ac = sym / a.N, act = sym % a.N;
    const cf *src = a.s1 + ((long long)s * a.nac + ac) * M;
    for (int n = tid; n < a.L; n += NT) {
      const int i = n < a.cp ? M - a.cp + n : n - a.cp;
      dst[n] = s == act ? cscale(src[i], a.gain) : mk(0.f, 0.f);
    }
    return;
  }
  // frequency-domain symbol, conjugated: FFTW_BACKWARD = conj(fwd(conj(X)))
  if (training) {
    const float *sg = a.sgn + ((long long)s * a.nac + sym) * M;
    for (int k = tid; k < M; k += NT) X[k] = cconj(mk((k % a.P == s) ? sg[k] : 0.f, 0.f));
    scale = a.g1;
  } else {
    for (int k = tid; k < M; k += NT) X[k] = cconj(mk(0.f, 0.f));
    __syncthreads();
    const unsigned char *tx = a.tx_data + ((frame * a.N + s) * a.D + (sym - a.T)) * (long long)a.Mo;
    const int m = a.q / 2, P = 1 << m;
    for (int j = tid; j < a.Mo; j += NT) {
      const unsigned v = tx[j];
      const unsigned si = gray_decode(v >> m), sq = gray_decode(v & (unsigned)(P - 1));
      const cf p = mk((float)(2 * (int)si - P + 1) * a.alpha, (float)(2 * (int)sq - P + 1) * a.alpha);
      X[a.occ[j]] = cconj(p);
    }
  }
  __syncthreads();
  cf v[FF::PTS];
  FF::S0::template load<false>(tid, X, v);
  FF::S0::compute(tid, v, nullptr);
  FF::S0::template store<true, false>(tid, v, A, 1.f);
  __syncthreads();
  FF::S1::template load<true>(tid, A, v);
  FF::S1::compute(tid, v, a.tw + TW::OFF1);
  cf *R = nullptr;  // natural-order result
  if (PL::NSTG == 2) {
    FF::S1::template store<false, false>(tid, v, B, 1.f);
    R = B;
  } else {
    FF::S1::template store<true, false>(tid, v, B, 1.f);
    __syncthreads();
    FF::S2::template load<true>(tid, B, v);
    FF::S2::compute(tid, v, a.tw + TW::OFF2);
    FF::S2::template store<false, false>(tid, v, A, 1.f);
    R = A;
  }
  __syncthreads();
  // conj, dft_normalizer, BASEBAND_GAIN, cyclic prefix (mimo/framing.cc:225-232)
  for (int n = tid; n < a.L; n += NT) {
    const int i = n < a.cp ? M - a.cp + n : n - a.cp;
    dst[n] = cscale(cscale(cconj(R[i]), scale), a.gain);
  }
}

}  // namespace rub
