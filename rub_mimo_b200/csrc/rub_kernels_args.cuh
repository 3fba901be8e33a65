// rub_kernels_args.cuh — kernel argument blocks shared by the kernels and their host launchers.
#pragma once
#include "rub_fft.cuh"

namespace rub {

struct ChainArgs {
  // input
  const cf *iq;
  unsigned long long frame_stride, rx_stride, first_sample;
  const int *timing;         // [frame][rx][T] or null
  const int *payload_start;  // [frame] or null
  // geometry
  int M, cp, L, N, nac, D, T, Mo, q, P, n_frames, row_bytes;
  float dn, s_ls;
  unsigned flags;
  int estimator;
  // tables
  const cf *tw;            // packed stage twiddles
  const unsigned short *occ;  // j -> k
  const float *sgn;        // [tx][code][k] in {-1,0,+1}
  const unsigned char *scnull;  // [k] 1 = null carrier
  // scratch / outputs
  cf *Y;       // [frame][sym][rx][k]
  cf *G;       // [frame][rx][tx][k]
  cf *W;       // [frame][stream][rx][k]
  float *gain; // [frame][stream][k]
  float *isig; // [frame][stream][k]
  // wrec != 0: the W / gain / isig region (contiguous, starting at W) holds task records instead, one per (frame,
  // 64-carrier block, stream): W[rx][64] complex, gain[64], isig[64] = N*512 + 512 bytes, fetched by k_detect_lean
  // with one TMA bulk copy
  int wrec;
  cf *eq;
  float *llr;
  unsigned char *bits;
  unsigned char *rx_data;
  const unsigned char *tx_data;
  unsigned long long *counters;
};

// byte offset of carrier k's entry in the record region (see ChainArgs::wrec): W[s][r] for part = r < N, gain for
// part = N, isig for part = N + 1
RUB_HD size_t wrec_offset(int N, int M, long long frame, int s, int k, int part) {
  const size_t slot = (size_t)N * 512 + 512;
  const size_t rec = ((size_t)frame * (M >> 6) + (k >> 6)) * N + s;
  const size_t inner = part < N ? (size_t)part * 512 + (size_t)(k & 63) * 8 : (size_t)N * 512 + (size_t)(part - N) * 256 + (size_t)(k & 63) * 4;
  return rec * slot + inner;
}

struct FusedArgs {
  ChainArgs a;
  cf *scratchW;      // [grid][N*N][M]   (G accumulates here, then W in place)
  float *scratchG;   // [grid][2][N][M]  gain, isig
  int llr_stage_bytes;  // 256*q
  WeightMode wm;
  // warp-specialised kernel only (rub_kernels_ws.cuh)
  cf *scratchAcc;            // [grid][N*N][M] LS estimate of the next frame, written by the FFT warps
  const unsigned char *sgn8; // [tx][code][M/R2] sign bits of the access codes in last-stage register order
};

}  // namespace rub
