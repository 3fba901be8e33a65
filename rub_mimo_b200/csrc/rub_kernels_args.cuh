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
  cf *eq;
  float *llr;
  unsigned char *bits;
  unsigned char *rx_data;
  const unsigned char *tx_data;
  unsigned long long *counters;
};

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
