// rub_kernels_sync.cuh — synchronisation rows f1/f2 of SURVEY.md 8: the Schmidl & Cox timing
// metric (mimo/framing.cc:626-637) and the access-code timing search (mimo/framing.cc:702-744).
// Both are the reference's real CPU hot spots (an O(M) dot product per sample, and
// symbol_len * N * (1 + nac*N) FFTs per frame); on the GPU they are brute-force data-parallel.
#pragma once
#include <cuda_runtime.h>

#include "rub_arith.cuh"

namespace rub {

// One thread per output sample n.  The CTA stages x[n0 - (M + M/2) .. n0 + blockDim) in shared
// memory; each thread then evaluates
//   P[n] = sum_{u=n-M/2+1..n} -1 * conj(x[u-M/2]) * x[u]      (firfilt_crcf, taps -1.0, :342)
//   R[n] = sum_{u=n-M+1..n}   0.5 * |x[u]|^2                   (firfilt_rrrf, taps 0.5,  :344)
// oldest to newest — the oracle's summation order, so y = |P|^2 / R^2 is bit-identical.
__global__ void __launch_bounds__(256) k_sc_metric(const cf *__restrict__ x, unsigned long long n_total, int M,
                                                   float *__restrict__ y) {
  extern __shared__ __align__(16) unsigned char sm_raw[];
  cf *xs = reinterpret_cast<cf *>(sm_raw);
  const int M2 = M / 2, halo = M + M2;
  const long long n0 = (long long)blockIdx.x * blockDim.x;
  for (int i = threadIdx.x; i < halo + (int)blockDim.x; i += blockDim.x) {
    const long long g = n0 - halo + i;
    xs[i] = (g >= 0 && g < (long long)n_total) ? x[g] : mk(0.f, 0.f);
  }
  __syncthreads();
  const long long n = n0 + threadIdx.x;
  if (n >= (long long)n_total) return;
  const int c = halo + threadIdx.x;  // position of sample n in xs
  cf P = mk(0.f, 0.f);
  float R = 0.f;
  // windows are clipped at the start of the capture exactly as the oracle clips them
  const int lenP = (int)((n + 1 < M2) ? n + 1 : M2), lenR = (int)((n + 1 < M) ? n + 1 : M);
  for (int i = lenP - 1; i >= 0; i--) {
    const cf xv = xs[c - i], d = xs[c - i - M2];
    const float cdx = d.x, cdy = -d.y;
    const float px = cdx * xv.x - cdy * xv.y, py = cdx * xv.y + cdy * xv.x;
    P.x += -1.0f * px;
    P.y += -1.0f * py;
  }
  for (int i = lenR - 1; i >= 0; i--) {
    const cf xv = xs[c - i];
    const float pw = xv.x * xv.x + xv.y * xv.y;
    R += 0.5f * pw;
  }
  y[n] = (P.x * P.x + P.y * P.y) / (R * R);
}

// Timing search: blockIdx.y = rx * (nac*N + 1) + code slot (slot 0 = S0, slot a+1 = access code
// a = code*N + tx); blockIdx.x = tile of 256 candidate offsets, thread = one offset i in [0, L).
// corr(i) = |sum_n w[base+i+n] * conj(tpl[n])|^2 with the template and the M+256 window samples
// of the tile staged in shared memory.  The winner of (rx, slot) is kept as a 64-bit key
// (corr bits << 32 | ~i) under atomicMax: the largest correlation wins and, among equals, the
// smallest offset — the reference's strict '>' scan in ascending i (framing.cc:717, :736).
// Key 0 (nothing above 0) leaves the reference's initial index 0.
// blockIdx.z = frame of a batch: its window starts win_off[z] samples into every rx row (rows are
// rx_stride apart); keys are [frame][rx][slot].
__global__ void __launch_bounds__(256) k_timing_search(const cf *__restrict__ window, unsigned long long wlen,
                                                       unsigned long long rx_stride, const long long *__restrict__ win_off,
                                                       const cf *__restrict__ s1, const cf *__restrict__ s0, int M,
                                                       int L, int N, int nac, unsigned long long *__restrict__ keys) {
  extern __shared__ __align__(16) unsigned char sm_raw[];
  cf *tpl = reinterpret_cast<cf *>(sm_raw);  // [M]
  cf *xs = tpl + M;                          // [M + 256]
  const int max_ac = nac * N, slots = max_ac + 1;
  const int r = blockIdx.y / slots, slot = blockIdx.y % slots;
  const cf *t = nullptr;
  long long base;
  if (slot == 0) { t = s0; base = 0; }
  else {
    const int ac = slot - 1, code = ac / N, tx = ac % N;
    t = s1 + ((size_t)tx * nac + code) * M;
    base = (long long)L * (ac + 1);
  }
  if (slot == 0 && !s0) return;
  const int i0 = blockIdx.x * 256;
  const cf *w = window + (size_t)r * rx_stride + (win_off ? win_off[blockIdx.z] : 0) + base + i0;
  const long long avail = (long long)wlen - base - i0;  // samples of this row from w on
  for (int i = threadIdx.x; i < M; i += 256) tpl[i] = t[i];
  for (int i = threadIdx.x; i < M + 256; i += 256) xs[i] = i < avail ? w[i] : mk(0.f, 0.f);
  __syncthreads();
  const int i = i0 + threadIdx.x;
  if (i >= L) return;
  float ax = 0.f, ay = 0.f;
  const cf *p = xs + threadIdx.x;
#pragma unroll 4
  for (int n = 0; n < M; n++) {
    const cf xv = p[n], tv = tpl[n];  // x * conj(t)
    ax = fmaf(xv.x, tv.x, ax); ax = fmaf(xv.y, tv.y, ax);
    ay = fmaf(xv.y, tv.x, ay); ay = fmaf(-xv.x, tv.y, ay);
  }
  const float v = ax * ax + ay * ay;
  if (v > 0.f) {
    const unsigned long long key = ((unsigned long long)__float_as_uint(v) << 32) | (unsigned long long)(0xffffffffu - (unsigned)i);
    atomicMax(keys + (size_t)blockIdx.z * gridDim.y + blockIdx.y, key);
  }
}

}  // namespace rub
