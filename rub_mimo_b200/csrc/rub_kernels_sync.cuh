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
// a = code*N + tx); thread = candidate offset i in [0, L).  corr(i) = |sum_n w[base+i+n] *
// conj(tpl[n])|^2; the first maximum wins (strict '>' in ascending i, framing.cc:717, :736).
__global__ void __launch_bounds__(256) k_timing_search(const cf *__restrict__ window, unsigned long long wlen,
                                                       const cf *__restrict__ s1, const cf *__restrict__ s0, int M,
                                                       int L, int N, int nac, int *__restrict__ corr_indices,
                                                       int *__restrict__ s0_index) {
  extern __shared__ __align__(16) unsigned char sm_raw[];
  cf *tpl = reinterpret_cast<cf *>(sm_raw);
  __shared__ float best_v[256];
  __shared__ int best_i[256];
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
  for (int i = threadIdx.x; i < M; i += blockDim.x) tpl[i] = t[i];
  __syncthreads();
  const cf *w = window + (size_t)r * wlen + base;
  float bv = -1.f;
  int bi = 0;
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    float ax = 0.f, ay = 0.f;
    const cf *p = w + i;
    for (int n = 0; n < M; n++) {
      const cf xv = p[n], tv = tpl[n];  // x * conj(t)
      ax = fmaf(xv.x, tv.x, ax); ax = fmaf(xv.y, tv.y, ax);
      ay = fmaf(xv.y, tv.x, ay); ay = fmaf(-xv.x, tv.y, ay);
    }
    const float v = ax * ax + ay * ay;
    if (v > bv) { bv = v; bi = i; }  // ascending i per thread: first maximum
  }
  best_v[threadIdx.x] = bv;
  best_i[threadIdx.x] = bi;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      const float ov = best_v[threadIdx.x + s];
      const int oi = best_i[threadIdx.x + s];
      if (ov > best_v[threadIdx.x] || (ov == best_v[threadIdx.x] && oi < best_i[threadIdx.x])) {
        best_v[threadIdx.x] = ov;
        best_i[threadIdx.x] = oi;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    // the reference starts from max = 0 and index 0 and only moves on a strictly larger value
    const int idx = best_v[0] > 0.f ? best_i[0] : 0;
    if (slot == 0) s0_index[r] = idx;
    else corr_indices[r * max_ac + slot - 1] = (int)(base + idx) * (best_v[0] > 0.f ? 1 : 0);
  }
}

}  // namespace rub
