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
                                                       int L, int N, int nac, unsigned long long *__restrict__ keys,
                                                       float *__restrict__ corr_out, float corr_scale_s0) {
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
  // debug sink (corr_%d_%d.dat, framing.cc:873-883): the reference stores |X . conj(S)|^2 / M^2.  The time-domain
  // templates are IFFT(S) * g with g = 1/sqrt(M) for the access codes and 1/sqrt(M_S0) for S0, so that is
  // v / (g^2 M^2) = v / M, or v * M_S0 / M^2 for S0 (corr_scale_s0)
  if (corr_out) corr_out[((size_t)blockIdx.z * gridDim.y + blockIdx.y) * L + i] = slot == 0 ? v * corr_scale_s0 : v / (float)M;
  if (v > 0.f) {
    const unsigned long long key = ((unsigned long long)__float_as_uint(v) << 32) | (unsigned long long)(0xffffffffu - (unsigned)i);
    atomicMax(keys + (size_t)blockIdx.z * gridDim.y + blockIdx.y, key);
  }
}

// ---------------------------------------------------------------------------------------------
// Schmidl & Cox metric as sliding sums (row f2 of SURVEY.md 8): P(n) = P(n-1) + c(n) - c(n - M/2),
// R(n) = R(n-1) + e(n) - e(n - M), c(u) = -conj(x[u - M/2]) x[u], e(u) = 0.5 |x[u]|^2.  A CTA owns a tile of
// 256 * PER outputs of one stream (blockIdx.y): it sums the windows ending just before the tile directly
// (block reduction), then scans the per-sample differences (thread-local prefix + block scan), so the work is
// O(1) per sample and the rounding error cannot grow beyond a tile.  The sums are added in another order than
// liquid's firfilt dot products, so y differs from k_sc_metric in the last bits: the plateau start can move
// where y sits within rounding of the threshold.  k_sc_metric is the form that reproduces the reference's
// threshold crossings bit for bit; this one is the fast path of rub_rx_process_capture.
template <int PER>
__global__ void __launch_bounds__(256) k_sc_metric_scan(const cf *__restrict__ x, unsigned long long n_total,
                                                        unsigned long long row_stride, int M, float *__restrict__ y) {
  extern __shared__ __align__(16) unsigned char sm_raw[];
  cf *xs = reinterpret_cast<cf *>(sm_raw);
  __shared__ float red[3][8];
  __shared__ float base[3];
  constexpr int TILE = 256 * PER;
  const int M2 = M / 2, halo = M + M2, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const long long n0 = (long long)blockIdx.x * TILE;
  x += (size_t)blockIdx.y * row_stride;
  y += (size_t)blockIdx.y * row_stride;
  for (int i = tid; i < halo + TILE; i += 256) {
    const long long g = n0 - halo + i;
    xs[i] = (g >= 0 && g < (long long)n_total) ? x[g] : mk(0.f, 0.f);
  }
  __syncthreads();
  auto cterm = [&](int i) {  // c(u) for the sample at xs[i]
    const cf xv = xs[i], d = xs[i - M2];
    return mk(-(d.x * xv.x + d.y * xv.y), -(d.x * xv.y - d.y * xv.x));
  };
  auto eterm = [&](int i) { const cf xv = xs[i]; return 0.5f * (xv.x * xv.x + xv.y * xv.y); };
  // windows ending at n0 - 1: xs[halo - M2 .. halo) for P, xs[halo - M .. halo) for R
  float px = 0.f, py = 0.f, r = 0.f;
  for (int i = halo - M2 + tid; i < halo; i += 256) { const cf c = cterm(i); px += c.x; py += c.y; }
  for (int i = halo - M + tid; i < halo; i += 256) r += eterm(i);
  auto block_sum3 = [&](float &a, float &b, float &c) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    if (lane == 0) { red[0][wid] = a; red[1][wid] = b; red[2][wid] = c; }
    __syncthreads();
    if (tid < 3) { float t = 0.f; for (int w = 0; w < 8; w++) t += red[tid][w]; base[tid] = t; }
    __syncthreads();
  };
  block_sum3(px, py, r);
  const float bpx = base[0], bpy = base[1], br = base[2];
  __syncthreads();
  // per-thread inclusive prefix of the differences of its PER consecutive outputs
  float dpx[PER], dpy[PER], dr[PER];
  float tx = 0.f, ty = 0.f, tr = 0.f;
#pragma unroll
  for (int k = 0; k < PER; k++) {
    const int i = halo + tid * PER + k;
    const cf cn = cterm(i), co = cterm(i - M2);
    tx += cn.x - co.x; ty += cn.y - co.y; tr += eterm(i) - eterm(i - M);
    dpx[k] = tx; dpy[k] = ty; dr[k] = tr;
  }
  // exclusive block scan of the thread totals
  float sx = tx, sy = ty, sr = tr;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float ax = __shfl_up_sync(0xffffffffu, sx, o), ay = __shfl_up_sync(0xffffffffu, sy, o), ar = __shfl_up_sync(0xffffffffu, sr, o);
    if (lane >= o) { sx += ax; sy += ay; sr += ar; }
  }
  if (lane == 31) { red[0][wid] = sx; red[1][wid] = sy; red[2][wid] = sr; }
  __syncthreads();
  float ox = sx - tx, oy = sy - ty, orr = sr - tr;  // exclusive within the warp
  for (int w = 0; w < wid; w++) { ox += red[0][w]; oy += red[1][w]; orr += red[2][w]; }
#pragma unroll
  for (int k = 0; k < PER; k++) {
    const long long n = n0 + tid * PER + k;
    if (n < (long long)n_total) {
      const float Px = bpx + (ox + dpx[k]), Py = bpy + (oy + dpy[k]), R = br + (orr + dr[k]);
      y[n] = (Px * Px + Py * Py) / (R * R);
    }
  }
}

// ok[n] = 1 when y[n - cp - 1 .. n] are all above the threshold, i.e. the reference's plateau rule
// (framing.cc:601-616: in_plateau && plateau_end - plateau_start > cp_len) holds for this stream at sample n.
// Tile + look-back of cp + 1 samples per CTA; "index of the last sample at or below the threshold" is a block
// max-scan.  Samples before the capture count as below.
__global__ void __launch_bounds__(256) k_plateau_ok(const float *__restrict__ y, unsigned long long n_total,
                                                    unsigned long long row_stride, int cp, float threshold,
                                                    unsigned char *__restrict__ ok) {
  __shared__ long long wmax[8];
  constexpr int TILE = 1024;
  const int halo = cp + 1, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const long long n0 = (long long)blockIdx.x * TILE, first = n0 - halo;
  y += (size_t)blockIdx.y * row_stride;
  ok += (size_t)blockIdx.y * row_stride;
  const int total = halo + TILE, per = (total + 255) / 256;
  // thread-local last-below index over its contiguous chunk
  const int c0 = tid * per, c1 = min(total, c0 + per);
  long long last = first - 1;  // "the sample before the region is below": harmless, the region reaches cp + 1 back
  for (int i = c0; i < c1; i++) {
    const long long g = first + i;
    const bool above = g >= 0 && g < (long long)n_total && y[g] > threshold;
    if (!above) last = g;
  }
  // inclusive block max-scan of the chunk results
  long long s = last;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const long long a = __shfl_up_sync(0xffffffffu, s, o);
    if (lane >= o && a > s) s = a;
  }
  if (lane == 31) wmax[wid] = s;
  __syncthreads();
  long long before = __shfl_up_sync(0xffffffffu, s, 1);  // exclusive within the warp
  if (lane == 0) before = first - 1;
  for (int w = 0; w < wid; w++) before = max(before, wmax[w]);
  long long run = before;
  for (int i = c0; i < c1; i++) {
    const long long g = first + i;
    const bool above = g >= 0 && g < (long long)n_total && y[g] > threshold;
    if (!above) run = g;
    if (i >= halo && g < (long long)n_total) ok[g] = (g - run >= (long long)cp + 2) ? 1 : 0;
  }
}

// first[t] = first sample of tile t (1024 samples) at which every stream's plateau rule holds, or -1: lets the
// walk below skip the gaps between bursts a tile at a time
__global__ void __launch_bounds__(1024) k_plateau_first(const unsigned char *__restrict__ ok, long long n,
                                                        unsigned long long row_stride, int N, long long *__restrict__ first) {
  __shared__ long long s_min;
  const long long i = (long long)blockIdx.x * 1024 + threadIdx.x;
  if (threadIdx.x == 0) s_min = 0x7fffffffffffffffLL;
  __syncthreads();
  bool all = i < n;
  for (int s = 0; s < N && all; s++) all = ok[(size_t)s * row_stride + i] != 0;
  long long best = all ? i : 0x7fffffffffffffffLL;
  for (int o = 16; o; o >>= 1) { const long long t = __shfl_xor_sync(0xffffffffu, best, o); if (t < best) best = t; }
  if ((threadIdx.x & 31) == 0 && best != 0x7fffffffffffffffLL) atomicMin((unsigned long long *)&s_min, (unsigned long long)best);
  __syncthreads();
  if (threadIdx.x == 0) first[blockIdx.x] = s_min == 0x7fffffffffffffffLL ? -1 : s_min;
}

// The receive loop's state machine over a whole capture (framing.cc:591-651), one CTA: find the first sample
// at which every stream's plateau rule holds, take the streams' plateau starts (walking back over the metric),
// sync_index = their mean, skip the access codes and the payload, and search again behind the burst.
struct PlateauWalk {
  long long n, L, acb_len, tx_sig_len, Wlen;
  int N, cp;
  float threshold;
  unsigned max_frames;
};
__global__ void __launch_bounds__(1024) k_plateau_walk(const unsigned char *__restrict__ ok, const long long *__restrict__ first,
                                                       const float *__restrict__ y, unsigned long long row_stride, PlateauWalk w,
                                                       long long *__restrict__ win_off, unsigned long long *__restrict__ syncs,
                                                       unsigned *__restrict__ count) {
  __shared__ long long s_found, s_pos, s_pstart[8];
  __shared__ int s_done;
  const int tid = threadIdx.x;
  if (tid == 0) { s_pos = 0; s_done = 0; *count = 0; }
  __syncthreads();
  unsigned cnt = 0;
  while (cnt < w.max_frames) {
    const long long pos = s_pos;
    long long found = -1;
    const long long b0 = pos + w.cp + 1, ntiles = (w.n + 1023) / 1024;
    if (b0 >= w.n) break;
    // the rest of the tile the search starts in, sample by sample; then whole tiles through the first[] table
    for (int phase = 0; phase < 2 && found < 0; phase++) {
      const long long t0 = b0 / 1024;
      for (long long c0 = phase ? t0 + 1 : t0; c0 < (phase ? ntiles : t0 + 1) && found < 0; c0 += 1024) {
        if (tid == 0) s_found = 0x7fffffffffffffffLL;
        __syncthreads();
        long long best = 0x7fffffffffffffffLL;
        if (phase == 0) {
          const long long i = t0 * 1024 + tid;
          bool all = i >= b0 && i < w.n;
          for (int s = 0; s < w.N && all; s++) all = ok[(size_t)s * row_stride + i] != 0;
          if (all) best = i;
        } else {
          const long long t = c0 + tid;
          if (t < ntiles && first[t] >= 0) best = first[t];
        }
        for (int o = 16; o; o >>= 1) { const long long t = __shfl_xor_sync(0xffffffffu, best, o); if (t < best) best = t; }
        if ((tid & 31) == 0 && best != 0x7fffffffffffffffLL) atomicMin((unsigned long long *)&s_found, (unsigned long long)best);
        __syncthreads();
        if (s_found != 0x7fffffffffffffffLL) found = s_found;
        __syncthreads();
      }
    }
    if (found < 0) break;
    // plateau start of every stream: the run of above-threshold samples that contains found, clipped at pos
    if (tid < w.N) {
      long long g = found - w.cp - 1;
      const float *ys = y + (size_t)tid * row_stride;
      while (g - 1 >= pos && ys[g - 1] > w.threshold) g--;
      s_pstart[tid] = g;
    }
    __syncthreads();
    if (tid == 0) {
      unsigned long long si = 0;
      for (int s = 0; s < w.N; s++) si += (unsigned long long)s_pstart[s];
      si /= (unsigned)w.N;
      const long long i_switch = (long long)si + w.tx_sig_len + w.acb_len - w.L;  // first sample that is not buffered
      if (i_switch > w.n || i_switch < w.Wlen) s_done = 1;                        // the burst runs past the capture
      else { win_off[cnt] = i_switch - w.Wlen; syncs[cnt] = si; *count = cnt + 1; s_pos = i_switch + 1; }
    }
    __syncthreads();
    if (s_done) break;
    cnt++;
  }
}

// keys of the batched timing search -> per-link FFT window starts (quirk Q2) and the payload start taken from
// rx stream 1's last access code (quirk Q4, framing.cc:857), in capture coordinates
__global__ void k_timing_tables(const unsigned long long *__restrict__ keys, const long long *__restrict__ win_off, int F,
                                int N, int max_ac, int L, int M, int *__restrict__ timing, int *__restrict__ pay) {
  const int slots = max_ac + 1;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= F * N * max_ac) return;
  const int ac = idx % max_ac, r = (idx / max_ac) % N, f = idx / (max_ac * N);
  const unsigned long long k = keys[((size_t)f * N + r) * slots + ac + 1];
  const unsigned off = (k >> 32) ? (unsigned)(L * (ac + 1)) + (0xffffffffu - (unsigned)(k & 0xffffffffu)) : 0u;
  const int t = (int)(win_off[f] + off);
  timing[((size_t)f * N + r) * max_ac + ac] = t;
  if (r == (N > 1 ? 1 : 0) && ac == max_ac - 1) pay[f] = t + M;
}

}  // namespace rub
