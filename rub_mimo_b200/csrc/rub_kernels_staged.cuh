// rub_kernels_staged.cuh — the staged receive path: FFT -> Y (HBM) -> LS estimate -> weights ->
// detect/demap.  Works for every supported configuration (any N <= 8, M <= 4096, ragged
// carrier allocations, per-link timing tables) and is the fallback for configurations the
// fused kernel (rub_kernels_fused.cuh) is not eligible for.
#pragma once
#include <cuda_runtime.h>

#include "rub_internal.h"
#include "rub_kernels_args.cuh"

namespace rub {

__device__ __forceinline__ long long window_start(const ChainArgs &a, int frame, int r, int sym) {
  if (sym < a.T) {
    if (a.timing) return a.timing[((long long)frame * a.N + r) * a.T + sym];
    return (long long)a.first_sample + (long long)sym * a.L + a.cp;
  }
  const long long pay0 = a.payload_start ? (long long)a.payload_start[frame]
                                         : (long long)a.first_sample + (long long)a.T * a.L;
  return pay0 + (long long)(sym - a.T) * a.L + a.cp;
}

// ---- K1: batched CP-strip + FFT (+ 1/sqrt(Mo) scale on payload symbols) -----------------
// CTA = F FFTs x NT threads; stage 0 reads the M window samples straight from HBM (coalesced
// 8-byte loads, the cp prefix is simply never touched), later stages ping-pong in smem.
template <int LOG2M, int F>
__global__ void __launch_bounds__(FftPlan<LOG2M>::NT *F) k_fft_staged(ChainArgs a) {
  using FF = Fft<LOG2M>;
  using PL = FftPlan<LOG2M>;
  using TW = FftTw<LOG2M>;
  constexpr int NT = FF::NT, M = FF::M, PAD = fft_padded_size(M);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  cf *smem = reinterpret_cast<cf *>(smem_raw);
  const int slot = threadIdx.x / NT, tid = threadIdx.x % NT;
  const long long nsym = a.T + a.D;
  const long long total = (long long)a.n_frames * nsym * a.N;
  long long fid = (long long)blockIdx.x * F + slot;
  const bool valid = fid < total;
  if (!valid) fid = total - 1;
  const int r = (int)(fid % a.N);
  const int sym = (int)((fid / a.N) % nsym);
  const int frame = (int)(fid / (a.N * nsym));
  const cf *in = a.iq + (long long)frame * a.frame_stride + (long long)r * a.rx_stride +
                 window_start(a, frame, r, sym);
  cf *out = a.Y + fid * M;
  const float scale = sym >= a.T ? a.dn : 1.0f;
  cf *A = smem + (size_t)slot * 2 * PAD, *B = A + PAD;
  cf v[FF::PTS];
  FF::S0::template load<false>(tid, in, v);
  FF::S0::compute(tid, v, nullptr);
  FF::S0::template store<true, false>(tid, v, A, 1.f);
  __syncthreads();
  FF::S1::template load<true>(tid, A, v);
  FF::S1::compute(tid, v, a.tw + TW::OFF1);
  if (PL::NSTG == 2) {
    if (valid) FF::S1::template store<false, true>(tid, v, out, scale);
  } else {
    FF::S1::template store<true, false>(tid, v, B, 1.f);
    __syncthreads();
    FF::S2::template load<true>(tid, B, v);
    FF::S2::compute(tid, v, a.tw + TW::OFF2);
    if (valid) FF::S2::template store<false, true>(tid, v, out, scale);
  }
}

// ---- K2: LS channel estimate (mimo/framing.cc:801-824) ---------------------------------
// thread per (frame, rx, tx, k)
__global__ void k_ls_fullband(ChainArgs a) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)a.n_frames * a.N * a.N * a.M;
  if (i >= total) return;
  const int k = (int)(i % a.M);
  const int t = (int)((i / a.M) % a.N);
  const int r = (int)((i / ((long long)a.M * a.N)) % a.N);
  const int frame = (int)(i / ((long long)a.M * a.N * a.N));
  const bool nul = a.scnull[k];
  // quirk Q1: G starts as identity on non-null carriers (framing.cc:302-319)
  cf acc = mk(((a.flags & RUB_FLAG_Q1_IDENTITY_INIT) && r == t && !nul) ? 1.0f : 0.0f, 0.f);
  if (!nul) {
    const long long nsym = a.T + a.D;
    for (int c = 0; c < a.nac; c++) {
      const int ts = c * a.N + t;
      const cf X = a.Y[(((long long)frame * nsym + ts) * a.N + r) * a.M + k];
      const float s = a.sgn[((long long)t * a.nac + c) * a.M + k];
      acc.x = acc.x + X.x * s;  // X / S1 with S1 = +-1: exact sign flip
      acc.y = acc.y + X.y * s;
    }
  }
  a.G[i] = cscale(acc, a.s_ls);
}

// comb pilots k = t (mod P) + linear interpolation, hold at the band edges
__device__ __forceinline__ cf comb_pilot(const ChainArgs &a, int frame, int r, int t, int kp) {
  cf acc = mk(((a.flags & RUB_FLAG_Q1_IDENTITY_INIT) && r == t) ? 1.0f : 0.0f, 0.f);
  const long long nsym = a.T + a.D;
  for (int c = 0; c < a.nac; c++) {
    const cf X = a.Y[(((long long)frame * nsym + c) * a.N + r) * a.M + kp];
    const float s = a.sgn[((long long)t * a.nac + c) * a.M + kp];
    acc.x = acc.x + X.x * s;
    acc.y = acc.y + X.y * s;
  }
  return cscale(acc, a.s_ls);
}
// thread per (frame, rx, tx, pilot i): evaluates the two pilots that bracket the P carriers
// [k0, k0+P), k0 = t + i*P, and writes them (plus the held band edges)
__global__ void k_ls_comb(ChainArgs a) {
  const int P = a.P, np = a.M / P;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)a.n_frames * a.N * a.N * np;
  if (idx >= total) return;
  const int i = (int)(idx % np);
  const int t = (int)((idx / np) % a.N);
  const int r = (int)((idx / ((long long)np * a.N)) % a.N);
  const int frame = (int)(idx / ((long long)np * a.N * a.N));
  cf *g = a.G + (((long long)frame * a.N + r) * a.N + t) * a.M;
  const int k0 = t + i * P;
  const cf ga = comb_pilot(a, frame, r, t, k0);
  if (i == 0)
    for (int k = 0; k < t; k++) g[k] = ga;             // hold below the first pilot
  g[k0] = ga;
  if (i + 1 < np) {
    const cf gb = comb_pilot(a, frame, r, t, k0 + P);
    const float invP = 1.0f / (float)P;
    for (int d = 1; d < P; d++) {
      const float f = (float)d * invP;
      g[k0 + d] = mk(fmaf(f, gb.x - ga.x, ga.x), fmaf(f, gb.y - ga.y, ga.y));
    }
  } else {
    for (int k = k0 + 1; k < a.M; k++) g[k] = ga;      // hold above the last pilot
  }
}

// W / gain / isig of carrier k (zero = null carrier): classic arrays or task records (ChainArgs::wrec)
template <int N>
__device__ __forceinline__ void store_weights(const ChainArgs &a, long long frame, int k, const cf *W, const float *gain,
                                              const float *isig, bool zero) {
  if (a.wrec) {
    unsigned char *base = reinterpret_cast<unsigned char *>(a.W);
#pragma unroll
    for (int e = 0; e < N * N; e++)
      *reinterpret_cast<cf *>(base + wrec_offset(N, a.M, frame, e / N, k, e % N)) = zero ? mk(0.f, 0.f) : W[e];
#pragma unroll
    for (int s = 0; s < N; s++) {
      *reinterpret_cast<float *>(base + wrec_offset(N, a.M, frame, s, k, N)) = zero ? 0.f : gain[s];
      *reinterpret_cast<float *>(base + wrec_offset(N, a.M, frame, s, k, N + 1)) = zero ? 0.f : isig[s];
    }
    return;
  }
  cf *Wf = a.W + frame * N * N * a.M;
  float *gf = a.gain + frame * N * a.M, *sf = a.isig + frame * N * a.M;
#pragma unroll
  for (int e = 0; e < N * N; e++) Wf[(long long)e * a.M + k] = zero ? mk(0.f, 0.f) : W[e];
#pragma unroll
  for (int s = 0; s < N; s++) { gf[(long long)s * a.M + k] = zero ? 0.f : gain[s]; sf[(long long)s * a.M + k] = zero ? 0.f : isig[s]; }
}

// ---- K3: per-subcarrier ZF / MMSE weights (mimo/framing.cc:826-832, :1344-1367) ---------
template <int N>
__global__ void k_weights(ChainArgs a, WeightMode wm) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)a.n_frames * a.M;
  if (i >= total) return;
  const int k = (int)(i % a.M);
  const long long frame = i / a.M;
  const cf *Gf = a.G + frame * N * N * a.M;
  cf G[N * N], W[N * N];
  float gain[N], isig[N];
  if (a.scnull[k]) {
    store_weights<N>(a, frame, k, W, gain, isig, true);
    return;
  }
#pragma unroll
  for (int e = 0; e < N * N; e++) G[e] = Gf[(long long)e * a.M + k];
  compute_weights<N>(wm, G, W, gain, isig);
  store_weights<N>(a, frame, k, W, gain, isig, false);
}

// ---- K2+K3 for the comb estimator in one pass: G never goes to HBM (it is written only when the caller asked
// for it).  CTA = 128 consecutive carriers of one frame: the pilots that bracket them (128 / P + 2 per antenna
// pair) are evaluated once into shared memory with comb_pilot(), every thread then interpolates its carrier's N*N
// entries with the expressions of k_ls_comb (bit-identical) and goes on with compute_weights<N>.
template <int N>
__global__ void __launch_bounds__(128) k_lscomb_weights(ChainArgs a, WeightMode wm, int write_G) {
  constexpr int TPB = 128, U = (N >= 8 ? 9 : N >= 4 ? 6 : 3);  // pilot entries per thread whose loads are in flight together
  extern __shared__ __align__(16) unsigned char sm_raw[];
  cf *pil = reinterpret_cast<cf *>(sm_raw);  // [N*N][J]
  const int P = a.P, lp = 31 - __clz(P), np = a.M >> lp, J = (TPB >> lp) + 2;  // P | M and M = 2^m: P is a power of two
  const int tiles = (a.M + TPB - 1) / TPB;
  const int frame = blockIdx.x / tiles, kt0 = (blockIdx.x % tiles) * TPB;
  const int E = N * N * J;
  const float invJ = 1.0f / (float)J;
  const long long nsym = a.T + a.D;
  const bool q1 = (a.flags & RUB_FLAG_Q1_IDENTITY_INIT) != 0;
  // pilots of the tile, U entries per thread and two access codes at a time so that their loads are in flight
  // together (one or two round trips to L2 per CTA instead of one per entry and code); the sums are
  // comb_pilot()'s, in its order
  for (int e0 = threadIdx.x; e0 < E; e0 += U * TPB) {
    int r[U], t[U], kp[U];
    bool ok[U];
    cf acc[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int e = e0 + u * TPB;
      int pair = (int)(((float)e + 0.5f) * invJ);
      if (pair * J > e) pair--;            // guard the float quotient
      if ((pair + 1) * J <= e) pair++;
      const int j = e - pair * J;
      r[u] = pair / N; t[u] = pair - r[u] * N;
      const int i = (kt0 >= t[u] ? (kt0 - t[u]) >> lp : 0) + j;
      kp[u] = t[u] + (i << lp);
      ok[u] = e < E && i < np;
      acc[u] = mk((q1 && r[u] == t[u]) ? 1.0f : 0.0f, 0.f);
    }
    for (int c0 = 0; c0 < a.nac; c0 += 2) {
      cf X[2][U];
      float sg[2][U];
#pragma unroll
      for (int cc = 0; cc < 2; cc++)
#pragma unroll
        for (int u = 0; u < U; u++)
          if (ok[u] && c0 + cc < a.nac) {
            X[cc][u] = a.Y[(((long long)frame * nsym + c0 + cc) * a.N + r[u]) * a.M + kp[u]];
            sg[cc][u] = a.sgn[((long long)t[u] * a.nac + c0 + cc) * a.M + kp[u]];
          }
#pragma unroll
      for (int cc = 0; cc < 2; cc++)
#pragma unroll
        for (int u = 0; u < U; u++)
          if (ok[u] && c0 + cc < a.nac) { acc[u].x = acc[u].x + X[cc][u].x * sg[cc][u]; acc[u].y = acc[u].y + X[cc][u].y * sg[cc][u]; }
    }
#pragma unroll
    for (int u = 0; u < U; u++)
      if (ok[u]) pil[e0 + u * TPB] = cscale(acc[u], a.s_ls);
  }
  __syncthreads();
  const int k = kt0 + threadIdx.x;
  if (k >= a.M) return;
  cf G[N * N], W[N * N];
  float gain[N], isig[N];
  const float invP = 1.0f / (float)P;
#pragma unroll
  for (int t = 0; t < N; t++) {
    // position of carrier k in tx t's comb: the same for every rx antenna
    const int kt = k - t;
    const bool below = kt < 0;                                  // hold below the first pilot
    const int i = below ? 0 : kt >> lp, d = below ? 0 : kt & (P - 1);
    const int j = i - (kt0 >= t ? (kt0 - t) >> lp : 0);
    const bool hold = d == 0 || i + 1 >= np;                    // a pilot itself, or held above the last one
    const float f = (float)d * invP;
#pragma unroll
    for (int r = 0; r < N; r++) {
      const cf *pp = pil + (r * N + t) * J + j;
      const cf ga = pp[0], gb = pp[1];                          // gb is unused (possibly unwritten) when hold
      const cf gi = mk(fmaf(f, gb.x - ga.x, ga.x), fmaf(f, gb.y - ga.y, ga.y));
      G[r * N + t] = hold ? ga : gi;
    }
  }
  if (write_G) {
    cf *Gf = a.G + (long long)frame * N * N * a.M;
#pragma unroll
    for (int e = 0; e < N * N; e++) Gf[(long long)e * a.M + k] = G[e];
  }
  if (a.scnull[k]) {
    store_weights<N>(a, frame, k, W, gain, isig, true);
    return;
  }
  compute_weights<N>(wm, G, W, gain, isig);
  store_weights<N>(a, frame, k, W, gain, isig, false);
}

// ---- K4: detect + demap + count (mimo/framing.cc:569-586, mimo/main.cc:1403-1410) --------
// blockIdx.x = (frame, symbol, stream); each thread owns 8 consecutive occupied carriers so
// its packed hard bits are whole bytes for every modulation.
template <int N>
__global__ void __launch_bounds__(128) k_detect(ChainArgs a, DemapConst lutp) {
  __shared__ DemapConst lut;
  __shared__ unsigned long long red[3];
  if (threadIdx.x == 0) { lut = lutp; red[0] = red[1] = red[2] = 0; }
  __syncthreads();
  const int s = blockIdx.x % N;
  const int d = (blockIdx.x / N) % a.D;
  const long long frame = blockIdx.x / (N * a.D);
  const int jg = blockIdx.y * blockDim.x + threadIdx.x;
  const int m = lut.m, q = a.q;
  const float alpha = lut.alpha;
  const long long nsym = a.T + a.D;
  const cf *Yf = a.Y + ((frame * nsym + a.T + d) * N) * a.M;
  const cf *Wf = a.W + (frame * N * N + (long long)s * N) * a.M;
  const float *gf = a.gain + (frame * N + s) * a.M, *sf = a.isig + (frame * N + s) * a.M;
  const long long orow = ((frame * N + s) * a.D + d);
  unsigned long long word = 0;
  unsigned be = 0, se = 0, ns = 0;
  const int j0 = jg * 8;
  for (int u = 0; u < 8; u++) {
    const int j = j0 + u;
    if (j >= a.Mo) break;
    const int k = a.occ[j];
    cf wv[N], yv[N];
#pragma unroll
    for (int r = 0; r < N; r++) { wv[r] = Wf[(long long)r * a.M + k]; yv[r] = Yf[(long long)r * a.M + k]; }
    const cf z = cscale(wy_dot<N>(wv, yv), gf[k]);
    const unsigned si = slice_axis_rt(z.x, m, alpha), sq = slice_axis_rt(z.y, m, alpha);
    const unsigned sym = (gray_encode(si) << m) + gray_encode(sq);
    const long long o = orow * a.Mo + j;
    if (a.eq) a.eq[o] = z;
    if (a.rx_data) a.rx_data[o] = (unsigned char)sym;
    if (a.llr) {
      const float is = sf[k];
      float *lp = a.llr + o * q;
      float l[8];
      const float kk = lut.k4 * is;
      llr_axis_rt(z.x, kk, m, lut, l);
      llr_axis_rt(z.y, kk, m, lut, l + m);
      for (int b = 0; b < 2 * m; b++) lp[b] = l[b];
    }
    word = (word << q) | sym;
    ns++;
    if (a.tx_data) {
      const unsigned ts = a.tx_data[o];
      be += __popc(ts ^ sym);
      se += (ts != sym);
    }
  }
  if (a.bits && ns) {
    const int nbits = ns * q, nbytes = (nbits + 7) / 8;
    word <<= (64 - nbits);
    unsigned char *bp = a.bits + orow * a.row_bytes + (long long)jg * q;
    for (int i = 0; i < nbytes; i++) bp[i] = (unsigned char)(word >> (56 - 8 * i));
  }
  if (a.tx_data && a.counters) {
    // warp reduce, then one shared atomic per warp, one global atomic per CTA and counter
    for (int off = 16; off; off >>= 1) {
      be += __shfl_xor_sync(0xffffffffu, be, off);
      se += __shfl_xor_sync(0xffffffffu, se, off);
      ns += __shfl_xor_sync(0xffffffffu, ns, off);
    }
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(&red[0], (unsigned long long)be);
      atomicAdd(&red[1], (unsigned long long)se);
      atomicAdd(&red[2], (unsigned long long)ns);
    }
    __syncthreads();
    if (threadIdx.x == 0 && red[2]) {
      atomicAdd(&a.counters[s * 4 + 0], red[0]);
      atomicAdd(&a.counters[s * 4 + 1], red[2] * (unsigned long long)q);
      atomicAdd(&a.counters[s * 4 + 2], red[1]);
      atomicAdd(&a.counters[s * 4 + 3], red[2]);
    }
  }
}

}  // namespace rub
