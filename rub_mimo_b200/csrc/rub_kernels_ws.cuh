// rub_kernels_ws.cuh — the fused receive kernel, warp specialised (one persistent CTA per SM):
//
//   producer warp   one thread: TMA bulk loads of the next OFDM symbol (CP strip by address) into a
//                   two-deep ring of landing buffers as soon as the detect warps release one
//   FFT warps       NT threads (one warpgroup): in-place FFT of the N antennas of a symbol, software
//                   pipelined over (stage, antenna) pairs; stage twiddles live in registers
//   detect warps    LS accumulate / weights on training symbols; W*y -> gain -> slicer -> max-log
//                   LLR -> packed bits -> error count on payload symbols, one symbol behind the FFT
//
// replacing framesync::execute_mimo_decode (mimo/framing.cc:535-589), the LS/invert part of
// estimate_channel (:801-832) and the demod/count loop of mimo/main.cc:1403-1410.
//
// Why specialise: the monolithic kernel (rub_kernels_fused.cuh) runs FFT and detection one after the
// other in the same 16 warps, so the store path idles during the FFT and the FMA path during
// detection, and every thread carries the FFT's register footprint.  Here the two phases of
// neighbouring symbols overlap, registers are re-partitioned with setmaxnreg (FFT threads keep
// two antennas' points and all their twiddles in registers, detect threads need far fewer), and
// the hand-offs are mbarriers (full -> y_ready -> empty) instead of CTA-wide barriers.
#pragma once
#include <type_traits>

#include "rub_kernels_fused.cuh"

namespace rub {

template <int LOG2M, int N>
struct WsTraits {
  using FF = Fft<LOG2M>;
  using PL = FftPlan<LOG2M>;
  static constexpr int M = FF::M, NT = FF::NT, PAD = fft_padded_size(M);
  static constexpr int FFT_WARPS = NT / 32;
  static constexpr int BLOCKS = M / 64;                         // 64-carrier blocks per OFDM symbol
  static constexpr int DET_WARPS = BLOCKS < 16 ? BLOCKS : 16;
  static constexpr int KPW = BLOCKS / DET_WARPS;                // blocks per detect warp and symbol
  static constexpr int DET_THREADS = DET_WARPS * 32;
  static constexpr int THREADS = NT + DET_THREADS + 128;        // + the producer's warpgroup (one thread works)
  static constexpr int BUF_ELEMS = N * PAD;
  // register split (setmaxnreg acts on whole warpgroups, and ptxas sizes the launch allocation for
  // the thread count rounded up to warpgroups): 768 threads launch with 80 registers each, the
  // producer's warpgroup drops to 24 and hands 7168 registers to the FFT warpgroup
  static constexpr int LAUNCH_REGS = 65536 / THREADS / 8 * 8;
  static constexpr int AUX_REGS = 24;
  static constexpr int FFT_REGS = (LAUNCH_REGS + (LAUNCH_REGS - AUX_REGS) * 128 / NT) / 8 * 8;
  static constexpr int TW_ELEMS = FftTw<LOG2M>::TOTAL;
  static_assert(PL::NSTG == 3, "three-stage plans only");
  static_assert(NT % 128 == 0 && DET_THREADS % 128 == 0, "roles are whole warpgroups (setmaxnreg)");
  static_assert(N >= 2, "the pipelined FFT schedule needs two antenna regions");
  static_assert(KPW * DET_WARPS == BLOCKS, "block split");
  static size_t smem_bytes(int q) {
    return (size_t)2 * BUF_ELEMS * sizeof(cf) + (size_t)DET_WARPS * 2 * (256 * q) + (size_t)2 * N * M /* tx_data */ +
           (size_t)TW_ELEMS * sizeof(cf) + 64 /* mbarriers */ + 64;
  }
};

__device__ __forceinline__ void named_bar(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
template <int R>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R)); }
template <int R>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R)); }

// W*y and gain of one detection task: stream s of a 64-carrier block (lane = 2 adjacent carriers)
template <int N>
__device__ __forceinline__ void ws_dot(const TaskRegs<N> &t, const float4 *y4, cf &z0, cf &z1) {
  cf w0[N], w1[N], y0[N], y1[N];
#pragma unroll
  for (int r = 0; r < N; r++) {
    w0[r] = mk(t.w[r].x, t.w[r].y); w1[r] = mk(t.w[r].z, t.w[r].w);
    y0[r] = mk(y4[r].x, y4[r].y); y1[r] = mk(y4[r].z, y4[r].w);
  }
  z0 = cscale(wy_dot<N>(w0, y0), t.g.x);
  z1 = cscale(wy_dot<N>(w1, y1), t.g.y);
}
// the rest of the task: slicer, max-log LLRs (staged in [k][bit] order at lp), packed bits, error count
template <int MB>
__device__ __forceinline__ void ws_demap(const ChainArgs &a, const DemapConst &dc, const float *refs, cf z0, cf z1, float2 is,
                                         long long o, float *lp, unsigned txv, unsigned &ec,
                                         unsigned long long pol_stream) {
  constexpr int Q = 2 * MB;
  const int lane = threadIdx.x & 31;
  if (a.eq) st_hint4(a.eq + o, make_float4(z0.x, z0.y, z1.x, z1.y), pol_stream);
  const unsigned si0 = slice_axis_refs<MB>(z0.x, refs), sq0 = slice_axis_refs<MB>(z0.y, refs);
  const unsigned si1 = slice_axis_refs<MB>(z1.x, refs), sq1 = slice_axis_refs<MB>(z1.y, refs);
  // (gray(si) << MB) | gray(sq) in one pass: the bit shifted from si into sq's top position is masked off
  const unsigned c0 = (si0 << MB) | sq0, c1 = (si1 << MB) | sq1;
  const unsigned sym0 = c0 ^ ((c0 >> 1) & ~(1u << (MB - 1))), sym1 = c1 ^ ((c1 >> 1) & ~(1u << (MB - 1)));
  if (a.llr) {
    float l[2 * Q];  // [k][bit] order
    const float k0 = dc.k4 * is.x, k1 = dc.k4 * is.y;
    llr_axis<MB>(z0.x, k0, dc, l);
    llr_axis<MB>(z0.y, k0, dc, l + MB);
    llr_axis<MB>(z1.x, k1, dc, l + Q);
    llr_axis<MB>(z1.y, k1, dc, l + Q + MB);
#pragma unroll
    for (int v = 0; v < 2 * Q / 4; v++)
      *reinterpret_cast<float4 *>(lp + 4 * v) = make_float4(l[4 * v], l[4 * v + 1], l[4 * v + 2], l[4 * v + 3]);
  }
  const unsigned rx2 = sym0 | (sym1 << 8);
  if (a.rx_data) *reinterpret_cast<unsigned short *>(a.rx_data + o) = (unsigned short)rx2;
  if (a.bits) {
    // 4 lanes = 8 symbols = Q bytes, MSB first, written by the first lane of each quad
    const unsigned v2 = (sym0 << Q) | sym1;                                   // 2Q bits
    const unsigned p1 = __shfl_xor_sync(0xffffffffu, v2, 1);
    const unsigned v4 = (v2 << (2 * Q)) | p1;                                 // even lanes: 4Q bits
    const unsigned p2 = __shfl_xor_sync(0xffffffffu, v4, 2);
    if ((lane & 3) == 0) {
      const unsigned long long v8 = ((unsigned long long)v4 << (4 * Q)) | p2;  // 8Q bits = Q bytes
      unsigned short *bp = reinterpret_cast<unsigned short *>(a.bits + ((o - 2 * lane) >> 3) * Q + (lane >> 2) * Q);
#pragma unroll
      for (int i = 0; i < Q / 2; i++) {
        const unsigned hw = (unsigned)(v8 >> (16 * (Q / 2 - 1 - i))) & 0xffffu;
        bp[i] = (unsigned short)__byte_perm(hw, 0, 0x4401);  // big-endian halfword
      }
    }
  }
  if (a.tx_data) {
    const unsigned x = rx2 ^ txv;
    ec += (unsigned)__popc(x) + (((x & 0xffu) != 0u ? 1u : 0u) << 16) + (((x >> 8) != 0u ? 1u : 0u) << 16);
  }
}

template <int LOG2M, int N>
__global__ void __launch_bounds__(WsTraits<LOG2M, N>::THREADS, 1) k_rx_ws(FusedArgs fa, DemapConst dc) {
  using TR = WsTraits<LOG2M, N>;
  using FF = Fft<LOG2M>;
  using TW = FftTw<LOG2M>;
  constexpr int M = TR::M, NT = TR::NT, PAD = TR::PAD, DET_WARPS = TR::DET_WARPS, DET_THREADS = TR::DET_THREADS, KPW = TR::KPW;
  const ChainArgs &a = fa.a;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  cf *buf0 = reinterpret_cast<cf *>(smem_raw);
  cf *buf1 = buf0 + TR::BUF_ELEMS;
  unsigned char *stage_base = reinterpret_cast<unsigned char *>(buf1 + TR::BUF_ELEMS);
  const int q = a.q;
  const int stage_stride = 256 * q;  // 64 carriers x q LLRs
  unsigned char *txbuf = stage_base + (size_t)DET_WARPS * 2 * stage_stride;              // [2][N][M] tx symbols
  cf *tw_s = reinterpret_cast<cf *>(txbuf + 2 * N * M);                                   // stage twiddles, copied once
  unsigned long long *mbar = reinterpret_cast<unsigned long long *>(tw_s + TR::TW_ELEMS);  // full[2], yrdy[2], empty[2]

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int nsym = a.T + a.D;
  if (tid == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    mbar_init(&mbar[2], TR::FFT_WARPS);
    mbar_init(&mbar[3], TR::FFT_WARPS);
    mbar_init(&mbar[4], DET_WARPS);
    mbar_init(&mbar[5], DET_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_async_smem();
  }
  for (int i = tid; i < TR::TW_ELEMS; i += TR::THREADS) tw_s[i] = a.tw[i];
  __syncthreads();

  const int nf_cta = (a.n_frames - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int total = nf_cta * nsym;  // flat (frame, symbol) sequence of this CTA
  const unsigned long long pol_stream = policy_evict_first(), pol_keep = policy_evict_last();

  if (warp >= TR::FFT_WARPS + DET_WARPS) {
    // ================================ producer ================================
    reg_dec<TR::AUX_REGS>();
    if (tid != NT + DET_THREADS) return;
    const unsigned sym_bytes = (unsigned)(M * sizeof(cf));
    int fl = 0, sym = 0;
    for (int g = 0; g < total; g++) {
      const int b = g & 1;
      if (g >= 2) mbar_wait(&mbar[4 + b], (unsigned)(((g - 2) >> 1) & 1));  // the detect warps are done with symbol g-2
      const long long frame = (long long)blockIdx.x + (long long)fl * gridDim.x;
      cf *dst = b ? buf1 : buf0;
      const bool with_tx = a.tx_data && sym >= a.T;
      mbar_expect_tx(&mbar[b], sym_bytes * N + (with_tx ? N * M : 0));
      const cf *src = a.iq + frame * a.frame_stride + a.first_sample + (long long)sym * a.L + a.cp;
#pragma unroll
      for (int r = 0; r < N; r++)
        bulk_load(dst + (size_t)r * PAD, src + (long long)r * a.rx_stride, sym_bytes, &mbar[b], pol_stream);
      if (with_tx) {
        // the transmitted symbol indices of this OFDM symbol ride on the same mbarrier
        const unsigned char *tsrc = a.tx_data + (frame * N * a.D + (sym - a.T)) * (long long)M;
#pragma unroll
        for (int s = 0; s < N; s++) bulk_load(txbuf + (b * N + s) * M, tsrc + (long long)s * a.D * M, M, &mbar[b], pol_stream);
      }
      if (++sym == nsym) { sym = 0; fl++; }
    }
    return;
  }

  if (warp < TR::FFT_WARPS) {
    // ================================ FFT warps ================================
    reg_inc<TR::FFT_REGS>();
    const int ft = tid;
    int sym = 0;
    for (int g = 0; g < total; g++) {
      const int b = g & 1;
      cf *buf = b ? buf1 : buf0;
      const float scale = (sym >= a.T) ? a.dn : 1.0f;
      mbar_wait(&mbar[b], (unsigned)((g >> 1) & 1));
      // pair p = (stage p / N, antenna p % N); loads of pair p+1 are issued before pair p is computed.
      // In place: a barrier separates every thread's loads of a pair from any thread's stores of it,
      // and the stores of a pair from the next stage's loads of the same antenna (N >= 2 pairs later).
      // the twiddles of a stage depend on the thread only: read once per symbol, used for all N antennas
      cf va[FF::PTS], vb[FF::PTS], tw[FF::S1::NTW > FF::S2::NTW ? FF::S1::NTW : FF::S2::NTW];
      auto load_pair = [&](int p, cf *v) {
        cf *reg = buf + (size_t)(p % N) * PAD;
        if (p / N == 0) FF::S0::template load<false>(ft, reg, v);
        else if (p / N == 1) FF::S1::template load<true>(ft, reg, v);
        else FF::S2::template load<true>(ft, reg, v);
      };
      auto finish_pair = [&](int p, cf *v) {
        cf *reg = buf + (size_t)(p % N) * PAD;
        if (p / N == 0) { FF::S0::compute(ft, v, nullptr); FF::S0::template store<true, false>(ft, v, reg, 1.f); }
        else if (p / N == 1) { FF::S1::compute_pre(v, tw); FF::S1::template store<true, false>(ft, v, reg, 1.f); }
        else { FF::S2::compute_pre(v, tw); FF::S2::template store<false, true>(ft, v, reg, scale); }
      };
      load_pair(0, va);
#pragma unroll
      for (int p = 0; p < 3 * N; p++) {
        named_bar(1, NT);
        if (p == N) FF::S1::load_twiddles(ft, tw_s + TW::OFF1, tw);
        if (p == 2 * N) FF::S2::load_twiddles(ft, tw_s + TW::OFF2, tw);
        if (p + 1 < 3 * N) load_pair(p + 1, (p & 1) ? va : vb);
        finish_pair(p, (p & 1) ? vb : va);
      }
      // Y complete: every lane's stores are ordered before lane 0's release-arrive
      __syncwarp();
      if (lane == 0) mbar_arrive(&mbar[2 + b]);
      if (++sym == nsym) sym = 0;
    }
    return;
  }

  // ================================ detect warps ================================
  const int dtid = tid - NT, dwarp = warp - TR::FFT_WARPS;
  cf *Wc = fa.scratchW + (size_t)blockIdx.x * N * N * M;
  float *gc = fa.scratchG + (size_t)blockIdx.x * 2 * N * M, *ic = gc + (size_t)N * M;
  float refs[4];  // liquid ref[k] = 2^k * alpha, most significant first
#pragma unroll
  for (int i = 0; i < 4; i++) refs[i] = (i < q / 2) ? (float)(1u << (q / 2 - 1 - i)) * dc.alpha : 0.f;
  const int koff = dwarp * 64 + 2 * lane;                // first carrier of this lane in block kb = 0
  constexpr int KSTEP = 64 * DET_WARPS;                  // carrier distance between a warp's blocks
  unsigned char *slot0 = stage_base + (size_t)(dwarp * 2) * stage_stride;
  unsigned ec[N];  // per-lane error counts of the current frame: bit errors | symbol errors << 16
#pragma unroll
  for (int s = 0; s < N; s++) ec[s] = 0;
  // the packed 16-bit fields must survive the warp sum: flush before 32 lanes x bit errors can reach 65536
  const int flush_every = max(1, 2047 / (KPW * 2 * q));
  int since_flush = 0;
  auto flush_counts = [&](int nsyms_flushed) {
#pragma unroll
    for (int s = 0; s < N; s++) {
      const unsigned v = __reduce_add_sync(0xffffffffu, ec[s]);
      ec[s] = 0;
      if (lane == 0 && a.counters) {
        atomicAdd(&a.counters[s * 4 + 0], (unsigned long long)(v & 0xffffu));
        atomicAdd(&a.counters[s * 4 + 1], (unsigned long long)nsyms_flushed * KPW * 64 * q);
        atomicAdd(&a.counters[s * 4 + 2], (unsigned long long)(v >> 16));
        atomicAdd(&a.counters[s * 4 + 3], (unsigned long long)nsyms_flushed * KPW * 64);
      }
    }
  };

  int fl = 0, sym = 0;
  for (int g = 0; g < total; g++) {
    const int b = g & 1;
    const long long frame = (long long)blockIdx.x + (long long)fl * gridDim.x;
    const cf *buf = b ? buf1 : buf0;
    const bool payload = sym >= a.T;
    TaskRegs<N> w;
    if (payload) task_load<N, M>(w, WarpCtx{Wc + koff, gc + koff, nullptr, 0, koff, 0}, 0, 0, pol_keep);  // before Y is needed
    if (sym == 0 && fl > 0) named_bar(2, DET_THREADS);  // every warp is done reading the previous frame's W
    mbar_wait(&mbar[2 + b], (unsigned)((g >> 1) & 1));
    auto release_buf = [&]() {
      // this warp's last generic-proxy access to `buf` is done: the producer may refill it by TMA
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&mbar[4 + b]);
    };
    if (!payload) {
      // ---------------- LS accumulate (mimo/framing.cc:801-815) ----------------
      const int c = sym / N, t = sym % N;
      const bool q1 = (a.flags & RUB_FLAG_Q1_IDENTITY_INIT) != 0;
      constexpr int LS_IT = (N * M / 2) / DET_THREADS;
      float4 accv[LS_IT];
#pragma unroll
      for (int i = 0; i < LS_IT; i++) {
        const int e = dtid + i * DET_THREADS, r = e / (M / 2), k = 2 * (e % (M / 2));
        if (c == 0) { const float d = (q1 && r == t) ? 1.0f : 0.0f; accv[i] = make_float4(d, 0.f, d, 0.f); }
        else accv[i] = ld_hint4(Wc + (size_t)(r * N + t) * M + k, pol_keep);
      }
#pragma unroll
      for (int i = 0; i < LS_IT; i++) {
        const int e = dtid + i * DET_THREADS, r = e / (M / 2), k = 2 * (e % (M / 2));
        const float4 x = *reinterpret_cast<const float4 *>(buf + (size_t)r * PAD + k);
        const float2 sg = __ldg(reinterpret_cast<const float2 *>(a.sgn + ((size_t)t * a.nac + c) * M + k));
        float4 acc = accv[i];
        acc.x = acc.x + x.x * sg.x; acc.y = acc.y + x.y * sg.x;
        acc.z = acc.z + x.z * sg.y; acc.w = acc.w + x.w * sg.y;
        st_hint4(Wc + (size_t)(r * N + t) * M + k, acc, pol_keep);
      }
      release_buf();
      if (sym == a.T - 1) {
        // ---------------- weights (mimo/framing.cc:817-832) ----------------
        named_bar(2, DET_THREADS);
        for (int k = dtid; k < M; k += DET_THREADS) {
          cf G[N * N], W[N * N];
          float gain[N], isig[N];
#pragma unroll
          for (int e = 0; e < N * N; e++) {
            const float2 t2 = ld_hint2(Wc + (size_t)e * M + k, pol_keep);
            G[e] = cscale(mk(t2.x, t2.y), a.s_ls);
          }
          if (a.G) {
#pragma unroll
            for (int e = 0; e < N * N; e++) a.G[(frame * N * N + e) * M + k] = G[e];
          }
          compute_weights<N>(fa.wm, G, W, gain, isig);
#pragma unroll
          for (int e = 0; e < N * N; e++) st_hint2(Wc + (size_t)e * M + k, make_float2(W[e].x, W[e].y), pol_keep);
#pragma unroll
          for (int s = 0; s < N; s++) { st_hint1(gc + (size_t)s * M + k, gain[s], pol_keep); st_hint1(ic + (size_t)s * M + k, isig[s], pol_keep); }
        }
        named_bar(2, DET_THREADS);  // W complete before any warp reads it
      }
    } else {
      // ---------------- detect + demap + count ----------------
      const long long symbase = (frame * N * a.D + (sym - a.T)) * (long long)M + koff;  // stream 0, block 0, this lane
      const int DM = a.D * M;
      const unsigned char *txl = txbuf + b * N * M + koff;
      const WarpCtx wc{Wc + koff, gc + koff, nullptr, 0, koff, 0};
      auto detect = [&](auto mbtag) {
        constexpr int MB = decltype(mbtag)::value, Q = 2 * MB;
#pragma unroll
        for (int kb = 0; kb < KPW; kb++) {
          float4 y4[N];
#pragma unroll
          for (int r = 0; r < N; r++) y4[r] = *reinterpret_cast<const float4 *>(buf + (size_t)r * PAD + koff + kb * KSTEP);
          unsigned txv[N];
#pragma unroll
          for (int s = 0; s < N; s++)
            txv[s] = a.tx_data ? (unsigned)*reinterpret_cast<const unsigned short *>(txl + s * M + kb * KSTEP) : 0u;
          if (kb == KPW - 1) release_buf();  // Y and the reference symbols are in registers
#pragma unroll
          for (int s = 0; s < N; s++) {
            const int it = kb * N + s;
            cf z0, z1;
            const float2 is = w.is;
            ws_dot<N>(w, y4, z0, z1);
            // W of the next task lands in the registers the products just released
            if (it + 1 < KPW * N) task_load<N, M>(w, wc, (s + 1) % N, ((s + 1 == N) ? kb + 1 : kb) * KSTEP, pol_keep);
            unsigned char *slot = slot0 + (it & 1) * stage_stride;
            if (a.llr) {
              // the bulk store issued two tasks ago from this staging slot must have drained
              if (lane == 0) bulk_wait_read<1>();
              __syncwarp();
            }
            const long long o = symbase + (long long)s * DM + kb * KSTEP;
            ws_demap<MB>(a, dc, refs, z0, z1, is, o, reinterpret_cast<float *>(slot) + lane * 2 * Q, txv[s], ec[s], pol_stream);
            if (a.llr) {
              fence_async_smem();
              __syncwarp();
              if (lane == 0) {
                bulk_store(a.llr + (o - 2 * lane) * Q, slot, (unsigned)(64 * Q * 4), pol_stream);
                bulk_commit();
              }
            }
          }
        }
      };
      switch (q) {
        case 2: detect(std::integral_constant<int, 1>{}); break;
        case 4: detect(std::integral_constant<int, 2>{}); break;
        case 6: detect(std::integral_constant<int, 3>{}); break;
        default: detect(std::integral_constant<int, 4>{}); break;
      }
      if (a.tx_data && (++since_flush == flush_every || sym == nsym - 1)) { flush_counts(since_flush); since_flush = 0; }
    }
    if (++sym == nsym) { sym = 0; fl++; }
  }
  if (lane == 0) bulk_wait_all();
}

}  // namespace rub
