// rub_kernels_ws.cuh — the fused receive kernel, warp specialised (one persistent CTA per SM):
//
//   FFT warps       NT threads (one warpgroup): in-place FFT of the N antennas of a symbol in a two-deep
//                   ring of landing buffers, stage by stage over the antennas (one barrier per
//                   (stage, antenna)); the twiddles of a stage are read once per symbol into registers
//   detect warps    LS accumulate / weights on training symbols; W*y -> gain -> slicer -> max-log
//                   LLR -> packed bits -> error count on payload symbols, one symbol behind the FFT.
//                   The last detect warp to finish a symbol refills the buffer it frees: TMA bulk
//                   loads of the symbol two ahead (CP strip by address), completed on an mbarrier
//
// replacing framesync::execute_mimo_decode (mimo/framing.cc:535-589), the LS/invert part of
// estimate_channel (:801-832) and the demod/count loop of mimo/main.cc:1403-1410.
//
// Why specialise: the monolithic kernel (rub_kernels_fused.cuh) runs FFT and detection one after the
// other in the same 16 warps, so the store path idles during the FFT and the FMA path during
// detection, and every thread carries the FFT's register footprint.  Here the two phases of
// neighbouring symbols overlap, registers are re-partitioned with setmaxnreg, and the hand-offs are
// mbarriers (full -> y_ready) and a shared-memory arrival counter instead of CTA-wide barriers.
// Both loops are kept small on purpose (rolled over antennas / tasks): the SM has one 32 KB
// instruction cache for both roles and a single FFT warp per scheduler hides no fetch latency.
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
  static constexpr int THREADS = NT + DET_THREADS;
  static constexpr int BUF_ELEMS = N * PAD;
  // register split (setmaxnreg acts on whole warpgroups; ptxas sizes the launch allocation for the
  // thread count rounded up to warpgroups): 640 threads launch with 96 registers each, the detect
  // warpgroups drop to 88 and hand 4096 registers to the FFT warpgroup (128 each)
  static constexpr int LAUNCH_REGS = 65536 / THREADS / 8 * 8;
  static constexpr int DET_REGS = LAUNCH_REGS - 8;
  static constexpr int FFT_REGS = (LAUNCH_REGS + 8 * DET_THREADS / NT) / 8 * 8;
  static constexpr int TW_ELEMS = FftTw<LOG2M>::TOTAL;
  static_assert(PL::NSTG == 3, "three-stage plans only");
  static_assert(NT % 128 == 0 && DET_THREADS % 128 == 0, "roles are whole warpgroups (setmaxnreg)");
  static_assert(N >= 2 && N <= 4, "FFT barrier scheme needs two antenna regions; packed counters hold four streams");
  static_assert(KPW * DET_WARPS == BLOCKS, "block split");
  static size_t smem_bytes(int q) {
    return (size_t)2 * BUF_ELEMS * sizeof(cf) + (size_t)DET_WARPS * 2 * (256 * q) + (size_t)2 * N * M /* tx_data */ +
           (size_t)TW_ELEMS * sizeof(cf) + 64 /* mbarriers, counters */ + 64;
  }
};

__device__ __forceinline__ void named_bar(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
template <int R>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R)); }
template <int R>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R)); }
__device__ __forceinline__ unsigned atom_add_acq_rel_smem(unsigned *p, unsigned v) {
  unsigned old;
  asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(smem_u32(p)), "r"(v) : "memory");
  return old;
}

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
// the rest of the task: slicer, max-log LLRs (staged in [k][bit] order at lp), packed bits; returns the two
// demodulated symbols (sym0 | sym1 << 8)
template <int MB>
__device__ __forceinline__ unsigned ws_demap(const ChainArgs &a, const DemapConst &dc, const float *refs, cf z0, cf z1, float2 is,
                                             long long o, float *lp, unsigned long long pol_stream) {
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
  return rx2;
}

template <int LOG2M, int N, int MB>
__global__ void __launch_bounds__(WsTraits<LOG2M, N>::THREADS, 1) k_rx_ws(FusedArgs fa, DemapConst dc) {
  using TR = WsTraits<LOG2M, N>;
  using FF = Fft<LOG2M>;
  using TW = FftTw<LOG2M>;
  constexpr int M = TR::M, NT = TR::NT, PAD = TR::PAD, DET_WARPS = TR::DET_WARPS, DET_THREADS = TR::DET_THREADS, KPW = TR::KPW;
  constexpr int Q = 2 * MB;
  const ChainArgs &a = fa.a;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  cf *buf0 = reinterpret_cast<cf *>(smem_raw);
  cf *buf1 = buf0 + TR::BUF_ELEMS;
  unsigned char *stage_base = reinterpret_cast<unsigned char *>(buf1 + TR::BUF_ELEMS);
  constexpr int stage_stride = 256 * Q;  // 64 carriers x Q LLRs
  unsigned char *txbuf = stage_base + (size_t)DET_WARPS * 2 * stage_stride;              // [2][N][M] tx symbols
  cf *tw_s = reinterpret_cast<cf *>(txbuf + 2 * N * M);                                   // stage twiddles, copied once
  unsigned long long *mbar = reinterpret_cast<unsigned long long *>(tw_s + TR::TW_ELEMS);  // full[2], yrdy[2]
  unsigned *done = reinterpret_cast<unsigned *>(mbar + 4);                                // detect warps done with buffer [2]

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int nsym = a.T + a.D;
  const int nf_cta = (a.n_frames - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int total = nf_cta * nsym;  // flat (frame, symbol) sequence of this CTA
  const unsigned long long pol_stream = policy_evict_first(), pol_keep = policy_evict_last();

  // TMA loads of symbol `sym` of this CTA's frame number `fl` into ring slot b (one thread)
  auto issue_load = [&](int b, int fl, int sym) {
    const long long frame = (long long)blockIdx.x + (long long)fl * gridDim.x;
    cf *dst = b ? buf1 : buf0;
    constexpr unsigned sym_bytes = (unsigned)(M * sizeof(cf));
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
  };

  if (tid == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    mbar_init(&mbar[2], TR::FFT_WARPS);
    mbar_init(&mbar[3], TR::FFT_WARPS);
    done[0] = 0;
    done[1] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_async_smem();
    if (total > 0) issue_load(0, 0, 0);
    if (total > 1) issue_load(1, nsym > 1 ? 0 : 1, nsym > 1 ? 1 : 0);
  }
  for (int i = tid; i < TR::TW_ELEMS; i += TR::THREADS) tw_s[i] = a.tw[i];
  __syncthreads();

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
      // In place, stage by stage over the antennas.  The barrier between a pair's loads and its
      // stores also orders the stores of the previous pair before the next stage's loads of that
      // antenna (N >= 2 pairs later), so one barrier per (stage, antenna) suffices.
      cf v[FF::PTS], tw[FF::S1::NTW > FF::S2::NTW ? FF::S1::NTW : FF::S2::NTW];
#pragma unroll 1
      for (int r = 0; r < N; r++) {
        cf *reg = buf + (size_t)r * PAD;
        FF::S0::template load<false>(ft, reg, v);
        named_bar(1, NT);
        FF::S0::compute(ft, v, nullptr);
        FF::S0::template store<true, false>(ft, v, reg, 1.f);
      }
      FF::S1::load_twiddles(ft, tw_s + TW::OFF1, tw);
#pragma unroll 1
      for (int r = 0; r < N; r++) {
        cf *reg = buf + (size_t)r * PAD;
        FF::S1::template load<true>(ft, reg, v);
        named_bar(1, NT);
        FF::S1::compute_pre(v, tw);
        FF::S1::template store<true, false>(ft, v, reg, 1.f);
      }
      FF::S2::load_twiddles(ft, tw_s + TW::OFF2, tw);
#pragma unroll 1
      for (int r = 0; r < N; r++) {
        cf *reg = buf + (size_t)r * PAD;
        FF::S2::template load<true>(ft, reg, v);
        named_bar(1, NT);
        FF::S2::compute_pre(v, tw);
        FF::S2::template store<false, true>(ft, v, reg, scale);
      }
      // Y complete: every lane's stores are ordered before lane 0's release-arrive
      __syncwarp();
      if (lane == 0) mbar_arrive(&mbar[2 + b]);
      if (++sym == nsym) sym = 0;
    }
    return;
  }

  // ================================ detect warps ================================
  reg_dec<TR::DET_REGS>();
  const int dtid = tid - NT, dwarp = warp - TR::FFT_WARPS;
  cf *Wc = fa.scratchW + (size_t)blockIdx.x * N * N * M;
  float *gc = fa.scratchG + (size_t)blockIdx.x * 2 * N * M, *ic = gc + (size_t)N * M;
  float refs[4];  // liquid ref[k] = 2^k * alpha, most significant first
#pragma unroll
  for (int i = 0; i < 4; i++) refs[i] = (i < MB) ? (float)(1u << (MB - 1 - i)) * dc.alpha : 0.f;
  const int koff = dwarp * 64 + 2 * lane;                // first carrier of this lane in block kb = 0
  constexpr int KSTEP = 64 * DET_WARPS;                  // carrier distance between a warp's blocks
  unsigned char *slot0 = stage_base + (size_t)(dwarp * 2) * stage_stride;
  // per-lane error counts, 16 bits per stream: bit errors in eb, symbol errors in es
  unsigned long long eb = 0, es = 0;
  // the 16-bit fields must survive the warp sum: flush before 32 lanes x bit errors can reach 65536
  const int flush_every = max(1, 2047 / (KPW * 2 * Q));
  int since_flush = 0;
  auto flush_counts = [&](int nsyms_flushed) {
    const unsigned b0 = __reduce_add_sync(0xffffffffu, (unsigned)eb), b1 = __reduce_add_sync(0xffffffffu, (unsigned)(eb >> 32));
    const unsigned s0 = __reduce_add_sync(0xffffffffu, (unsigned)es), s1 = __reduce_add_sync(0xffffffffu, (unsigned)(es >> 32));
    eb = 0; es = 0;
    if (lane < N && a.counters) {
      const unsigned bw = (lane & 2) ? b1 : b0, sw = (lane & 2) ? s1 : s0;
      atomicAdd(&a.counters[lane * 4 + 0], (unsigned long long)((bw >> (16 * (lane & 1))) & 0xffffu));
      atomicAdd(&a.counters[lane * 4 + 1], (unsigned long long)nsyms_flushed * KPW * 64 * Q);
      atomicAdd(&a.counters[lane * 4 + 2], (unsigned long long)((sw >> (16 * (lane & 1))) & 0xffffu));
      atomicAdd(&a.counters[lane * 4 + 3], (unsigned long long)nsyms_flushed * KPW * 64);
    }
  };

  int fl = 0, sym = 0;
  for (int g = 0; g < total; g++) {
    const int b = g & 1;
    const long long frame = (long long)blockIdx.x + (long long)fl * gridDim.x;
    const cf *buf = b ? buf1 : buf0;
    const bool payload = sym >= a.T;
    const cf *wp = Wc + koff;         // W[s][0][k] of the current task (this lane's carriers)
    const float *gp = gc + koff;      // gain[s][k]; isig follows N*M floats later
    TaskRegs<N> w;
    auto load_w = [&]() {
#pragma unroll
      for (int r = 0; r < N; r++) w.w[r] = ld_hint4(wp + r * M, pol_keep);
      w.g = ld_hint2(gp, pol_keep);
      w.is = ld_hint2(gp + N * M, pol_keep);
    };
    if (payload) load_w();  // W of the first task: requested before Y is needed
    if (sym == 0 && fl > 0) named_bar(2, DET_THREADS);  // every warp is done reading the previous frame's W
    mbar_wait(&mbar[2 + b], (unsigned)((g >> 1) & 1));
    // This warp's last generic-proxy access to the ring slot is done.  The last warp to say so refills the
    // slot with the symbol two ahead.
    auto release_buf = [&]() {
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        if (atom_add_acq_rel_smem(&done[b], 1u) == (unsigned)(DET_WARPS - 1)) {
          done[b] = 0;
          if (g + 2 < total) {
            int s2 = sym + 2, f2 = fl;
            if (s2 >= nsym) { s2 -= nsym; f2++; }
            issue_load(b, f2, s2);
          }
        }
      }
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
#pragma unroll 1
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
      long long o = (frame * N * a.D + (sym - a.T)) * (long long)M + koff;  // stream 0, block 0, this lane
      const long long DM = (long long)a.D * M;
      const unsigned char *txl = txbuf + b * N * M + koff;
      int it = 0;
#pragma unroll 1
      for (int kb = 0; kb < KPW; kb++) {
        float4 y4[N];
#pragma unroll
        for (int r = 0; r < N; r++) y4[r] = *reinterpret_cast<const float4 *>(buf + (size_t)r * PAD + koff + kb * KSTEP);
        unsigned long long txp = 0;  // reference symbols of the N streams, 16 bits each
        if (a.tx_data) {
#pragma unroll
          for (int s = 0; s < N; s++)
            txp |= (unsigned long long)*reinterpret_cast<const unsigned short *>(txl + s * M + kb * KSTEP) << (16 * s);
        }
        if (kb == KPW - 1) release_buf();  // Y and the reference symbols are in registers
#pragma unroll 1
        for (int s = 0; s < N; s++, it++) {
          cf z0, z1;
          const float2 is = w.is;
          ws_dot<N>(w, y4, z0, z1);
          // W of the next task lands in the registers the products just released
          if (s + 1 < N) { wp += N * M; gp += M; }
          else { wp += KSTEP - (N - 1) * N * M; gp += KSTEP - (N - 1) * M; }
          if (it + 1 < KPW * N) load_w();
          unsigned char *slot = slot0 + (it & 1) * stage_stride;
          if (a.llr) {
            // the bulk store issued two tasks ago from this staging slot must have drained
            if (lane == 0) bulk_wait_read<1>();
            __syncwarp();
          }
          const unsigned rx2 = ws_demap<MB>(a, dc, refs, z0, z1, is, o, reinterpret_cast<float *>(slot) + lane * 2 * Q, pol_stream);
          if (a.llr) {
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
              bulk_store(a.llr + (o - 2 * lane) * Q, slot, (unsigned)(64 * Q * 4), pol_stream);
              bulk_commit();
            }
          }
          if (a.tx_data) {
            const unsigned x = rx2 ^ ((unsigned)(txp >> (16 * s)) & 0xffffu);
            eb += (unsigned long long)__popc(x) << (16 * s);
            es += (unsigned long long)(((x & 0xffu) != 0u ? 1u : 0u) + ((x >> 8) != 0u ? 1u : 0u)) << (16 * s);
          }
          o += (s + 1 < N) ? DM : (long long)KSTEP - (N - 1) * DM;
        }
      }
      if (a.tx_data && (++since_flush == flush_every || sym == nsym - 1)) { flush_counts(since_flush); since_flush = 0; }
    }
    if (++sym == nsym) { sym = 0; fl++; }
  }
  if (lane == 0) bulk_wait_all();
}

}  // namespace rub
