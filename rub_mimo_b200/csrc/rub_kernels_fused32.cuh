// rub_kernels_fused32.cuh — the fused receive kernel at twice the warps per SM.
//
// Same chain, same arithmetic (bit for bit) and same data movement as k_rx_fused
// (rub_kernels_fused.cuh: TMA bulk loads with CP strip -> FFT in shared memory -> LS / weights per
// frame -> W*y -> slicer -> max-log LLR -> packed bits -> error count -> TMA bulk stores; replaces
// mimo/framing.cc:535-589, :801-832 and mimo/main.cc:1403-1410), but every thread holds half as
// much so that a CTA is 8 points x M/8 threads per antenna at <= 64 registers:
//   * FFT: a thread owns 8 points.  A radix-8 stage is one butterfly per thread; a radix-16 stage
//     is shared by two lanes (l, l^16): each does two of the four column 4-point DFTs of the 4x4
//     decomposition, the lanes swap half of their values by shuffle and each does two of the four
//     row DFTs — the operations and their order are exactly those of bfly16();
//   * detection: one carrier per lane, a 64-carrier task is two half-tasks; W/gain/isig of the
//     next half-task are in flight while the current one is computed; one staging slot per warp.
// Both phases of k_rx_fused are latency bound at 16 warps/SM; this kernel runs
// them at 32.  Status: bit-exact, but the lane-role selects of the split radix-16 stages and the
// per-half-task overheads add ~30 % instructions, so at 62 % issue utilisation it is 8 % slower than
// k_rx_fused on C3 (1.71 vs 1.58 ms).  Opt-in (rub_rx_set_path(h, RUB_PATH_FUSED32)) until it wins.
#pragma once
#include "rub_kernels_fused.cuh"

namespace rub {

template <int LOG2M, int N>
struct Fused32Traits {
  using PL = FftPlan<LOG2M>;
  static constexpr int M = PL::M, NT = M / 8, THREADS = N * NT, PAD = fft_padded_size(M);
  static constexpr int NWARPS = THREADS / 32;
  static constexpr int BLOCKS = M / 64;         // 64-carrier blocks per OFDM symbol
  static constexpr int KPW = BLOCKS / NWARPS;   // blocks per warp (all N streams of a block stay in one warp)
  static constexpr int KSTEP = 64 * NWARPS;
  static constexpr int BUF_ELEMS = N * PAD;
  static_assert(PL::NSTG == 3 && PL::R2 == 8 && (PL::R0 == 16 || PL::R0 == 8) && (PL::R1 == 16 || PL::R1 == 8), "plan");
  static_assert(NT % 32 == 0 && THREADS <= 1024 && KPW >= 1 && KPW * NWARPS == BLOCKS, "shape");
  static size_t smem_bytes(int q) {
    return (size_t)2 * BUF_ELEMS * sizeof(cf) + (size_t)NWARPS * (256 * q + 64) + 64 * sizeof(float2) +
           (size_t)FftTw<LOG2M>::TOTAL * sizeof(cf) + (size_t)2 * N * M /* tx_data */ + 64 /* mbarriers */ + 8 * N * 2 + 64;
  }
};

// ------------------------------------------------------------------ 8-point-per-thread FFT ----
// One stage for thread `ft` of the M/8 threads of an FFT.  Index conventions are FftStage's:
// butterfly j reads elements j + t*(M/R), multiplies element t >= 1 by tws[(t-1)*NS + j%NS] and
// writes (j/NS)*NS*R + j%NS + t*NS.
template <int M, int R, int NS, bool PAD_IN, bool PAD_OUT, bool SCALE>
struct Stage8 {
  static constexpr int Q = M / R;
  template <bool PADDED>
  __device__ __forceinline__ static int at(int idx) { return PADDED ? pad_idx(idx) : idx; }

  __device__ __forceinline__ static void run(int ft, cf *buf, const cf *tws, float scale, int bar_id, int nthr) {
    if (R == 8) {
      const int j = ft, k = j % NS;
      cf v[8];
#pragma unroll
      for (int t = 0; t < 8; t++) v[t] = buf[at<PAD_IN>(j + t * Q)];
      asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(nthr) : "memory");
      if (NS > 1) {
#pragma unroll
        for (int t = 1; t < 8; t++) v[t] = cmul(v[t], tws[(t - 1) * NS + k]);
      }
      bfly8(v);
      const int base = (j / NS) * NS * R + k;
#pragma unroll
      for (int t = 0; t < 8; t++) buf[at<PAD_OUT>(base + t * NS)] = SCALE ? cscale(v[t], scale) : v[t];
    } else {
      // radix 16 shared by lanes l and l^16
      const int l = ft & 31, c = l >> 4;
      const int j = (ft >> 5) * 16 + (l & 15), k = j % NS;
      const bool hi = c != 0;
      cf r[2][4];  // r[i][m] = element (2c+i) + 4m
#pragma unroll
      for (int i = 0; i < 2; i++)
#pragma unroll
        for (int m = 0; m < 4; m++) r[i][m] = buf[at<PAD_IN>(j + (2 * c + i + 4 * m) * Q)];
      asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(nthr) : "memory");
      if (NS > 1) {
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
          for (int m = 0; m < 4; m++) {
            const int t = 2 * c + i + 4 * m;                       // t == 0 only for c = i = m = 0
            const cf w = tws[(t > 0 ? t - 1 : 0) * NS + k];
            const cf p = cmul(r[i][m], w);
            if (i == 0 && m == 0) r[i][m] = hi ? p : r[i][m];
            else r[i][m] = p;
          }
      }
      // column DFTs of rows n0 = 2c, 2c+1 and the 16-point twiddles of bfly16()
      bfly4(r[0][0], r[0][1], r[0][2], r[0][3]);
      bfly4(r[1][0], r[1][1], r[1][2], r[1][3]);
      {
        // row n0 = 2c: nothing for c = 0; (w8^1, -i, w8^3) for c = 1 (n0 = 2)
        const cf a1 = r[0][1], a2 = r[0][2], a3 = r[0][3];
        const cf w1 = mul_w8_1(a1), w2 = mul_mi(a2), w3 = mul_w8_3(a3);
        r[0][1] = hi ? w1 : a1;
        r[0][2] = hi ? w2 : a2;
        r[0][3] = hi ? w3 : a3;
        // row n0 = 2c+1: (w16^1, w8^1, w16^3) for c = 0 (n0 = 1); (w16^3, w8^3, w16^9) for c = 1 (n0 = 3)
        const cf c1 = hi ? mk(RUB_S16, -RUB_C16) : mk(RUB_C16, -RUB_S16);
        const cf c3 = hi ? mk(-RUB_C16, RUB_S16) : mk(RUB_S16, -RUB_C16);
        const cf b2 = r[1][2];
        const cf m1 = mul_w8_1(b2), m3 = mul_w8_3(b2);
        r[1][1] = cmul(r[1][1], c1);
        r[1][2] = hi ? m3 : m1;
        r[1][3] = cmul(r[1][3], c3);
      }
      // swap: this lane keeps columns k1 = 2c, 2c+1 and sends the other two of both rows
      cf own[2][2], got[2][2];
#pragma unroll
      for (int i = 0; i < 2; i++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
          own[i][e] = hi ? r[i][2 + e] : r[i][e];
          const cf snd = hi ? r[i][e] : r[i][2 + e];
          got[i][e].x = __shfl_xor_sync(0xffffffffu, snd.x, 16);
          got[i][e].y = __shfl_xor_sync(0xffffffffu, snd.y, 16);
        }
      // row DFTs for k1 = 2c+e over u[0..3][k1]: rows 2c, 2c+1 are own, the others came from the partner
      const int base = (j / NS) * NS * R + k;
#pragma unroll
      for (int e = 0; e < 2; e++) {
        cf a0 = hi ? got[0][e] : own[0][e], a1 = hi ? got[1][e] : own[1][e];
        cf a2 = hi ? own[0][e] : got[0][e], a3 = hi ? own[1][e] : got[1][e];
        bfly4(a0, a1, a2, a3);
        const int t0 = 2 * c + e;  // outputs are elements k1 + 4m
        buf[at<PAD_OUT>(base + (t0 + 0) * NS)] = SCALE ? cscale(a0, scale) : a0;
        buf[at<PAD_OUT>(base + (t0 + 4) * NS)] = SCALE ? cscale(a1, scale) : a1;
        buf[at<PAD_OUT>(base + (t0 + 8) * NS)] = SCALE ? cscale(a2, scale) : a2;
        buf[at<PAD_OUT>(base + (t0 + 12) * NS)] = SCALE ? cscale(a3, scale) : a3;
      }
    }
    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(nthr) : "memory");
  }
};

template <int N> struct HalfRegs { float2 w[N]; float g, is; };

// detection of one payload OFDM symbol by the whole CTA, one carrier per lane.  A warp owns KPW
// blocks of 64 carriers and walks them stream by stream, each 64-carrier task as two half-tasks
// of 32 carriers; W/gain/isig of the next half-task are in flight while the current one is
// computed.  LLRs and packed bits of a task are staged in the warp's slot and leave by TMA.
template <int LOG2M, int N, int MB>
__device__ __forceinline__ void detect_symbol32(const FusedArgs &fa, const cf *Wc, const float *gc, int kw,
                                                unsigned char *slot, const cf *buf, const unsigned char *txs,
                                                long long symbase, const float2 *lut, const float *refs,
                                                unsigned long long pol_keep, unsigned long long pol_stream,
                                                unsigned *cnt) {
  using TR = Fused32Traits<LOG2M, N>;
  constexpr int M = TR::M, PAD = TR::PAD, KPW = TR::KPW, KSTEP = TR::KSTEP, Q = 2 * MB, PL = 1 << MB;
  constexpr int NHALF = 2 * KPW * N;
  const ChainArgs &a = fa.a;
  const int lane = threadIdx.x & 31;
  const int DM = a.D * M;
  const int k1 = kw + lane;  // carrier of half 0 of block 0
  const cf *Wl = Wc + k1;
  const float *gl = gc + k1;
  const cf *Yl = buf + k1;
  const unsigned char *txl = txs + k1;
  const bool want_llr = a.llr != nullptr;
  unsigned eb[ErrWords<N>::NW] = {}, es[ErrWords<N>::NW] = {};
  // half-task j: block kb = j / (2N), stream s = (j / 2) % N, half h = j % 2
  auto load_half = [&](HalfRegs<N> &t, int j) {
    const int kb = j / (2 * N), s = (j >> 1) % N, h = j & 1, kadd = kb * KSTEP + 32 * h;
#pragma unroll
    for (int r = 0; r < N; r++) t.w[r] = ld_hint2(Wl + (s * N + r) * M + kadd, pol_keep);
    t.g = ld_hint1(gl + s * M + kadd, pol_keep);
    t.is = ld_hint1(gl + (N + s) * M + kadd, pol_keep);
  };
  HalfRegs<N> cur, nxt;
  load_half(cur, 0);
  unsigned symh[2] = {0u, 0u};
#pragma unroll
  for (int j = 0; j < NHALF; j++) {
    const int kb = j / (2 * N), s = (j >> 1) % N, h = j & 1;
    if (j + 1 < NHALF) load_half(nxt, j + 1);
    float2 y[N];
#pragma unroll
    for (int r = 0; r < N; r++) y[r] = *reinterpret_cast<const float2 *>(Yl + r * PAD + kb * KSTEP + 32 * h);
    cf wv[N], yv[N];
#pragma unroll
    for (int r = 0; r < N; r++) { wv[r] = mk(cur.w[r].x, cur.w[r].y); yv[r] = mk(y[r].x, y[r].y); }
    const cf z = cscale(wy_dot<N>(wv, yv), cur.g);
    const unsigned si = slice_axis_refs<MB>(z.x, refs), sq = slice_axis_refs<MB>(z.y, refs);
    const unsigned c = (si << MB) | sq;
    symh[h] = c ^ ((c >> 1) & ~(1u << (MB - 1)));
    const long long o = symbase + (long long)s * DM + kb * KSTEP + 32 * h + k1;
    if (a.eq) st_hint2(a.eq + o, make_float2(z.x, z.y), pol_stream);
    if (a.rx_data) a.rx_data[o] = (unsigned char)symh[h];
    if (h == 0 && (a.llr || a.bits)) {
      // single staging slot: the bulk store of the previous task must have read it out
      if (lane == 0) bulk_wait_read<0>();
      __syncwarp();
    }
    if (want_llr) {
      const float2 *li = lut + si, *lq = lut + sq;
      float2 *lp = reinterpret_cast<float2 *>(slot) + (32 * h + lane) * MB;
      float l[Q];
#pragma unroll
      for (int b = 0; b < MB; b++) {
        const float2 ci = li[b * PL], cq = lq[b * PL];
        l[b] = fmaf(ci.x, z.x, ci.y) * cur.is;
        l[MB + b] = fmaf(cq.x, z.y, cq.y) * cur.is;
      }
#pragma unroll
      for (int b = 0; b < MB; b++) lp[b] = make_float2(l[2 * b], l[2 * b + 1]);
    }
    if (a.tx_data) {
      const unsigned x = symh[h] ^ (unsigned)txl[s * M + kb * KSTEP + 32 * h];
      eb[s >> 2] += (unsigned)__popc(x) << (8 * (s & 3));
      es[s >> 2] += (unsigned)(x != 0u) << (8 * (s & 3));
    }
    if (h == 1 && (a.llr || a.bits)) {
      if (a.bits) {
        // lane L holds symbols L and 32+L of the block; 8 consecutive symbols = Q bytes, MSB first
#pragma unroll
        for (int hh = 0; hh < 2; hh++) {
          const unsigned v1 = symh[hh];
          const unsigned p1 = __shfl_down_sync(0xffffffffu, v1, 1);
          const unsigned v2 = (v1 << Q) | p1;                       // lanes 0 mod 2: 2 symbols
          const unsigned p2 = __shfl_down_sync(0xffffffffu, v2, 2);
          const unsigned v4 = (v2 << (2 * Q)) | p2;                 // lanes 0 mod 4: 4 symbols
          const unsigned p4 = __shfl_down_sync(0xffffffffu, v4, 4);
          if ((lane & 7) == 0) store_packed_bits<Q>(slot + fa.llr_stage_bytes + (4 * hh + (lane >> 3)) * Q, v4, p4);
        }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        const long long ob = symbase + kw + (long long)s * DM + kb * KSTEP;  // warp-uniform
        if (a.llr) bulk_store(a.llr + ob * Q, slot, (unsigned)(64 * Q * 4), pol_stream);
        if (a.bits) bulk_store(a.bits + (ob >> 3) * Q, slot + fa.llr_stage_bytes, (unsigned)(8 * Q), pol_stream);
        bulk_commit();
      }
    }
    if (j + 1 < NHALF) cur = nxt;
  }
  if (a.tx_data) flush_counts<N>(eb, es, cnt);
}

template <int LOG2M, int N>
__global__ void __launch_bounds__(Fused32Traits<LOG2M, N>::THREADS, 1) k_rx_fused32(FusedArgs fa, DemapLut lutp) {
  using TR = Fused32Traits<LOG2M, N>;
  using PL = FftPlan<LOG2M>;
  using TW = FftTw<LOG2M>;
  constexpr int M = TR::M, NT = TR::NT, PAD = TR::PAD, THREADS = TR::THREADS, NWARPS = TR::NWARPS;
  const ChainArgs &a = fa.a;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  cf *buf0 = reinterpret_cast<cf *>(smem_raw);
  cf *buf1 = buf0 + TR::BUF_ELEMS;
  unsigned char *stage_base = reinterpret_cast<unsigned char *>(buf1 + TR::BUF_ELEMS);
  const int stage_stride = fa.llr_stage_bytes + 64;  // llr block followed by 64 B of packed bits
  float2 *lut = reinterpret_cast<float2 *>(stage_base + (size_t)NWARPS * stage_stride);
  cf *tw_s = reinterpret_cast<cf *>(lut + 64);  // stage twiddles, copied once
  unsigned char *txbuf = reinterpret_cast<unsigned char *>(tw_s + TW::TOTAL);  // [2][N][M] tx symbols
  unsigned long long *mbar = reinterpret_cast<unsigned long long *>(txbuf + 2 * N * M);  // full[2], empty[2]
  unsigned *cnt = reinterpret_cast<unsigned *>(mbar + 4);  // [N][2] bit errors, symbol errors

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // provably warp-uniform for the compiler
  const int ant = tid / NT, ft = tid % NT;
  const int nsym = a.T + a.D;
  const int q = a.q;
  const unsigned long long pol_stream = policy_evict_first(), pol_keep = policy_evict_last();

  if (tid < 64) lut[tid] = make_float2(lutp.slope[tid], lutp.icpt[tid]);
  for (int i = tid; i < TW::TOTAL; i += THREADS) tw_s[i] = a.tw[i];
  if (tid < 2 * N) cnt[tid] = 0;
  if (tid == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    mbar_init(&mbar[2], NWARPS);
    mbar_init(&mbar[3], NWARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_async_smem();
  }
  __syncthreads();

  const int nf_cta = (a.n_frames - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int total = nf_cta * nsym;  // flat (frame, symbol) sequence of this CTA
  cf *Wc = fa.scratchW + (size_t)blockIdx.x * N * N * M;
  float *gc = fa.scratchG + (size_t)blockIdx.x * 2 * N * M, *ic = gc + (size_t)N * M;
  const unsigned sym_bytes = (unsigned)(M * sizeof(cf));
  const int kw = warp * 64;
  unsigned char *slot = stage_base + (size_t)warp * stage_stride;
  float refs[4];  // liquid ref[k] = 2^k * alpha, most significant first
#pragma unroll
  for (int i = 0; i < 4; i++) refs[i] = (i < q / 2) ? (float)(1u << (q / 2 - 1 - i)) * lutp.alpha : 0.f;

  auto issue_load = [&](int g) {  // thread 0 only
    const int fl = g / nsym, sym = g - fl * nsym;
    const long long frame = (long long)blockIdx.x + (long long)fl * gridDim.x;
    cf *dst = (g & 1) ? buf1 : buf0;
    unsigned long long *bar = &mbar[g & 1];
    const bool with_tx = a.tx_data && sym >= a.T;
    mbar_expect_tx(bar, sym_bytes * N + (with_tx ? N * M : 0));
    const cf *src = a.iq + frame * a.frame_stride + a.first_sample + (long long)sym * a.L + a.cp;
#pragma unroll
    for (int r = 0; r < N; r++)
      bulk_load(dst + (size_t)r * PAD, src + (long long)r * a.rx_stride, sym_bytes, bar, pol_stream);
    if (with_tx) {
      const unsigned char *tsrc = a.tx_data + (frame * N * a.D + (sym - a.T)) * (long long)M;
#pragma unroll
      for (int s = 0; s < N; s++)
        bulk_load(txbuf + ((g & 1) * N + s) * M, tsrc + (long long)s * a.D * M, M, bar, pol_stream);
    }
  };
  if (tid == 0) {
    if (total > 0) issue_load(0);
    if (total > 1) issue_load(1);
  }

  int fl = 0, sym = 0;
  for (int g = 0; g < total; g++) {
    const long long frame = (long long)blockIdx.x + (long long)fl * gridDim.x;
    cf *buf = (g & 1) ? buf1 : buf0;
    cf *mine = buf + (size_t)ant * PAD;
    const bool payload = sym >= a.T;
    const long long symbase = (frame * N * a.D + (sym - a.T)) * (long long)M;
    auto release_buf = [&]() {
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&mbar[2 + (g & 1)]);
    };
    mbar_wait(&mbar[g & 1], (unsigned)((g >> 1) & 1));

    // ---------------- FFT of the N antennas, in place ----------------
    {
      const float scale = payload ? a.dn : 1.0f;
      const int bar_id = 1 + ant;
      Stage8<M, PL::R0, 1, false, true, false>::run(ft, mine, nullptr, 1.f, bar_id, NT);
      if (tid == 0 && g >= 1 && g + 1 < total) {
        // the other buffer held symbol g-1: refill it once every warp has released it
        mbar_wait(&mbar[2 + ((g + 1) & 1)], (unsigned)(((g - 1) >> 1) & 1));
        issue_load(g + 1);
      }
      Stage8<M, PL::R1, PL::R0, true, true, false>::run(ft, mine, tw_s + TW::OFF1, 1.f, bar_id, NT);
      Stage8<M, PL::R2, PL::R0 * PL::R1, true, false, true>::run(ft, mine, tw_s + TW::OFF2, scale, bar_id, NT);
    }
    __syncthreads();

    if (!payload) {
      // ---------------- LS accumulate (mimo/framing.cc:801-815) ----------------
      const int c = sym / N, t = sym % N;
      const bool q1 = (a.flags & RUB_FLAG_Q1_IDENTITY_INIT) != 0;
      constexpr int LS_IT = (N * M / 2) / THREADS;
      float4 accv[LS_IT];
#pragma unroll
      for (int i = 0; i < LS_IT; i++) {
        const int e = tid + i * THREADS, r = e / (M / 2), k = 2 * (e % (M / 2));
        if (c == 0) { const float d = (q1 && r == t) ? 1.0f : 0.0f; accv[i] = make_float4(d, 0.f, d, 0.f); }
        else accv[i] = ld_hint4(Wc + (size_t)(r * N + t) * M + k, pol_keep);
      }
#pragma unroll
      for (int i = 0; i < LS_IT; i++) {
        const int e = tid + i * THREADS, r = e / (M / 2), k = 2 * (e % (M / 2));
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
        __syncthreads();
        for (int k = tid; k < M; k += THREADS) {
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
        __syncthreads();  // W complete before any warp prefetches it for the first payload symbol
      }
    } else {
      // ---------------- detect + demap + count ----------------
#define RUB_DETECT32(MBV) detect_symbol32<LOG2M, N, MBV>(fa, Wc, gc, kw, slot, buf, txbuf + (g & 1) * N * M, symbase, lut, refs, pol_keep, pol_stream, cnt)
      switch (q) {
        case 2: RUB_DETECT32(1); break;
        case 4: RUB_DETECT32(2); break;
        case 6: RUB_DETECT32(3); break;
        default: RUB_DETECT32(4); break;
      }
#undef RUB_DETECT32
      release_buf();
    }
    const bool frame_end = sym == nsym - 1;
    if (frame_end && a.tx_data) __syncthreads();  // shared counters complete
    if (frame_end && a.tx_data && a.counters && tid < N) {
      atomicAdd(&a.counters[tid * 4 + 0], (unsigned long long)cnt[2 * tid]);
      atomicAdd(&a.counters[tid * 4 + 1], (unsigned long long)a.D * M * q);
      atomicAdd(&a.counters[tid * 4 + 2], (unsigned long long)cnt[2 * tid + 1]);
      atomicAdd(&a.counters[tid * 4 + 3], (unsigned long long)a.D * M);
      cnt[2 * tid] = 0;
      cnt[2 * tid + 1] = 0;
    }
    if (++sym == nsym) { sym = 0; fl++; }
  }
  if (lane == 0) bulk_wait_all();
}

}  // namespace rub
