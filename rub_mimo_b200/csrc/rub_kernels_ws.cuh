// rub_kernels_ws.cuh — the fused receive kernel, warp specialised (one persistent CTA per SM):
//
//   FFT warps       NT threads (one warpgroup, 128 registers).  Payload symbols: in-place FFT of the N antennas
//                   in a two-deep ring of landing slots, stage by stage over the antennas, handed to the detect
//                   warps through an mbarrier.  Training symbols never reach the detect warps: they arrive one
//                   antenna at a time in a landing row, are transformed in a work row, and the LS estimate is
//                   accumulated straight from the last stage's registers into the frame's G scratch.  The
//                   training symbols of frame f+1 are interleaved with the payload symbols of frame f (quota +
//                   work-conserving test), so the FFT load is even over time.
//   detect warps    16 warps, 88 registers.  Weights once per frame (written as task records: W, gain and
//                   1/sigma^2 of one stream's 64 carriers, 2560 B contiguous), then per payload symbol
//                   W*y -> gain -> slicer -> max-log LLR -> packed bits -> error count, one symbol behind the
//                   FFT.  A warp fetches the record of its next task with one TMA bulk copy while it finishes
//                   the current one; the last detect warp to finish a symbol refills the ring slot it frees (TMA
//                   bulk loads of the payload symbol two ahead, CP strip by address, completed on an mbarrier).
//
// replacing framesync::execute_mimo_decode (mimo/framing.cc:535-589), the LS/invert part of
// estimate_channel (:801-832) and the demod/count loop of mimo/main.cc:1403-1410.
//
// Why specialise: the monolithic kernel (rub_kernels_fused.cuh) runs FFT and detection one after the
// other in the same 16 warps, so the store path idles during the FFT and the FMA path during
// detection, and every thread carries the FFT's register footprint.  Here the two phases overlap,
// registers are re-partitioned with setmaxnreg, and the hand-offs are mbarriers and a shared-memory
// arrival counter instead of CTA-wide barriers.  Both loops are kept small on purpose (rolled over
// antennas / tasks): the SM has one 32 KB instruction cache for both roles and a single FFT warp
// per scheduler hides no fetch latency.  Frame f+1 is estimated while frame f is detected: its LS
// estimate goes to a scratch of its own, the task records that the detect warps re-read D times per frame
// stay a single evict_last set per CTA.  DESIGN.md 4.1 has the measurements behind each of these choices.
#pragma once
#include <type_traits>

#include "rub_kernels_fused.cuh"

#ifndef RUB_WS_ABLATE
#define RUB_WS_ABLATE 0  // measurement builds only (wrong results): 1 FFT warps skip the transform, 2 detect warps skip
                         // detection, 4 no G scratch stores, 8 no W/gain/isig scratch stores, 16 no gain/isig/tx loads
#endif
#ifndef RUB_WS_BACKOFF_NS
#define RUB_WS_BACKOFF_NS 100
#endif

namespace rub {

template <int LOG2M, int N>
struct WsTraits {
  using FF = Fft<LOG2M>;
  using PL = FftPlan<LOG2M>;
  static constexpr int M = FF::M, NT = FF::NT, PAD = M;  // row stride: the stage-0 -> stage-1 exchange is XOR-swizzled, not padded
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
  static constexpr int WREC = N * 64 + 64;                      // task record in cf units: W[r][64], gain[64] f32, isig[64] f32
  static_assert(PL::R0 == 16 && PL::R1 == 16 && M / NT == 16, "the swizzled exchange below is written for 16-point threads");
  static_assert(PL::NSTG == 3, "three-stage plans only");
  static_assert(NT % 128 == 0 && DET_THREADS % 128 == 0, "roles are whole warpgroups (setmaxnreg)");
  static_assert(N >= 2 && N <= 4, "FFT barrier scheme needs two antenna regions; packed counters hold four streams");
  static_assert(KPW * DET_WARPS == BLOCKS, "block split");
  static size_t smem_bytes(int q) {
    return (size_t)2 * BUF_ELEMS * sizeof(cf) /* payload ring */ + (size_t)(M + PAD) * sizeof(cf) /* training: landing + work row */ +
           (size_t)DET_WARPS * (256 * q) /* LLR staging */ + (size_t)DET_WARPS * WREC * sizeof(cf) /* task records, one per detect warp */ +
           (size_t)(8 + DET_WARPS) * 8 + 16 /* mbarriers, counters */;
  }
};

// Host side: sign bytes of the access codes in the register order of the last FFT stage.  Thread ft, butterfly
// b of that stage ends up with the carriers k = (j / NS2) * NS2 * R2 + j % NS2 + t2 * NS2, j = ft + b * NT,
// t2 = 0..R2-1: byte [tx][code][j] carries their signs, bit t2 set = S1 is -1.
template <int LOG2M>
void ws_pack_signs(int N, int nac, const float *sgn /* [tx][code][M] */, unsigned char *out /* [tx][code][M/R2] */) {
  using FF = Fft<LOG2M>;
  constexpr int M = FF::M, NS2 = FftTw<LOG2M>::NS2, R2 = FF::S2::P / FF::S2::B;
  static_assert(R2 <= 8, "one byte per butterfly");
  for (int tc = 0; tc < N * nac; tc++)
    for (int j = 0; j < M / R2; j++) {
      unsigned v = 0;
      for (int t2 = 0; t2 < R2; t2++) {
        const int k = (j / NS2) * NS2 * R2 + (j % NS2) + t2 * NS2;
        if (sgn[(size_t)tc * M + k] < 0.f) v |= 1u << t2;
      }
      out[(size_t)tc * (M / R2) + j] = (unsigned char)v;
    }
}

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

// Exchange between the first and the second FFT stage without padding: element idx sits at
// (idx & ~15) | ((idx ^ (idx >> 4)) & 15).  Stage 0 (radix 16, stride 1): thread j writes idx = 16 j + t, i.e. the
// byte address (16 j + (j & 15)) * 8 ^ (t * 8) of a 128-byte aligned row: one LOP3 per element, and 16 consecutive
// lanes hit 16 different 8-byte banks for every t.  Stage 1 (radix 16, stride 16) reads idx = j + 128 t: the low
// nibble is (j & 15) ^ ((j >> 4) | ((t & 1) << 3)), two base addresses and immediate offsets.
template <int M>
__device__ __forceinline__ void ws_s0_store(int j, const cf *v, cf *row) {
  const unsigned a = smem_u32(row) + (unsigned)((16 * j + (j & 15)) * 8);
#pragma unroll
  for (int t = 0; t < 16; t++)
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a ^ (unsigned)(t * 8)), "f"(v[t].x), "f"(v[t].y) : "memory");
}
template <int M>
__device__ __forceinline__ void ws_s1_load(int j, const cf *row, cf *v) {
  static_assert(M == 2048, "Q = M / 16 = 128 and j < 128");
  const int a0 = (j & ~15) + ((j & 15) ^ (j >> 4));
  const cf *p0 = row + a0, *p1 = row + (a0 ^ 8);
#pragma unroll
  for (int t = 0; t < 16; t++) v[t] = (t & 1) ? p1[128 * t] : p0[128 * t];
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
// hard decisions of the task: equalised symbols out, slicer, packed bits; returns the two demodulated symbols
// (sym0 | sym1 << 8).  Runs BEFORE the next task's W loads are requested: its warp shuffles must not share a
// scoreboard with loads that take an L2 round trip.
template <int MB, bool FULL>
__device__ __forceinline__ unsigned ws_hard(const float *refs, cf z0, cf z1, cf *eqp, unsigned char *rxp,
                                            unsigned short *bp, unsigned long long pol_stream, int lane) {
  constexpr int Q = 2 * MB;
  if (FULL || eqp) st_hint4(eqp, make_float4(z0.x, z0.y, z1.x, z1.y), pol_stream);
  const unsigned si0 = slice_axis_refs<MB>(z0.x, refs), sq0 = slice_axis_refs<MB>(z0.y, refs);
  const unsigned si1 = slice_axis_refs<MB>(z1.x, refs), sq1 = slice_axis_refs<MB>(z1.y, refs);
  // (gray(si) << MB) | gray(sq) in one pass: the bit shifted from si into sq's top position is masked off
  const unsigned c0 = (si0 << MB) | sq0, c1 = (si1 << MB) | sq1;
  const unsigned sym0 = c0 ^ ((c0 >> 1) & ~(1u << (MB - 1))), sym1 = c1 ^ ((c1 >> 1) & ~(1u << (MB - 1)));
  const unsigned rx2 = sym0 | (sym1 << 8);
  if (!FULL && rxp) *reinterpret_cast<unsigned short *>(rxp) = (unsigned short)rx2;
  if (FULL || bp) {
    // 4 lanes = 8 symbols = Q bytes, MSB first, written by the first lane of each quad
    const unsigned v2 = (sym0 << Q) | sym1;                                   // 2Q bits
    const unsigned p1 = __shfl_xor_sync(0xffffffffu, v2, 1);
    const unsigned v4 = (v2 << (2 * Q)) | p1;                                 // even lanes: 4Q bits
    const unsigned p2 = __shfl_xor_sync(0xffffffffu, v4, 2);
    if ((lane & 3) == 0) {
      const unsigned long long v8 = ((unsigned long long)v4 << (4 * Q)) | p2;  // 8Q bits = Q bytes
#pragma unroll
      for (int i = 0; i < Q / 2; i++) {
        const unsigned hw = (unsigned)(v8 >> (16 * (Q / 2 - 1 - i))) & 0xffffu;
        bp[i] = (unsigned short)__byte_perm(hw, 0, 0x4401);  // big-endian halfword
      }
    }
  }
  return rx2;
}
// max-log LLRs of the task, staged in [k][bit] order at lp
template <int MB>
__device__ __forceinline__ void ws_llr(const DemapConst &dc, cf z0, cf z1, float2 is, float *lp) {
  constexpr int Q = 2 * MB;
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

// FULL: the caller asked for eq + LLR + bits + counters and no rx_data (the usual set): the per-task null checks
// of the output pointers are compiled out
template <int LOG2M, int N, int MB, bool FULL>
__global__ void __launch_bounds__(WsTraits<LOG2M, N>::THREADS, 1) k_rx_ws(FusedArgs fa, DemapConst dc) {
  using TR = WsTraits<LOG2M, N>;
  using FF = Fft<LOG2M>;
  using TW = FftTw<LOG2M>;
  constexpr int M = TR::M, NT = TR::NT, PAD = TR::PAD, DET_WARPS = TR::DET_WARPS, DET_THREADS = TR::DET_THREADS, KPW = TR::KPW;
  constexpr int Q = 2 * MB;
  const ChainArgs &a = fa.a;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  cf *ring = reinterpret_cast<cf *>(smem_raw);                       // [2][N][PAD] payload symbols
  cf *land = ring + 2 * TR::BUF_ELEMS;                               // [M]   training symbol as it lands (one antenna)
  cf *work = land + M;                                               // [PAD] the training symbol being transformed
  unsigned char *stage_base = reinterpret_cast<unsigned char *>(work + PAD);
  constexpr int stage_stride = 256 * Q;  // 64 carriers x Q LLRs: one LLR staging slot per detect warp
  cf *wbuf_base = reinterpret_cast<cf *>(stage_base + (size_t)DET_WARPS * stage_stride);  // [DET_WARPS][WREC] W of the task
  // mbarriers: full[2] (ring slot loaded), yrdy[2] (ring slot transformed), tfull at 4 (landing row loaded),
  // gdone at 6 (a frame's G is complete), wdone at 7 (the weights of a frame are computed), wrdy[DET_WARPS] from 8 on
  // (a detect warp's W record has landed)
  unsigned long long *mbar = reinterpret_cast<unsigned long long *>(wbuf_base + (size_t)DET_WARPS * TR::WREC);
  unsigned *done = reinterpret_cast<unsigned *>(mbar + 8 + DET_WARPS);  // detect warps done with ring slot [2]
  volatile unsigned *sched = done + 2;                                // FFT warps: verdict of the work-conserving test

  const int tid = threadIdx.x;
  int lane;  // kept in a register: ptxas would otherwise re-read SR_TID.X (a long-latency S2R) at every use
  asm volatile("mov.u32 %0, %1;" : "=r"(lane) : "r"(tid & 31));
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int D = a.D;                 // payload symbols per frame
  const int TU = N * N * a.nac;      // training units ((symbol, antenna) transforms) per frame
  const int nf = (a.n_frames - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // frames of this CTA
  const unsigned long long pol_stream = policy_evict_first(), pol_keep = policy_evict_last();
  constexpr unsigned sym_bytes = (unsigned)(M * sizeof(cf));

  // TMA loads of payload symbol number pg (counted over this CTA's frames) into ring slot pg & 1 (one thread)
  auto issue_payload = [&](int pg) {
    const int f = pg / D, d = pg - f * D;
    if (f >= nf) return;
    const long long frame = (long long)blockIdx.x + (long long)f * gridDim.x;
    cf *dst = ring + (size_t)(pg & 1) * TR::BUF_ELEMS;
    mbar_expect_tx(&mbar[pg & 1], sym_bytes * N);
    const cf *src = a.iq + frame * a.frame_stride + a.first_sample + (long long)(a.T + d) * a.L + a.cp;
#pragma unroll
    for (int r = 0; r < N; r++)
      bulk_load(dst + (size_t)r * PAD, src + (long long)r * a.rx_stride, sym_bytes, &mbar[pg & 1], pol_stream);
  };
  // Training unit number ug of this CTA: frame ug / TU, then (tx t, rx antenna r, code c) with the code fastest,
  // i.e. OFDM symbol c * N + t of antenna r: the nac symbols that accumulate into G[r][t] follow each other.  One
  // TMA load into the landing row (free again as soon as the first stage of the previous unit has read it).
  auto issue_training = [&](int ug) {
    const int f = ug / TU, u = ug - f * TU;
    if (f >= nf) return;
    const long long frame = (long long)blockIdx.x + (long long)f * gridDim.x;
    const int c = u % a.nac, r = (u / a.nac) % N, t = u / (a.nac * N);
    mbar_expect_tx(&mbar[4], sym_bytes);
    bulk_load(land, a.iq + frame * a.frame_stride + a.first_sample + (long long)(c * N + t) * a.L + a.cp +
              (long long)r * a.rx_stride, sym_bytes, &mbar[4], pol_stream);
  };

  if (tid == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    mbar_init(&mbar[2], TR::FFT_WARPS);
    mbar_init(&mbar[3], TR::FFT_WARPS);
    mbar_init(&mbar[4], 1);
    mbar_init(&mbar[5], 1);
    mbar_init(&mbar[6], TR::FFT_WARPS);
    mbar_init(&mbar[7], DET_WARPS);
    for (int w = 0; w < DET_WARPS; w++) mbar_init(&mbar[8 + w], 1);
    done[0] = 0;
    done[1] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_async_smem();
    issue_training(0);
    issue_payload(0);
    issue_payload(1);
  }
  __syncthreads();

  if (warp < TR::FFT_WARPS) {
    // ================================ FFT warps ================================
    reg_inc<TR::FFT_REGS>();
    const int ft = tid;
    const bool q1 = (a.flags & RUB_FLAG_Q1_IDENTITY_INIT) != 0;
    constexpr int NS2 = TW::NS2, B2 = FF::S2::B, R2 = FF::S2::P / B2;
    const int nac = a.nac;
    int ug = 0;               // next training unit (counted over this CTA's frames)
    int target = TU;          // training units to have finished before the next payload symbol
    float2 acc[FF::PTS];      // LS accumulators of this thread's 16 carriers of G[r][t], kept over the nac codes
#pragma unroll
    for (int i = 0; i < FF::PTS; i++) acc[i] = make_float2(0.f, 0.f);
    // One loop serves both kinds of work (a single copy of the stage code in the instruction cache): a
    // payload job transforms the N antennas of a ring slot in place; a training job transforms the one
    // antenna of a landing slot, keeps the last stage's outputs in registers and accumulates them over
    // the nac codes into G[r][t] (mimo/framing.cc:801-815).  After each payload symbol the training units
    // of the next frame that are due are worked off.
    int pg = -1;  // current payload symbol number (-1: prologue, the training units of the first frame)
    while (true) {
      bool training = ug < target;
      if (!training) {
        const int nx = pg + 1;
        if (nx >= nf * D) break;
        // Work conserving: if the ring slot of the next payload symbol is not loaded yet (the detect warps are
        // behind) and training units of the next frame are left, work one off instead of waiting.  From the third
        // payload symbol of a frame on the G scratch is known to be free (see the quota below).  Thread 0 tests
        // the mbarrier, the group barrier publishes its verdict.
        const int fn = nx / D, dn = nx - fn * D;
        if (dn >= 2 && ug < (fn + 2) * TU && ug < nf * TU) {
          if (tid == 0) *sched = mbar_test_wait(&mbar[nx & 1], (unsigned)((nx >> 1) & 1)) ? 0u : 1u;
          named_bar(1, NT);
          training = *sched != 0u;
          named_bar(1, NT);  // everyone has read the verdict before thread 0 can overwrite it
        }
        if (!training) pg = nx;
      }
      cf v[FF::PTS], tw[FF::S1::NTW > FF::S2::NTW ? FF::S1::NTW : FF::S2::NTW];
      const int nr = training ? 1 : N;
      int f, c = 0, r0 = 0, t = 0;
      cf *buf;
      if (training) {
        f = ug / TU;
        const int u = ug - f * TU;
        c = u % nac; r0 = (u / nac) % N; t = u / (nac * N);
        buf = work;

        // the G scratch is free once the weights of the previous frame have been computed from it
        if (u == 0 && f >= 1) mbar_wait_parked(&mbar[7], (unsigned)((f - 1) & 1));
        mbar_wait_parked(&mbar[4], (unsigned)(ug & 1));
      } else {
        f = pg / D;
        buf = ring + (size_t)(pg & 1) * TR::BUF_ELEMS;
        mbar_wait_parked(&mbar[pg & 1], (unsigned)((pg >> 1) & 1));
      }
      // training: the sign bytes of this thread's carriers (one per last-stage butterfly), requested early
      unsigned sgb[B2];
#pragma unroll
      for (int b = 0; b < B2; b++) sgb[b] = training ? (unsigned)fa.sgn8[((size_t)t * nac + c) * (M / R2) + ft + b * NT] : 0u;
      // In place, stage by stage over the job's antennas (a training unit: from the landing row into the work row);
      // the twiddles of a stage depend on the thread only and are read once per job.  The barrier between an
      // antenna's loads and its stores also orders the stores of the previous antenna before the next stage's
      // loads of it (N antennas later); a one-antenna job needs a barrier of its own between the stages.  The
      // barriers sit AFTER the butterflies: a barrier only orders the issue of loads, and a load still queued in
      // another scheduler's LSU pipe can be overtaken by the async proxy (the TMA refill of the landing row);
      // having consumed the loaded values, every thread's loads have been performed when it arrives.
#if (RUB_WS_ABLATE & 1)  // measurement only: no transform at all (results are wrong)
#pragma unroll
      for (int i = 0; i < FF::PTS; i++) v[i] = mk(0.f, 0.f);
      if (training) { named_bar(1, NT); if (tid == 0) issue_training(ug + 1); }
#else
#pragma unroll 1
      for (int r = 0; r < nr; r++) {
        cf *reg = buf + (size_t)r * PAD;
        FF::S0::template load<false>(ft, training ? land : reg, v);
        FF::S0::compute(ft, v, nullptr);
        named_bar(1, NT);
        if (training && tid == 0) issue_training(ug + 1);
        ws_s0_store<M>(ft, v, reg);
      }
      FF::S1::load_twiddles(ft, a.tw + TW::OFF1, tw);
      if (training) named_bar(1, NT);
#pragma unroll 1
      for (int r = 0; r < nr; r++) {
        cf *reg = buf + (size_t)r * PAD;
        ws_s1_load<M>(ft, reg, v);
        named_bar(1, NT);
        FF::S1::compute_pre(v, tw);
        // unpadded: NS1 consecutive lanes write consecutive elements (conflict free without the pad), and the last
        // stage then reads and writes the same M/R2-strided positions per butterfly, i.e. it is in place per thread
        FF::S1::template store<false, false>(ft, v, reg, 1.f);
      }
      // last-stage twiddles: tw[b * (R2 - 1) + t - 1] = table[(t - 1) * NS2 + ft + b * NT] (global memory / L1: shared
      // memory is spent on the rings and the W records)
#pragma unroll
      for (int b = 0; b < B2; b++)
#pragma unroll
        for (int t2 = 1; t2 < R2; t2++)
          tw[b * (R2 - 1) + t2 - 1] = ld_tw(a.tw + TW::OFF2 + (t2 - 1) * NS2 + (ft + b * NT) % NS2);
      named_bar(1, NT);  // the stage-1 stores of the last antenna are visible
#pragma unroll 1
      for (int r = 0; r < nr; r++) {
        cf *reg = buf + (size_t)r * PAD;
        FF::S2::template load<false>(ft, reg, v);
        FF::S2::compute_pre(v, tw);
        // payload: in place per butterfly, no barrier; training: the outputs stay in registers (the work row is
        // rewritten by the next unit's first stage behind that stage's barrier)
        if (!training) FF::S2::template store<false, true>(ft, v, reg, a.dn);
      }
#endif
      if (!training) {
        // Y complete: every lane's stores are ordered before lane 0's release-arrive
        __syncwarp();
        if (lane == 0) mbar_arrive(&mbar[2 + (pg & 1)]);
        // training units of the next frame are spread over this frame's payload symbols from the third on
        const int d = pg - f * D;
        if (d == D - 1) target = (f + 2) * TU;
        else if (d >= 2) target = (f + 1) * TU + (int)(((long long)TU * (d - 1) + (D - 3)) / (D - 2));
        if (target > nf * TU) target = nf * TU;
      } else {
        // X / S1 with S1 = +-1 is a sign flip (mimo/framing.cc:809-812): the signs of this thread's carriers
        // come as one byte per butterfly, bit t2 = carrier j + t2 * NS2
        const float dinit = (q1 && r0 == t) ? 1.0f : 0.0f;
#pragma unroll
        for (int b = 0; b < B2; b++) {
          const unsigned sb = sgb[b];
#pragma unroll
          for (int t2 = 0; t2 < R2; t2++) {
            const int i = b * R2 + t2;
            const unsigned neg = (sb << (31 - t2)) & 0x80000000u;
            float2 ac = (c == 0) ? make_float2(dinit, 0.f) : acc[i];
            ac.x = ac.x + u2f(f2u(v[i].x) ^ neg); ac.y = ac.y + u2f(f2u(v[i].y) ^ neg);
            acc[i] = ac;
          }
        }
        if (c == nac - 1 && !(RUB_WS_ABLATE & 4)) {
          cf *Gf = fa.scratchAcc + (size_t)blockIdx.x * N * N * M + (size_t)(r0 * N + t) * M;
#pragma unroll
          for (int b = 0; b < B2; b++)
#pragma unroll
            for (int t2 = 0; t2 < R2; t2++) {
              const int j = ft + b * NT, k = (j / NS2) * NS2 * R2 + (j % NS2) + t2 * NS2;
              st_hint2(Gf + k, acc[b * R2 + t2], pol_stream);
            }
        }
        ug++;
        if (ug - f * TU == TU) {
          // G(f) complete.  The detect warps read it from L2 (L1::no_allocate loads): a CTA-scope release does not
          // wait for global stores to get there, so every thread fences its own stores at GPU scope first
          __threadfence();
          __syncwarp();
          if (lane == 0) mbar_arrive(&mbar[6]);
        }
      }
    }
    return;
  }

  // ================================ detect warps ================================
  reg_dec<TR::DET_REGS>();
  const int dtid = tid - NT, dwarp = warp - TR::FFT_WARPS;
  float refs[4];  // liquid ref[k] = 2^k * alpha, most significant first
#pragma unroll
  for (int i = 0; i < 4; i++) refs[i] = (i < MB) ? (float)(1u << (MB - 1 - i)) * dc.alpha : 0.f;
  const int koff = dwarp * 64 + 2 * lane;                // first carrier of this lane in block kb = 0
  constexpr int KSTEP = 64 * DET_WARPS;                  // carrier distance between a warp's blocks
  unsigned char *slot = stage_base + (size_t)dwarp * stage_stride;  // this warp's LLR staging slot
  cf *wbuf = wbuf_base + (size_t)dwarp * TR::WREC;                  // this warp's W record (TMA destination)
  unsigned long long *wrdy = &mbar[8 + dwarp];
  constexpr int KB = M / 64;                                         // 64-carrier blocks per symbol
  unsigned wn = 0;                                                   // W records consumed so far (mbarrier phase)
  // per-lane error counts, 16 bits per stream (streams 0,1 in word 0; 2,3 in word 1)
  unsigned eb0 = 0, eb1 = 0, es0 = 0, es1 = 0;
  // the 16-bit fields must survive the warp sum: flush before 32 lanes x bit errors can reach 65536
  const int flush_every = max(1, 2047 / (KPW * 2 * Q));
  int since_flush = 0;
  auto flush_counts = [&](int nsyms_flushed) {
    const unsigned b0 = __reduce_add_sync(0xffffffffu, eb0), b1 = __reduce_add_sync(0xffffffffu, eb1);
    const unsigned s0 = __reduce_add_sync(0xffffffffu, es0), s1 = __reduce_add_sync(0xffffffffu, es1);
    eb0 = eb1 = es0 = es1 = 0;
    if (lane < N && a.counters) {
      const unsigned bw = (lane & 2) ? b1 : b0, sw = (lane & 2) ? s1 : s0;
      atomicAdd(&a.counters[lane * 4 + 0], (unsigned long long)((bw >> (16 * (lane & 1))) & 0xffffu));
      atomicAdd(&a.counters[lane * 4 + 1], (unsigned long long)nsyms_flushed * KPW * 64 * Q);
      atomicAdd(&a.counters[lane * 4 + 2], (unsigned long long)((sw >> (16 * (lane & 1))) & 0xffffu));
      atomicAdd(&a.counters[lane * 4 + 3], (unsigned long long)nsyms_flushed * KPW * 64);
    }
  };

  int pg = 0;
  for (int f = 0; f < nf; f++) {
    const long long frame = (long long)blockIdx.x + (long long)f * gridDim.x;
    // task records of this CTA: [stream][64-carrier block]{ W[rx][64] complex, gain[64], isig[64] }
    cf *Wc = fa.scratchW + (size_t)blockIdx.x * N * KB * TR::WREC;
    const cf *Gc = fa.scratchAcc + (size_t)blockIdx.x * N * N * M;
    // this frame's outputs; inside a frame 32-bit element offsets do ((stream * D + symbol) * M + carrier)
    const long long fbase = frame * N * (long long)D * M;
    cf *const eqf = (FULL || a.eq) ? a.eq + fbase : nullptr;
    unsigned char *const rxf = (!FULL && a.rx_data) ? a.rx_data + fbase : nullptr;
    unsigned char *const bitsf = (FULL || a.bits) ? a.bits + fbase / 8 * Q : nullptr;
    unsigned char *const llrf = (FULL || a.llr) ? reinterpret_cast<unsigned char *>(a.llr) + fbase * (Q * 4) : nullptr;
    const unsigned char *const txf = (FULL || a.tx_data) ? a.tx_data + fbase : nullptr;
    const int DMi = D * M;
    // ---------------- weights (mimo/framing.cc:817-832) ----------------
    mbar_wait_parked(&mbar[6], (unsigned)(f & 1));  // G(f) is complete (written by the FFT warps)
    if (f > 0) named_bar(2, DET_THREADS);     // every detect warp is done reading the previous frame's W
#pragma unroll 1
    for (int k = dtid; k < M; k += DET_THREADS) {
      cf G[N * N], W[N * N];
      float gain[N], isig[N];
#pragma unroll
      for (int e = 0; e < N * N; e++) {
        const float2 t2 = ld_hint2(Gc + (size_t)e * M + k, pol_stream);
        G[e] = cscale(mk(t2.x, t2.y), a.s_ls);
      }
      if (a.G) {
#pragma unroll
        for (int e = 0; e < N * N; e++) a.G[(frame * N * N + e) * M + k] = G[e];
      }
      compute_weights<N>(fa.wm, G, W, gain, isig);
      // task records: one bulk copy per detection task brings W, gain and 1/sigma^2 of a stream's 64 carriers
#if (RUB_WS_ABLATE & 8)  // measurement only: no scratch stores
      if (k < 0)
#endif
#pragma unroll
      for (int e = 0; e < N * N; e++)
        st_hint2(Wc + (size_t)((e / N) * KB + (k >> 6)) * TR::WREC + (e % N) * 64 + (k & 63), make_float2(W[e].x, W[e].y), pol_keep);
#if (RUB_WS_ABLATE & 8)
      if (k < 0)
#endif
#pragma unroll
      for (int s = 0; s < N; s++) {
        float *rec = reinterpret_cast<float *>(Wc + (size_t)(s * KB + (k >> 6)) * TR::WREC + N * 64);
        st_hint1(rec + (k & 63), gain[s], pol_keep);
        st_hint1(rec + 64 + (k & 63), isig[s], pol_keep);
      }
    }
    // the W records are read through the async proxy (TMA): order this thread's generic-proxy stores before it
    __threadfence();
    asm volatile("fence.proxy.async;" ::: "memory");
    // the G scratch has been consumed: the FFT warps may start on the frame after this one
    __syncwarp();
    if (lane == 0) mbar_arrive(&mbar[7]);
    named_bar(2, DET_THREADS);  // W complete before any warp reads it; every warp is past the previous frame
    // W record of detection task (stream s, block kb of this warp) -> this warp's buffer (one lane)
    auto issue_w = [&](int s_, int kb_) {
      mbar_expect_tx(wrdy, (unsigned)(TR::WREC * sizeof(cf)));
      bulk_load(wbuf, Wc + (size_t)(s_ * KB + dwarp + kb_ * DET_WARPS) * TR::WREC, (unsigned)(TR::WREC * sizeof(cf)), wrdy, pol_keep);
    };
    if (lane == 0) issue_w(0, 0);

    // ---------------- detect + demap + count, payload symbol by payload symbol ----------------
#pragma unroll 1
    for (int d = 0; d < D; d++, pg++) {
      const int b = pg & 1;
      const cf *buf = ring + (size_t)b * TR::BUF_ELEMS;
      int ob = 0;                       // element offset of the task's block from the symbol's (stream 0, block 0)
      // this symbol's output positions of the lane / the warp's block, kept in registers (the compiler would
      // otherwise rebuild them from the kernel arguments in every task)
      const int d0 = d * M + dwarp * 64;
      cf *eqd = eqf + d0 + 2 * lane;
      const unsigned char *txd = txf + d0 + 2 * lane;
      unsigned char *rxd = rxf ? rxf + d0 + 2 * lane : nullptr;
      unsigned char *bitd = bitsf + ((d0 >> 3) + (lane >> 2)) * Q;
      unsigned char *llrd = llrf + (size_t)d0 * (Q * 4);
      asm volatile("" : "+l"(eqd), "+l"(txd), "+l"(bitd), "+l"(llrd));
      // transmitted symbols of the lane's two carriers: the only plain load of the task loop, requested a whole
      // task before its use and AFTER the LLR store's proxy fence (which waits for loads in flight)
      unsigned txv = 0;
      auto load_tx = [&]() {
#if (RUB_WS_ABLATE & 16)  // measurement only: no tx loads
        return;
#endif
        if (FULL || txf) txv = ld_hint_u16(txd + ob, pol_stream);
      };
      load_tx();  // first task: requested before Y is needed
      if ((FULL || txf) && lane < KPW * N) {
        // the reference symbols of this warp's tasks of the NEXT payload symbol: pull their lines into L2 now so
        // that the 2-byte loads riding with the gain loads never wait for HBM
        const unsigned char *tp = txf + d0 + (lane % N) * DMi + (lane / N) * KSTEP;
        if (d == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(tp));
        if (d + 1 < D) asm volatile("prefetch.global.L2 [%0];" ::"l"(tp + M));
      }
      mbar_wait_parked(&mbar[2 + b], (unsigned)((pg >> 1) & 1));
      int it = 0;
#pragma unroll 1
      for (int kb = 0; kb < KPW; kb++) {
        float4 y4[N];
#pragma unroll
        for (int r = 0; r < N; r++) y4[r] = *reinterpret_cast<const float4 *>(buf + (size_t)r * PAD + koff + kb * KSTEP);
        if (kb == KPW - 1) {
          // This warp's last generic-proxy access to the ring slot is done (Y is in registers).  The last
          // warp to say so refills the slot with the payload symbol two ahead.
          fence_async_smem();
          __syncwarp();
          if (lane == 0 && atom_add_acq_rel_smem(&done[b], 1u) == (unsigned)(DET_WARPS - 1)) {
            done[b] = 0;
            issue_payload(pg + 2);
          }
        }
#if (RUB_WS_ABLATE & 2)  // measurement only: no detection (results are wrong)
        if (y4[0].x == 123.456f) eb0++;
        continue;
#endif
#pragma unroll 1
        for (int s = 0; s < N; s++, it++) {
          cf z0, z1;
          TaskRegs<N> w;
          const int obc = ob;
          mbar_wait_parked(wrdy, wn & 1u);
          wn++;
#pragma unroll
          for (int r = 0; r < N; r++) w.w[r] = *reinterpret_cast<const float4 *>(wbuf + r * 64 + 2 * lane);
          w.g = *reinterpret_cast<const float2 *>(reinterpret_cast<const float *>(wbuf + N * 64) + 2 * lane);
          const float2 is = *reinterpret_cast<const float2 *>(reinterpret_cast<const float *>(wbuf + N * 64) + 64 + 2 * lane);
          ws_dot<N>(w, y4, z0, z1);
          // The record of the next task of this frame (the next symbol starts over at task 0).  The proxy fence
          // waits for this lane's loads of the buffer (their values went into the products above) before the
          // copy may overwrite it.
          fence_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (it + 1 < KPW * N) issue_w(s + 1 < N ? s + 1 : 0, s + 1 < N ? kb : kb + 1);
            else if (d + 1 < D) issue_w(0, 0);
          }
          const unsigned rx2 = ws_hard<MB, FULL>(refs, z0, z1, (FULL || eqf) ? eqd + obc : nullptr, rxd ? rxd + obc : nullptr,
                                                 (FULL || bitsf) ? reinterpret_cast<unsigned short *>(bitd + (obc >> 3) * Q) : nullptr,
                                                 pol_stream, lane);
          if (s + 1 < N) ob += DMi;
          else ob += KSTEP - (N - 1) * DMi;
          if (FULL || llrf) {
            // the bulk store of the previous task must have read the staging slot
            if (lane == 0) bulk_wait_read<0>();
            __syncwarp();
            ws_llr<MB>(dc, z0, z1, is, reinterpret_cast<float *>(slot) + lane * 2 * Q);
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
              bulk_store(llrd + obc * (Q * 4), slot, (unsigned)(64 * Q * 4), pol_stream);
              bulk_commit();
            }
          }
          if (FULL || txf) {
            const unsigned x = rx2 ^ txv;
            const unsigned vb = (unsigned)__popc(x) << (16 * (s & 1));
            const unsigned vs = (((x & 0xffu) != 0u ? 1u : 0u) + ((x >> 8) != 0u ? 1u : 0u)) << (16 * (s & 1));
            if (s & 2) { eb1 += vb; es1 += vs; } else { eb0 += vb; es0 += vs; }
          }
          if (it + 1 < KPW * N) load_tx();  // the next task's (ob has moved on)
        }
      }
      if ((FULL || a.tx_data) && (++since_flush == flush_every || d == D - 1)) { flush_counts(since_flush); since_flush = 0; }
    }
  }
  if (lane == 0) bulk_wait_all();
}

}  // namespace rub
