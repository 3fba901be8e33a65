// rub_kernels_fused.cuh — the fused receive kernel: one persistent CTA walks whole frames
//   TMA bulk load (CP strip) -> FFT in shared memory -> LS estimate / weights (per frame)
//   -> W*y -> gain -> slicer -> max-log LLR -> packed bits -> error count -> TMA bulk store
// replacing the per-symbol loop of framesync::execute_mimo_decode (mimo/framing.cc:535-589),
// the LS/invert part of estimate_channel (:801-832) and the demod/count loop of
// mimo/main.cc:1403-1410.
//
// Data movement (DESIGN.md "Fused kernel"):
//   * input samples: cp.async.bulk global->shared (one 8*M byte copy per rx antenna and OFDM
//     symbol, the cp prefix is skipped by the source address), double buffered over symbols
//     and completed through an mbarrier, so HBM reads overlap the previous symbol's math;
//   * FFT: in place in the landing buffer, 3 register-radix stages, padded exchange layout;
//   * W/gain/isig: per-CTA scratch in global memory, rewritten per frame and re-read for each
//     of the D payload symbols; kept L2 resident with an evict_last policy while every
//     streaming access (samples in, LLR/bits/eq out, tx_data) carries evict_first;
//   * detection is software pipelined: the W/gain/isig/tx_data registers of task i+1 are
//     loaded while task i is computed;
//   * LLRs and packed bits: staged per warp in shared memory in their final byte order and
//     written with cp.async.bulk shared->global; equalised symbols go out as 16-byte
//     coalesced stores.
#pragma once
#include <cuda_runtime.h>

#include "rub_internal.h"
#include "rub_kernels_args.cuh"

namespace rub {

// ---------------------------------------------------------------- PTX helpers ---------
__device__ __forceinline__ unsigned smem_u32(const void *p) {
  return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long *bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// non-blocking: has the phase with this parity completed?
__device__ __forceinline__ bool mbar_test_wait(unsigned long long *bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// try_wait with a suspend-time hint: the warp is parked by the hardware until the phase completes or the hint
// (nanoseconds) expires, instead of re-issuing the test every few cycles
__device__ __forceinline__ bool mbar_try_wait_hint(unsigned long long *bar, unsigned parity, unsigned ns) {
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(ns)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_parked(unsigned long long *bar, unsigned parity) {
  while (!mbar_try_wait_hint(bar, parity, 20000u)) {
  }
}
// A consumer that is expected to wait for a while (the detect warps waiting for the FFT warps): a bare
// try_wait loop returns after a few cycles and its polling takes issue slots from the very warps it waits for
// (measured: a quarter of the kernel's issued instructions), so sleep between polls.
__device__ __forceinline__ void mbar_wait_backoff(unsigned long long *bar, unsigned parity, unsigned ns) {
  while (!mbar_try_wait(bar, parity)) __nanosleep(ns);
}
// L2 eviction policies: streams are read/written once (evict_first), the per-CTA W scratch
// is re-read D times per frame (evict_last)
__device__ __forceinline__ unsigned long long policy_evict_first() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ unsigned long long policy_evict_last() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void bulk_load(void *dst_smem, const void *src_gmem, unsigned bytes,
                                          unsigned long long *bar, unsigned long long pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void bulk_store(void *dst_gmem, const void *src_smem, unsigned bytes,
                                           unsigned long long pol) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst_gmem),
               "r"(smem_u32(src_smem)), "r"(bytes), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
template <int NTHR>
__device__ __forceinline__ void group_sync(int id) {
  if (NTHR % 32 == 0) asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(NTHR) : "memory");
  else __syncthreads();
}
__device__ __forceinline__ float4 ld_hint4(const void *p, unsigned long long pol) {
  float4 v;
  asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ float2 ld_hint2(const void *p, unsigned long long pol) {
  float2 v;
  asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v2.f32 {%0, %1}, [%2], %3;"
               : "=f"(v.x), "=f"(v.y)
               : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ float ld_hint1(const void *p, unsigned long long pol) {
  float v;
  asm volatile("ld.global.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ unsigned ld_hint_u16(const void *p, unsigned long long pol) {
  unsigned short v;
  asm volatile("ld.global.L1::no_allocate.L2::cache_hint.u16 %0, [%1], %2;" : "=h"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ void st_hint4(void *p, float4 v, unsigned long long pol) {
  asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void st_hint2(void *p, float2 v, unsigned long long pol) {
  asm volatile("st.global.L1::no_allocate.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(p), "f"(v.x), "f"(v.y),
               "l"(pol)
               : "memory");
}
__device__ __forceinline__ void st_hint1(void *p, float v, unsigned long long pol) {
  asm volatile("st.global.L1::no_allocate.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(pol) : "memory");
}

template <int LOG2M, int N>
struct FusedTraits {
  using FF = Fft<LOG2M>;
  static constexpr int M = FF::M, NT = FF::NT, THREADS = N * NT, PAD = fft_padded_size(M);
  static constexpr int NWARPS = THREADS / 32;
  static constexpr int TASKS = N * M / 64;        // (stream, 64-carrier block) per OFDM symbol
  static constexpr int TPW = TASKS / NWARPS;      // tasks per warp per symbol (= M / (2 NT))
  static constexpr int KPW = TPW / N;             // 64-carrier blocks per warp: all N streams of a
                                                  // block are served by one warp (Y is read once)
  static constexpr int KSTEP = 64 * NWARPS;       // carrier distance between a warp's blocks
  static constexpr int BUF_ELEMS = N * PAD;
  // small CTAs share an SM: cap their registers so that 512 threads fit (4 x 128 or 2 x 256 threads x 128 registers)
  static constexpr int MIN_CTAS = THREADS <= 128 ? 4 : (THREADS <= 256 ? 2 : 1);
  // the 4096-point stage tables (32 KB) stay in global memory / L1 for the 256-thread instance so that
  // two of its CTAs fit an SM; everywhere else they are copied to shared memory once
  static constexpr bool TW_SMEM = !(LOG2M == 12 && THREADS <= 256);
  static constexpr int TW_SMEM_ELEMS = TW_SMEM ? FftTw<LOG2M>::TOTAL : 0;
  static_assert(NT % 32 == 0, "fused path needs NT to be whole warps");
  static_assert(KPW >= 1 && KPW * N * NWARPS == TASKS && (TPW % 2) == 0, "task split");
  // WTMA variant (two-stream instances): W, gain and 1/sigma^2 of a detection task come as one TMA record into a
  // two-deep per-warp ring, two tasks ahead of their use; it pays for the ring with the second LLR staging slot and
  // the shared-memory copy of the stage twiddles (read through L1 instead)
  // (measured: 2-4 % faster at M = 1024 / 2048; slower at 512, where a CTA is two warps, and at 4096, whose 32 KB of
  // stage twiddles then come through L1)
  static constexpr bool HAS_WTMA = N == 2 && (LOG2M == 10 || LOG2M == 11);
  static constexpr int KB = M / 64;                 // 64-carrier blocks per symbol
  static constexpr int REC = N * 64 + 64;           // task record in cf units: W[rx][64], gain[64] f32, isig[64] f32
  static size_t smem_bytes(int q, bool wtma = false) {
    return (size_t)2 * BUF_ELEMS * sizeof(cf) + (size_t)NWARPS * (wtma ? 1 : 2) * (256 * q + 64) +
           (wtma ? (size_t)NWARPS * 2 * REC * sizeof(cf) + (size_t)NWARPS * 2 * 8 : (size_t)TW_SMEM_ELEMS * sizeof(cf)) +
           (size_t)2 * N * M /* tx_data */ + 64 /* mbarriers */ + 8 * N * 2 + 64;
  }
  // position (cf index) of W[e = stream * N + rx][k] in the per-CTA scratch: classic [e][k] or task records
  template <bool WTMA>
  RUB_HD static int wpos(int e, int k) {
    return WTMA ? ((e / N) * KB + (k >> 6)) * REC + (e % N) * 64 + (k & 63) : e * M + k;
  }
};

// registers of one detection task: stream s, carriers [k0, k0+64); lane owns k0+2*lane, +1
template <int N>
struct TaskRegs {
  float4 w[N];   // W[s][r][k], W[s][r][k+1]
  float2 g, is;  // gain, 1/sigma_eff^2 of the two carriers
};

// per-warp constants of the detection phase (fixed for the whole kernel)
struct WarpCtx {
  const cf *Wp;          // W scratch at (s=0, r=0, first carrier of this lane)
  const float *gp;       // gain scratch at (s=0, first carrier); isig follows N*M floats later
  unsigned char *slot0;  // this warp's first staging slot (the second follows stage_stride bytes later)
  int stage_stride;
  int koff;              // first carrier of this lane: warp*64 + 2*lane
  int kw;                // first carrier of this warp (warp-uniform)
  // WTMA variant: this warp's two-deep record ring, its mbarriers and the number of records consumed so far
  cf *wring;
  unsigned long long *wrdy;
  const cf *Wrec;        // the CTA's record scratch
  unsigned wn;
};

template <int N, int M>
__device__ __forceinline__ void task_load(TaskRegs<N> &t, const WarpCtx &c, int s, int kadd,
                                          unsigned long long pol_keep) {
#pragma unroll
  for (int r = 0; r < N; r++) t.w[r] = ld_hint4(c.Wp + (s * N + r) * M + kadd, pol_keep);
  t.g = ld_hint2(c.gp + s * M + kadd, pol_keep);
  t.is = ld_hint2(c.gp + (N + s) * M + kadd, pol_keep);
}

// demap of one equalised symbol: returns the symbol index, writes 2*MB LLRs
template <int MB>
__device__ __forceinline__ unsigned demap_one(cf z, float isig, const float *refs, const DemapConst &dc, float *llr,
                                              bool want_llr) {
  const unsigned si = slice_axis_refs<MB>(z.x, refs), sq = slice_axis_refs<MB>(z.y, refs);
  if (want_llr) {
    const float k = dc.k4 * isig;
    llr_axis<MB>(z.x, k, dc, llr);
    llr_axis<MB>(z.y, k, dc, llr + MB);
  }
  // (gray(si) << MB) | gray(sq) in one pass: the shifted-in bit that crosses from si into sq's
  // top position is masked off
  const unsigned c = (si << MB) | sq;
  return c ^ ((c >> 1) & ~(1u << (MB - 1)));
}

template <int N> struct ErrWords;
template <int N>
__device__ __forceinline__ void flush_counts(const unsigned *eb, const unsigned *es, unsigned *cnt);

template <int N, int MB>
__device__ __forceinline__ void task_compute(const TaskRegs<N> &t, const ChainArgs &a, const float4 *y4,
                                             long long o, float *lp, unsigned char *bp, const DemapConst &lut,
                                             const float *refs, unsigned long long pol_stream,
                                             unsigned txv, unsigned &ebw, unsigned &esw, int cshift) {
  constexpr int Q = 2 * MB;
  const int lane = threadIdx.x & 31;
  cf w0[N], w1[N], y0[N], y1[N];
#pragma unroll
  for (int r = 0; r < N; r++) {
    w0[r] = mk(t.w[r].x, t.w[r].y); w1[r] = mk(t.w[r].z, t.w[r].w);
    y0[r] = mk(y4[r].x, y4[r].y); y1[r] = mk(y4[r].z, y4[r].w);
  }
  const cf acc0 = wy_dot<N>(w0, y0), acc1 = wy_dot<N>(w1, y1);
  const cf z0 = cscale(acc0, t.g.x), z1 = cscale(acc1, t.g.y);
  float l0[Q], l1[Q];
  const bool want_llr = a.llr != nullptr;
  const unsigned sym0 = demap_one<MB>(z0, t.is.x, refs, lut, l0, want_llr);
  const unsigned sym1 = demap_one<MB>(z1, t.is.y, refs, lut, l1, want_llr);
  const unsigned rx2 = sym0 | (sym1 << 8);
  if (a.eq) st_hint4(a.eq + o, make_float4(z0.x, z0.y, z1.x, z1.y), pol_stream);
  if (a.rx_data) *reinterpret_cast<unsigned short *>(a.rx_data + o) = (unsigned short)rx2;
  if (want_llr) {
    float tmp[2 * Q];  // [k][bit] order, 2*Q floats per lane
#pragma unroll
    for (int b = 0; b < Q; b++) { tmp[b] = l0[b]; tmp[Q + b] = l1[b]; }
#pragma unroll
    for (int v = 0; v < 2 * Q / 4; v++)
      *reinterpret_cast<float4 *>(lp + 4 * v) = make_float4(tmp[4 * v], tmp[4 * v + 1], tmp[4 * v + 2], tmp[4 * v + 3]);
  }
  if (a.bits) {
    // 4 lanes = 8 symbols = Q bytes, MSB first
    const unsigned v2 = (sym0 << Q) | sym1;                                   // 2Q bits
    const unsigned p1 = __shfl_xor_sync(0xffffffffu, v2, 1);
    const unsigned v4 = (v2 << (2 * Q)) | p1;                                 // even lanes: 4Q bits
    const unsigned p2 = __shfl_xor_sync(0xffffffffu, v4, 2);
    if ((lane & 3) == 0) {
      const unsigned long long v8 = ((unsigned long long)v4 << (4 * Q)) | p2;  // 8Q bits = Q bytes
      // big-endian byte order, written as Q/2 byte-swapped halfwords (bp is 2-byte aligned)
#pragma unroll
      for (int i = 0; i < Q / 2; i++) {
        const unsigned hw = (unsigned)(v8 >> (16 * (Q / 2 - 1 - i))) & 0xffffu;
        reinterpret_cast<unsigned short *>(bp)[i] = (unsigned short)__byte_perm(hw, 0, 0x4401);
      }
    }
  }
  if (a.tx_data) {
    // per-lane packed counters (one byte per stream), reduced once per OFDM symbol by the caller
    const unsigned x = rx2 ^ txv;
    ebw += (unsigned)__popc(x) << cshift;
    esw += ((unsigned)((x & 0xffu) != 0u) + (unsigned)((x >> 8) != 0u)) << cshift;
  }
}

// Q bytes (8 symbols, MSB first) from hi = the first 4 symbols (4Q/2 bits each... 4*Q/2 = 2Q bits... see callers)
// and lo = the next 4: byte order fixed with PRMT
template <int Q>
__device__ __forceinline__ void store_packed_bits(unsigned char *bp, unsigned hi, unsigned lo) {
  if (Q == 2) {         // 8 + 8 bits
    *reinterpret_cast<unsigned short *>(bp) = (unsigned short)__byte_perm(hi, lo, 0x4440);   // [hi.b0, lo.b0]
  } else if (Q == 4) {  // 16 + 16 bits
    *reinterpret_cast<unsigned *>(bp) = __byte_perm(hi, lo, 0x4501);                         // [hi.b1, hi.b0, lo.b1, lo.b0]
  } else if (Q == 6) {  // 24 + 24 bits, 2-byte aligned
    unsigned short *b16 = reinterpret_cast<unsigned short *>(bp);
    b16[0] = (unsigned short)__byte_perm(hi, lo, 0x4412);  // [hi.b2, hi.b1]
    b16[1] = (unsigned short)__byte_perm(hi, lo, 0x4460);  // [hi.b0, lo.b2]
    b16[2] = (unsigned short)__byte_perm(hi, lo, 0x4445);  // [lo.b1, lo.b0]
  } else {              // 32 + 32 bits
    *reinterpret_cast<uint2 *>(bp) = make_uint2(__byte_perm(hi, 0, 0x0123), __byte_perm(lo, 0, 0x0123));
  }
}

// packed per-lane error counters: one byte per stream, four streams per word
template <int N> struct ErrWords { static constexpr int NW = (N + 3) / 4; };

// adds this warp's packed per-lane error counts to the CTA counters cnt[N][2] = {bit errors, symbol errors}
template <int N>
__device__ __forceinline__ void flush_counts(const unsigned *eb, const unsigned *es, unsigned *cnt) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int w = 0; w < ErrWords<N>::NW; w++) {
    // bytes -> 16-bit fields so that the sum over 32 lanes cannot overflow
    const unsigned b_even = __reduce_add_sync(0xffffffffu, eb[w] & 0x00ff00ffu);         // streams 4w, 4w+2
    const unsigned s_even = __reduce_add_sync(0xffffffffu, es[w] & 0x00ff00ffu);
    unsigned b_odd = 0, s_odd = 0;
    if (N > 1) {
      b_odd = __reduce_add_sync(0xffffffffu, (eb[w] >> 8) & 0x00ff00ffu);                // streams 4w+1, 4w+3
      s_odd = __reduce_add_sync(0xffffffffu, (es[w] >> 8) & 0x00ff00ffu);
    }
    const int first = 8 * w;  // cnt index of stream 4w
    if (lane < 8 && first + lane < 2 * N) {
      const int st = lane >> 1;
      const unsigned even = (lane & 1) ? s_even : b_even, odd = (lane & 1) ? s_odd : b_odd;
      const unsigned v = (((st & 1) ? odd : even) >> (16 * (st >> 1))) & 0xffffu;
      atomicAdd(cnt + first + lane, v);
    }
  }
}

// detection of one payload OFDM symbol by the whole CTA; `cur` already holds task 0.
// A warp walks KPW blocks of 64 carriers; for each block it reads Y once (N x 16 B per lane) and
// serves the N streams one after the other.
// record of task number `task` of this warp (stream task % N, block warp + (task / N) * NWARPS) -> ring slot
// `slot` (one lane)
template <int LOG2M, int N>
__device__ __forceinline__ void issue_record(const WarpCtx &wc, int warp, int task, int slot, unsigned long long pol_keep) {
  using TR = FusedTraits<LOG2M, N>;
  const int s = task % N, blk = warp + (task / N) * TR::NWARPS;
  unsigned long long *bar = wc.wrdy + slot;
  mbar_expect_tx(bar, (unsigned)(TR::REC * sizeof(cf)));
  bulk_load(wc.wring + (size_t)slot * TR::REC, wc.Wrec + (size_t)(s * TR::KB + blk) * TR::REC, (unsigned)(TR::REC * sizeof(cf)), bar, pol_keep);
}

template <int LOG2M, int N, int MB, bool WTMA>
__device__ __forceinline__ void detect_symbol(const FusedArgs &fa, TaskRegs<N> &cur, WarpCtx &wc,
                                              const cf *buf, const unsigned char *txs, long long symbase,
                                              const DemapConst &lut,
                                              const float *refs, unsigned long long pol_keep,
                                              unsigned long long pol_stream, unsigned *cnt, bool more_symbols) {
  using TR = FusedTraits<LOG2M, N>;
  constexpr int M = TR::M, PAD = TR::PAD, KPW = TR::KPW, KSTEP = TR::KSTEP, Q = 2 * MB;
  const ChainArgs &a = fa.a;
  const int lane = threadIdx.x & 31;
  const int DM = a.D * M;
  const long long obase = symbase + wc.koff;  // stream 0, block 0, this lane
  const cf *Yl = buf + wc.koff;
  const unsigned char *txl = txs + wc.koff;   // transmitted symbols of this OFDM symbol, [stream][k] in smem
  TaskRegs<N> nxt;
  static_assert(KPW * 2 * 8 <= 255, "a symbol's bit errors of one lane and stream fit a byte");
  unsigned eb[ErrWords<N>::NW] = {}, es[ErrWords<N>::NW] = {};
#pragma unroll
  for (int kb = 0; kb < KPW; kb++) {
    float4 y4[N];
#pragma unroll
    for (int r = 0; r < N; r++) y4[r] = *reinterpret_cast<const float4 *>(Yl + r * PAD + kb * KSTEP);
#pragma unroll
    for (int s = 0; s < N; s++) {
      constexpr int dummy = 0; (void)dummy;
      const int it = kb * N + s;
      if (WTMA) {
        // this task's record has landed in ring slot wn & 1 (requested two tasks ago)
        const int rs = (int)(wc.wn & 1u);
        mbar_wait_parked(wc.wrdy + rs, (wc.wn >> 1) & 1u);
        const cf *rec = wc.wring + (size_t)rs * TR::REC;
#pragma unroll
        for (int r = 0; r < N; r++) cur.w[r] = *reinterpret_cast<const float4 *>(rec + r * 64 + 2 * lane);
        cur.g = *reinterpret_cast<const float2 *>(reinterpret_cast<const float *>(rec + N * 64) + 2 * lane);
        cur.is = *reinterpret_cast<const float2 *>(reinterpret_cast<const float *>(rec + N * 64) + 64 + 2 * lane);
      } else if (it + 1 < KPW * N) {
        const int sn = (s + 1) % N, kn = (s + 1 == N) ? kb + 1 : kb;
        task_load<N, M>(nxt, wc, sn, kn * KSTEP, pol_keep);
      }
      // the bulk store issued from this staging slot (two tasks ago; WTMA: the previous task) must have drained
      if (lane == 0) { if (WTMA) bulk_wait_read<0>(); else bulk_wait_read<1>(); }
      __syncwarp();
      const long long o = obase + (long long)s * DM + kb * KSTEP;
      unsigned char *slot = wc.slot0 + (WTMA ? 0 : (it & 1) * wc.stage_stride);
      task_compute<N, MB>(cur, a, y4, o, reinterpret_cast<float *>(slot) + lane * 2 * Q,
                          slot + fa.llr_stage_bytes + (lane >> 2) * Q, lut, refs, pol_stream,
                          a.tx_data ? (unsigned)*reinterpret_cast<const unsigned short *>(txl + s * M + kb * KSTEP) : 0u,
                          eb[s >> 2], es[s >> 2], 8 * (s & 3));
      // (the proxy fence also waits for this lane's reads of the record: the record two tasks ahead may land there)
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        const long long ob = o - 2 * lane;  // first symbol of the 64-carrier block
        if (a.llr) bulk_store(a.llr + ob * Q, slot, (unsigned)(64 * Q * 4), pol_stream);
        if (a.bits) bulk_store(a.bits + (ob >> 3) * Q, slot + fa.llr_stage_bytes, (unsigned)(8 * Q), pol_stream);
        bulk_commit();
        if (WTMA && (it + 2 < KPW * N || more_symbols))
          issue_record<LOG2M, N>(wc, (int)(threadIdx.x >> 5), (it + 2) % (KPW * N), (int)(wc.wn & 1u), pol_keep);
      }
      if (WTMA) wc.wn++;
      else if (it + 1 < KPW * N) cur = nxt;
    }
  }
  if (a.tx_data) flush_counts<N>(eb, es, cnt);
}

template <int LOG2M, int N, bool WTMA = false>
__global__ void __launch_bounds__(FusedTraits<LOG2M, N>::THREADS, FusedTraits<LOG2M, N>::MIN_CTAS) k_rx_fused(FusedArgs fa, DemapConst lutp) {
  using TR = FusedTraits<LOG2M, N>;
  using FF = Fft<LOG2M>;
  using PL = FftPlan<LOG2M>;
  using TW = FftTw<LOG2M>;
  constexpr int M = TR::M, NT = TR::NT, PAD = TR::PAD, THREADS = TR::THREADS, NWARPS = TR::NWARPS;
  const ChainArgs &a = fa.a;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  cf *buf0 = reinterpret_cast<cf *>(smem_raw);
  cf *buf1 = buf0 + TR::BUF_ELEMS;
  unsigned char *stage_base = reinterpret_cast<unsigned char *>(buf1 + TR::BUF_ELEMS);
  const int stage_stride = fa.llr_stage_bytes + 64;  // llr block followed by 64 B of packed bits
  // classic: two LLR staging slots per warp, then the stage twiddles; WTMA: one slot per warp, then the record rings
  cf *tw_s = reinterpret_cast<cf *>(stage_base + (size_t)NWARPS * (WTMA ? 1 : 2) * stage_stride);  // stage twiddles, copied once
  cf *rings = tw_s;                                                                                  // [NWARPS][2][REC] (WTMA)
  unsigned char *txbuf = reinterpret_cast<unsigned char *>(WTMA ? rings + (size_t)NWARPS * 2 * TR::REC : tw_s + TR::TW_SMEM_ELEMS);  // [2][N][M] tx symbols
  unsigned long long *mbar = reinterpret_cast<unsigned long long *>(txbuf + 2 * N * M);  // full[2], empty[2], WTMA: + wrdy[NWARPS][2]
  unsigned *cnt = reinterpret_cast<unsigned *>(mbar + 4 + (WTMA ? NWARPS * 2 : 0));  // [N][2] bit errors, symbol errors

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // provably warp-uniform for the compiler
  const int ant = tid / NT, ft = tid % NT;
  const int nsym = a.T + a.D;
  const int q = a.q;
  const unsigned long long pol_stream = policy_evict_first(), pol_keep = policy_evict_last();

  constexpr bool TW_S = TR::TW_SMEM && !WTMA;
  if (TW_S)
    for (int i = tid; i < TR::TW_SMEM_ELEMS; i += THREADS) tw_s[i] = a.tw[i];
  const cf *tws = TW_S ? tw_s : a.tw;
  if (tid < 2 * N) cnt[tid] = 0;
  if (tid == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    mbar_init(&mbar[2], NWARPS);
    mbar_init(&mbar[3], NWARPS);
    if (WTMA)
      for (int i = 0; i < NWARPS * 2; i++) mbar_init(&mbar[4 + i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_async_smem();
  }
  __syncthreads();

  const int nf_cta = (a.n_frames - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int total = nf_cta * nsym;  // flat (frame, symbol) sequence of this CTA
  cf *Wc = fa.scratchW + (size_t)blockIdx.x * (WTMA ? N * TR::KB * TR::REC : N * N * M);
  float *gc = fa.scratchG + (size_t)blockIdx.x * 2 * N * M, *ic = gc + (size_t)N * M;
  const unsigned sym_bytes = (unsigned)(M * sizeof(cf));

  // detection constants of this warp / lane
  WarpCtx wc;
  wc.kw = warp * 64;
  wc.koff = wc.kw + 2 * lane;
  wc.Wp = Wc + wc.koff;
  wc.gp = gc + wc.koff;
  wc.slot0 = stage_base + (size_t)(warp * (WTMA ? 1 : 2)) * stage_stride;
  wc.stage_stride = stage_stride;
  wc.wring = rings + (size_t)warp * 2 * TR::REC;
  wc.wrdy = mbar + 4 + warp * 2;
  wc.Wrec = Wc;
  wc.wn = 0;
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
      // the transmitted symbol indices of this OFDM symbol ride on the same mbarrier
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
    TaskRegs<N> cur;
    auto prefetch_task0 = [&]() {
      // W/gain/isig/tx_data of this warp's first detection task, issued before the last FFT
      // stage so the L2 latency hides behind it
      if (payload && !WTMA) {
        task_load<N, M>(cur, wc, 0, 0, pol_keep);
      }
    };
    // every warp releases `buf` (its last generic-proxy access is done) so that thread 0 may
    // refill it by TMA one iteration later
    auto release_buf = [&]() {
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&mbar[2 + (g & 1)]);
    };
    mbar_wait(&mbar[g & 1], (unsigned)((g >> 1) & 1));

    // ---------------- FFT of the N antennas, in place ----------------
    {
      cf v[FF::PTS];
      const float scale = payload ? a.dn : 1.0f;
      FF::S0::template load<false>(ft, mine, v);
      group_sync<NT>(1 + ant);
      FF::S0::compute(ft, v, nullptr);
      FF::S0::template store<true, false>(ft, v, mine, 1.f);
      group_sync<NT>(1 + ant);
      if (tid == 0 && g >= 1 && g + 1 < total) {
        // the other buffer held symbol g-1: refill it once every warp has released it
        mbar_wait(&mbar[2 + ((g + 1) & 1)], (unsigned)(((g - 1) >> 1) & 1));
        issue_load(g + 1);
      }
      if (PL::NSTG == 2) prefetch_task0();
      FF::S1::template load<true>(ft, mine, v);
      group_sync<NT>(1 + ant);
      FF::S1::template compute<TW_S>(ft, v, tws + TW::OFF1);
      if (PL::NSTG == 2) {
        FF::S1::template store<false, true>(ft, v, mine, scale);
      } else {
        FF::S1::template store<true, false>(ft, v, mine, 1.f);
        group_sync<NT>(1 + ant);
        prefetch_task0();
        FF::S2::template load<true>(ft, mine, v);
        group_sync<NT>(1 + ant);
        FF::S2::template compute<TW_S>(ft, v, tws + TW::OFF2);
        FF::S2::template store<false, true>(ft, v, mine, scale);
      }
    }
    __syncthreads();

    if (!payload) {
      // ---------------- LS accumulate (mimo/framing.cc:801-815) ----------------
      const int c = sym / N, t = sym % N;
      const bool q1 = (a.flags & RUB_FLAG_Q1_IDENTITY_INIT) != 0;
      // all read-modify-write loads first (they are ordered asm volatile: interleaving them with
      // the stores would serialise one L2 round trip per element), then the adds and the stores
      constexpr int LS_IT = (N * M / 2) / THREADS;
      float4 accv[LS_IT];
#pragma unroll
      for (int i = 0; i < LS_IT; i++) {
        const int e = tid + i * THREADS, r = e / (M / 2), k = 2 * (e % (M / 2));
        if (c == 0) { const float d = (q1 && r == t) ? 1.0f : 0.0f; accv[i] = make_float4(d, 0.f, d, 0.f); }
        else accv[i] = ld_hint4(Wc + TR::template wpos<WTMA>(r * N + t, k), pol_keep);
      }
#pragma unroll
      for (int i = 0; i < LS_IT; i++) {
        const int e = tid + i * THREADS, r = e / (M / 2), k = 2 * (e % (M / 2));
        const float4 x = *reinterpret_cast<const float4 *>(buf + (size_t)r * PAD + k);
        const float2 sg = __ldg(reinterpret_cast<const float2 *>(a.sgn + ((size_t)t * a.nac + c) * M + k));
        float4 acc = accv[i];
        acc.x = acc.x + x.x * sg.x; acc.y = acc.y + x.y * sg.x;
        acc.z = acc.z + x.z * sg.y; acc.w = acc.w + x.w * sg.y;
        st_hint4(Wc + TR::template wpos<WTMA>(r * N + t, k), acc, pol_keep);
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
            const float2 t2 = ld_hint2(Wc + TR::template wpos<WTMA>(e, k), pol_keep);
            G[e] = cscale(mk(t2.x, t2.y), a.s_ls);
          }
          if (a.G) {
#pragma unroll
            for (int e = 0; e < N * N; e++) a.G[(frame * N * N + e) * M + k] = G[e];
          }
          compute_weights<N>(fa.wm, G, W, gain, isig);
#pragma unroll
          for (int e = 0; e < N * N; e++) st_hint2(Wc + TR::template wpos<WTMA>(e, k), make_float2(W[e].x, W[e].y), pol_keep);
#pragma unroll
          for (int s = 0; s < N; s++) {
            if (WTMA) {  // gain and isig ride in the task record
              float *rec = reinterpret_cast<float *>(Wc + (size_t)(s * TR::KB + (k >> 6)) * TR::REC + N * 64);
              st_hint1(rec + (k & 63), gain[s], pol_keep);
              st_hint1(rec + 64 + (k & 63), isig[s], pol_keep);
            } else {
              st_hint1(gc + (size_t)s * M + k, gain[s], pol_keep);
              st_hint1(ic + (size_t)s * M + k, isig[s], pol_keep);
            }
          }
        }
        if (WTMA) {  // the records are read through the async proxy (TMA)
          __threadfence();
          asm volatile("fence.proxy.async;" ::: "memory");
        }
        __syncthreads();  // W complete before any warp prefetches it for the first payload symbol
        if (WTMA && lane == 0) {
          // the first two task records of the frame; every later one is requested two tasks ahead by detect_symbol
          issue_record<LOG2M, N>(wc, warp, 0, (int)(wc.wn & 1u), pol_keep);
          issue_record<LOG2M, N>(wc, warp, 1, (int)((wc.wn + 1) & 1u), pol_keep);
        }
      }
    } else {
      // ---------------- detect + demap + count ----------------
#define DETECT(MBV) detect_symbol<LOG2M, N, MBV, WTMA>(fa, cur, wc, buf, txbuf + (g & 1) * N * M, symbase, lutp, refs, pol_keep, pol_stream, cnt, sym + 1 < nsym)
      switch (q) {
        case 2: DETECT(1); break;
        case 4: DETECT(2); break;
        case 6: DETECT(3); break;
        default: DETECT(4); break;
      }
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


// ---------------------------------------------------------------------------------------------
// k_detect_lean: detection for the staged path when every carrier is occupied (e.g. C4, 8x8 /
// 4096, whose 256 KB of FFT output per symbol does not fit one CTA's shared memory).  Y, W, gain,
// isig and tx_data come from HBM/L2 (written by k_fft_staged / k_weights).  One warp per (frame,
// symbol, 64-carrier block); it reads Y once and serves the N streams in turn, one carrier per
// lane (two half-tasks of 32 carriers per stream): a thread then needs ~80 registers for N = 4,
// an SM holds 24 warps instead of the 16 a two-carriers-per-lane mapping allows.  W, gain and 1/sigma^2 of a
// stream's 64 carriers come as one TMA record (one bulk copy into a two-deep per-warp ring, two streams
// ahead of their use, completed on an mbarrier): no W registers are held across half-tasks and an L2 round trip
// (the stall that dominated the register-prefetch version) has two half-tasks of work to hide behind.
// LLRs and packed bits are staged per warp in their final byte order and leave by TMA bulk store.
template <int N> struct DetectLeanTmaW { static constexpr bool value = N >= 4; };
template <int N, int MB>
__global__ void __launch_bounds__(256, (N <= 4 ? 3 : 2)) k_detect_lean(ChainArgs a, DemapConst lutp, int llr_stage_bytes) {
  constexpr int Q = 2 * MB, WARPS = 8;
  // TMAW: W / gain / isig come as TMA task records (ChainArgs::wrec).  With one or two streams a warp's whole job
  // is two or four half-tasks and the ring's prologue would be most of it: those keep the register prefetch.
  constexpr bool TMAW = DetectLeanTmaW<N>::value;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ unsigned cnt[2 * N];
  __shared__ __align__(8) unsigned long long wbar[WARPS][2];
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  if (tid < 2 * N) cnt[tid] = 0;
  if (TMAW && lane == 0) {
    mbar_init(&wbar[warp][0], 1);
    mbar_init(&wbar[warp][1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_async_smem();
  }
  __syncthreads();
  const int M = a.M;
  constexpr int WSLOT = N * 512 + 512;  // W[r][64] complex, gain[64], isig[64] of one stream
  constexpr int RSLOT = WSLOT + 64;     // ring slot: the record + the 64 transmitted symbols of the block (tx_data)
  const int blocks_per_sym = M / 64;
  const long long wid = (long long)blockIdx.x * WARPS + warp;            // (frame, symbol, block)
  const long long nwork = (long long)a.n_frames * a.D * blocks_per_sym;
  const unsigned long long pol_stream = policy_evict_first(), pol_keep = policy_evict_last();
  float refs[4];
#pragma unroll
  for (int i = 0; i < 4; i++) refs[i] = (i < MB) ? (float)(1u << (MB - 1 - i)) * lutp.alpha : 0.f;
  if (wid < nwork) {
    const int kb = (int)(wid % blocks_per_sym);
    const int d = (int)((wid / blocks_per_sym) % a.D);
    const long long frame = wid / ((long long)blocks_per_sym * a.D);
    const int k0 = kb * 64;
    const long long nsym = a.T + a.D;
    const cf *Yf = a.Y + ((frame * nsym + a.T + d) * N) * M + k0 + lane;
    const long long DM = (long long)a.D * M;
    const long long obase = (frame * N * a.D + d) * (long long)M + k0;   // warp-uniform
    // this block's task records, one per stream (ChainArgs::wrec), or the classic arrays at this lane's carrier
    const unsigned char *wrec0 = reinterpret_cast<const unsigned char *>(a.W) + (TMAW ? wrec_offset(N, M, frame, 0, k0, 0) : 0);
    const cf *Wf = a.W + frame * N * N * M + k0 + lane;
    const float *gf = a.gain + frame * N * M + k0 + lane, *sf = a.isig + frame * N * M + k0 + lane;
    const int stage_stride = llr_stage_bytes + 64;
    unsigned char *wring = smem_raw + (size_t)WARPS * 2 * stage_stride + (size_t)warp * 2 * RSLOT;
    // record of stream s_ -> ring slot s_ & 1 (one lane)
    auto issue_w = [&](int s_) {
      unsigned char *dst = wring + (s_ & 1) * RSLOT;
      unsigned long long *bar = &wbar[warp][s_ & 1];
      mbar_expect_tx(bar, (unsigned)(a.tx_data ? RSLOT : WSLOT));
      bulk_load(dst, wrec0 + (size_t)s_ * WSLOT, (unsigned)WSLOT, bar, pol_keep);
      if (a.tx_data) bulk_load(dst + WSLOT, a.tx_data + obase + s_ * DM, 64u, bar, pol_stream);
    };
    if (TMAW && lane == 0) {
      issue_w(0);
      if (N > 1) issue_w(1);
    }
    float2 y[2][N];
#pragma unroll
    for (int h = 0; h < 2; h++)
#pragma unroll
      for (int r = 0; r < N; r++) y[h][r] = ld_hint2(Yf + (long long)r * M + 32 * h, pol_stream);
    unsigned eb[ErrWords<N>::NW] = {}, es[ErrWords<N>::NW] = {};
    const bool want_llr = a.llr != nullptr;
    // half-task j = (stream j/2, carriers k0 + 32*(j%2) + lane); only the transmitted symbol of the next
    // half-task is prefetched through a register
    // (!TMAW: W, gain and isig of the next half-task are prefetched through registers as well)
    struct HalfRegs { float2 w[TMAW ? 1 : N]; float g, is; unsigned tx; };
    auto load_half = [&](HalfRegs &t, int j) {
      const int s = j >> 1, h = j & 1;
      if (!TMAW) {
#pragma unroll
        for (int r = 0; r < N; r++) t.w[TMAW ? 0 : r] = ld_hint2(Wf + (long long)(s * N + r) * M + 32 * h, pol_keep);
        t.g = ld_hint1(gf + (long long)s * M + 32 * h, pol_keep);
        t.is = ld_hint1(sf + (long long)s * M + 32 * h, pol_keep);
      }
      t.tx = (a.tx_data && !TMAW) ? (unsigned)a.tx_data[obase + s * DM + 32 * h + lane] : 0u;
    };
    HalfRegs cur, nxt;
    load_half(cur, 0);
    unsigned symh[2] = {0u, 0u};
#pragma unroll
    for (int j = 0; j < 2 * N; j++) {
      const int s = j >> 1, h = j & 1;
      if (j + 1 < 2 * N) load_half(nxt, j + 1);
      unsigned char *slot = smem_raw + (size_t)(warp * 2 + (s & 1)) * stage_stride;
      const unsigned char *wrec = wring + (s & 1) * RSLOT;
      if (h == 0) {
        if (lane == 0) bulk_wait_read<1>();
        __syncwarp();
        if (TMAW) mbar_wait(&wbar[warp][s & 1], (unsigned)((s >> 1) & 1));  // this stream's record has landed
      }
      cf wv[N], yv[N];
#pragma unroll
      for (int r = 0; r < N; r++) {
        const float2 t2 = TMAW ? *reinterpret_cast<const float2 *>(wrec + r * 512 + (32 * h + lane) * 8) : cur.w[TMAW ? 0 : r];
        wv[r] = mk(t2.x, t2.y);
        yv[r] = mk(y[h][r].x, y[h][r].y);
      }
      if (TMAW) {
        cur.g = *reinterpret_cast<const float *>(wrec + N * 512 + (32 * h + lane) * 4);
        cur.is = *reinterpret_cast<const float *>(wrec + N * 512 + 256 + (32 * h + lane) * 4);
        if (a.tx_data) cur.tx = wrec[WSLOT + 32 * h + lane];
      }
      const cf z = cscale(wy_dot<N>(wv, yv), cur.g);
      const unsigned si = slice_axis_refs<MB>(z.x, refs), sq = slice_axis_refs<MB>(z.y, refs);
      const unsigned c = (si << MB) | sq;
      symh[h] = c ^ ((c >> 1) & ~(1u << (MB - 1)));
      const long long o = obase + s * DM + 32 * h + lane;
      if (a.eq) st_hint2(a.eq + o, make_float2(z.x, z.y), pol_stream);
      if (a.rx_data) a.rx_data[o] = (unsigned char)symh[h];
      if (want_llr) {
        float2 *lp = reinterpret_cast<float2 *>(slot) + (32 * h + lane) * MB;
        float l[Q];
        const float kk = lutp.k4 * cur.is;
        llr_axis<MB>(z.x, kk, lutp, l);
        llr_axis<MB>(z.y, kk, lutp, l + MB);
        if (MB % 2 == 0) {  // 16-byte stores: half the shared-memory wavefronts of the 8-byte ones at this lane stride
#pragma unroll
          for (int b = 0; b < MB / 2; b++)
            reinterpret_cast<float4 *>(lp)[b] = make_float4(l[4 * b], l[4 * b + 1], l[4 * b + 2], l[4 * b + 3]);
        } else {
#pragma unroll
          for (int b = 0; b < MB; b++) lp[b] = make_float2(l[2 * b], l[2 * b + 1]);
        }
      }
      if (a.tx_data) {
        const unsigned x = symh[h] ^ cur.tx;
        eb[s >> 2] += (unsigned)__popc(x) << (8 * (s & 3));
        es[s >> 2] += (unsigned)(x != 0u) << (8 * (s & 3));
      }
      if (h == 1) {
        if (a.bits) {
          // lane L holds symbols k0+L and k0+32+L; 8 consecutive symbols = Q bytes, MSB first
#pragma unroll
          for (int hh = 0; hh < 2; hh++) {
            const unsigned v1 = symh[hh];
            const unsigned p1 = __shfl_down_sync(0xffffffffu, v1, 1);
            const unsigned v2 = (v1 << Q) | p1;                       // lanes 0 mod 2: 2 symbols
            const unsigned p2 = __shfl_down_sync(0xffffffffu, v2, 2);
            const unsigned v4 = (v2 << (2 * Q)) | p2;                 // lanes 0 mod 4: 4 symbols
            const unsigned p4 = __shfl_down_sync(0xffffffffu, v4, 4);
            if ((lane & 7) == 0) store_packed_bits<Q>(slot + llr_stage_bytes + (4 * hh + (lane >> 3)) * Q, v4, p4);
          }
        }
        // the proxy fence also waits for this lane's reads of the W record (their values went into the products):
        // the record two streams ahead may overwrite it
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          const long long ob = obase + s * DM;
          if (a.llr) bulk_store(a.llr + ob * Q, slot, (unsigned)(64 * Q * 4), pol_stream);
          if (a.bits) bulk_store(a.bits + (ob >> 3) * Q, slot + llr_stage_bytes, (unsigned)(8 * Q), pol_stream);
          bulk_commit();
          if (TMAW && s + 2 < N) issue_w(s + 2);
        }
      }
      if (j + 1 < 2 * N) cur = nxt;
    }
    if (a.tx_data) flush_counts<N>(eb, es, cnt);
    if (lane == 0) bulk_wait_all();
  }
  __syncthreads();
  if (a.tx_data && a.counters && tid < N) {
    const long long w0 = (long long)blockIdx.x * WARPS;
    const long long nw = nwork - w0 < WARPS ? nwork - w0 : WARPS;  // warps of this CTA that had work
    if (nw > 0) {
      atomicAdd(&a.counters[tid * 4 + 0], (unsigned long long)cnt[2 * tid]);
      atomicAdd(&a.counters[tid * 4 + 1], (unsigned long long)nw * 64 * a.q);
      atomicAdd(&a.counters[tid * 4 + 2], (unsigned long long)cnt[2 * tid + 1]);
      atomicAdd(&a.counters[tid * 4 + 3], (unsigned long long)nw * 64);
    }
  }
}

}  // namespace rub
