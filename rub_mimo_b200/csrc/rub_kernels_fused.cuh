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
//     of the D payload symbols (L2 resident);
//   * LLRs and packed bits: staged per warp in shared memory in their final byte order and
//     written with cp.async.bulk shared->global; equalised symbols go out as 16-byte
//     coalesced stores.
#pragma once
#include <cuda_runtime.h>

#include "rub_kernels_staged.cuh"

namespace rub {

struct FusedArgs {
  ChainArgs a;
  cf *scratchW;      // [grid][N*N][M]   (G accumulates here, then W in place)
  float *scratchG;   // [grid][2][N][M]  gain, isig
  int llr_stage_bytes;  // 256*q
  WeightMode wm;
};

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
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void bulk_load(void *dst_smem, const void *src_gmem, unsigned bytes,
                                          unsigned long long *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void bulk_store(void *dst_gmem, const void *src_smem, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
               "r"(smem_u32(src_smem)), "r"(bytes)
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
__device__ __forceinline__ float4 ldcg4(const void *p) { return __ldcg(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ float2 ldcg2(const void *p) { return __ldcg(reinterpret_cast<const float2 *>(p)); }

// demap of one equalised symbol: returns the symbol index, writes 2*MB LLRs
template <int MB>
__device__ __forceinline__ unsigned demap_one(cf z, float isig, float alpha, const float *lut_slope,
                                              const float *lut_icpt, float *llr, bool want_llr) {
  constexpr int PL = 1 << MB;
  const unsigned si = slice_axis<MB>(z.x, alpha), sq = slice_axis<MB>(z.y, alpha);
  if (want_llr) {
#pragma unroll
    for (int b = 0; b < MB; b++) {
      llr[b] = fmaf(lut_slope[b * PL + si], z.x, lut_icpt[b * PL + si]) * isig;
      llr[MB + b] = fmaf(lut_slope[b * PL + sq], z.y, lut_icpt[b * PL + sq]) * isig;
    }
  }
  return (gray_encode(si) << MB) + gray_encode(sq);
}

template <int LOG2M, int N>
struct FusedTraits {
  using FF = Fft<LOG2M>;
  static constexpr int M = FF::M, NT = FF::NT, THREADS = N * NT, PAD = fft_padded_size(M);
  static constexpr int NWARPS = THREADS / 32;
  static constexpr int TASKS = N * M / 64;  // (stream, 64-carrier block) per OFDM symbol
  static constexpr int BUF_ELEMS = N * PAD;
  static_assert(THREADS % 32 == 0, "fused path needs whole warps");
  static size_t smem_bytes(int q) {
    return (size_t)2 * BUF_ELEMS * sizeof(cf) + (size_t)NWARPS * 2 * (256 * q + 64) + 2 * 64 * sizeof(float) +
           64 /* mbarriers + counters */ + 8 * N * 2;
  }
};

// one detection task: stream s, carriers [k0, k0+64); lane owns carriers k0+2*lane, +1
template <int N, int MB>
__device__ __forceinline__ void detect_task(const FusedArgs &fa, const cf *Y, int PAD, const cf *Wc,
                                            const float *gc, const float *ic, int M, int s, int k0,
                                            long long orow, float *llr_stage, unsigned char *bit_stage,
                                            const float *lut_slope, const float *lut_icpt, float alpha,
                                            unsigned &be, unsigned &se) {
  constexpr int Q = 2 * MB;
  const ChainArgs &a = fa.a;
  const int lane = threadIdx.x & 31;
  const int k = k0 + 2 * lane;
  cf acc0 = mk(0.f, 0.f), acc1 = mk(0.f, 0.f);
#pragma unroll
  for (int r = 0; r < N; r++) {
    const float4 w = ldcg4(Wc + ((long long)(s * N + r)) * M + k);
    const float4 y = *reinterpret_cast<const float4 *>(Y + (long long)r * PAD + k);
    acc0 = cmac(acc0, mk(w.x, w.y), mk(y.x, y.y));
    acc1 = cmac(acc1, mk(w.z, w.w), mk(y.z, y.w));
  }
  const float2 g = ldcg2(gc + (long long)s * M + k);
  const float2 is = ldcg2(ic + (long long)s * M + k);
  const cf z0 = cscale(acc0, g.x), z1 = cscale(acc1, g.y);
  float l0[Q], l1[Q];
  const bool want_llr = a.llr != nullptr;
  const unsigned sym0 = demap_one<MB>(z0, is.x, alpha, lut_slope, lut_icpt, l0, want_llr);
  const unsigned sym1 = demap_one<MB>(z1, is.y, alpha, lut_slope, lut_icpt, l1, want_llr);
  const long long o = orow * M + k;
  if (a.eq) *reinterpret_cast<float4 *>(a.eq + o) = make_float4(z0.x, z0.y, z1.x, z1.y);
  if (a.rx_data) *reinterpret_cast<uchar2 *>(a.rx_data + o) = make_uchar2((unsigned char)sym0, (unsigned char)sym1);
  if (want_llr) {
    float *lp = llr_stage + lane * 2 * Q;  // [k][bit] order, 2*Q floats per lane
    if (Q == 2) {
      *reinterpret_cast<float4 *>(lp) = make_float4(l0[0], l0[1], l1[0], l1[1]);
    } else {
      float tmp[2 * Q];
#pragma unroll
      for (int b = 0; b < Q; b++) { tmp[b] = l0[b]; tmp[Q + b] = l1[b]; }
#pragma unroll
      for (int v = 0; v < 2 * Q / 4; v++)
        *reinterpret_cast<float4 *>(lp + 4 * v) = make_float4(tmp[4 * v], tmp[4 * v + 1], tmp[4 * v + 2], tmp[4 * v + 3]);
    }
  }
  if (a.bits) {
    // 4 lanes = 8 symbols = Q bytes, MSB first
    unsigned v2 = (sym0 << Q) | sym1;                       // 2Q bits
    const unsigned p1 = __shfl_xor_sync(0xffffffffu, v2, 1);
    unsigned long long v4 = ((unsigned long long)v2 << (2 * Q)) | p1;  // valid on even lanes, 4Q bits
    const unsigned long long p2 = __shfl_xor_sync(0xffffffffu, v4, 2);
    if ((lane & 3) == 0) {
      const unsigned long long v8 = (v4 << (4 * Q)) | p2;   // 8Q bits = Q bytes
      unsigned char *bp = bit_stage + (lane >> 2) * Q;
#pragma unroll
      for (int i = 0; i < Q; i++) bp[i] = (unsigned char)(v8 >> (8 * (Q - 1 - i)));
    }
  }
  if (a.tx_data) {
    const uchar2 t = *reinterpret_cast<const uchar2 *>(a.tx_data + o);
    be += __popc((unsigned)t.x ^ sym0) + __popc((unsigned)t.y ^ sym1);
    se += ((unsigned)t.x != sym0) + ((unsigned)t.y != sym1);
  }
}

template <int LOG2M, int N>
__global__ void __launch_bounds__(FusedTraits<LOG2M, N>::THREADS) k_rx_fused(FusedArgs fa, DemapLut lutp) {
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
  float *lut_slope = reinterpret_cast<float *>(stage_base + (size_t)NWARPS * 2 * stage_stride);
  float *lut_icpt = lut_slope + 64;
  unsigned long long *mbar = reinterpret_cast<unsigned long long *>(lut_icpt + 64);
  unsigned *cnt = reinterpret_cast<unsigned *>(mbar + 2);  // [N][2] bit errors, symbol errors

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ant = tid / NT, ft = tid % NT;
  const int nsym = a.T + a.D;
  const int q = a.q;

  if (tid < 64) { lut_slope[tid] = lutp.slope[tid]; lut_icpt[tid] = lutp.icpt[tid]; }
  if (tid < 2 * N) cnt[tid] = 0;
  if (tid == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_async_smem();
  }
  __syncthreads();

  const int nf_cta = (a.n_frames - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const long long total = (long long)nf_cta * nsym;  // flat (frame, symbol) sequence of this CTA
  cf *Wc = fa.scratchW + (size_t)blockIdx.x * N * N * M;
  float *gc = fa.scratchG + (size_t)blockIdx.x * 2 * N * M, *ic = gc + (size_t)N * M;
  const unsigned sym_bytes = (unsigned)(M * sizeof(cf));

  auto issue_load = [&](long long g) {  // thread 0 only
    const int fl = (int)(g / nsym), sym = (int)(g % nsym);
    const long long frame = (long long)blockIdx.x + (long long)fl * gridDim.x;
    cf *dst = (g & 1) ? buf1 : buf0;
    unsigned long long *bar = &mbar[g & 1];
    mbar_expect_tx(bar, sym_bytes * N);
    const cf *src = a.iq + frame * a.frame_stride + a.first_sample + (long long)sym * a.L + a.cp;
#pragma unroll
    for (int r = 0; r < N; r++) bulk_load(dst + (size_t)r * PAD, src + (long long)r * a.rx_stride, sym_bytes, bar);
  };
  if (tid == 0) {
    if (total > 0) issue_load(0);
    if (total > 1) issue_load(1);
  }

  for (long long g = 0; g < total; g++) {
    const int fl = (int)(g / nsym), sym = (int)(g % nsym);
    const long long frame = (long long)blockIdx.x + (long long)fl * gridDim.x;
    cf *buf = (g & 1) ? buf1 : buf0;
    cf *mine = buf + (size_t)ant * PAD;
    mbar_wait(&mbar[g & 1], (unsigned)((g >> 1) & 1));

    // ---------------- FFT of the N antennas, in place ----------------
    {
      cf v[FF::PTS];
      FF::S0::template load<false>(ft, mine, v);
      group_sync<NT>(1 + ant);
      FF::S0::compute(ft, v, nullptr);
      FF::S0::template store<true, false>(ft, v, mine, 1.f);
      group_sync<NT>(1 + ant);
      FF::S1::template load<true>(ft, mine, v);
      group_sync<NT>(1 + ant);
      FF::S1::compute(ft, v, a.tw + TW::OFF1);
      const float scale = sym >= a.T ? a.dn : 1.0f;
      if (PL::NSTG == 2) {
        FF::S1::template store<false, true>(ft, v, mine, scale);
      } else {
        FF::S1::template store<true, false>(ft, v, mine, 1.f);
        group_sync<NT>(1 + ant);
        FF::S2::template load<true>(ft, mine, v);
        group_sync<NT>(1 + ant);
        FF::S2::compute(ft, v, a.tw + TW::OFF2);
        FF::S2::template store<false, true>(ft, v, mine, scale);
      }
    }
    __syncthreads();

    if (sym < a.T) {
      // ---------------- LS accumulate (mimo/framing.cc:801-815) ----------------
      const int c = sym / N, t = sym % N;
      const bool q1 = (a.flags & RUB_FLAG_Q1_IDENTITY_INIT) != 0;
      for (int e = tid; e < N * M / 2; e += THREADS) {
        const int r = e / (M / 2), k = 2 * (e % (M / 2));
        const float4 x = *reinterpret_cast<const float4 *>(buf + (size_t)r * PAD + k);
        const float2 sg = __ldg(reinterpret_cast<const float2 *>(a.sgn + ((size_t)t * a.nac + c) * M + k));
        float4 *gp = reinterpret_cast<float4 *>(Wc + (size_t)(r * N + t) * M + k);
        float4 acc;
        if (c == 0) { const float d = (q1 && r == t) ? 1.0f : 0.0f; acc = make_float4(d, 0.f, d, 0.f); }
        else acc = __ldcg(gp);
        acc.x = acc.x + x.x * sg.x; acc.y = acc.y + x.y * sg.x;
        acc.z = acc.z + x.z * sg.y; acc.w = acc.w + x.w * sg.y;
        __stcg(gp, acc);
      }
      if (sym == a.T - 1) {
        // ---------------- weights (mimo/framing.cc:817-832) ----------------
        __syncthreads();
        for (int k = tid; k < M; k += THREADS) {
          cf G[N * N], W[N * N];
          float gain[N], isig[N];
#pragma unroll
          for (int e = 0; e < N * N; e++) {
            const float2 t2 = ldcg2(Wc + (size_t)e * M + k);
            G[e] = cscale(mk(t2.x, t2.y), a.s_ls);
          }
          if (a.G) {
#pragma unroll
            for (int e = 0; e < N * N; e++) a.G[(frame * N * N + e) * M + k] = G[e];
          }
          compute_weights<N>(fa.wm, G, W, gain, isig);
#pragma unroll
          for (int e = 0; e < N * N; e++) __stcg(reinterpret_cast<float2 *>(Wc + (size_t)e * M + k), make_float2(W[e].x, W[e].y));
#pragma unroll
          for (int s = 0; s < N; s++) { __stcg(gc + (size_t)s * M + k, gain[s]); __stcg(ic + (size_t)s * M + k, isig[s]); }
        }
      }
    } else {
      // ---------------- detect + demap + count ----------------
      const int d = sym - a.T;
      for (int tsk = warp; tsk < TR::TASKS; tsk += NWARPS) {
        const int s = tsk % N, k0 = (tsk / N) * 64;
        const int it = (tsk / NWARPS) & 1;
        unsigned char *stg = stage_base + (size_t)(warp * 2 + it) * stage_stride;
        float *llr_stage = reinterpret_cast<float *>(stg);
        unsigned char *bit_stage = stg + fa.llr_stage_bytes;
        // the bulk store issued two tasks ago from this staging slot must have drained
        if (lane == 0) bulk_wait_read<1>();
        __syncwarp();
        const long long orow = (frame * N + s) * a.D + d;
        unsigned tbe = 0, tse = 0;
        switch (q) {
          case 2: detect_task<N, 1>(fa, buf, PAD, Wc, gc, ic, M, s, k0, orow, llr_stage, bit_stage, lut_slope, lut_icpt, lutp.alpha, tbe, tse); break;
          case 4: detect_task<N, 2>(fa, buf, PAD, Wc, gc, ic, M, s, k0, orow, llr_stage, bit_stage, lut_slope, lut_icpt, lutp.alpha, tbe, tse); break;
          case 6: detect_task<N, 3>(fa, buf, PAD, Wc, gc, ic, M, s, k0, orow, llr_stage, bit_stage, lut_slope, lut_icpt, lutp.alpha, tbe, tse); break;
          default: detect_task<N, 4>(fa, buf, PAD, Wc, gc, ic, M, s, k0, orow, llr_stage, bit_stage, lut_slope, lut_icpt, lutp.alpha, tbe, tse); break;
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (a.llr) bulk_store(a.llr + (orow * M + k0) * q, llr_stage, (unsigned)(64 * q * 4));
          if (a.bits) bulk_store(a.bits + orow * a.row_bytes + (long long)k0 * q / 8, bit_stage, (unsigned)(8 * q));
          bulk_commit();
        }
        if (a.tx_data) {
          for (int off = 16; off; off >>= 1) {
            tbe += __shfl_xor_sync(0xffffffffu, tbe, off);
            tse += __shfl_xor_sync(0xffffffffu, tse, off);
          }
          if (lane == 0) { atomicAdd(&cnt[2 * s], tbe); atomicAdd(&cnt[2 * s + 1], tse); }
        }
      }
    }
    __syncthreads();  // every read of `buf` is done: it may be refilled
    if (tid == 0 && g + 2 < total) { fence_async_smem(); issue_load(g + 2); }
    if (sym == nsym - 1 && a.tx_data && a.counters && tid < N) {
      atomicAdd(&a.counters[tid * 4 + 0], (unsigned long long)cnt[2 * tid]);
      atomicAdd(&a.counters[tid * 4 + 1], (unsigned long long)a.D * M * q);
      atomicAdd(&a.counters[tid * 4 + 2], (unsigned long long)cnt[2 * tid + 1]);
      atomicAdd(&a.counters[tid * 4 + 3], (unsigned long long)a.D * M);
      cnt[2 * tid] = 0;
      cnt[2 * tid + 1] = 0;
    }
  }
  if (lane == 0) bulk_wait_all();
}

}  // namespace rub
