// rub_arith.cuh — scalar arithmetic of the receive chain, shared by every kernel.
//
// Arithmetic contract (DESIGN.md "Arithmetic contract"): IEEE-754 binary32, operations in
// source order, no implicit contraction (nvcc -fmad=false / g++ -ffp-contract=off); the only
// fused operations are the explicit fmaf() calls.  The CPU oracle (oracle/rub_oracle.c)
// restates the same sequence independently in C, which is what makes hard decisions and error
// counters comparable bit for bit.
//
// Everything here is __host__ __device__ so that (a) the host-only transmit-side helpers
// (framegen, preamble construction) reuse the FFT and (b) tests can replay the device
// arithmetic on the CPU.  The receive path itself only exists as CUDA kernels.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define RUB_HD __host__ __device__ __forceinline__
#else
#define RUB_HD inline
#endif

namespace rub {

struct __attribute__((aligned(8))) cf {
  float x, y;
};

RUB_HD cf mk(float re, float im) { cf r; r.x = re; r.y = im; return r; }
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
// Blackwell packed fp32: one FADD2 / FMUL2 does both halves of a complex add / scale.  Each half is an IEEE
// round-to-nearest operation, so the results are bit-identical to the scalar forms the oracle uses.
__device__ __forceinline__ unsigned long long pk2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ cf up2(unsigned long long v) {
  cf r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
__device__ __forceinline__ cf cadd(cf a, cf b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a.x, a.y)), "l"(pk2(b.x, b.y)));
  return up2(r);
}
__device__ __forceinline__ cf csub(cf a, cf b) {
  unsigned long long r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a.x, a.y)), "l"(pk2(b.x, b.y)));
  return up2(r);
}
__device__ __forceinline__ cf cscale(cf a, float s) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a.x, a.y)), "l"(pk2(s, s)));
  return up2(r);
}
#else
RUB_HD cf cadd(cf a, cf b) { return mk(a.x + b.x, a.y + b.y); }
RUB_HD cf csub(cf a, cf b) { return mk(a.x - b.x, a.y - b.y); }
RUB_HD cf cscale(cf a, float s) { return mk(a.x * s, a.y * s); }
#endif
RUB_HD cf cneg(cf a) { return mk(-a.x, -a.y); }
RUB_HD cf cconj(cf a) { return mk(a.x, -a.y); }
// complex product: re = fma(a.x,b.x,-(a.y*b.y)); im = fma(a.x,b.y,a.y*b.x)
RUB_HD cf cmul(cf a, cf b) {
  return mk(fmaf(a.x, b.x, -(a.y * b.y)), fmaf(a.x, b.y, a.y * b.x));
}
// acc += a*b
// std::complex operator* as the reference's decode evaluates it (libgcc __mulsc3: four rounded
// products, one rounded difference, one rounded sum; no fusion)
RUB_HD cf cmul_ref(cf a, cf b) { return mk(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
RUB_HD cf cmac(cf acc, cf a, cf b) {
  acc.x = fmaf(a.x, b.x, acc.x);
  acc.x = fmaf(-a.y, b.y, acc.x);
  acc.y = fmaf(a.x, b.y, acc.y);
  acc.y = fmaf(a.y, b.x, acc.y);
  return acc;
}
// acc += conj(a)*b
RUB_HD cf cmac_conj_a(cf acc, cf a, cf b) {
  acc.x = fmaf(a.x, b.x, acc.x);
  acc.x = fmaf(a.y, b.y, acc.x);
  acc.y = fmaf(a.x, b.y, acc.y);
  acc.y = fmaf(-a.y, b.x, acc.y);
  return acc;
}
// acc += a*conj(b)
RUB_HD cf cmac_conj_b(cf acc, cf a, cf b) {
  acc.x = fmaf(a.x, b.x, acc.x);
  acc.x = fmaf(a.y, b.y, acc.x);
  acc.y = fmaf(a.y, b.x, acc.y);
  acc.y = fmaf(-a.x, b.y, acc.y);
  return acc;
}
// acc -= a*conj(b)
RUB_HD cf cmsub_conj_b(cf acc, cf a, cf b) {
  acc.x = fmaf(-a.x, b.x, acc.x);
  acc.x = fmaf(-a.y, b.y, acc.x);
  acc.y = fmaf(-a.y, b.x, acc.y);
  acc.y = fmaf(a.x, b.y, acc.y);
  return acc;
}

// z = sum_r W[r] * y[r] for one stream and carrier.  N = 2 follows the reference to the bit
// (mimo/framing.cc:573-576: W[sc][s][0]*X[0][sc] + W[sc][s][1]*X[1][sc], std::complex arithmetic);
// the reference has no N > 2 code, there the sum is an fmaf chain in r order.
template <int N>
RUB_HD cf wy_dot(const cf *w, const cf *y) {
  if (N == 2) return cadd(cmul_ref(w[0], y[0]), cmul_ref(w[1], y[1]));
  cf acc = mk(0.f, 0.f);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int r = 0; r < N; r++) acc = cmac(acc, w[r], y[r]);
  return acc;
}

// ------------------------------------------------------------------ butterflies -------
#define RUB_H8 0.70710678118654752440f   // sqrt(1/2)
#define RUB_C16 0.92387953251128675613f  // cos(pi/8)
#define RUB_S16 0.38268343236508977173f  // sin(pi/8)

// forward 4-point DFT (w4 = -i)
RUB_HD void bfly4(cf &a0, cf &a1, cf &a2, cf &a3) {
  cf t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = csub(a1, a3);
  a0 = cadd(t0, t2);
  a2 = csub(t0, t2);
  a1 = mk(t1.x + t3.y, t1.y - t3.x);
  a3 = mk(t1.x - t3.y, t1.y + t3.x);
}
RUB_HD cf mul_w8_1(cf a) { return mk((a.x + a.y) * RUB_H8, (a.y - a.x) * RUB_H8); }
RUB_HD cf mul_mi(cf a) { return mk(a.y, -a.x); }
RUB_HD cf mul_w8_3(cf a) { return mk((a.y - a.x) * RUB_H8, -((a.x + a.y) * RUB_H8)); }

// forward 8-point DFT, natural order in and out
RUB_HD void bfly8(cf *v) {
  cf e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
  cf o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];
  bfly4(e0, e1, e2, e3);
  bfly4(o0, o1, o2, o3);
  o1 = mul_w8_1(o1);
  o2 = mul_mi(o2);
  o3 = mul_w8_3(o3);
  v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
  v[1] = cadd(e1, o1); v[5] = csub(e1, o1);
  v[2] = cadd(e2, o2); v[6] = csub(e2, o2);
  v[3] = cadd(e3, o3); v[7] = csub(e3, o3);
}
// forward 16-point DFT as 4x4, natural order in and out
RUB_HD void bfly16(cf *v) {
  cf u[4][4];
#pragma unroll
  for (int n0 = 0; n0 < 4; n0++) {
    u[n0][0] = v[n0]; u[n0][1] = v[n0 + 4]; u[n0][2] = v[n0 + 8]; u[n0][3] = v[n0 + 12];
    bfly4(u[n0][0], u[n0][1], u[n0][2], u[n0][3]);
  }
  u[1][1] = cmul(u[1][1], mk(RUB_C16, -RUB_S16));
  u[1][2] = mul_w8_1(u[1][2]);
  u[1][3] = cmul(u[1][3], mk(RUB_S16, -RUB_C16));
  u[2][1] = mul_w8_1(u[2][1]);
  u[2][2] = mul_mi(u[2][2]);
  u[2][3] = mul_w8_3(u[2][3]);
  u[3][1] = cmul(u[3][1], mk(RUB_S16, -RUB_C16));
  u[3][2] = mul_w8_3(u[3][2]);
  u[3][3] = cmul(u[3][3], mk(-RUB_C16, RUB_S16));
#pragma unroll
  for (int k1 = 0; k1 < 4; k1++) {
    bfly4(u[0][k1], u[1][k1], u[2][k1], u[3][k1]);
    v[k1] = u[0][k1]; v[k1 + 4] = u[1][k1]; v[k1 + 8] = u[2][k1]; v[k1 + 12] = u[3][k1];
  }
}
template <int R>
RUB_HD void bfly(cf *v) {
  if (R == 16) bfly16(v); else bfly8(v);
}

// ------------------------------------------------------------------ modem -------------
// liquid-dsp square QAM (replaces modem_modulate/modem_demodulate, mimo/main.cc:1237, :1405)
RUB_HD uint32_t gray_encode(uint32_t s) { return s ^ (s >> 1); }
RUB_HD uint32_t gray_decode(uint32_t g) {
  uint32_t s = g;
  s ^= s >> 1; s ^= s >> 2; s ^= s >> 4; s ^= s >> 8; s ^= s >> 16;
  return s;
}
// successive-comparison slicer of one axis (liquid modem_demodulate_linear_array_ref):
// returns the level index 0..2^m-1; ref_top = 2^(m-1)*alpha
template <int MBITS>
RUB_HD uint32_t slice_axis(float v, float alpha) {
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < MBITS; k++) {
    float ref = (float)(1u << (MBITS - k - 1)) * alpha;
    s <<= 1;
    if (v > 0) { s |= 1; v -= ref; } else { v += ref; }
  }
  return s;
}
// Same decisions, branch-free: refs[k] = 2^(m-1-k)*alpha precomputed by the caller.  v - ref and
// v + (-ref) are the same IEEE operation, so the residual chain is bit-identical.
template <int MBITS>
RUB_HD uint32_t slice_axis_refs(float v, const float *refs) {
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < MBITS; k++) {
#if defined(__CUDA_ARCH__)
    // one compare, two predicated adds, one predicated or (instead of compare/select/add/select)
    if (k + 1 < MBITS)
      asm("{\n .reg .pred p;\n setp.gt.f32 p, %1, 0f00000000;\n @p or.b32 %0, %0, %4;\n"
          " @p add.f32 %1, %1, %2;\n @!p add.f32 %1, %1, %3;\n}"
          : "+r"(s), "+f"(v)
          : "f"(-refs[k]), "f"(refs[k]), "r"(1u << (MBITS - 1 - k)));
    else
      asm("{\n .reg .pred p;\n setp.gt.f32 p, %1, 0f00000000;\n @p or.b32 %0, %0, 1;\n}" : "+r"(s) : "f"(v));
#else
    const bool p = v > 0.f;
    s |= p ? (1u << (MBITS - 1 - k)) : 0u;
    if (k + 1 < MBITS) v = v + (p ? -refs[k] : refs[k]);
#endif
  }
  return s;
}
RUB_HD uint32_t slice_axis_rt(float v, int m, float alpha) {
  uint32_t s = 0;
  for (int k = 0; k < m; k++) {
    float ref = (float)(1u << (m - k - 1)) * alpha;
    s <<= 1;
    if (v > 0) { s |= 1; v -= ref; } else { v += ref; }
  }
  return s;
}

// ------------------------------------------------------------------ max-log LLR -------
// Closed form of the max-log LLR of one axis of Gray square QAM (the oracle's orc_llr restates
// it): folded residuals t_0 = x, t_j = |t_{j-1}| - 2^(m-j) alpha; axis bit j is decided by the
// sign of t_j inside a sub-constellation of n_j = 2^(m-1-j) levels per side, where the metric is
// F_j(w) = max_{i=1..n_j}(i w - i(i-1) alpha), w = |t_j|.  LLR_j = (F_j * k) with the sign of -x
// (j = 0) or of t_j (j >= 1) xor-ed in; k = (4 alpha) / sigma_eff^2; positive => bit 0.
struct DemapConst {
  float alpha;   // level spacing / 2
  int m;         // bits per axis
  float h[4];    // h[j] = 2^(m-j) * alpha, j = 1..m-1: fold offsets = the slicer's ref[] values
  float nc[9];   // nc[i] = -(float)(i(i-1) * (double)alpha), i = 2..8
  float k4;      // 4 * alpha
};
RUB_HD uint32_t f2u(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  union { float f; uint32_t u; } c; c.f = f; return c.u;
#endif
}
RUB_HD float u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}
template <int MB>
RUB_HD void llr_axis(float x, float k, const DemapConst &dc, float *llr) {
  float t = x;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int j = 0; j < MB; j++) {
    if (j > 0) t = fabsf(t) - dc.h[j];
    const int n = 1 << (MB - 1 - j);
    if (n == 1) {
      llr[j] = (j == 0 ? -t : t) * k;   // F = |t|: (|t| k) with the sign of -/+t is (-/+t) k
    } else {
      const float w = fabsf(t);
      float F = w;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
      for (int i = 2; i <= n; i++) F = fmaxf(F, fmaf((float)i, w, dc.nc[i]));
      const uint32_t sg = (j == 0 ? ~f2u(t) : f2u(t)) & 0x80000000u;
      llr[j] = u2f(f2u(F * k) ^ sg);
    }
  }
}
RUB_HD void llr_axis_rt(float x, float k, int m, const DemapConst &dc, float *llr) {
  switch (m) {
    case 1: llr_axis<1>(x, k, dc, llr); break;
    case 2: llr_axis<2>(x, k, dc, llr); break;
    case 3: llr_axis<3>(x, k, dc, llr); break;
    default: llr_axis<4>(x, k, dc, llr); break;
  }
}

// ------------------------------------------------------------------ weights -----------
// invert(), mimo/framing.cc:1344-1367 (INVERT_TO_UNITY false): W = conj(det)*adj(G),
// returns 1/|det|^2.  Row-major 2x2.  Every product is the reference's std::complex operator*
// (cmul_ref), so W and the gain match the reference's own output bit for bit
// (tests/golden/ref_*.npz).
RUB_HD float invert_2x2(cf *W, const cf *G) {
  cf det = csub(cmul_ref(G[0], G[3]), cmul_ref(G[1], G[2]));
  cf di = cconj(det);
  W[0] = cmul_ref(di, G[3]);
  W[3] = cmul_ref(di, G[0]);
  W[2] = cmul_ref(cneg(di), G[2]);
  W[1] = cmul_ref(cneg(di), G[1]);
  return 1.0f / (det.x * det.x + det.y * det.y);
}

struct WeightMode {
  int zf2_adjugate;  // N==2 && ZF && !RUB_FLAG_ZF_CHOLESKY
  int mmse;          // detector==MMSE && noise_var>0
  int unbiased;      // RUB_FLAG_MMSE_UNBIASED
  float nv;          // noise_var
};

// G row-major [rx][tx]; W row-major [stream][rx]; gain[N]; isig[N]
template <int N>
RUB_HD void compute_weights(const WeightMode &wm, const cf *G, cf *W, float *gain, float *isig) {
  const float nv = wm.nv;
  if (N == 2 && wm.zf2_adjugate) {
    float g = invert_2x2(W, G);
#pragma unroll
    for (int s = 0; s < 2; s++) {
      float t = W[2 * s].x * W[2 * s].x;
      t = fmaf(W[2 * s].y, W[2 * s].y, t);
      t = fmaf(W[2 * s + 1].x, W[2 * s + 1].x, t);
      t = fmaf(W[2 * s + 1].y, W[2 * s + 1].y, t);
      gain[s] = g;
      isig[s] = nv > 0.f ? 1.0f / (nv * ((g * g) * t)) : 1.0f;
    }
    return;
  }
  cf A[N][N], L[N][N], Li[N][N];
  float inv[N];
#pragma unroll
  for (int i = 0; i < N; i++)
#pragma unroll
    for (int j = 0; j <= i; j++) {
      cf acc = mk(0.f, 0.f);
#pragma unroll
      for (int r = 0; r < N; r++) acc = cmac_conj_a(acc, G[r * N + i], G[r * N + j]);
      if (i == j) { acc.y = 0.f; if (wm.mmse) acc.x = acc.x + nv; }
      A[i][j] = acc;
    }
#pragma unroll
  for (int j = 0; j < N; j++) {
    float d = A[j][j].x;
#pragma unroll
    for (int p = 0; p < j; p++) { d = fmaf(-L[j][p].x, L[j][p].x, d); d = fmaf(-L[j][p].y, L[j][p].y, d); }
    float ljj = sqrtf(d);
    inv[j] = 1.0f / ljj;
    L[j][j] = mk(ljj, 0.f);
#pragma unroll
    for (int i = j + 1; i < N; i++) {
      cf s = A[i][j];
#pragma unroll
      for (int p = 0; p < j; p++) s = cmsub_conj_b(s, L[i][p], L[j][p]);
      L[i][j] = mk(s.x * inv[j], s.y * inv[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < N; j++) {
    Li[j][j] = mk(inv[j], 0.f);
#pragma unroll
    for (int i = j + 1; i < N; i++) {
      cf s = mk(0.f, 0.f);
#pragma unroll
      for (int p = j; p < i; p++) s = cmac(s, L[i][p], Li[p][j]);
      Li[i][j] = mk(-s.x * inv[i], -s.y * inv[i]);
    }
  }
  // Ai = Li^H Li (reuse A as full Hermitian storage)
#pragma unroll
  for (int i = 0; i < N; i++)
#pragma unroll
    for (int j = 0; j <= i; j++) {
      cf s = mk(0.f, 0.f);
#pragma unroll
      for (int p = i; p < N; p++) s = cmac_conj_a(s, Li[p][i], Li[p][j]);
      if (i == j) s.y = 0.f;
      A[i][j] = s;
      if (i != j) A[j][i] = cconj(s);
    }
#pragma unroll
  for (int s = 0; s < N; s++)
#pragma unroll
    for (int r = 0; r < N; r++) {
      cf acc = mk(0.f, 0.f);
#pragma unroll
      for (int j = 0; j < N; j++) acc = cmac_conj_b(acc, A[s][j], G[r * N + j]);
      W[s * N + r] = acc;
    }
#pragma unroll
  for (int s = 0; s < N; s++) {
    float ass = A[s][s].x;
    if (nv > 0.f) {
      float e = nv * ass;
      if (wm.mmse && wm.unbiased) {
        float mu = 1.0f - e;
        gain[s] = 1.0f / mu;
        isig[s] = mu / e;
      } else {
        gain[s] = 1.0f;
        isig[s] = 1.0f / e;
      }
    } else {
      gain[s] = 1.0f;
      isig[s] = 1.0f;
    }
  }
}

}  // namespace rub
