/*
 * rub_oracle.c — CPU ORACLE (test infrastructure, see rub_oracle.h).  Plain C99.
 *
 * Each function cites the reference lines it restates (paths relative to /root/reference).
 * Third-party arithmetic the reference calls but does not vendor is restated from its
 * published semantics: FFTW3f (unnormalised DFT), VOLK (element-wise ops), liquid-dsp
 * (msequence, square-QAM modem, windowcf/wdelay/firfilt).  None of them is version-pinned
 * upstream (mimo/makefile:8-13 only has -l flags).
 */
#include "rub_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXN 8

/* ------------------------------------------------------------ complex helpers ------- */
static inline ocf c_make(float re, float im) { ocf r; r.re = re; r.im = im; return r; }
static inline ocf c_add(ocf a, ocf b) { return c_make(a.re + b.re, a.im + b.im); }
static inline ocf c_sub(ocf a, ocf b) { return c_make(a.re - b.re, a.im - b.im); }
static inline ocf c_neg(ocf a) { return c_make(-a.re, -a.im); }
static inline ocf c_conj(ocf a) { return c_make(a.re, -a.im); }
/* complex product, contract: re = fma(a.re,b.re,-(a.im*b.im)); im = fma(a.re,b.im,a.im*b.re) */
static inline ocf c_mul(ocf a, ocf b) {
  ocf r;
  r.re = fmaf(a.re, b.re, -(a.im * b.im));
  r.im = fmaf(a.re, b.im, a.im * b.re);
  return r;
}
/* acc += a*b */
static inline ocf c_mac(ocf acc, ocf a, ocf b) {
  acc.re = fmaf(a.re, b.re, acc.re);
  acc.re = fmaf(-a.im, b.im, acc.re);
  acc.im = fmaf(a.re, b.im, acc.im);
  acc.im = fmaf(a.im, b.re, acc.im);
  return acc;
}
/* acc += conj(a)*b */
static inline ocf c_mac_conj_a(ocf acc, ocf a, ocf b) {
  acc.re = fmaf(a.re, b.re, acc.re);
  acc.re = fmaf(a.im, b.im, acc.re);
  acc.im = fmaf(a.re, b.im, acc.im);
  acc.im = fmaf(-a.im, b.re, acc.im);
  return acc;
}
/* acc += a*conj(b) */
static inline ocf c_mac_conj_b(ocf acc, ocf a, ocf b) {
  acc.re = fmaf(a.re, b.re, acc.re);
  acc.re = fmaf(a.im, b.im, acc.re);
  acc.im = fmaf(a.im, b.re, acc.im);
  acc.im = fmaf(-a.re, b.im, acc.im);
  return acc;
}
/* acc -= a*conj(b) */
static inline ocf c_msub_conj_b(ocf acc, ocf a, ocf b) {
  acc.re = fmaf(-a.re, b.re, acc.re);
  acc.re = fmaf(-a.im, b.im, acc.re);
  acc.im = fmaf(-a.im, b.re, acc.im);
  acc.im = fmaf(a.re, b.im, acc.im);
  return acc;
}

uint32_t orc_num_occupied(const orc_config *c) {
  if (!c->sctype) return c->M;
  uint32_t n = 0;
  for (uint32_t i = 0; i < c->M; i++) n += (c->sctype[i] != ORC_SC_NULL);
  return n;
}
uint32_t orc_num_training(const orc_config *c) {
  return c->estimator == ORC_EST_COMB ? c->nac : c->nac * c->N;
}
const char *orc_build_info(void) {
#ifdef __FMA__
  return "rub_oracle fp32 mirror, hw-fma";
#else
  return "rub_oracle fp32 mirror, libm-fmaf";
#endif
}

/* ------------------------------------------------------------ msequence ------------- */
/* liquid-dsp msequence_create / advance / generate_symbol (liquid <= 1.3):
 * g is stored shifted right by one, the initial state is the bit-reversed `a`, the output
 * bit is parity(v & g), and the register shifts left taking the new bit.               */
void orc_mseq_init(orc_mseq *ms, uint32_t m, uint32_t g, uint32_t a) {
  ms->m = m;
  ms->g = g >> 1;
  ms->a = 0;
  for (uint32_t i = 0; i < m; i++) { ms->a <<= 1; ms->a |= (a & 1u); a >>= 1; }
  ms->n = (1u << m) - 1u;
  ms->v = ms->a;
  ms->b = 0;
}
void orc_mseq_reset(orc_mseq *ms) { ms->v = ms->a; }
uint32_t orc_mseq_advance(orc_mseq *ms) {
  uint32_t x = ms->v & ms->g, b = 0;
  while (x) { b ^= (x & 1u); x >>= 1; }
  ms->b = b;
  ms->v <<= 1; ms->v |= b; ms->v &= ms->n;
  return b;
}
uint32_t orc_mseq_symbol(orc_mseq *ms, uint32_t bps) {
  uint32_t s = 0;
  for (uint32_t i = 0; i < bps; i++) { s <<= 1; s |= orc_mseq_advance(ms); }
  return s;
}

/* ------------------------------------------------------------ sctype ---------------- */
/* mimo/framing.cc:949-998 (both USE_ALL_CARRIERS branches) */
void orc_init_default_sctype(uint8_t *p, uint32_t M, int use_all, int add_null) {
  if (use_all) { for (uint32_t i = 0; i < M; i++) p[i] = ORC_SC_DATA; return; }
  uint32_t M2 = M / 2, G = 0;
  if (add_null) { G = M / 10; if (G < 2) G = 2; }
  uint32_t P = (M > 34) ? 8 : 4, P2 = P / 2;
  for (uint32_t i = 0; i < M; i++) p[i] = ORC_SC_NULL;
  for (uint32_t i = 1; i < M2 - G; i++) p[i] = (((i + P2) % P) == 0) ? ORC_SC_PILOT : ORC_SC_DATA;
  for (uint32_t i = 1; i < M2 - G; i++) {
    uint32_t k = M - i;
    p[k] = (((i + P2) % P) == 0) ? ORC_SC_PILOT : ORC_SC_DATA;
  }
}
/* mimo/framing.cc:1000-1030; returns nonzero on an invalid type instead of exit(1) */
int orc_validate_sctype(const uint8_t *p, uint32_t M, uint32_t *Mn, uint32_t *Mp, uint32_t *Md) {
  uint32_t n = 0, pl = 0, d = 0;
  for (uint32_t i = 0; i < M; i++) {
    if (p[i] == ORC_SC_NULL) n++;
    else if (p[i] == ORC_SC_PILOT) pl++;
    else if (p[i] == ORC_SC_DATA) d++;
    else return 1;
  }
  *Mn = n; *Mp = pl; *Md = d;
  return 0;
}

/* ------------------------------------------------------------ FFT ------------------- */
/* FFTW's internal operation order cannot be reproduced (and FFTW is absent), so the oracle
 * fixes its own: a Stockham autosort DIT with radix plan {16,8} per size, twiddles from one
 * master table tw[i] = (float)cos(2*pi*i/M), (float)(-sin(2*pi*i/M)) computed in double.
 * The result is the unnormalised DFT X[k] = sum_n x[n] exp(-2*pi*i*k*n/M) that
 * fftwf_plan_dft_1d(.., FFTW_FORWARD, ..) defines (mimo/framing.cc:368-372, :560).       */
void orc_fft_twiddles(uint32_t M, ocf *tw) {
  for (uint32_t i = 0; i < M; i++) {
    double a = 2.0 * 3.14159265358979323846 * (double)i / (double)M;
    tw[i].re = (float)cos(a);
    tw[i].im = (float)(-sin(a));
  }
}
static int fft_plan(uint32_t M, uint32_t rad[4]) {
  switch (M) {
    case 64:   rad[0] = 8;  rad[1] = 8;  return 2;
    case 128:  rad[0] = 16; rad[1] = 8;  return 2;
    case 256:  rad[0] = 16; rad[1] = 16; return 2;
    case 512:  rad[0] = 8;  rad[1] = 8;  rad[2] = 8;  return 3;
    case 1024: rad[0] = 16; rad[1] = 8;  rad[2] = 8;  return 3;
    case 2048: rad[0] = 16; rad[1] = 16; rad[2] = 8;  return 3;
    case 4096: rad[0] = 16; rad[1] = 16; rad[2] = 16; return 3;
    default: return 0;
  }
}
#define H8 0.70710678118654752440f  /* sqrt(1/2) */
#define C16 0.92387953251128675613f /* cos(pi/8) */
#define S16 0.38268343236508977173f /* sin(pi/8) */
/* forward 4-point DFT */
static inline void bfly4(ocf *a0, ocf *a1, ocf *a2, ocf *a3) {
  ocf t0 = c_add(*a0, *a2), t1 = c_sub(*a0, *a2), t2 = c_add(*a1, *a3), t3 = c_sub(*a1, *a3);
  *a0 = c_add(t0, t2);
  *a2 = c_sub(t0, t2);
  *a1 = c_make(t1.re + t3.im, t1.im - t3.re); /* t1 + (-i) t3 */
  *a3 = c_make(t1.re - t3.im, t1.im + t3.re); /* t1 - (-i) t3 */
}
/* a * w8^1, w8^1 = (1-i)/sqrt2 */
static inline ocf mul_w8_1(ocf a) { return c_make((a.re + a.im) * H8, (a.im - a.re) * H8); }
/* a * (-i) */
static inline ocf mul_mi(ocf a) { return c_make(a.im, -a.re); }
/* a * w8^3, w8^3 = (-1-i)/sqrt2 */
static inline ocf mul_w8_3(ocf a) { return c_make((a.im - a.re) * H8, -((a.re + a.im) * H8)); }
/* forward 8-point DFT, natural order in/out */
static void bfly8(ocf *v) {
  ocf e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
  ocf o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];
  bfly4(&e0, &e1, &e2, &e3);
  bfly4(&o0, &o1, &o2, &o3);
  o1 = mul_w8_1(o1); o2 = mul_mi(o2); o3 = mul_w8_3(o3);
  v[0] = c_add(e0, o0); v[4] = c_sub(e0, o0);
  v[1] = c_add(e1, o1); v[5] = c_sub(e1, o1);
  v[2] = c_add(e2, o2); v[6] = c_sub(e2, o2);
  v[3] = c_add(e3, o3); v[7] = c_sub(e3, o3);
}
/* forward 16-point DFT as 4x4, natural order in/out */
static void bfly16(ocf *v) {
  ocf u[4][4];
  for (int n0 = 0; n0 < 4; n0++) {
    u[n0][0] = v[n0]; u[n0][1] = v[n0 + 4]; u[n0][2] = v[n0 + 8]; u[n0][3] = v[n0 + 12];
    bfly4(&u[n0][0], &u[n0][1], &u[n0][2], &u[n0][3]);
  }
  /* u[n0][k1] *= w16^(n0*k1) */
  u[1][1] = c_mul(u[1][1], c_make(C16, -S16)); /* e=1 */
  u[1][2] = mul_w8_1(u[1][2]);                 /* e=2 */
  u[1][3] = c_mul(u[1][3], c_make(S16, -C16)); /* e=3 */
  u[2][1] = mul_w8_1(u[2][1]);                 /* e=2 */
  u[2][2] = mul_mi(u[2][2]);                   /* e=4 */
  u[2][3] = mul_w8_3(u[2][3]);                 /* e=6 */
  u[3][1] = c_mul(u[3][1], c_make(S16, -C16)); /* e=3 */
  u[3][2] = mul_w8_3(u[3][2]);                 /* e=6 */
  u[3][3] = c_mul(u[3][3], c_make(-C16, S16)); /* e=9 */
  for (int k1 = 0; k1 < 4; k1++) {
    bfly4(&u[0][k1], &u[1][k1], &u[2][k1], &u[3][k1]);
    v[k1] = u[0][k1]; v[k1 + 4] = u[1][k1]; v[k1 + 8] = u[2][k1]; v[k1 + 12] = u[3][k1];
  }
}
static void fft_forward_tw(uint32_t M, const ocf *in, ocf *out, const ocf *tw) {
  uint32_t rad[4];
  int ns = fft_plan(M, rad);
  ocf bufA[4096], bufB[4096];
  const ocf *src = in;
  uint32_t Ns = 1;
  for (int s = 0; s < ns; s++) {
    uint32_t R = rad[s], Q = M / R;
    ocf *dst = (s == ns - 1) ? out : ((s & 1) ? bufB : bufA);
    for (uint32_t j = 0; j < Q; j++) {
      uint32_t k = j % Ns;
      ocf v[16];
      for (uint32_t t = 0; t < R; t++) v[t] = src[j + t * Q];
      if (Ns > 1)
        for (uint32_t t = 1; t < R; t++) v[t] = c_mul(v[t], tw[t * k * (M / (Ns * R))]);
      if (R == 16) bfly16(v); else bfly8(v);
      uint32_t base = (j / Ns) * Ns * R + k;
      for (uint32_t t = 0; t < R; t++) dst[base + t * Ns] = v[t];
    }
    Ns *= R;
    src = dst;
  }
}
/* cached master twiddle tables, one per log2 size */
static ocf *g_tw[16];
static const ocf *get_tw(uint32_t M) {
  int l = 0; while ((1u << l) < M) l++;
  ocf *t;
#pragma omp critical(orc_tw)
  {
    if (!g_tw[l]) { ocf *n = (ocf *)malloc(sizeof(ocf) * M); orc_fft_twiddles(M, n); g_tw[l] = n; }
    t = g_tw[l];
  }
  return t;
}
void orc_fft_forward(uint32_t M, const ocf *in, ocf *out) { fft_forward_tw(M, in, out, get_tw(M)); }
/* FFTW_BACKWARD (mimo/framing.cc:135-139): unnormalised inverse = conj(fwd(conj(x))) */
void orc_fft_backward(uint32_t M, const ocf *in, ocf *out) {
  ocf *t = (ocf *)malloc(sizeof(ocf) * M);
  for (uint32_t i = 0; i < M; i++) t[i] = c_conj(in[i]);
  orc_fft_forward(M, t, out);
  for (uint32_t i = 0; i < M; i++) out[i] = c_conj(out[i]);
  free(t);
}

/* ------------------------------------------------------------ preambles ------------- */
/* mimo/framing.cc:1053-1111 (USE_NEW_INIT_S0): one LFSR bit per bin incl. nulls/odd (:1075),
 * even non-null bins +-1 (:1081-1088), s0 = IFFT(S0)*sqrt(1/M_S0) (:1100-1107).          */
int orc_init_S0(const uint8_t *p, uint32_t M, ocf *S0, ocf *s0, orc_mseq *ms) {
  uint32_t M_S0 = 0;
  for (uint32_t i = 0; i < M; i++) {
    uint32_t s = orc_mseq_symbol(ms, 1) & 1u;
    int is_null = p ? (p[i] == ORC_SC_NULL) : 0;
    if (is_null) S0[i] = c_make(0.f, 0.f);
    else if ((i % 2) == 0) { S0[i] = c_make(s ? 1.0f : -1.0f, 0.f); M_S0++; }
    else S0[i] = c_make(0.f, 0.f);
  }
  if (M_S0 == 0) return 1;
  float g = (float)sqrt(1.0 / (double)(float)M_S0);
  orc_fft_backward(M, S0, s0);
  for (uint32_t i = 0; i < M; i++) s0[i] = c_make(s0[i].re * g, s0[i].im * g);
  return 0;
}
/* mimo/framing.cc:1214-1262 (USE_NEW_INIT_S1, MAKE_S1_QPSK false): per code one bit per bin
 * (:1240), non-null -> BPSK_CONSTELLATION[s] = {-1,+1} (:35-39, :1246), s1 = IFFT*sqrt(1/M). */
int orc_init_S1(const uint8_t *p, uint32_t M, uint32_t nac, ocf *S1, ocf *s1, orc_mseq *ms) {
  float g = (float)sqrt(1.0 / (double)(float)M);
  for (uint32_t j = 0; j < nac; j++) {
    for (uint32_t i = 0; i < M; i++) {
      uint32_t s = orc_mseq_symbol(ms, 1) & 1u;
      int is_null = p ? (p[i] == ORC_SC_NULL) : 0;
      S1[(size_t)M * j + i] = is_null ? c_make(0.f, 0.f) : c_make(s ? 1.0f : -1.0f, 0.f);
    }
    if (s1) {
      orc_fft_backward(M, S1 + (size_t)M * j, s1 + (size_t)M * j);
      for (uint32_t i = 0; i < M; i++) {
        ocf *x = &s1[(size_t)M * j + i];
        *x = c_make(x->re * g, x->im * g);
      }
    }
  }
  return 0;
}

/* ------------------------------------------------------------ framegen -------------- */
/* framegen::write_sync_words, mimo/framing.cc:169-208 */
uint32_t orc_write_sync_words(const orc_config *c, const ocf *s0, const ocf *s1, ocf *const *tx) {
  uint32_t M = c->M, cp = c->cp_len, L = M + cp, N = c->N, nac = c->nac;
  uint32_t total = (nac * N + 1) * L, idx = 0;
  for (uint32_t s = 0; s < N; s++) memset(tx[s], 0, sizeof(ocf) * total);
  memcpy(tx[0] + idx, s0 + M - cp, sizeof(ocf) * cp); idx += cp;
  memcpy(tx[0] + idx, s0, sizeof(ocf) * M); idx += M;
  for (uint32_t ac = 0; ac < nac; ac++)
    for (uint32_t s = 0; s < N; s++) {
      const ocf *sym = s1 + ((size_t)s * nac + ac) * M;
      memcpy(tx[s] + idx, sym + M - cp, sizeof(ocf) * cp); idx += cp;
      memcpy(tx[s] + idx, sym, sizeof(ocf) * M); idx += M;
    }
  return idx;
}
/* comb training (extension): code c is one OFDM symbol on which tx t sends
 * S1[t][c][k] on bins k = t (mod P) and zero elsewhere; time domain = IFFT * sqrt(1/M)
 * exactly as ofdmframe_init_S1 scales s1 (mimo/framing.cc:1228, :1254-1257).             */
uint32_t orc_write_comb_words(const orc_config *c, const ocf *S1, ocf *const *tx) {
  uint32_t M = c->M, cp = c->cp_len, L = M + cp, N = c->N, nac = c->nac, P = c->P ? c->P : 8;
  float g = (float)sqrt(1.0 / (double)(float)M);
  ocf *X = (ocf *)malloc(sizeof(ocf) * M), *x = (ocf *)malloc(sizeof(ocf) * M);
  for (uint32_t ac = 0; ac < nac; ac++)
    for (uint32_t s = 0; s < N; s++) {
      const ocf *S = S1 + ((size_t)s * nac + ac) * M;
      for (uint32_t k = 0; k < M; k++) X[k] = (k % P == s) ? S[k] : c_make(0.f, 0.f);
      orc_fft_backward(M, X, x);
      for (uint32_t i = 0; i < M; i++) x[i] = c_make(x[i].re * g, x[i].im * g);
      memcpy(tx[s] + (size_t)ac * L, x + M - cp, sizeof(ocf) * cp);
      memcpy(tx[s] + (size_t)ac * L + cp, x, sizeof(ocf) * M);
    }
  free(X); free(x);
  return nac * L;
}
/* framegen::assemble_mimo_packet, mimo/framing.cc:210-235; dft_normalizer :115 */
uint32_t orc_assemble_mimo_packet(const orc_config *c, ocf *const *tx, const ocf *const *in) {
  uint32_t M = c->M, cp = c->cp_len, N = c->N;
  float dn = 1.0f / sqrtf((float)orc_num_occupied(c));
  ocf *X = (ocf *)malloc(sizeof(ocf) * M), *x = (ocf *)malloc(sizeof(ocf) * M);
  for (uint32_t s = 0; s < N; s++) {
    for (uint32_t i = 0, j = 0; i < M; i++) {
      int is_null = c->sctype ? (c->sctype[i] == ORC_SC_NULL) : 0;
      X[i] = is_null ? c_make(0.f, 0.f) : in[s][j++];
    }
    orc_fft_backward(M, X, x);
    for (uint32_t i = 0; i < M; i++) x[i] = c_make(x[i].re * dn, x[i].im * dn);
    memcpy(tx[s], x + M - cp, sizeof(ocf) * cp);
    memcpy(tx[s] + cp, x, sizeof(ocf) * M);
  }
  free(X); free(x);
  return M + cp;
}

/* ------------------------------------------------------------ modem ----------------- */
/* liquid-dsp modem_create_qam / modem_modulate_qam / modem_demodulate_qam for square
 * constellations: index split into I (MSBs) and Q (LSBs) halves, each gray-decoded to a
 * level, level value (2*s - P + 1)*alpha; demod = successive comparison against
 * ref[k] = 2^k*alpha (modem_demodulate_linear_array_ref), then gray-encode.              */
static float qam_alpha(uint32_t q) {
  switch (q) {
    case 2: return (float)(1.0 / sqrt(2.0));
    case 4: return (float)(1.0 / sqrt(10.0));
    case 6: return (float)(1.0 / sqrt(42.0));
    case 8: return (float)(1.0 / sqrt(170.0));
    default: return 0.f;
  }
}
static uint32_t gray_encode(uint32_t s) { return s ^ (s >> 1); }
static uint32_t gray_decode(uint32_t g) {
  uint32_t s = g;
  for (uint32_t sh = 1; sh < 32; sh <<= 1) s ^= (s >> sh);
  return s;
}
ocf orc_modulate(uint32_t q, uint32_t sym) {
  uint32_t m = q / 2, P = 1u << m;
  float alpha = qam_alpha(q);
  uint32_t s_i = gray_decode(sym >> m), s_q = gray_decode(sym & (P - 1));
  return c_make((float)(2 * (int)s_i - (int)P + 1) * alpha, (float)(2 * (int)s_q - (int)P + 1) * alpha);
}
/* returns the level index (not gray coded) of one axis */
static uint32_t slice_axis(float v, uint32_t m, float alpha) {
  uint32_t s = 0;
  for (uint32_t k = 0; k < m; k++) {
    float ref = (float)(1u << (m - k - 1)) * alpha;
    s <<= 1;
    if (v > 0) { s |= 1; v -= ref; } else { v += ref; }
  }
  return s;
}
uint32_t orc_demodulate(uint32_t q, ocf x) {
  uint32_t m = q / 2;
  float alpha = qam_alpha(q);
  uint32_t s_i = slice_axis(x.re, m, alpha), s_q = slice_axis(x.im, m, alpha);
  return (gray_encode(s_i) << m) + gray_encode(s_q);
}
/* Max-log LLR (extension): LLR_b = (min_{a:b=1}|z-a|^2 - min_{a:b=0}|z-a|^2)
 * / sigma_eff^2, positive => bit 0.  For Gray square QAM it separates per axis and has a closed
 * form in the folded residuals of the successive-comparison slicer: t_0 = x,
 * t_j = |t_{j-1}| - 2^(m-j) alpha.  Axis bit j (0 = MSB) is decided by the sign of t_j inside a
 * sub-constellation of n_j = 2^(m-1-j) levels per side, where the max-log metric is the convex
 * piecewise-linear function
 *     F_j(w) = max_{i=1..n_j} ( i*w - i(i-1) alpha ),  w = |t_j|,
 * (inside the i-th cell from the boundary the nearest opposite-bit level is i cells away), so
 *     LLR_j = -/+ 4 alpha F_j(|t_j|) / sigma_eff^2
 * with the sign of -x for j = 0 (positive levels carry gray MSB 1) and the sign of t_j for
 * j >= 1 (the outer half carries gray bit 0).  fp32 evaluation order (the arithmetic contract):
 * k = (4 alpha) * isig; terms fmaf((float)i, w, -c_i) with c_i = (float)(i(i-1) * (double)alpha),
 * max taken in ascending i starting from w; LLR = (F * k) with the sign bit xor-ed in.        */
float orc_llr_coef(uint32_t q, uint32_t i) { return (float)((double)(i * (i - 1u)) * (double)qam_alpha(q)); }
static void llr_axis(uint32_t m, float alpha, uint32_t q, float x, float k, float *llr) {
  float t = x;
  for (uint32_t j = 0; j < m; j++) {
    if (j > 0) t = fabsf(t) - (float)(1u << (m - j)) * alpha;
    float w = fabsf(t), F = w;
    uint32_t n = 1u << (m - 1 - j);
    for (uint32_t i = 2; i <= n; i++) F = fmaxf(F, fmaf((float)i, w, -orc_llr_coef(q, i)));
    float v = F * k;
    int neg = (j == 0) ? !signbit(t) : signbit(t);
    llr[j] = neg ? -v : v;
  }
}
void orc_llr(uint32_t q, ocf x, float isig, float *llr) {
  uint32_t m = q / 2;
  float alpha = qam_alpha(q);
  float k = (4.0f * alpha) * isig;
  llr_axis(m, alpha, q, x.re, k, llr);
  llr_axis(m, alpha, q, x.im, k, llr + m);
}

/* ------------------------------------------------------------ weights --------------- */
/* invert(), mimo/framing.cc:1344-1367 with INVERT_TO_UNITY false (config.h:103):
 * W = conj(det)*adj(G), returns 1/|det|^2.  G, W row-major 2x2.                          */
/* std::complex operator* as libgcc's __mulsc3 evaluates it: four rounded products, a rounded
 * difference and a rounded sum */
static inline ocf c_mul_std(ocf a, ocf b) { return c_make(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
float orc_invert_2x2(ocf W[4], const ocf G[4]) {
  ocf det = c_sub(c_mul_std(G[0], G[3]), c_mul_std(G[1], G[2]));
  ocf di = c_conj(det);
  W[0] = c_mul_std(di, G[3]);
  W[3] = c_mul_std(di, G[0]);
  W[2] = c_mul_std(c_neg(di), G[2]);
  W[1] = c_mul_std(c_neg(di), G[1]);
  return 1.0f / (det.re * det.re + det.im * det.im);
}
void orc_weights(const orc_config *c, const ocf *G, ocf *W, float *gain, float *isig) {
  int N = (int)c->N;
  float nv = c->noise_var;
  int mmse = (c->detector == ORC_DET_MMSE) && nv > 0.f;
  if (N == 2 && c->detector == ORC_DET_ZF && !(c->flags & ORC_FLAG_ZF_CHOLESKY)) {
    float g = orc_invert_2x2(W, G);
    for (int s = 0; s < 2; s++) {
      float t = W[2 * s].re * W[2 * s].re;
      t = fmaf(W[2 * s].im, W[2 * s].im, t);
      t = fmaf(W[2 * s + 1].re, W[2 * s + 1].re, t);
      t = fmaf(W[2 * s + 1].im, W[2 * s + 1].im, t);
      gain[s] = g;
      isig[s] = nv > 0.f ? 1.0f / (nv * ((g * g) * t)) : 1.0f;
    }
    return;
  }
  ocf A[MAXN][MAXN], L[MAXN][MAXN], Li[MAXN][MAXN], Ai[MAXN][MAXN];
  float inv[MAXN];
  /* A = G^H G (+ nv I), lower triangle */
  for (int i = 0; i < N; i++)
    for (int j = 0; j <= i; j++) {
      ocf acc = c_make(0.f, 0.f);
      for (int r = 0; r < N; r++) acc = c_mac_conj_a(acc, G[r * N + i], G[r * N + j]);
      if (i == j) { acc.im = 0.f; if (mmse) acc.re = acc.re + nv; }
      A[i][j] = acc;
    }
  /* Cholesky A = L L^H */
  for (int j = 0; j < N; j++) {
    float d = A[j][j].re;
    for (int p = 0; p < j; p++) { d = fmaf(-L[j][p].re, L[j][p].re, d); d = fmaf(-L[j][p].im, L[j][p].im, d); }
    float ljj = sqrtf(d);
    inv[j] = 1.0f / ljj;
    L[j][j] = c_make(ljj, 0.f);
    for (int i = j + 1; i < N; i++) {
      ocf s = A[i][j];
      for (int p = 0; p < j; p++) s = c_msub_conj_b(s, L[i][p], L[j][p]);
      L[i][j] = c_make(s.re * inv[j], s.im * inv[j]);
    }
  }
  /* Li = L^-1 (lower) */
  for (int j = 0; j < N; j++) {
    Li[j][j] = c_make(inv[j], 0.f);
    for (int i = j + 1; i < N; i++) {
      ocf s = c_make(0.f, 0.f);
      for (int p = j; p < i; p++) s = c_mac(s, L[i][p], Li[p][j]);
      Li[i][j] = c_make(-s.re * inv[i], -s.im * inv[i]);
    }
  }
  /* Ai = Li^H Li */
  for (int i = 0; i < N; i++)
    for (int j = 0; j <= i; j++) {
      ocf s = c_make(0.f, 0.f);
      for (int p = i; p < N; p++) s = c_mac_conj_a(s, Li[p][i], Li[p][j]);
      if (i == j) s.im = 0.f;
      Ai[i][j] = s;
      if (i != j) Ai[j][i] = c_conj(s);
    }
  /* W = Ai G^H */
  for (int s = 0; s < N; s++)
    for (int r = 0; r < N; r++) {
      ocf acc = c_make(0.f, 0.f);
      for (int j = 0; j < N; j++) acc = c_mac_conj_b(acc, Ai[s][j], G[r * N + j]);
      W[s * N + r] = acc;
    }
  for (int s = 0; s < N; s++) {
    float ass = Ai[s][s].re;
    if (nv > 0.f) {
      float e = nv * ass;
      if (mmse && (c->flags & ORC_FLAG_UNBIASED)) {
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

/* ------------------------------------------------------------ receive chain --------- */
static int is_null_sc(const orc_config *c, uint32_t k) {
  return c->sctype ? (c->sctype[k] == ORC_SC_NULL) : 0;
}
int orc_rx_frame(const orc_config *c, const ocf *S1, const ocf *const *rx, uint64_t first_sample,
                 const int32_t *timing, int64_t payload_start, const uint8_t *tx_data,
                 orc_frame_out *out) {
  const uint32_t M = c->M, cp = c->cp_len, L = M + cp, N = c->N, nac = c->nac, D = c->D, q = c->q;
  const uint32_t Mo = orc_num_occupied(c), T = orc_num_training(c);
  const uint32_t P = c->P ? c->P : 8;
  uint32_t rad[4];
  if (!fft_plan(M, rad) || N < 1 || N > MAXN || qam_alpha(q) == 0.f) return 1;
  if (c->estimator == ORC_EST_COMB && (N > P || Mo != M || M % P)) return 1;
  const float dn = 1.0f / sqrtf((float)Mo);  /* dft_normalizer, mimo/framing.cc:330 */
  const float s_ls = dn / (float)nac;        /* mimo/framing.cc:821 */
  ocf *X = (ocf *)malloc(sizeof(ocf) * (size_t)M * N);      /* [r][k] */
  ocf *G = (ocf *)malloc(sizeof(ocf) * (size_t)M * N * N);  /* [r][t][k] */
  ocf *W = (ocf *)malloc(sizeof(ocf) * (size_t)M * N * N);  /* [s][r][k] */
  float *gain = (float *)malloc(sizeof(float) * (size_t)M * N); /* [s][k] */
  float *isig = (float *)malloc(sizeof(float) * (size_t)M * N);
  /* --- LS estimate: G starts as identity on non-null k when Q1 (mimo/framing.cc:302-319),
   *     accumulates X/S1 per code (:801-815), scaled by dft_normalizer/nac (:817-824) --- */
  for (uint32_t r = 0; r < N; r++)
    for (uint32_t t = 0; t < N; t++)
      for (uint32_t k = 0; k < M; k++)
        G[((size_t)r * N + t) * M + k] =
            c_make(((c->flags & ORC_FLAG_Q1) && r == t && !is_null_sc(c, k)) ? 1.0f : 0.0f, 0.f);
  if (c->estimator == ORC_EST_FULLBAND) {
    for (uint32_t code = 0; code < nac; code++)
      for (uint32_t r = 0; r < N; r++)
        for (uint32_t t = 0; t < N; t++) {
          uint32_t ac = code * N + t;
          uint64_t start = timing ? (uint64_t)timing[r * T + ac] : first_sample + (uint64_t)ac * L + cp;
          orc_fft_forward(M, rx[r] + start, X);
          const ocf *S = S1 + ((size_t)t * nac + code) * M;
          for (uint32_t k = 0; k < M; k++) {
            if (is_null_sc(c, k)) continue;
            ocf *g = &G[((size_t)r * N + t) * M + k];
            /* X/S1 with S1 = +-1+0i is an exact sign flip */
            g->re = g->re + X[k].re * S[k].re;
            g->im = g->im + X[k].im * S[k].re;
          }
        }
    for (size_t i = 0; i < (size_t)M * N * N; i++) G[i] = c_make(G[i].re * s_ls, G[i].im * s_ls);
  } else {
    /* comb: LS on bins k = t (mod P), scaled like the full-band estimate, then linear
     * interpolation between a tx's pilot bins and hold at the band edges.               */
    for (uint32_t code = 0; code < nac; code++)
      for (uint32_t r = 0; r < N; r++) {
        uint64_t start = timing ? (uint64_t)timing[r * T + code] : first_sample + (uint64_t)code * L + cp;
        orc_fft_forward(M, rx[r] + start, X);
        for (uint32_t t = 0; t < N; t++) {
          const ocf *S = S1 + ((size_t)t * nac + code) * M;
          for (uint32_t k = t; k < M; k += P) {
            ocf *g = &G[((size_t)r * N + t) * M + k];
            g->re = g->re + X[k].re * S[k].re;
            g->im = g->im + X[k].im * S[k].re;
          }
        }
      }
    const float invP = 1.0f / (float)P;
    for (uint32_t r = 0; r < N; r++)
      for (uint32_t t = 0; t < N; t++) {
        ocf *g = &G[((size_t)r * N + t) * M];
        uint32_t last = t + (M / P - 1) * P;
        for (uint32_t k = t; k < M; k += P) g[k] = c_make(g[k].re * s_ls, g[k].im * s_ls);
        for (uint32_t k = 0; k < M; k++) {
          if (k % P == t) continue;
          if (k < t) { g[k] = g[t]; continue; }
          if (k > last) { g[k] = g[last]; continue; }
          uint32_t k0 = t + ((k - t) / P) * P;
          ocf a = g[k0], b = g[k0 + P];
          float f = (float)(k - k0) * invP;
          g[k] = c_make(fmaf(f, b.re - a.re, a.re), fmaf(f, b.im - a.im, a.im));
        }
      }
  }
  /* --- weights per non-null carrier (mimo/framing.cc:826-832) --- */
  for (uint32_t k = 0; k < M; k++) {
    ocf Gk[MAXN * MAXN], Wk[MAXN * MAXN];
    float gk[MAXN], ik[MAXN];
    if (is_null_sc(c, k)) {
      for (uint32_t i = 0; i < N * N; i++) W[(size_t)i * M + k] = c_make(0.f, 0.f);
      for (uint32_t s = 0; s < N; s++) { gain[(size_t)s * M + k] = 0.f; isig[(size_t)s * M + k] = 0.f; }
      continue;
    }
    for (uint32_t i = 0; i < N * N; i++) Gk[i] = G[(size_t)i * M + k];
    orc_weights(c, Gk, Wk, gk, ik);
    for (uint32_t i = 0; i < N * N; i++) W[(size_t)i * M + k] = Wk[i];
    for (uint32_t s = 0; s < N; s++) { gain[(size_t)s * M + k] = gk[s]; isig[(size_t)s * M + k] = ik[s]; }
  }
  if (out->G) memcpy(out->G, G, sizeof(ocf) * (size_t)M * N * N);
  if (out->W) memcpy(out->W, W, sizeof(ocf) * (size_t)M * N * N);
  if (out->gain) memcpy(out->gain, gain, sizeof(float) * (size_t)M * N);
  if (out->isig) memcpy(out->isig, isig, sizeof(float) * (size_t)M * N);
  /* --- payload: CP strip, FFT, scale, W*y, gain, demap (mimo/framing.cc:535-589;
   *     demod + count mimo/main.cc:1403-1410) --- */
  const uint64_t pay0 = payload_start >= 0 ? (uint64_t)payload_start : first_sample + (uint64_t)T * L;
  const uint32_t row_bytes = (Mo * q + 7) / 8;
  float llr[8];
  for (uint32_t d = 0; d < D; d++) {
    for (uint32_t r = 0; r < N; r++) {
      orc_fft_forward(M, rx[r] + pay0 + (uint64_t)d * L + cp, X + (size_t)r * M);
      for (uint32_t k = 0; k < M; k++) {
        ocf *x = &X[(size_t)r * M + k];
        *x = c_make(x->re * dn, x->im * dn);
      }
    }
    if (out->bits)
      for (uint32_t s = 0; s < N; s++) memset(out->bits + ((size_t)s * D + d) * row_bytes, 0, row_bytes);
    uint32_t j = 0;
    for (uint32_t k = 0; k < M; k++) {
      if (is_null_sc(c, k)) continue;
      for (uint32_t s = 0; s < N; s++) {
        ocf acc = c_make(0.f, 0.f);
        if (N == 2) {
          /* mimo/framing.cc:573-576: W[sc][s][0]*X[0][sc] + W[sc][s][1]*X[1][sc] in std::complex
           * arithmetic (four rounded products, a rounded difference and sum per product, then
           * the complex add); pinned by tests/golden/ref_*.npz, the reference's own output */
          const ocf w0 = W[((size_t)s * N + 0) * M + k], w1 = W[((size_t)s * N + 1) * M + k];
          const ocf x0 = X[(size_t)0 * M + k], x1 = X[(size_t)1 * M + k];
          const ocf p0 = c_make(w0.re * x0.re - w0.im * x0.im, w0.re * x0.im + w0.im * x0.re);
          const ocf p1 = c_make(w1.re * x1.re - w1.im * x1.im, w1.re * x1.im + w1.im * x1.re);
          acc = c_add(p0, p1);
        } else {
          for (uint32_t r = 0; r < N; r++) acc = c_mac(acc, W[((size_t)s * N + r) * M + k], X[(size_t)r * M + k]);
        }
        float g = gain[(size_t)s * M + k];
        ocf z = c_make(acc.re * g, acc.im * g);
        size_t o = ((size_t)s * D + d) * Mo + j;
        uint32_t sym = orc_demodulate(q, z);
        if (out->eq) out->eq[o] = z;
        if (out->rx_data) out->rx_data[o] = (uint8_t)sym;
        if (out->llr) {
          orc_llr(q, z, isig[(size_t)s * M + k], llr);
          memcpy(out->llr + o * q, llr, sizeof(float) * q);
        }
        if (out->bits) {
          uint8_t *row = out->bits + ((size_t)s * D + d) * row_bytes;
          for (uint32_t b = 0; b < q; b++) {
            uint32_t bit = (sym >> (q - 1 - b)) & 1u, pos = j * q + b;
            row[pos >> 3] |= (uint8_t)(bit << (7 - (pos & 7)));
          }
        }
        if (tx_data && out->counters) {
          uint32_t ts = tx_data[o], x = ts ^ sym, pc = 0;
          while (x) { pc += x & 1u; x >>= 1; }
          out->counters[s * 4 + 0] += pc;
          out->counters[s * 4 + 1] += q;
          out->counters[s * 4 + 2] += (ts != sym);
          out->counters[s * 4 + 3] += 1;
        }
      }
      j++;
    }
  }
  free(X); free(G); free(W); free(gain); free(isig);
  return 0;
}

int orc_rx_batch(const orc_config *c, const ocf *S1, const ocf *iq, uint64_t frame_stride,
                 uint64_t rx_stride, uint64_t first_sample, uint32_t n_frames,
                 const uint8_t *tx_data, ocf *eq, float *llr, uint8_t *bits, uint8_t *rx_data,
                 ocf *G, uint64_t *counters, int n_threads) {
  const uint32_t N = c->N, D = c->D, q = c->q, M = c->M, Mo = orc_num_occupied(c);
  const uint32_t T = orc_num_training(c), L = M + c->cp_len;
  if (!rx_stride) rx_stride = (uint64_t)(T + D) * L + first_sample;
  if (!frame_stride) frame_stride = rx_stride * N;
  const size_t per = (size_t)N * D * Mo, row_bytes = (Mo * q + 7) / 8;
  int err = 0;
  (void)get_tw(M);
#ifdef _OPENMP
  if (n_threads < 1) n_threads = 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads) if (n_threads > 1)
#endif
  for (uint32_t f = 0; f < n_frames; f++) {
    const ocf *rx[MAXN];
    uint64_t cnt[MAXN * 4];
    memset(cnt, 0, sizeof(cnt));
    for (uint32_t r = 0; r < N; r++) rx[r] = iq + (size_t)f * frame_stride + (size_t)r * rx_stride;
    orc_frame_out o;
    memset(&o, 0, sizeof(o));
    o.eq = eq ? eq + f * per : 0;
    o.llr = llr ? llr + f * per * q : 0;
    o.bits = bits ? bits + (size_t)f * N * D * row_bytes : 0;
    o.rx_data = rx_data ? rx_data + f * per : 0;
    o.G = G ? G + (size_t)f * M * N * N : 0;
    o.counters = cnt;
    int e = orc_rx_frame(c, S1, rx, first_sample, 0, -1, tx_data ? tx_data + f * per : 0, &o);
    if (e) {
#pragma omp atomic write
      err = e;
    }
    if (counters && tx_data)
      for (uint32_t i = 0; i < N * 4; i++) {
#pragma omp atomic
        counters[i] += cnt[i];
      }
  }
  return err;
}

/* ------------------------------------------------------------ faithful framesync ---- */
/* Schmidl & Cox metric, framesync::execute_sc_sync(x, stream), mimo/framing.cc:626-637:
 * wdelay(M/2) returns the sample pushed M/2 pushes ago; firfilt_crcf with M/2 taps of -1.0
 * (:342) over conj(delayed)*x; firfilt_rrrf with M taps of 0.5 (:344) over |x|^2; metric
 * |P|^2 / R^2 (:636).  liquid evaluates each FIR output as a full dot product per sample;
 * the summation order of its SIMD dotprod is not specified, this restatement sums oldest to
 * newest.                                                                                 */
void orc_sc_metric(uint32_t M, const ocf *x, uint64_t n, float *y) {
  const uint32_t M2 = M / 2;
  ocf *prod = (ocf *)calloc(n, sizeof(ocf));
  float *pw = (float *)calloc(n, sizeof(float));
  for (uint64_t i = 0; i < n; i++) {
    ocf d = (i >= M2) ? x[i - M2] : c_make(0.f, 0.f);
    ocf cd = c_conj(d);
    /* std::complex operator*: (a+bi)(c+di) = (ac-bd) + (ad+bc)i */
    prod[i] = c_make(cd.re * x[i].re - cd.im * x[i].im, cd.re * x[i].im + cd.im * x[i].re);
    pw[i] = x[i].re * x[i].re + x[i].im * x[i].im;
    ocf Pn = c_make(0.f, 0.f);
    float Rn = 0.f;
    for (uint64_t u = (i + 1 >= M2) ? i + 1 - M2 : 0; u <= i; u++) { Pn.re += -1.0f * prod[u].re; Pn.im += -1.0f * prod[u].im; }
    for (uint64_t u = (i + 1 >= M) ? i + 1 - M : 0; u <= i; u++) Rn += 0.5f * pw[u];
    y[i] = (Pn.re * Pn.re + Pn.im * Pn.im) / (Rn * Rn);
  }
  free(prod); free(pw);
}

int orc_framesync_execute(const orc_config *c, const ocf *S0, const ocf *S1,
                          const ocf *const *in, uint64_t num_samples, float threshold,
                          orc_sync_result *res, ocf *eq, ocf *G_ref, ocf *W_ref, float *gain_ref) {
  const uint32_t M = c->M, cp = c->cp_len, L = M + cp, N = c->N, nac = c->nac, D = c->D;
  const uint32_t Mo = orc_num_occupied(c), max_ac = nac * N;
  const uint64_t acb_len = (uint64_t)L * (nac * N + 4), tx_sig_len = (uint64_t)D * L; /* :284-285 */
  const uint64_t Wlen = acb_len + tx_sig_len;
  memset(res->plateau_start, 0, sizeof(res->plateau_start));
  memset(res->plateau_end, 0, sizeof(res->plateau_end));
  res->sync_index = 0; res->num_samples_processed = 0; res->state = 0; res->symbols_decoded = 0;
  /* --- STATE_SEEK_PLATEAU, execute_sc_sync(_x[]) mimo/framing.cc:591-624 --- */
  float *metric[MAXN];
  for (uint32_t s = 0; s < N; s++) { metric[s] = (float *)malloc(sizeof(float) * num_samples); orc_sc_metric(M, in[s], num_samples, metric[s]); }
  int in_plateau[MAXN] = {0};
  uint64_t i = 0, pushed = 0;
  int state = 0; /* 0 seek, 1 save, 3 mimo (framesync_states_t, mimo/framing.h:34-39) */
  for (; i < num_samples && state == 0; i++) {
    int proceed = 1;
    for (uint32_t s = 0; s < N; s++) {
      if (metric[s][i] > threshold) {
        if (in_plateau[s]) res->plateau_end[s] = res->num_samples_processed;
        else { in_plateau[s] = 1; res->plateau_start[s] = res->num_samples_processed; res->plateau_end[s] = res->num_samples_processed; }
      } else in_plateau[s] = 0;
      proceed = proceed && (res->plateau_end[s] - res->plateau_start[s] > cp) && in_plateau[s];
    }
    pushed++;
    if (proceed) {
      for (uint32_t s = 0; s < N; s++) res->sync_index += res->plateau_start[s];
      res->sync_index /= N;
      state = 1;
    }
    res->num_samples_processed++;
  }
  for (uint32_t s = 0; s < N; s++) free(metric[s]);
  if (state == 0) { res->state = 0; return 1; }
  /* --- STATE_SAVE_ACCESS_CODES, mimo/framing.cc:639-651 --- */
  for (; i < num_samples && state == 1; i++) {
    if (res->num_samples_processed - res->sync_index < tx_sig_len + acb_len - L) pushed++;
    else state = 3;
    res->num_samples_processed++;
  }
  if (state != 3) { res->state = state; return 1; }
  /* execute() processes one more sample in STATE_MIMO and breaks (mimo/framing.cc:494-504) */
  if (i < num_samples) res->num_samples_processed++;
  res->state = 3;
  /* window buffer = the Wlen most recent pushed samples, zero-filled at the front (:699-700) */
  ocf *buf[MAXN];
  for (uint32_t s = 0; s < N; s++) {
    buf[s] = (ocf *)calloc(Wlen, sizeof(ocf));
    if (pushed >= Wlen) memcpy(buf[s], in[s] + (pushed - Wlen), sizeof(ocf) * Wlen);
    else memcpy(buf[s] + (Wlen - pushed), in[s], sizeof(ocf) * pushed);
  }
  res->window_start = pushed >= Wlen ? pushed - Wlen : 0;
  /* --- timing search, mimo/framing.cc:702-744 (USE_NEW_CHANNEL_EST) --- */
  ocf *X = (ocf *)malloc(sizeof(ocf) * M);
  float max_corr[MAXN][MAXN * 32], max_s0[MAXN];
  for (uint32_t r = 0; r < N; r++) { max_s0[r] = 0.f; res->s0_corr_index[r] = 0; for (uint32_t a = 0; a < max_ac; a++) { max_corr[r][a] = 0.f; res->corr_indices[r * max_ac + a] = 0; } }
  const float MM = (float)(M * M);
  for (uint32_t off = 0; off < L; off++)
    for (uint32_t r = 0; r < N; r++) {
      orc_fft_forward(M, buf[r] + off, X);
      ocf acc = c_make(0.f, 0.f);
      for (uint32_t k = 0; k < M; k++) acc = c_add(acc, c_make(X[k].re * S0[k].re + X[k].im * S0[k].im, X[k].im * S0[k].re - X[k].re * S0[k].im));
      float v = (acc.re * acc.re + acc.im * acc.im) / MM;
      if (v > max_s0[r]) { max_s0[r] = v; res->s0_corr_index[r] = (int32_t)off; }
      for (uint32_t code = 0; code < nac; code++)
        for (uint32_t t = 0; t < N; t++) {
          uint32_t ac = code * N + t;
          uint64_t sample = off + (uint64_t)L * (ac + 1);
          orc_fft_forward(M, buf[r] + sample, X);
          const ocf *S = S1 + ((size_t)t * nac + code) * M;
          acc = c_make(0.f, 0.f);
          for (uint32_t k = 0; k < M; k++) acc = c_add(acc, c_make(X[k].re * S[k].re + X[k].im * S[k].im, X[k].im * S[k].re - X[k].re * S[k].im));
          v = (acc.re * acc.re + acc.im * acc.im) / MM;
          if (v > max_corr[r][ac]) { max_corr[r][ac] = v; res->corr_indices[r * max_ac + ac] = (int32_t)sample; }
        }
    }
  free(X);
  /* --- LS + invert + decode in the buffer (mimo/framing.cc:801-868): same arithmetic as
   *     orc_rx_frame with the per-link timing table (Q2), identity init (Q1) and the
   *     payload start taken from rx stream 1's last access code (Q4, :857) --- */
  res->payload_start = (int64_t)res->corr_indices[(N > 1 ? 1 : 0) * max_ac + max_ac - 1] + M;
  uint32_t nsym = (uint32_t)((Wlen - (uint64_t)res->payload_start) / L);
  res->symbols_decoded = nsym;
  orc_config cc = *c;
  cc.flags |= ORC_FLAG_Q1;
  cc.estimator = ORC_EST_FULLBAND;
  cc.D = nsym < D ? nsym : D; /* callback keeps only PID_MAX packets, mimo/main.cc:105-108 */
  ocf *Gp = (ocf *)malloc(sizeof(ocf) * (size_t)M * N * N), *Wp = (ocf *)malloc(sizeof(ocf) * (size_t)M * N * N);
  float *gp = (float *)malloc(sizeof(float) * (size_t)M * N);
  orc_frame_out o;
  memset(&o, 0, sizeof(o));
  ocf *eqt = (ocf *)malloc(sizeof(ocf) * (size_t)N * cc.D * Mo);
  o.eq = eqt; o.G = Gp; o.W = Wp; o.gain = gp;
  const ocf *rxp[MAXN];
  for (uint32_t s = 0; s < N; s++) rxp[s] = buf[s];
  int e = orc_rx_frame(&cc, S1, rxp, 0, res->corr_indices, res->payload_start, 0, &o);
  if (eq) {
    memset(eq, 0, sizeof(ocf) * (size_t)N * D * Mo);
    for (uint32_t s = 0; s < N; s++) memcpy(eq + (size_t)s * D * Mo, eqt + (size_t)s * cc.D * Mo, sizeof(ocf) * (size_t)cc.D * Mo);
  }
  /* reference layouts: G[k][rx][tx], W[k][rx][tx], normalize_gain[j] (mimo/framing.h:137-139) */
  for (uint32_t k = 0, j = 0; k < M; k++) {
    for (uint32_t a = 0; a < N * N; a++) {
      if (G_ref) G_ref[(size_t)k * N * N + a] = Gp[(size_t)a * M + k];
      if (W_ref) W_ref[(size_t)k * N * N + a] = Wp[(size_t)a * M + k];
    }
    if (!is_null_sc(c, k)) { if (gain_ref) gain_ref[j] = gp[k]; j++; }
  }
  free(Gp); free(Wp); free(gp); free(eqt);
  for (uint32_t s = 0; s < N; s++) free(buf[s]);
  return e;
}
