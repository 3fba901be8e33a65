/* stand-in for liquid/liquid.h (liquid-dsp 1.2/1.3 semantics) limited to what mimo/framing.cc and
 * mimo/framing.h use: msequence, wdelaycf, windowcf, firfilt_crcf / firfilt_rrrf and the
 * OFDMFRAME_SCTYPE_* constants. */
#ifndef RUB_SHIM_LIQUID_H
#define RUB_SHIM_LIQUID_H
#include <complex>
#include <stdlib.h>
#include <string.h>
typedef std::complex<float> liquid_float_complex;
#define OFDMFRAME_SCTYPE_NULL 0
#define OFDMFRAME_SCTYPE_PILOT 1
#define OFDMFRAME_SCTYPE_DATA 2

/* ---- msequence: Fibonacci LFSR, generator g given with the leading and trailing one; liquid
 * stores g >> 1 and shifts the parity of (v & g) in at the LSB ---- */
struct msequence_s { unsigned int m, g, a, n, v, b; };
typedef struct msequence_s *msequence;
inline msequence msequence_create(unsigned int m, unsigned int g, unsigned int a) {
  msequence ms = (msequence)malloc(sizeof(struct msequence_s));
  ms->m = m; ms->g = g >> 1;              /* generator polynomial without its most significant bit */
  ms->a = 0;                              /* initial state, bit order reversed: 0001 -> 1000 */
  for (unsigned int i = 0; i < m; i++) { ms->a <<= 1; ms->a |= (a & 1u); a >>= 1; }
  ms->n = (1u << m) - 1; ms->v = ms->a; ms->b = 0;
  return ms;
}
inline void msequence_destroy(msequence ms) { free(ms); }
inline void msequence_reset(msequence ms) { ms->v = ms->a; }
inline unsigned int msequence_advance(msequence ms) {
  ms->b = (unsigned int)__builtin_parity(ms->v & ms->g);
  ms->v <<= 1;
  ms->v |= ms->b;
  ms->v &= ms->n;
  return ms->b;
}
inline unsigned int msequence_generate_symbol(msequence ms, unsigned int bps) {
  unsigned int s = 0;
  for (unsigned int i = 0; i < bps; i++) { s <<= 1; s |= msequence_advance(ms); }
  return s;
}

/* ---- wdelaycf: read returns the sample pushed `delay` pushes ago ---- */
struct wdelaycf_s { unsigned int delay, idx; liquid_float_complex *v; };
typedef struct wdelaycf_s *wdelaycf;
inline wdelaycf wdelaycf_create(unsigned int delay) {
  wdelaycf q = (wdelaycf)malloc(sizeof(struct wdelaycf_s));
  q->delay = delay; q->idx = 0;
  q->v = (liquid_float_complex *)calloc(delay ? delay : 1, sizeof(liquid_float_complex));
  return q;
}
inline void wdelaycf_destroy(wdelaycf q) { free(q->v); free(q); }
inline void wdelaycf_read(wdelaycf q, liquid_float_complex *v) { *v = q->v[q->idx]; }
inline void wdelaycf_push(wdelaycf q, liquid_float_complex v) { q->v[q->idx] = v; q->idx = (q->idx + 1) % (q->delay ? q->delay : 1); }

/* ---- windowcf: the n most recent samples, oldest first, zero filled; read returns a
 * contiguous view (2n backing store, compacted every n pushes) ---- */
struct windowcf_s { unsigned int n, b; liquid_float_complex *v; };
typedef struct windowcf_s *windowcf;
inline windowcf windowcf_create(unsigned int n) {
  windowcf q = (windowcf)malloc(sizeof(struct windowcf_s));
  q->n = n; q->b = 0;
  q->v = (liquid_float_complex *)calloc(2 * (size_t)n, sizeof(liquid_float_complex));
  return q;
}
inline void windowcf_destroy(windowcf q) { free(q->v); free(q); }
inline void windowcf_push(windowcf q, liquid_float_complex v) {
  if (q->b == q->n) { memmove(q->v, q->v + q->n, sizeof(liquid_float_complex) * q->n); q->b = 0; }
  q->v[q->b + q->n] = v;
  q->b++;
}
inline void windowcf_read(windowcf q, liquid_float_complex **r) { *r = q->v + q->b; }

/* ---- firfilt: y[n] = sum_i h[i] x[n-i], a full dot product per execute, accumulated from the
 * oldest sample to the newest ---- */
struct firfilt_crcf_s { unsigned int n; float *h; liquid_float_complex *w; };
typedef struct firfilt_crcf_s *firfilt_crcf;
inline firfilt_crcf firfilt_crcf_create(float *h, unsigned int n) {
  firfilt_crcf q = (firfilt_crcf)malloc(sizeof(struct firfilt_crcf_s));
  q->n = n;
  q->h = (float *)malloc(sizeof(float) * n);
  memcpy(q->h, h, sizeof(float) * n);
  q->w = (liquid_float_complex *)calloc(n, sizeof(liquid_float_complex));
  return q;
}
inline void firfilt_crcf_destroy(firfilt_crcf q) { free(q->h); free(q->w); free(q); }
inline void firfilt_crcf_push(firfilt_crcf q, liquid_float_complex x) {
  memmove(q->w, q->w + 1, sizeof(liquid_float_complex) * (q->n - 1));
  q->w[q->n - 1] = x;
}
inline void firfilt_crcf_execute(firfilt_crcf q, liquid_float_complex *y) {
  float re = 0.f, im = 0.f;
  for (unsigned int i = 0; i < q->n; i++) {  /* w[i] is x[n-(N-1-i)], tap h[N-1-i] */
    re += q->h[q->n - 1 - i] * q->w[i].real();
    im += q->h[q->n - 1 - i] * q->w[i].imag();
  }
  *y = liquid_float_complex(re, im);
}
struct firfilt_rrrf_s { unsigned int n; float *h; float *w; };
typedef struct firfilt_rrrf_s *firfilt_rrrf;
inline firfilt_rrrf firfilt_rrrf_create(float *h, unsigned int n) {
  firfilt_rrrf q = (firfilt_rrrf)malloc(sizeof(struct firfilt_rrrf_s));
  q->n = n;
  q->h = (float *)malloc(sizeof(float) * n);
  memcpy(q->h, h, sizeof(float) * n);
  q->w = (float *)calloc(n, sizeof(float));
  return q;
}
inline void firfilt_rrrf_destroy(firfilt_rrrf q) { free(q->h); free(q->w); free(q); }
inline void firfilt_rrrf_push(firfilt_rrrf q, float x) {
  memmove(q->w, q->w + 1, sizeof(float) * (q->n - 1));
  q->w[q->n - 1] = x;
}
inline void firfilt_rrrf_execute(firfilt_rrrf q, float *y) {
  float acc = 0.f;
  for (unsigned int i = 0; i < q->n; i++) acc += q->h[q->n - 1 - i] * q->w[i];
  *y = acc;
}
#endif
