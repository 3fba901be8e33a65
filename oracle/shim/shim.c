/* oracle/shim/shim.c — the compiled part of the stand-ins: FFTW plans executed by the oracle's FFT
 * and the stdio redirections that keep the reference's debug dumps (/tmp/f_sc_*.dat, corr files,
 * printf tracing) out of the file system.  Test infrastructure only. */
#define _GNU_SOURCE
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../rub_oracle.h"
#include "fftw3.h"

struct rub_shim_plan_s { int n, sign; fftwf_complex *in, *out; };

void *fftwf_malloc(size_t n) { void *p = NULL; return posix_memalign(&p, 64, n ? n : 64) ? NULL : p; }
void fftwf_free(void *p) { free(p); }
fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex *in, fftwf_complex *out, int sign, unsigned flags) {
  (void)flags;
  if (n < 64 || n > 4096 || (n & (n - 1))) { fprintf(stderr, "shim fftw: unsupported size %d\n", n); abort(); }
  fftwf_plan p = (fftwf_plan)malloc(sizeof(*p));
  p->n = n; p->sign = sign; p->in = in; p->out = out;
  return p;
}
void fftwf_execute(const fftwf_plan p) {
  if (p->sign == FFTW_FORWARD) orc_fft_forward((uint32_t)p->n, (const ocf *)p->in, (ocf *)p->out);
  else orc_fft_backward((uint32_t)p->n, (const ocf *)p->in, (ocf *)p->out);
}
void fftwf_destroy_plan(fftwf_plan p) { free(p); }

/* mimo/framing.cc is compiled with -Dfopen=rub_shim_fopen -Dprintf=rub_shim_printf: its debug files
 * become in-memory streams and its tracing is dropped */
typedef struct { char *buf; size_t len; } memfile;
FILE *rub_shim_fopen(const char *path, const char *mode) {
  (void)path; (void)mode;
  memfile *m = (memfile *)calloc(1, sizeof(memfile));
  return open_memstream(&m->buf, &m->len);  /* freed at process exit; a handful per framesync */
}
int rub_shim_printf(const char *fmt, ...) { (void)fmt; return 0; }
