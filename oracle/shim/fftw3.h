/* stand-in for fftw3.h (single precision): plans over power-of-two sizes executed by the
 * oracle's FFT (oracle/rub_oracle.c, unnormalised DFT, FFTW sign convention). */
#ifndef RUB_SHIM_FFTW3_H
#define RUB_SHIM_FFTW3_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef float fftwf_complex[2];
typedef struct rub_shim_plan_s *fftwf_plan;
#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_MEASURE (0U)
#define FFTW_ESTIMATE (1U << 6)
#define FFTW_PATIENT (1U << 5)
void *fftwf_malloc(size_t n);
void fftwf_free(void *p);
fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex *in, fftwf_complex *out, int sign, unsigned flags);
void fftwf_execute(const fftwf_plan p);
void fftwf_destroy_plan(fftwf_plan p);
#ifdef __cplusplus
}
#endif
#endif
