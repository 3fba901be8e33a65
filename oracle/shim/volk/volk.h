/* stand-in for volk/volk.h: the generic (scalar, in-order) forms of the kernels
 * mimo/framing.cc calls, on std::complex<float>. */
#ifndef RUB_SHIM_VOLK_H
#define RUB_SHIM_VOLK_H
#include <complex>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
typedef std::complex<float> lv_32fc_t;
inline size_t volk_get_alignment() { return 64; }
inline bool volk_is_aligned(const void *p) { return ((uintptr_t)p % 64) == 0; }
inline void *volk_malloc(size_t size, size_t alignment) {
  void *p = nullptr;
  if (posix_memalign(&p, alignment < sizeof(void *) ? sizeof(void *) : alignment, size ? size : alignment)) return nullptr;
  return p;
}
inline void volk_free(void *p) { free(p); }
/* c[i] = a[i] * scalar, (ar + j ai)(sr + j si) = (ar sr - ai si) + j (ar si + ai sr) */
inline void volk_32fc_s32fc_multiply_32fc(lv_32fc_t *c, const lv_32fc_t *a, const lv_32fc_t scalar, unsigned int n) {
  const float sr = scalar.real(), si = scalar.imag();
  for (unsigned int i = 0; i < n; i++) {
    const float ar = a[i].real(), ai = a[i].imag();
    c[i] = lv_32fc_t(ar * sr - ai * si, ar * si + ai * sr);
  }
}
/* c[i] = a[i] * b[i] with real b */
inline void volk_32fc_32f_multiply_32fc(lv_32fc_t *c, const lv_32fc_t *a, const float *b, unsigned int n) {
  for (unsigned int i = 0; i < n; i++) c[i] = lv_32fc_t(a[i].real() * b[i], a[i].imag() * b[i]);
}
/* result = sum_i a[i] * conj(b[i]), accumulated in order */
inline void volk_32fc_x2_conjugate_dot_prod_32fc(lv_32fc_t *result, const lv_32fc_t *a, const lv_32fc_t *b, unsigned int n) {
  float re = 0.f, im = 0.f;
  for (unsigned int i = 0; i < n; i++) {
    const float ar = a[i].real(), ai = a[i].imag(), br = b[i].real(), bi = b[i].imag();
    re = re + (ar * br + ai * bi);
    im = im + (ai * br - ar * bi);
  }
  *result = lv_32fc_t(re, im);
}
#endif
