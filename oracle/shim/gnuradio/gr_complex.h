/* stand-in for gnuradio/gr_complex.h: gr_complex is std::complex<float> */
#ifndef RUB_SHIM_GR_COMPLEX_H
#define RUB_SHIM_GR_COMPLEX_H
#include <complex>
typedef std::complex<float> gr_complex;
typedef std::complex<double> gr_complexd;
#endif
