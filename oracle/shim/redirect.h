/* force-included (-include) when mimo/framing.cc is compiled: after the standard headers are in,
 * the reference's debug fopen()/printf() calls are routed to the stand-ins in shim.c so that its
 * /tmp dumps become in-memory streams and its tracing is dropped. */
#ifndef RUB_SHIM_REDIRECT_H
#define RUB_SHIM_REDIRECT_H
#include <cstdio>
#include <complex>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>
extern "C" FILE *rub_shim_fopen(const char *path, const char *mode);
extern "C" int rub_shim_printf(const char *fmt, ...);
#define fopen rub_shim_fopen
#define printf rub_shim_printf
#endif
