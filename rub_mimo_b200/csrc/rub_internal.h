// rub_internal.h — host-side structures shared by rub_host.cpp (pure host: tables, framegen,
// synthetic source) and rub_rx.cu (device handle + kernels).
#pragma once
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/rub_mimo/rub_mimo.h"
#include "rub_fft.cuh"

namespace rub {

struct HostCfg {
  rub_config c;                 // copy (sctype pointer replaced by `sctype` below)
  std::vector<uint8_t> sctype;  // always M entries
  uint32_t M, cp, L, N, nac, D, q, T, Mo, log2M, P;
  float dn;    // dft_normalizer = 1/sqrtf(Mo)          (mimo/framing.cc:330)
  float s_ls;  // dft_normalizer / nac                  (mimo/framing.cc:821)
  float alpha; // QAM level spacing / 2
  uint32_t row_bytes;  // ceil(Mo*q/8)
};

rub_status host_cfg_init(HostCfg &h, const rub_config *cfg);
void build_twiddles(uint32_t log2M, std::vector<cf> &master, std::vector<cf> &packed);
void build_demap_const(uint32_t q, DemapConst &dc);
float qam_alpha(uint32_t q);
// forward FFT on the host through the same stage code the kernels run (transmit side only)
void host_fft_forward(uint32_t log2M, const cf *in, cf *out, const cf *packed_tw);
void host_fft_backward(uint32_t log2M, const cf *in, cf *out, const cf *packed_tw);
void set_error(const char *fmt, ...);

}  // namespace rub
