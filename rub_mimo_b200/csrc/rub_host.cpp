// rub_host.cpp — host-only part of librubmimo_b200: configuration, table construction and the
// transmit-side / setup rows of the reference plus the
// synthetic IQ source that stands in for the USRP stream.  Nothing here is on the receive
// hot path; the receive chain exists only as CUDA kernels (rub_rx.cu).
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <thread>

#include "rub_internal.h"

namespace rub {

static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

float qam_alpha(uint32_t q) {
  switch (q) {  // liquid modem_create_qam alpha for square constellations
    case 2: return (float)(1.0 / sqrt(2.0));
    case 4: return (float)(1.0 / sqrt(10.0));
    case 6: return (float)(1.0 / sqrt(42.0));
    case 8: return (float)(1.0 / sqrt(170.0));
    default: return 0.f;
  }
}

static uint32_t ilog2(uint32_t v) { uint32_t l = 0; while ((1u << l) < v) l++; return l; }

rub_status host_cfg_init(HostCfg &h, const rub_config *cfg) {
  if (!cfg) { set_error("cfg is NULL"); return RUB_ERR_INVALID_ARG; }
  if (cfg->struct_size != sizeof(rub_config)) { set_error("rub_config.struct_size %u != %zu", cfg->struct_size, sizeof(rub_config)); return RUB_ERR_INVALID_ARG; }
  h.c = *cfg;
  h.M = cfg->M; h.cp = cfg->cp_len; h.N = cfg->num_streams; h.nac = cfg->num_access_codes;
  h.D = cfg->num_data_symbols; h.q = cfg->modulation;
  h.P = cfg->pilot_spacing ? cfg->pilot_spacing : 8;
  if (h.M < 64 || h.M > 4096 || (h.M & (h.M - 1))) { set_error("M=%u: need a power of two in 64..4096", h.M); return RUB_ERR_UNSUPPORTED; }
  // CP_LENGTH > M underflows s0 + M - cp_len in write_sync_words (quirk Q7, framing.cc:184)
  if (h.cp > h.M) { set_error("cp_len=%u > M=%u", h.cp, h.M); return RUB_ERR_INVALID_ARG; }
  if (h.N < 1 || h.N > 8) { set_error("num_streams=%u: need 1..8", h.N); return RUB_ERR_UNSUPPORTED; }
  if (h.nac < 1 || h.nac > 64) { set_error("num_access_codes=%u: need 1..64", h.nac); return RUB_ERR_INVALID_ARG; }
  if (h.D < 1) { set_error("num_data_symbols must be >= 1"); return RUB_ERR_INVALID_ARG; }
  if (qam_alpha(h.q) == 0.f) { set_error("modulation=%u: need 2, 4, 6 or 8 bits/symbol", h.q); return RUB_ERR_UNSUPPORTED; }
  if (cfg->detector > RUB_DET_MMSE) { set_error("detector=%u", cfg->detector); return RUB_ERR_INVALID_ARG; }
  if (cfg->estimator > RUB_EST_LS_COMB_INTERP) { set_error("estimator=%u", cfg->estimator); return RUB_ERR_INVALID_ARG; }
  h.log2M = ilog2(h.M);
  h.L = h.M + h.cp;
  h.sctype.assign(h.M, RUB_SCTYPE_DATA);
  if (cfg->sctype) {
    for (uint32_t i = 0; i < h.M; i++) {
      if (cfg->sctype[i] > RUB_SCTYPE_DATA) { set_error("invalid subcarrier type %u at %u", cfg->sctype[i], i); return RUB_ERR_INVALID_ARG; }
      h.sctype[i] = cfg->sctype[i];
    }
  }
  h.c.sctype = nullptr;
  h.Mo = 0;
  for (uint32_t i = 0; i < h.M; i++) h.Mo += (h.sctype[i] != RUB_SCTYPE_NULL);
  if (h.Mo == 0) { set_error("no subcarriers enabled"); return RUB_ERR_INVALID_ARG; }
  if (cfg->estimator == RUB_EST_LS_COMB_INTERP) {
    if (h.N > h.P || h.M % h.P || h.Mo != h.M) { set_error("comb estimator needs N <= P, P | M and all carriers enabled"); return RUB_ERR_UNSUPPORTED; }
    h.T = h.nac;
  } else {
    h.T = h.nac * h.N;
  }
  h.dn = 1.0f / sqrtf((float)h.Mo);
  h.s_ls = h.dn / (float)h.nac;
  h.alpha = qam_alpha(h.q);
  h.row_bytes = (h.Mo * h.q + 7) / 8;
  return RUB_OK;
}

void build_twiddles(uint32_t log2M, std::vector<cf> &master, std::vector<cf> &packed) {
  const uint32_t M = 1u << log2M;
  master.resize(M);
  for (uint32_t i = 0; i < M; i++) {
    double a = 2.0 * 3.14159265358979323846 * (double)i / (double)M;
    master[i] = mk((float)cos(a), (float)(-sin(a)));
  }
  uint32_t R[3] = {0, 0, 0}, nst = 0;
  switch (log2M) {
    case 6: R[0] = 8; R[1] = 8; nst = 2; break;
    case 7: R[0] = 16; R[1] = 8; nst = 2; break;
    case 8: R[0] = 16; R[1] = 16; nst = 2; break;
    case 9: R[0] = 8; R[1] = 8; R[2] = 8; nst = 3; break;
    case 10: R[0] = 16; R[1] = 8; R[2] = 8; nst = 3; break;
    case 11: R[0] = 16; R[1] = 16; R[2] = 8; nst = 3; break;
    case 12: R[0] = 16; R[1] = 16; R[2] = 16; nst = 3; break;
  }
  packed.clear();
  uint32_t Ns = R[0];
  for (uint32_t s = 1; s < nst; s++) {
    for (uint32_t t = 1; t < R[s]; t++)
      for (uint32_t k = 0; k < Ns; k++) packed.push_back(master[t * k * (M / (Ns * R[s]))]);
    Ns *= R[s];
  }
}

// Constants of the closed-form max-log demapper (rub_arith.cuh llr_axis, DESIGN.md "Demapper")
void build_demap_const(uint32_t q, DemapConst &dc) {
  const uint32_t m = q / 2;
  const float alpha = qam_alpha(q);
  memset(&dc, 0, sizeof(dc));
  dc.alpha = alpha;
  dc.m = (int)m;
  dc.k4 = 4.0f * alpha;
  for (uint32_t j = 1; j < m && j < 4; j++) dc.h[j] = (float)(1u << (m - j)) * alpha;
  for (uint32_t i = 2; i <= 8; i++) dc.nc[i] = -(float)((double)(i * (i - 1u)) * (double)alpha);
}

void host_fft_forward(uint32_t log2M, const cf *in, cf *out, const cf *tw) {
  switch (log2M) {
    case 6: fft_host<6>(in, out, tw, false, 1.f); break;
    case 7: fft_host<7>(in, out, tw, false, 1.f); break;
    case 8: fft_host<8>(in, out, tw, false, 1.f); break;
    case 9: fft_host<9>(in, out, tw, false, 1.f); break;
    case 10: fft_host<10>(in, out, tw, false, 1.f); break;
    case 11: fft_host<11>(in, out, tw, false, 1.f); break;
    case 12: fft_host<12>(in, out, tw, false, 1.f); break;
  }
}
// FFTW_BACKWARD (unnormalised inverse, mimo/framing.cc:135-139) = conj(fwd(conj(x)))
void host_fft_backward(uint32_t log2M, const cf *in, cf *out, const cf *tw) {
  const uint32_t M = 1u << log2M;
  std::vector<cf> t(M);
  for (uint32_t i = 0; i < M; i++) t[i] = cconj(in[i]);
  host_fft_forward(log2M, t.data(), out, tw);
  for (uint32_t i = 0; i < M; i++) out[i] = cconj(out[i]);
}

}  // namespace rub

using namespace rub;

// =============================================================== C ABI: misc ==========
extern "C" {

const char *rub_strerror(rub_status s) {
  switch (s) {
    case RUB_OK: return "ok";
    case RUB_ERR_INVALID_ARG: return "invalid argument";
    case RUB_ERR_UNSUPPORTED: return "unsupported configuration";
    case RUB_ERR_NO_DEVICE: return "no CUDA device (the receive path has no CPU fallback)";
    case RUB_ERR_CUDA: return "CUDA error";
    case RUB_ERR_NOMEM: return "out of memory";
    case RUB_ERR_NCCL: return "NCCL error";
    case RUB_ERR_IO: return "I/O error";
  }
  return "unknown";
}
const char *rub_last_error(void) { return g_err; }
uint32_t rub_abi_version(void) { return RUB_ABI_VERSION; }

uint32_t rub_config_num_training_symbols(const rub_config *cfg) {
  if (!cfg) return 0;
  return cfg->estimator == RUB_EST_LS_COMB_INTERP ? cfg->num_access_codes
                                                  : cfg->num_access_codes * cfg->num_streams;
}
uint32_t rub_config_num_occupied(const rub_config *cfg) {
  if (!cfg) return 0;
  if (!cfg->sctype) return cfg->M;
  uint32_t n = 0;
  for (uint32_t i = 0; i < cfg->M; i++) n += (cfg->sctype[i] != RUB_SCTYPE_NULL);
  return n;
}
rub_status rub_config_validate(const rub_config *cfg) {
  HostCfg h;
  return host_cfg_init(h, cfg);
}

void rub_shard_range(uint64_t n_frames, int rank, int world_size, uint64_t *begin, uint64_t *end) {
  if (world_size < 1) world_size = 1;
  if (rank < 0) rank = 0;
  const uint64_t base = n_frames / (uint64_t)world_size, rem = n_frames % (uint64_t)world_size;
  const uint64_t r = (uint64_t)rank;
  const uint64_t b = r * base + (r < rem ? r : rem);
  if (begin) *begin = b;
  if (end) *end = b + base + (r < rem ? 1 : 0);
}

// ------------------------------------------------------------- msequence --------------
// liquid-dsp msequence (liquid <= 1.3): generator stored shifted right by one, initial state =
// bit-reversed `a`, output bit = parity(v & g), register shifts left.
void rub_msequence_init(rub_msequence *ms, uint32_t m, uint32_t g, uint32_t a) {
  ms->m = m;
  ms->g = g >> 1;
  ms->a = 0;
  for (uint32_t i = 0; i < m; i++) { ms->a = (ms->a << 1) | (a & 1u); a >>= 1; }
  ms->n = (1u << m) - 1u;
  ms->v = ms->a;
  ms->b = 0;
}
void rub_msequence_reset(rub_msequence *ms) { ms->v = ms->a; }
uint32_t rub_msequence_advance(rub_msequence *ms) {
  ms->b = (uint32_t)__builtin_parity(ms->v & ms->g);
  ms->v = ((ms->v << 1) | ms->b) & ms->n;
  return ms->b;
}
uint32_t rub_msequence_generate_symbol(rub_msequence *ms, uint32_t bps) {
  uint32_t s = 0;
  for (uint32_t i = 0; i < bps; i++) s = (s << 1) | rub_msequence_advance(ms);
  return s;
}

// ------------------------------------------------------------- sctype -----------------
// mimo/framing.cc:949-998
void rub_ofdmframe_init_default_sctype(uint8_t *p, uint32_t M, int use_all, int add_null) {
  if (use_all) { memset(p, RUB_SCTYPE_DATA, M); return; }
  const uint32_t M2 = M / 2;
  uint32_t G = 0;
  if (add_null) G = std::max(M / 10, 2u);
  const uint32_t P = (M > 34) ? 8 : 4, P2 = P / 2;
  memset(p, RUB_SCTYPE_NULL, M);
  for (uint32_t i = 1; i < M2 - G; i++) {
    const uint8_t ty = (((i + P2) % P) == 0) ? RUB_SCTYPE_PILOT : RUB_SCTYPE_DATA;
    p[i] = ty;       // upper band
    p[M - i] = ty;   // lower band
  }
}
// mimo/framing.cc:1000-1030
rub_status rub_ofdmframe_validate_sctype(const uint8_t *p, uint32_t M, uint32_t *Mn, uint32_t *Mp,
                                         uint32_t *Md) {
  uint32_t cnt[3] = {0, 0, 0};
  for (uint32_t i = 0; i < M; i++) {
    if (p[i] > RUB_SCTYPE_DATA) { set_error("ofdmframe_validate_sctype: invalid subcarrier type (%u)", p[i]); return RUB_ERR_INVALID_ARG; }
    cnt[p[i]]++;
  }
  if (Mn) *Mn = cnt[0];
  if (Mp) *Mp = cnt[1];
  if (Md) *Md = cnt[2];
  return RUB_OK;
}

// ------------------------------------------------------------- preambles --------------
static bool pow2_ok(uint32_t M) { return M >= 64 && M <= 4096 && !(M & (M - 1)); }
// mimo/framing.cc:1053-1111
rub_status rub_ofdmframe_init_S0(const uint8_t *p, uint32_t M, float *S0f, float *s0f, rub_msequence *ms) {
  if (!pow2_ok(M) || !S0f || !ms) { set_error("init_S0: bad arguments"); return RUB_ERR_INVALID_ARG; }
  cf *S0 = reinterpret_cast<cf *>(S0f);
  uint32_t M_S0 = 0;
  for (uint32_t i = 0; i < M; i++) {
    const uint32_t s = rub_msequence_generate_symbol(ms, 1) & 1u;
    const bool nul = p && p[i] == RUB_SCTYPE_NULL;
    if (!nul && (i % 2) == 0) { S0[i] = mk(s ? 1.0f : -1.0f, 0.f); M_S0++; }
    else S0[i] = mk(0.f, 0.f);
  }
  if (M_S0 == 0) { set_error("ofdmframe_init_S0: no subcarriers enabled; check allocation"); return RUB_ERR_INVALID_ARG; }
  if (s0f) {
    cf *s0 = reinterpret_cast<cf *>(s0f);
    std::vector<cf> master, tw;
    const uint32_t l2 = ilog2(M);
    build_twiddles(l2, master, tw);
    const float g = (float)sqrt(1.0 / (double)(float)M_S0);
    host_fft_backward(l2, S0, s0, tw.data());
    for (uint32_t i = 0; i < M; i++) s0[i] = cscale(s0[i], g);
  }
  return RUB_OK;
}
// mimo/framing.cc:1214-1262
rub_status rub_ofdmframe_init_S1(const uint8_t *p, uint32_t M, uint32_t nac, float *S1f, float *s1f,
                                 rub_msequence *ms) {
  if (!pow2_ok(M) || !S1f || !ms) { set_error("init_S1: bad arguments"); return RUB_ERR_INVALID_ARG; }
  cf *S1 = reinterpret_cast<cf *>(S1f);
  cf *s1 = reinterpret_cast<cf *>(s1f);
  std::vector<cf> master, tw;
  const uint32_t l2 = ilog2(M);
  if (s1) build_twiddles(l2, master, tw);
  const float g = (float)sqrt(1.0 / (double)(float)M);
  for (uint32_t j = 0; j < nac; j++) {
    for (uint32_t i = 0; i < M; i++) {
      const uint32_t s = rub_msequence_generate_symbol(ms, 1) & 1u;
      const bool nul = p && p[i] == RUB_SCTYPE_NULL;
      S1[(size_t)M * j + i] = nul ? mk(0.f, 0.f) : mk(s ? 1.0f : -1.0f, 0.f);
    }
    if (s1) {
      host_fft_backward(l2, S1 + (size_t)M * j, s1 + (size_t)M * j, tw.data());
      for (uint32_t i = 0; i < M; i++) s1[(size_t)M * j + i] = cscale(s1[(size_t)M * j + i], g);
    }
  }
  return RUB_OK;
}
// LFSR_LARGE_0/1_GEN_POLY (mimo/config.h:74-75) for streams 0/1; the reference has only two
// (mimo/main.cc:1265-1267), streams 2..7 use the next degree-13 maximal-length polynomials.
uint32_t rub_default_lfsr_poly(uint32_t stream) {
  static const uint32_t polys[8] = {020033, 020047, 020065, 020123, 020145, 020157, 020213, 020215};
  return polys[stream & 7];
}
rub_status rub_default_S1(const rub_config *cfg, float *S1, float *s1) {
  HostCfg h;
  rub_status st = host_cfg_init(h, cfg);
  if (st) return st;
  for (uint32_t t = 0; t < h.N; t++) {
    rub_msequence ms;
    rub_msequence_init(&ms, 13, rub_default_lfsr_poly(t), 1);  // LFSR_LARGE_LENGTH, main.cc:1269
    const size_t off = (size_t)t * h.nac * h.M * 2;
    st = rub_ofdmframe_init_S1(h.sctype.data(), h.M, h.nac, S1 + off, s1 ? s1 + off : nullptr, &ms);
    if (st) return st;
  }
  return RUB_OK;
}
rub_status rub_default_S0(const rub_config *cfg, float *S0, float *s0) {
  HostCfg h;
  rub_status st = host_cfg_init(h, cfg);
  if (st) return st;
  rub_msequence ms;
  rub_msequence_init(&ms, 12, 010123, 1);  // LFSR_SMALL_LENGTH / LFSR_SMALL_0_GEN_POLY
  return rub_ofdmframe_init_S0(h.sctype.data(), h.M, S0, s0, &ms);
}

// ------------------------------------------------------------- invert / modem ---------
rub_status rub_invert_2x2(float *W, const float *G, float *gain) {
  if (!W || !G || !gain) return RUB_ERR_INVALID_ARG;
  *gain = invert_2x2(reinterpret_cast<cf *>(W), reinterpret_cast<const cf *>(G));
  return RUB_OK;
}
rub_status rub_modem_modulate(uint32_t q, uint32_t sym, float out[2]) {
  const float alpha = qam_alpha(q);
  if (alpha == 0.f || sym >= (1u << q)) { set_error("modulate: q=%u sym=%u", q, sym); return RUB_ERR_INVALID_ARG; }
  const uint32_t m = q / 2, P = 1u << m;
  const uint32_t s_i = gray_decode(sym >> m), s_q = gray_decode(sym & (P - 1));
  out[0] = (float)(2 * (int)s_i - (int)P + 1) * alpha;
  out[1] = (float)(2 * (int)s_q - (int)P + 1) * alpha;
  return RUB_OK;
}
rub_status rub_modem_demodulate(uint32_t q, const float in[2], uint32_t *sym) {
  const float alpha = qam_alpha(q);
  if (alpha == 0.f || !sym) return RUB_ERR_INVALID_ARG;
  const int m = (int)q / 2;
  const uint32_t s_i = slice_axis_rt(in[0], m, alpha), s_q = slice_axis_rt(in[1], m, alpha);
  *sym = (gray_encode(s_i) << m) + gray_encode(s_q);
  return RUB_OK;
}

// ------------------------------------------------------------- framegen ---------------
struct rub_framegen {
  HostCfg h;
  std::vector<cf> S0, s0, s1, S1;  // s1/S1 [N][nac][M]
  std::vector<cf> tw_master, tw;
  std::vector<cf> X, x;
};

rub_status rub_framegen_create(rub_framegen **out, const rub_config *cfg, const float *S0,
                               const float *s0, const float *s1) {
  if (!out) return RUB_ERR_INVALID_ARG;
  rub_framegen *fg = new rub_framegen();
  rub_status st = host_cfg_init(fg->h, cfg);
  if (st) { delete fg; return st; }
  const HostCfg &h = fg->h;
  build_twiddles(h.log2M, fg->tw_master, fg->tw);
  fg->S0.resize(h.M); fg->s0.resize(h.M);
  fg->S1.resize((size_t)h.N * h.nac * h.M); fg->s1.resize((size_t)h.N * h.nac * h.M);
  rub_config c2 = h.c;
  c2.sctype = h.sctype.data();
  if (s0) { memcpy(fg->s0.data(), s0, sizeof(cf) * h.M); if (S0) memcpy(fg->S0.data(), S0, sizeof(cf) * h.M); }
  else { st = rub_default_S0(&c2, (float *)fg->S0.data(), (float *)fg->s0.data()); if (st) { delete fg; return st; } }
  if (s1) memcpy(fg->s1.data(), s1, sizeof(cf) * fg->s1.size());
  else { st = rub_default_S1(&c2, (float *)fg->S1.data(), (float *)fg->s1.data()); if (st) { delete fg; return st; } }
  fg->X.resize(h.M); fg->x.resize(h.M);
  *out = fg;
  return RUB_OK;
}
void rub_framegen_destroy(rub_framegen *fg) { delete fg; }

static void write_cp_symbol(cf *dst, const cf *sym, uint32_t M, uint32_t cp) {
  memcpy(dst, sym + M - cp, sizeof(cf) * cp);
  memcpy(dst + cp, sym, sizeof(cf) * M);
}
// TDMA access codes, code-major / stream-minor (mimo/framing.cc:191-204)
static void write_access_codes(const rub_framegen *fg, cf *const *tx, uint64_t offset) {
  const HostCfg &h = fg->h;
  uint64_t idx = offset;
  for (uint32_t ac = 0; ac < h.nac; ac++)
    for (uint32_t s = 0; s < h.N; s++) {
      write_cp_symbol(tx[s] + idx, fg->s1.data() + ((size_t)s * h.nac + ac) * h.M, h.M, h.cp);
      idx += h.L;
    }
}
// framegen::write_sync_words, mimo/framing.cc:169-208
uint32_t rub_framegen_write_sync_words(rub_framegen *fg, float *const *tx_buff) {
  const HostCfg &h = fg->h;
  cf *const *tx = reinterpret_cast<cf *const *>(tx_buff);
  const uint32_t total = (h.nac * h.N + 1) * h.L;
  for (uint32_t s = 0; s < h.N; s++) memset(tx[s], 0, sizeof(cf) * total);
  write_cp_symbol(tx[0], fg->s0.data(), h.M, h.cp);  // S0 on stream 0 only (:183-190)
  write_access_codes(fg, tx, h.L);
  return total;
}
// comb training symbols (extension): tx t sends S1[t][c][k] on bins k = t (mod P)
static void comb_symbol(rub_framegen *fg, uint32_t s, uint32_t ac, cf *dst) {
  const HostCfg &h = fg->h;
  const cf *S = fg->S1.data() + ((size_t)s * h.nac + ac) * h.M;
  for (uint32_t k = 0; k < h.M; k++) fg->X[k] = (k % h.P == s) ? S[k] : mk(0.f, 0.f);
  host_fft_backward(h.log2M, fg->X.data(), fg->x.data(), fg->tw.data());
  const float g = (float)sqrt(1.0 / (double)(float)h.M);
  for (uint32_t i = 0; i < h.M; i++) fg->x[i] = cscale(fg->x[i], g);
  write_cp_symbol(dst, fg->x.data(), h.M, h.cp);
}
uint32_t rub_framegen_write_comb_words(rub_framegen *fg, float *const *tx_buff) {
  const HostCfg &h = fg->h;
  cf *const *tx = reinterpret_cast<cf *const *>(tx_buff);
  for (uint32_t ac = 0; ac < h.nac; ac++)
    for (uint32_t s = 0; s < h.N; s++) comb_symbol(fg, s, ac, tx[s] + (size_t)ac * h.L);
  return h.nac * h.L;
}
// framegen::assemble_mimo_packet, mimo/framing.cc:210-235 (dft_normalizer :115)
uint32_t rub_framegen_assemble_mimo_packet(rub_framegen *fg, float *const *tx_buff,
                                           const float *const *in_buff) {
  const HostCfg &h = fg->h;
  for (uint32_t s = 0; s < h.N; s++) {
    const cf *in = reinterpret_cast<const cf *>(in_buff[s]);
    for (uint32_t i = 0, j = 0; i < h.M; i++)
      fg->X[i] = (h.sctype[i] == RUB_SCTYPE_NULL) ? mk(0.f, 0.f) : in[j++];
    host_fft_backward(h.log2M, fg->X.data(), fg->x.data(), fg->tw.data());
    for (uint32_t i = 0; i < h.M; i++) fg->x[i] = cscale(fg->x[i], h.dn);
    write_cp_symbol(reinterpret_cast<cf *>(tx_buff[s]), fg->x.data(), h.M, h.cp);
  }
  return h.L;
}

// ------------------------------------------------------------- synthetic source -------
// Counter-based RNG: splitmix64 finaliser over (seed, frame, lane, index); any frame can be
// regenerated independently of how the batch is sharded.
static inline uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static inline uint64_t rnd_u64(uint64_t seed, uint64_t frame, uint64_t lane, uint64_t idx) {
  return mix64(mix64(mix64(seed ^ 0xA5A5A5A5DEADBEEFull) + frame * 0x100000001B3ull) + (lane << 40) + idx);
}
static inline double rnd_unit(uint64_t u) { return ((double)(u >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
static inline void rnd_gauss2(uint64_t seed, uint64_t frame, uint64_t lane, uint64_t idx, double *a, double *b) {
  const double u1 = rnd_unit(rnd_u64(seed, frame, lane, 2 * idx)), u2 = rnd_unit(rnd_u64(seed, frame, lane, 2 * idx + 1));
  const double r = sqrt(-2.0 * log(u1)), th = 2.0 * 3.14159265358979323846 * u2;
  *a = r * cos(th);
  *b = r * sin(th);
}

uint64_t rub_synth_row_samples(const rub_config *cfg, const rub_synth_params *sp) {
  HostCfg h;
  if (host_cfg_init(h, cfg)) return 0;
  uint64_t n = (uint64_t)(h.T + h.D) * h.L;
  if (sp && sp->include_s0) n += (uint64_t)h.L + 2ull * sp->lead_zeros;
  return n;
}

rub_status rub_synth_frames(const rub_config *cfg, const rub_synth_params *sp, const float *S0s0,
                            const float *S1, const float *s1, uint32_t n_frames, float *iq,
                            uint8_t *tx_data, float *noise_var_out) {
  HostCfg h;
  rub_status st = host_cfg_init(h, cfg);
  if (st) return st;
  if (!sp || !iq) { set_error("synth: NULL argument"); return RUB_ERR_INVALID_ARG; }
  if (sp->n_taps == 0 && !sp->fixed_H) { set_error("synth: n_taps == 0 needs fixed_H"); return RUB_ERR_INVALID_ARG; }
  if (sp->n_taps > h.cp + 1 && sp->n_taps > 1) { set_error("synth: n_taps %u exceeds cp_len+1", sp->n_taps); return RUB_ERR_INVALID_ARG; }
  const uint64_t row = rub_synth_row_samples(cfg, sp);
  const uint32_t N = h.N;
  const float g = sp->baseband_gain;
  // expected received signal power per time sample and the matching noise variance
  double psig;
  if (sp->n_taps) psig = (double)N * g * g;
  else {
    double s = 0;
    for (uint32_t i = 0; i < N * N; i++) s += (double)sp->fixed_H[2 * i] * sp->fixed_H[2 * i] + (double)sp->fixed_H[2 * i + 1] * sp->fixed_H[2 * i + 1];
    psig = s / N * g * g;
  }
  const double sig2_t = psig / pow(10.0, (double)sp->snr_db / 10.0);
  if (noise_var_out) *noise_var_out = (float)(sig2_t * (double)h.M / (double)h.Mo);
  unsigned nthr = sp->n_threads ? sp->n_threads : std::max(1u, std::thread::hardware_concurrency());
  nthr = std::min<unsigned>(nthr, std::max(1u, n_frames));
  std::vector<rub_status> results(nthr, RUB_OK);
  rub_config c2 = h.c;
  c2.sctype = h.sctype.data();
  auto worker = [&](unsigned w) {
    rub_framegen *fg = nullptr;
    rub_status r = rub_framegen_create(&fg, &c2, S0s0, S0s0 ? S0s0 + 2 * (size_t)h.M : nullptr, s1);
    if (r) { results[w] = r; return; }
    if (S1) memcpy(fg->S1.data(), S1, sizeof(cf) * fg->S1.size());
    std::vector<std::vector<cf>> tx(N, std::vector<cf>(row)), syms(N, std::vector<cf>(h.Mo));
    std::vector<cf *> txp(N);
    std::vector<const cf *> inp(N);
    std::vector<cf> taps((size_t)N * N * std::max(1u, sp->n_taps));
    std::vector<float> mod_tab(2u << h.q);
    for (uint32_t s = 0; s < (1u << h.q); s++) rub_modem_modulate(h.q, s, &mod_tab[2 * s]);
    for (uint32_t f = w; f < n_frames; f += nthr) {
      const uint64_t gf = sp->first_frame + f;
      for (uint32_t s = 0; s < N; s++) std::fill(tx[s].begin(), tx[s].end(), mk(0.f, 0.f));
      uint64_t pos = 0;
      if (sp->include_s0) {
        pos = sp->lead_zeros;
        write_cp_symbol(tx[0].data() + pos, fg->s0.data(), h.M, h.cp);
        pos += h.L;
      }
      for (uint32_t s = 0; s < N; s++) txp[s] = tx[s].data() + pos;
      if (h.c.estimator == RUB_EST_LS_COMB_INTERP) rub_framegen_write_comb_words(fg, (float *const *)txp.data());
      else write_access_codes(fg, txp.data(), 0);
      pos += (uint64_t)h.T * h.L;
      for (uint32_t d = 0; d < h.D; d++) {
        for (uint32_t s = 0; s < N; s++) {
          for (uint32_t j = 0; j < h.Mo; j++) {
            const uint64_t idx = ((uint64_t)d * h.Mo + j);
            const uint32_t sym = (uint32_t)(rnd_u64(sp->seed, gf, 1 + s, idx) >> (64 - h.q));  // uniform in [0, 2^q), main.cc:1236
            syms[s][j] = mk(mod_tab[2 * sym], mod_tab[2 * sym + 1]);
            if (tx_data) tx_data[(((size_t)f * N + s) * h.D + d) * h.Mo + j] = (uint8_t)sym;
          }
          inp[s] = syms[s].data();
          txp[s] = tx[s].data() + pos;
        }
        rub_framegen_assemble_mimo_packet(fg, (float *const *)txp.data(), (const float *const *)inp.data());
        pos += h.L;
      }
      // BASEBAND_GAIN (mimo/main.cc:1049, :1093)
      for (uint32_t s = 0; s < N; s++) for (auto &v : tx[s]) v = cscale(v, g);
      // channel
      const uint32_t nt = std::max(1u, sp->n_taps);
      if (sp->n_taps) {
        const double sc = sqrt(0.5 / (double)sp->n_taps);
        for (uint32_t i = 0; i < N * N * nt; i++) {
          double a, b;
          rnd_gauss2(sp->seed, gf, 100, i, &a, &b);
          taps[i] = mk((float)(a * sc), (float)(b * sc));
        }
      } else {
        for (uint32_t i = 0; i < N * N; i++) taps[i] = mk(sp->fixed_H[2 * i], sp->fixed_H[2 * i + 1]);
      }
      const double nsc = sqrt(0.5 * sig2_t);
      for (uint32_t r = 0; r < N; r++) {
        cf *dst = reinterpret_cast<cf *>(iq) + ((size_t)f * N + r) * row;
        for (uint64_t n = 0; n < row; n++) {
          double are = 0, aim = 0;
          for (uint32_t t = 0; t < N; t++) {
            const cf *hp = &taps[((size_t)r * N + t) * nt];
            const cf *xp = tx[t].data();
            const uint32_t lmax = (uint32_t)std::min<uint64_t>(nt - 1, n);
            for (uint32_t l = 0; l <= lmax; l++) {
              const cf xv = xp[n - l], hv = hp[l];
              are += (double)hv.x * xv.x - (double)hv.y * xv.y;
              aim += (double)hv.x * xv.y + (double)hv.y * xv.x;
            }
          }
          double na, nb;
          rnd_gauss2(sp->seed, gf, 200 + r, n, &na, &nb);
          dst[n] = mk((float)(are + nsc * na), (float)(aim + nsc * nb));
        }
      }
    }
    rub_framegen_destroy(fg);
  };
  std::vector<std::thread> th;
  for (unsigned w = 1; w < nthr; w++) th.emplace_back(worker, w);
  worker(0);
  for (auto &t : th) t.join();
  for (auto r : results) if (r) return r;
  return RUB_OK;
}

// ------------------------------------------------------------- file formats -----------
// raw fc32 as written by mimo/main.cc:831-833 / read back at :906-918
rub_status rub_file_read_fc32(const char *path, float *dst, uint64_t max_samples, uint64_t *n_read) {
  FILE *f = fopen(path, "rb");
  if (!f) { set_error("cannot open %s", path); return RUB_ERR_IO; }
  const size_t n = fread(dst, 2 * sizeof(float), max_samples, f);
  fclose(f);
  if (n_read) *n_read = n;
  return RUB_OK;
}
rub_status rub_file_write_fc32(const char *path, const float *src, uint64_t n) {
  FILE *f = fopen(path, "wb");
  if (!f) { set_error("cannot open %s", path); return RUB_ERR_IO; }
  const size_t w = fwrite(src, 2 * sizeof(float), n, f);
  fclose(f);
  return w == n ? RUB_OK : RUB_ERR_IO;
}
rub_status rub_file_write_u32(const char *path, const uint32_t *src, uint64_t n) {
  FILE *f = fopen(path, "wb");
  if (!f) { set_error("cannot open %s", path); return RUB_ERR_IO; }
  const size_t w = fwrite(src, sizeof(uint32_t), n, f);
  fclose(f);
  return w == n ? RUB_OK : RUB_ERR_IO;
}


// ------------------------------------------------------------- configuration front-end
// read_options, mimo/main.cc:174-240
static bool parse_double(const char *v, double *out) {
  if (!v || !*v) return false;
  char *end = nullptr;
  const double d = strtod(v, &end);
  if (end == v || *end) return false;
  *out = d;
  return true;
}
static bool parse_uint(const char *v, uint32_t *out) {
  if (!v || !*v || *v == '-') return false;
  char *end = nullptr;
  const unsigned long u = strtoul(v, &end, 10);
  if (end == v || *end || u > 0xfffffffful) return false;
  *out = (uint32_t)u;
  return true;
}
static void copy_str(char *dst, size_t cap, const char *src) {
  const size_t n = std::min(strlen(src), cap - 1);
  memcpy(dst, src, n);
  dst[n] = 0;
}

rub_status rub_config_from_args(int argc, const char *const *argv, rub_config *cfg, rub_frontend_options *fe) {
  if (!cfg || !fe || (argc > 0 && !argv)) { set_error("config_from_args: NULL argument"); return RUB_ERR_INVALID_ARG; }
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i] ? argv[i] : "";
    std::string val;
    bool has_val = false;
    const size_t eq = a.find('=');
    if (a.rfind("--", 0) == 0 && eq != std::string::npos) { val = a.substr(eq + 1); a = a.substr(0, eq); has_val = true; }
    if (a == "-h" || a == "--help") { fe->help = 1; continue; }
    if (a == "-v" || a == "--verbose") { fe->verbose = 1; continue; }
    if (a == "-q" || a == "--quite") { fe->verbose = 0; continue; }
    static const char *const valued[] = {"-f", "--freq", "-r", "--rate", "--dsp_gain", "--tx_gain", "--rx_gain",
                                         "--num_subcarriers", "--cp_len", "--rx_addr", "--tx_addr", "--tx_subdev",
                                         "--rx_subdev"};
    bool known = false;
    for (const char *k : valued) known = known || a == k;
    if (!known) { set_error("unrecognised option '%s'", a.c_str()); return RUB_ERR_INVALID_ARG; }
    if (!has_val) {
      if (i + 1 >= argc || !argv[i + 1]) { set_error("the required argument for option '%s' is missing", a.c_str()); return RUB_ERR_INVALID_ARG; }
      val = argv[++i];
    }
    bool ok = true;
    double d = 0;
    if (a == "-f" || a == "--freq") { ok = parse_double(val.c_str(), &d); if (ok) fe->cent_freq = d; }
    else if (a == "-r" || a == "--rate") { ok = parse_double(val.c_str(), &d); if (ok) fe->samp_rate = d; }
    else if (a == "--dsp_gain") { ok = parse_double(val.c_str(), &d); if (ok) fe->dsp_gain = (float)d; }
    else if (a == "--tx_gain") { ok = parse_double(val.c_str(), &d); if (ok) fe->txgain = d; }
    else if (a == "--rx_gain") { ok = parse_double(val.c_str(), &d); if (ok) fe->rxgain = d; }
    else if (a == "--num_subcarriers") ok = parse_uint(val.c_str(), &cfg->M);
    else if (a == "--cp_len") ok = parse_uint(val.c_str(), &cfg->cp_len);
    else if (a == "--rx_addr") copy_str(fe->rx_addr, sizeof(fe->rx_addr), val.c_str());
    else if (a == "--tx_addr") copy_str(fe->tx_addr, sizeof(fe->tx_addr), val.c_str());
    else if (a == "--tx_subdev") copy_str(fe->tx_subdev, sizeof(fe->tx_subdev), val.c_str());
    else if (a == "--rx_subdev") copy_str(fe->rx_subdev, sizeof(fe->rx_subdev), val.c_str());
    if (!ok) { set_error("the argument ('%s') for option '%s' is invalid", val.c_str(), a.c_str()); return RUB_ERR_INVALID_ARG; }
  }
  return RUB_OK;
}

// value text of "key" in a flat JSON object: number / string token without the quotes
static bool json_value(const char *json, const char *key, std::string *out) {
  const std::string pat = std::string("\"") + key + "\"";
  const char *p = strstr(json, pat.c_str());
  if (!p) return false;
  p += pat.size();
  while (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r') p++;
  if (*p != ':') return false;
  p++;
  while (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r') p++;
  out->clear();
  if (*p == '"') {
    for (p++; *p && *p != '"'; p++) { if (*p == '\\' && p[1]) p++; out->push_back(*p); }
    return *p == '"';
  }
  while (*p && *p != ',' && *p != '}' && *p != ' ' && *p != '\n' && *p != '\r' && *p != '\t') out->push_back(*p++);
  return !out->empty();
}

// Interface/usrp_device.cpp:13-29
rub_status rub_config_from_json(const char *json, rub_config *cfg, rub_frontend_options *fe) {
  if (!json || !cfg || !fe) { set_error("config_from_json: NULL argument"); return RUB_ERR_INVALID_ARG; }
  std::string v;
  double d;
  auto num = [&](const char *key, double *dst) -> bool {
    if (!json_value(json, key, &v)) return true;  // absent: unchanged
    if (!parse_double(v.c_str(), dst)) { set_error("JSON key \"%s\": '%s' is not a number", key, v.c_str()); return false; }
    return true;
  };
  double M = cfg->M, nn = fe->num_nullcarriers, cp = cfg->cp_len, ts = cfg->num_access_codes;
  if (!num("Number of Subcarriers", &M) || !num("Number of Nullcarriers", &nn) || !num("Prefix Length", &cp) ||
      !num("Training Sequences", &ts) || !num("TX Gain", &fe->txgain) || !num("RX Gain", &fe->rxgain) ||
      !num("Center Freq.", &fe->cent_freq) || !num("Samp. Rate", &fe->samp_rate))
    return RUB_ERR_INVALID_ARG;
  for (double x : {M, nn, cp, ts})
    if (x < 0 || x > 4294967295.0 || x != (double)(uint32_t)x) { set_error("JSON record: %g is not an unsigned integer", x); return RUB_ERR_INVALID_ARG; }
  cfg->M = (uint32_t)M; fe->num_nullcarriers = (uint32_t)nn; cfg->cp_len = (uint32_t)cp; cfg->num_access_codes = (uint32_t)ts;
  (void)d;
  if (json_value(json, "Adress", &v)) { copy_str(fe->rx_addr, sizeof(fe->rx_addr), v.c_str()); copy_str(fe->tx_addr, sizeof(fe->tx_addr), v.c_str()); }
  if (json_value(json, "Subdevice Specifications", &v)) { copy_str(fe->rx_subdev, sizeof(fe->rx_subdev), v.c_str()); copy_str(fe->tx_subdev, sizeof(fe->tx_subdev), v.c_str()); }
  return RUB_OK;
}

}  // extern "C"
