// tests/framing_driver.cc — a caller written against the framing.h class API, the way mimo/main.cc
// uses it (:1262-1300, the rx loop :1003-1013 and the callback :1384-1421), behind a C interface
// that Python loads with ctypes.  The SAME source is compiled twice:
//   * against the REFERENCE's own mimo/framing.h + framing.cc (oracle/Makefile `ref`, with the
//     stand-in headers of oracle/shim/) -> oracle/_ref/libref_framing.so, which produced
//     tests/golden/ref_*.npz (oracle/make_ref_fixtures.py);
//   * against the product's facade include/rub_mimo/framing.h + librubmimo_b200.so
//     (tests/test_ref_fixtures.py), which must give the same outputs: the drop-in claim.
// Test infrastructure only.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "framing.h"  // the reference's (-I/root/reference/mimo) or the facade (-Iinclude/rub_mimo)

namespace {
struct Sink { float *eq; unsigned max_syms, syms, Mo, N; } g_sink;
void *on_symbols(std::vector<gr_complex *> s, unsigned int occupied) {
  if (g_sink.syms < g_sink.max_syms)
    for (unsigned n = 0; n < g_sink.N; n++)
      memcpy(g_sink.eq + 2 * (((size_t)n * g_sink.max_syms + g_sink.syms) * g_sink.Mo), s[n], sizeof(gr_complex) * occupied);
  g_sink.syms++;
  return nullptr;
}
msequence make_ms(int which) {  // mimo/main.cc:1265-1267
  switch (which) {
    case -1: return msequence_create(LFSR_SMALL_LENGTH, LFSR_SMALL_0_GEN_POLY, 1);
    case 0: return msequence_create(LFSR_LARGE_LENGTH, LFSR_LARGE_0_GEN_POLY, 1);
    default: return msequence_create(LFSR_LARGE_LENGTH, LFSR_LARGE_1_GEN_POLY, 1);
  }
}
}  // namespace

extern "C" {

struct ref_sync_result {
  int32_t state;
  uint64_t sync_index, num_samples_processed;
  uint64_t plateau_start[2], plateau_end[2];
  uint32_t symbols;
};

// invert() (mimo/framing.cc:1344-1367) on n 2x2 matrices, row-major complex64 in and out
void ref_invert(const float *G, float *W, float *gain, unsigned n) {
  for (unsigned i = 0; i < n; i++) {
    std::vector<std::vector<gr_complex> > g(2, std::vector<gr_complex>(2)), w(2, std::vector<gr_complex>(2));
    for (unsigned r = 0; r < 2; r++)
      for (unsigned c = 0; c < 2; c++) g[r][c] = gr_complex(G[2 * (4 * i + 2 * r + c)], G[2 * (4 * i + 2 * r + c) + 1]);
    gain[i] = invert(w, g);
    for (unsigned r = 0; r < 2; r++)
      for (unsigned c = 0; c < 2; c++) { W[2 * (4 * i + 2 * r + c)] = w[r][c].real(); W[2 * (4 * i + 2 * r + c) + 1] = w[r][c].imag(); }
  }
}

// the reference's default allocation (mimo/framing.cc:949-1008) and its count
void ref_default_sctype(unsigned M, unsigned char *p, unsigned *n_null, unsigned *n_pilot, unsigned *n_data) {
  ofdmframe_init_default_sctype(p, M);
  ofdmframe_validate_sctype(p, M, n_null, n_pilot, n_data);
}

// framegen as main.cc uses it: tx[s] = write_sync_words (S0 + access codes) followed by D packets
unsigned ref_framegen(unsigned M, unsigned cp, unsigned nac, const unsigned char *p, const float *symbols /* [D][2][Mo] */,
                      unsigned D, unsigned Mo, float *tx /* [2][(nac*2+1+D)*(M+cp)] */) {
  const unsigned N = 2, L = M + cp;
  msequence ms0 = make_ms(-1);
  std::vector<msequence> ms1 = {make_ms(0), make_ms(1)};
  unsigned char *pp = const_cast<unsigned char *>(p);
  rx_beamforming::framegen fg(M, cp, N, nac, pp, ms0, ms1);
  const size_t row = (size_t)(nac * N + 1 + D) * L;
  std::vector<gr_complex *> out = {reinterpret_cast<gr_complex *>(tx), reinterpret_cast<gr_complex *>(tx) + row};
  unsigned n = fg.write_sync_words(out);
  for (unsigned d = 0; d < D; d++) {
    std::vector<gr_complex *> dst = {out[0] + n, out[1] + n};
    std::vector<gr_complex *> in = {
        const_cast<gr_complex *>(reinterpret_cast<const gr_complex *>(symbols)) + ((size_t)d * N + 0) * Mo,
        const_cast<gr_complex *>(reinterpret_cast<const gr_complex *>(symbols)) + ((size_t)d * N + 1) * Mo};
    n += fg.assemble_mimo_packet(dst, in);
  }
  return n;
}

// framesync fed a capture in chunks, as the rx worker does
int ref_framesync(unsigned M, unsigned cp, unsigned nac, const unsigned char *p, const float *cap0, const float *cap1,
                  uint64_t num_samples, unsigned chunk, ref_sync_result *res, float *G_out /* [M][2][2] */,
                  float *eq_out /* [2][max_syms][Mo] */, unsigned max_syms, unsigned Mo) {
  const unsigned N = 2;
  msequence ms0 = make_ms(-1);
  std::vector<msequence> ms1 = {make_ms(0), make_ms(1)};
  g_sink = Sink{eq_out, max_syms, 0, Mo, N};
  unsigned char *pp = const_cast<unsigned char *>(p);
  rx_beamforming::framesync fs(M, cp, N, nac, pp, ms0, ms1, on_symbols);
  framesync_states_t st = STATE_SEEK_PLATEAU;
  const gr_complex *c[2] = {reinterpret_cast<const gr_complex *>(cap0), reinterpret_cast<const gr_complex *>(cap1)};
  for (uint64_t off = 0; off < num_samples && st != STATE_MIMO; off += chunk) {
    const unsigned n = (unsigned)((num_samples - off < chunk) ? num_samples - off : chunk);
    std::vector<gr_complex *> in = {const_cast<gr_complex *>(c[0] + off), const_cast<gr_complex *>(c[1] + off)};
    st = fs.execute(in, n);
  }
  res->state = (int32_t)st;
  res->sync_index = fs.get_sync_index();
  res->num_samples_processed = fs.get_num_samples_processed();
  for (unsigned s = 0; s < N; s++) { res->plateau_start[s] = fs.get_plateau_start(s); res->plateau_end[s] = fs.get_plateau_end(s); }
  res->symbols = g_sink.syms;
  std::vector<std::vector<std::vector<gr_complex> > > G = fs.get_G();
  for (unsigned k = 0; k < M; k++)
    for (unsigned r = 0; r < N; r++)
      for (unsigned t = 0; t < N; t++) {
        G_out[2 * ((k * N + r) * N + t)] = G[k][r][t].real();
        G_out[2 * ((k * N + r) * N + t) + 1] = G[k][r][t].imag();
      }
  return 0;
}

}  // extern "C"
