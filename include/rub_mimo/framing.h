/*
 * framing.h — framing.h-compatible C++ facade over the C ABI of librubmimo_b200.so.
 *
 * Drop-in for /root/reference/mimo/framing.h: the same class names, constructor signatures
 * and entry points that mimo/main.cc uses (rx_beamforming::framegen :42-103,
 * rx_beamforming::framesync :105-213, mimo_callback :30-31, ofdmframe_* :226-270,
 * invert :278-280), with small stand-ins for the liquid-dsp types the reference header pulls
 * in (msequence, modem, OFDMFRAME_SCTYPE_*) and gr_complex = std::complex<float>
 * (gnuradio/gr_complex.h).  FFTW / VOLK / liquid / Boost / UHD are not needed.
 *
 * What runs where
 *   - framegen and the ofdmframe_* helpers: host code inside the library (rub_framegen_*).
 *   - framesync::execute: the sample-serial state machine stays on the host, but its two hot
 *     loops run on the GPU: the Schmidl & Cox metric (framing.cc:626-637, rub_rx_sc_metric,
 *     bit-exact) and the access-code timing search (framing.cc:702-744, rub_rx_timing_search,
 *     the time-domain form of the same correlation, sum_k X[k] conj(S[k]) = sqrt(M) sum_n x[n]
 *     conj(s1[n])).  Everything after it — CP strip, FFT, LS estimate, invert, W*y, gain
 *     (framing.cc:535-589, :801-832) — is one rub_rx_process_batch_host call on the GPU with
 *     the reference's quirks Q1 (identity-initialised G), Q2 (per-link timing index) and Q4
 *     (payload start from rx stream 1) switched on, after which mimo_callback fires once per
 *     OFDM symbol exactly as framing.cc:587 does.
 *   - Errors: the reference asserts / exit(1)s; the facade throws std::runtime_error carrying
 *     rub_last_error().
 */
#ifndef RUB_MIMO_FRAMING_H
#define RUB_MIMO_FRAMING_H

#include <algorithm>
#include <complex>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "config.h"
#include "rub_mimo.h"

typedef std::complex<float> gr_complex;

// ---- liquid-dsp stand-ins ------------------------------------------------------------
#define OFDMFRAME_SCTYPE_NULL RUB_SCTYPE_NULL
#define OFDMFRAME_SCTYPE_PILOT RUB_SCTYPE_PILOT
#define OFDMFRAME_SCTYPE_DATA RUB_SCTYPE_DATA

typedef rub_msequence *msequence;
inline msequence msequence_create(unsigned int m, unsigned int g, unsigned int a) {
  msequence ms = new rub_msequence;
  rub_msequence_init(ms, m, g, a);
  return ms;
}
inline void msequence_destroy(msequence ms) { delete ms; }
inline void msequence_reset(msequence ms) { rub_msequence_reset(ms); }
inline unsigned int msequence_advance(msequence ms) { return rub_msequence_advance(ms); }
inline unsigned int msequence_generate_symbol(msequence ms, unsigned int bps) {
  return rub_msequence_generate_symbol(ms, bps);
}

// square Gray QAM modems of liquid-dsp (LIQUID_MODEM_QAM4/16/64/256); value = bits per symbol
typedef enum { LIQUID_MODEM_QAM4 = 2, LIQUID_MODEM_QAM16 = 4, LIQUID_MODEM_QAM64 = 6, LIQUID_MODEM_QAM256 = 8 } modulation_scheme;
struct modem_s { unsigned int bps; };
typedef modem_s *modem;
inline modem modem_create(modulation_scheme s) { return new modem_s{(unsigned int)s}; }
inline void modem_destroy(modem m) { delete m; }
inline void modem_modulate(modem m, unsigned int sym, gr_complex *y) {
  float out[2];
  if (rub_modem_modulate(m->bps, sym, out)) throw std::runtime_error(rub_last_error());
  *y = gr_complex(out[0], out[1]);
}
inline void modem_demodulate(modem m, gr_complex x, unsigned int *sym) {
  const float in[2] = {x.real(), x.imag()};
  if (rub_modem_demodulate(m->bps, in, sym)) throw std::runtime_error(rub_last_error());
}

// ---- callback and receiver states (mimo/framing.h:30-39) ------------------------------
typedef void *(*mimo_callback)(std::vector<gr_complex *>, unsigned int occupied_carriers);

typedef enum { STATE_SEEK_PLATEAU = 0, STATE_SAVE_ACCESS_CODES, STATE_WAIT, STATE_MIMO } framesync_states_t;

namespace rub_detail {
inline void check(rub_status s) {
  if (s != RUB_OK) throw std::runtime_error(std::string(rub_strerror(s)) + ": " + rub_last_error());
}
inline rub_config make_config(unsigned M, unsigned cp, unsigned N, unsigned nac, unsigned D, unsigned bps,
                              const unsigned char *p) {
  rub_config c;
  std::memset(&c, 0, sizeof(c));
  c.struct_size = sizeof(c);
  c.M = M; c.cp_len = cp; c.num_streams = N; c.num_access_codes = nac; c.num_data_symbols = D;
  c.modulation = bps; c.detector = RUB_DET_ZF; c.estimator = RUB_EST_LS_FULLBAND;
  c.flags = RUB_FLAG_Q1_IDENTITY_INIT;  // framing.cc:302-319
  c.sctype = p;
  return c;
}
}  // namespace rub_detail

// ---- free functions (mimo/framing.h:226-280) ------------------------------------------
inline void ofdmframe_init_default_sctype(unsigned char *p, unsigned int M) {
  rub_ofdmframe_init_default_sctype(p, M, USE_ALL_CARRIERS, ADD_NULL_CARRIERS);
}
inline void ofdmframe_validate_sctype(const unsigned char *p, unsigned int M, unsigned int *M_null,
                                      unsigned int *M_pilot, unsigned int *M_data) {
  rub_detail::check(rub_ofdmframe_validate_sctype(p, M, M_null, M_pilot, M_data));
}
inline void ofdmframe_print_sctype(const unsigned char *p, unsigned int M) {  // framing.cc:1032-1051
  std::printf("[");
  for (unsigned int i = 0; i < M; i++) {
    const unsigned int k = (i + M / 2) % M;
    std::printf("%c", p[k] == OFDMFRAME_SCTYPE_NULL ? '.' : p[k] == OFDMFRAME_SCTYPE_PILOT ? '|' : '+');
  }
  std::printf("]\n");
}
inline void ofdmframe_init_S0(const unsigned char *p, unsigned int M, std::complex<float> *S0,
                              std::complex<float> *s0, msequence ms) {
  rub_detail::check(rub_ofdmframe_init_S0(p, M, reinterpret_cast<float *>(S0), reinterpret_cast<float *>(s0), ms));
}
inline void ofdmframe_init_S1(const unsigned char *p, unsigned int M, unsigned int num_access_codes,
                              std::complex<float> *S1, std::complex<float> *s1, msequence ms) {
  rub_detail::check(rub_ofdmframe_init_S1(p, M, num_access_codes, reinterpret_cast<float *>(S1),
                                          reinterpret_cast<float *>(s1), ms));
}
// currently only for 2 X 2 matrix (mimo/framing.h:278)
inline float invert(std::vector<std::vector<gr_complex> > &W, std::vector<std::vector<gr_complex> > const &G) {
  if (G.size() != 2 || W.size() != 2 || G[0].size() != 2 || G[1].size() != 2)
    throw std::runtime_error("invert: only 2x2 (mimo/framing.cc:1346-1351)");
  gr_complex g[4] = {G[0][0], G[0][1], G[1][0], G[1][1]}, w[4];
  float gain = 0.f;
  rub_detail::check(rub_invert_2x2(reinterpret_cast<float *>(w), reinterpret_cast<const float *>(g), &gain));
  W[0].resize(2); W[1].resize(2);
  W[0][0] = w[0]; W[0][1] = w[1]; W[1][0] = w[2]; W[1][1] = w[3];
  return gain;
}

namespace rx_beamforming {

// ---------------------------------------------------------------- framegen -------------
class framegen {
 private:
  unsigned int M, cp_len, symbol_len, num_streams, num_access_codes;
  unsigned int M_null, M_pilot, M_data;
  std::vector<unsigned char> p;
  rub_framegen *fg;

 public:
  framegen(unsigned int _M, unsigned int _cp_len, unsigned int _num_streams, unsigned int _num_access_codes,
           unsigned char *const &_p, msequence const &_ms_S0, std::vector<msequence> const &_ms_S1)
      : M(_M), cp_len(_cp_len), symbol_len(_M + _cp_len), num_streams(_num_streams),
        num_access_codes(_num_access_codes), p(_p, _p + _M), fg(nullptr) {
    ofdmframe_validate_sctype(p.data(), M, &M_null, &M_pilot, &M_data);  // framing.cc:104-108
    // the generators are borrowed and advanced, as in framing.cc:110-146
    std::vector<gr_complex> S0(M), s0(M), S1((size_t)num_streams * num_access_codes * M), s1(S1.size());
    ofdmframe_init_S0(p.data(), M, S0.data(), s0.data(), _ms_S0);
    for (unsigned int i = 0; i < num_streams; i++)
      ofdmframe_init_S1(p.data(), M, num_access_codes, S1.data() + (size_t)i * num_access_codes * M,
                        s1.data() + (size_t)i * num_access_codes * M, _ms_S1[i]);
    rub_config c = rub_detail::make_config(M, cp_len, num_streams, num_access_codes, 1, 2, p.data());
    rub_detail::check(rub_framegen_create(&fg, &c, reinterpret_cast<float *>(S0.data()),
                                          reinterpret_cast<float *>(s0.data()), reinterpret_cast<float *>(s1.data())));
  }
  ~framegen() { rub_framegen_destroy(fg); }
  framegen(const framegen &) = delete;
  framegen &operator=(const framegen &) = delete;
  void print() {
    std::printf("ofdmframegen:\n    num subcarriers     :   %-u\n      - NULL            :   %-u\n"
                "      - pilot           :   %-u\n      - data            :   %-u\n    cyclic prefix len   :   %-u\n    ",
                M, M_null, M_pilot, M_data, cp_len);
    ofdmframe_print_sctype(p.data(), M);
  }
  unsigned int write_sync_words(std::vector<std::complex<float> *> tx_buff) {
    if (tx_buff.size() != num_streams) throw std::runtime_error("write_sync_words: tx_buff.size() != num_streams");
    return rub_framegen_write_sync_words(fg, reinterpret_cast<float *const *>(tx_buff.data()));
  }
  unsigned int assemble_mimo_packet(std::vector<gr_complex *> tx_buff, std::vector<gr_complex *> in_buff) {
    if (tx_buff.size() != num_streams || in_buff.size() != num_streams)
      throw std::runtime_error("assemble_mimo_packet: buffer count != num_streams");
    return rub_framegen_assemble_mimo_packet(fg, reinterpret_cast<float *const *>(tx_buff.data()),
                                             reinterpret_cast<const float *const *>(in_buff.data()));
  }
  unsigned int get_num_streams() { return num_streams; }
};

// ---------------------------------------------------------------- framesync ------------
class framesync {
 private:
  unsigned int M, M2, cp_len, symbol_len, num_streams, num_access_codes;
  unsigned int M_null, M_pilot, M_data, M_occupied;
  std::vector<unsigned char> p;
  std::vector<gr_complex> S0, s0, S1, s1;  // S1/s1 [stream][code][k]
  std::vector<std::vector<std::vector<gr_complex> > > G, W;
  std::vector<float> normalize_gain;
  mimo_callback callback;
  unsigned int siso_tx, siso_rx;
  unsigned int pid_max;  // PID_MAX
  // sync
  unsigned long int sync_index;
  unsigned long long int num_samples_processed;
  framesync_states_t state;
  unsigned int access_code_buffer_len, tx_sig_len;
  std::vector<unsigned long int> plateau_start, plateau_end;
  std::vector<bool> in_plateau;
  // windowcf / wdelay stand-in: the most recent samples of every stream.  Only the last Wlen (estimate_channel)
  // or M + M/2 (metric look-back) are ever read, so execute() trims it to that bound (amortised O(1))
  std::vector<std::vector<gr_complex> > history;
  rub_rx *rx;       // decode handle (created once the number of buffered payload symbols is known)
  rub_rx *rx_sync;  // handle used for the synchronisation kernels
  std::vector<int32_t> corr_indices;  // [rx][ac_id]
  std::vector<std::vector<float> > sc_metric;  // S&C metric of the samples of the current execute() call
  size_t sc_metric_base;                       // index in the call of sc_metric[s][0]

  rub_rx *sync_handle() {
    if (!rx_sync) {
      rub_config c = rub_detail::make_config(M, cp_len, num_streams, num_access_codes, 1, 2, p.data());
      rub_detail::check(rub_rx_create(&rx_sync, &c, reinterpret_cast<const float *>(S1.data()), -1, nullptr));
    }
    return rx_sync;
  }
  // |P|^2 / R^2 (framing.cc:626-637) for samples [from, num_samples) of this call, every stream, on the
  // GPU (bit-exact with the O(M)-per-sample FIR dot products); the history supplies the look-back
  void compute_sc_metric(std::vector<gr_complex *> const &in_buff, size_t from, size_t num_samples) {
    sc_metric.assign(num_streams, std::vector<float>());
    sc_metric_base = from;
    for (unsigned int s = 0; s < num_streams; s++) {
      const std::vector<gr_complex> &h = history[s];
      const size_t look = std::min(h.size(), (size_t)(M + M2));
      std::vector<gr_complex> tmp(look + (num_samples - from));
      std::copy(h.end() - (long)look, h.end(), tmp.begin());
      std::copy(in_buff[s] + from, in_buff[s] + num_samples, tmp.begin() + (long)look);
      std::vector<float> y(tmp.size());
      rub_detail::check(rub_rx_sc_metric(sync_handle(), reinterpret_cast<const float *>(tmp.data()), tmp.size(), y.data()));
      sc_metric[s].assign(y.begin() + (long)look, y.end());
    }
  }
  void execute_sc_sync(const gr_complex *x, size_t i) {  // framing.cc:591-624
    bool proceed = true;
    for (unsigned int s = 0; s < num_streams; s++) {
      history[s].push_back(x[s]);
      const float y = sc_metric[s][i - sc_metric_base];
      if (y > PLATEAU_THREASHOLD) {
        if (in_plateau[s]) plateau_end[s] = num_samples_processed;
        else { in_plateau[s] = true; plateau_start[s] = num_samples_processed; plateau_end[s] = num_samples_processed; }
      } else in_plateau[s] = false;
      proceed = proceed && (plateau_end[s] - plateau_start[s] > cp_len) && in_plateau[s];
    }
    if (proceed) {
      for (unsigned int s = 0; s < num_streams; s++) sync_index += plateau_start[s];  // quirk Q6: not zeroed
      sync_index /= num_streams;
      std::printf("***** proceed to save access codes ***** \n");
      state = STATE_SAVE_ACCESS_CODES;
    }
  }
  void execute_save_access_codes(const gr_complex *x) {  // framing.cc:639-651
    if (num_samples_processed - sync_index < tx_sig_len + access_code_buffer_len - symbol_len) {
      for (unsigned int s = 0; s < num_streams; s++) history[s].push_back(x[s]);
      return;
    }
    std::printf("**** access codes saved *****\n");
    estimate_channel();
    state = STATE_MIMO;
  }

 public:
  framesync(unsigned int _M, unsigned int _cp_len, unsigned int _num_streams, unsigned int _num_access_codes,
            unsigned char *const &_p, msequence const &_ms_S0, std::vector<msequence> const &_ms_S1,
            mimo_callback _callback)
      : M(_M), M2(_M / 2), cp_len(_cp_len), symbol_len(_M + _cp_len), num_streams(_num_streams),
        num_access_codes(_num_access_codes), p(_p, _p + _M), callback(_callback), siso_tx(0), siso_rx(0),
        pid_max(PID_MAX), sync_index(0), num_samples_processed(0), state(STATE_SEEK_PLATEAU), rx(nullptr),
        rx_sync(nullptr), sc_metric_base(0) {
    ofdmframe_validate_sctype(p.data(), M, &M_null, &M_pilot, &M_data);
    M_occupied = M_data + M_pilot;
    S0.resize(M); s0.resize(M);
    S1.resize((size_t)num_streams * num_access_codes * M); s1.resize(S1.size());
    for (unsigned int i = 0; i < num_streams; i++)
      ofdmframe_init_S1(p.data(), M, num_access_codes, S1.data() + (size_t)i * num_access_codes * M,
                        s1.data() + (size_t)i * num_access_codes * M, _ms_S1[i]);
    ofdmframe_init_S0(p.data(), M, S0.data(), s0.data(), _ms_S0);
    // G and W start as identity on non-null carriers (framing.cc:302-319)
    G.assign(M, std::vector<std::vector<gr_complex> >(num_streams, std::vector<gr_complex>(num_streams)));
    for (unsigned int k = 0; k < M; k++)
      for (unsigned int r = 0; r < num_streams; r++)
        G[k][r][r] = (p[k] != OFDMFRAME_SCTYPE_NULL) ? gr_complex(1.0f, 0.0f) : gr_complex(0.0f, 0.0f);
    W = G;
    normalize_gain.assign(M_occupied, 1.0f);
    plateau_start.assign(num_streams, 0); plateau_end.assign(num_streams, 0); in_plateau.assign(num_streams, false);
    history.resize(num_streams);
    set_num_data_symbols(pid_max);
  }
  ~framesync() { rub_rx_destroy(rx); rub_rx_destroy(rx_sync); }
  framesync(const framesync &) = delete;
  framesync &operator=(const framesync &) = delete;

  // PID_MAX is a compile-time macro upstream (config.h:92); here it can also be set at run time
  void set_num_data_symbols(unsigned int d) {
    pid_max = d;
    access_code_buffer_len = symbol_len * (num_access_codes * num_streams + 4);  // framing.cc:284
    tx_sig_len = pid_max * symbol_len;                                            // framing.cc:285
    rub_rx_destroy(rx);
    rx = nullptr;
  }
  void print() {
    std::printf("ofdmframegen:\n    num subcarriers     :   %-u\n      - NULL            :   %-u\n"
                "      - pilot           :   %-u\n      - data            :   %-u\n    cyclic prefix len   :   %-u\n    ",
                M, M_null, M_pilot, M_data, cp_len);
    ofdmframe_print_sctype(p.data(), M);
  }
  unsigned long int get_sync_index() { return sync_index; }
  std::vector<std::vector<std::vector<gr_complex> > > get_G() { return G; }  // by value, as upstream
  unsigned long long int get_num_samples_processed() { return num_samples_processed; }
  unsigned long int get_plateau_start(unsigned int s) { return plateau_start.at(s); }
  unsigned long int get_plateau_end(unsigned int s) { return plateau_end.at(s); }
  void reset() { state = STATE_SEEK_PLATEAU; }  // framing.cc:461-464
  void compute_receive_beamformer() {}          // empty upstream, framing.cc:898-900
  void set_siso_tx(unsigned int t) { siso_tx = t; }
  void set_siso_rx(unsigned int r) { siso_rx = r; }

  // framing.cc:471-506
  framesync_states_t execute(std::vector<gr_complex *> const &in_buff, unsigned int num_samples) {
    std::vector<gr_complex> x(num_streams);
    bool break_loop = false, have_metric = false;
    for (unsigned int i = 0; i < num_samples; i++) {
      for (unsigned int s = 0; s < num_streams; s++) x[s] = in_buff[s][i];
      switch (state) {
        case STATE_SEEK_PLATEAU:
          if (!have_metric) { compute_sc_metric(in_buff, i, num_samples); have_metric = true; }
          execute_sc_sync(x.data(), i);
          break;
        case STATE_SAVE_ACCESS_CODES: execute_save_access_codes(x.data()); break;
        case STATE_MIMO: break_loop = true; break;
        default: throw std::runtime_error("framesync: state not handled");
      }
      num_samples_processed++;
      if (break_loop) break;
    }
    trim_history();
    return state;
  }
  void trim_history() {
    const size_t keep = std::max((size_t)access_code_buffer_len + tx_sig_len, (size_t)(M + M2));
    for (unsigned int s = 0; s < num_streams; s++)
      if (history[s].size() > 2 * keep) history[s].erase(history[s].begin(), history[s].end() - (long)keep);
  }

  // framing.cc:653-886
  void estimate_channel() {
    const unsigned int max_ac_id = num_streams * num_access_codes;
    const size_t Wlen = (size_t)access_code_buffer_len + tx_sig_len;
    // windowcf_read: the Wlen most recent samples, oldest first, zero-filled at the front
    std::vector<std::vector<gr_complex> > buf(num_streams, std::vector<gr_complex>(Wlen, gr_complex(0, 0)));
    for (unsigned int s = 0; s < num_streams; s++) {
      const std::vector<gr_complex> &h = history[s];
      if (h.size() >= Wlen) std::copy(h.end() - (long)Wlen, h.end(), buf[s].begin());
      else std::copy(h.begin(), h.end(), buf[s].begin() + (long)(Wlen - h.size()));
    }
    // timing search (framing.cc:702-744) on the GPU: argmax over i in [0, symbol_len) of the
    // access-code correlation at i + symbol_len*(ac_id+1)
    std::vector<gr_complex> iq((size_t)num_streams * Wlen);
    for (unsigned int s = 0; s < num_streams; s++) std::copy(buf[s].begin(), buf[s].end(), iq.begin() + (size_t)s * Wlen);
    corr_indices.assign((size_t)num_streams * max_ac_id, 0);
    rub_detail::check(rub_rx_timing_search(sync_handle(), reinterpret_cast<const float *>(iq.data()), Wlen,
                                           corr_indices.data(), nullptr));
    // LS + invert + decode on the GPU: one frame, per-link windows (Q2), payload start from rx
    // stream 1's last access code (Q4, framing.cc:857), identity-initialised G (Q1)
    const int32_t payload_start = corr_indices[(size_t)(num_streams > 1 ? 1 : 0) * max_ac_id + max_ac_id - 1] + (int32_t)M;
    // every whole symbol between the payload start and the end of the window is decoded and handed
    // to the callback, as upstream does (framing.cc:856-868): a couple more than PID_MAX, which
    // main.cc's callback ignores (quirk Q14, main.cc:105-108)
    const unsigned int D = (unsigned int)((Wlen - (size_t)payload_start) / symbol_len);
    if (D == 0) return;
    rub_config c = rub_detail::make_config(M, cp_len, num_streams, num_access_codes, D, 2, p.data());
    rub_rx_destroy(rx);
    rx = nullptr;
    rub_detail::check(rub_rx_create(&rx, &c, reinterpret_cast<const float *>(S1.data()), -1, nullptr));
    std::vector<gr_complex> eq((size_t)num_streams * D * M_occupied), Gd((size_t)num_streams * num_streams * M);
    rub_rx_io io;
    std::memset(&io, 0, sizeof(io));
    io.iq = reinterpret_cast<const float *>(iq.data());
    io.layout.frame_stride = (uint64_t)num_streams * Wlen;
    io.layout.rx_stride = Wlen;
    io.timing = corr_indices.data();
    io.payload_start = &payload_start;
    io.eq = reinterpret_cast<float *>(eq.data());
    io.G = reinterpret_cast<float *>(Gd.data());
    io.out_mask = RUB_OUT_EQ | RUB_OUT_G;
    rub_detail::check(rub_rx_process_batch_host(rx, &io, 1));
    // expose G / W / normalize_gain in the reference's layouts (framing.h:137-139)
    unsigned int j = 0;
    for (unsigned int k = 0; k < M; k++) {
      for (unsigned int r = 0; r < num_streams; r++)
        for (unsigned int t = 0; t < num_streams; t++) G[k][r][t] = Gd[((size_t)r * num_streams + t) * M + k];
      if (p[k] != OFDMFRAME_SCTYPE_NULL) {
        if (num_streams == 2) normalize_gain[j] = invert(W[k], G[k]);  // framing.cc:826-832
        j++;
      }
    }
    // callback per OFDM symbol with pointers valid only during the call (framing.cc:587)
    std::vector<gr_complex *> X(num_streams);
    for (unsigned int d = 0; d < D; d++) {
      for (unsigned int s = 0; s < num_streams; s++) X[s] = eq.data() + ((size_t)s * D + d) * M_occupied;
      callback(X, M_occupied);
    }
  }
};

}  // namespace rx_beamforming

namespace tx_beamforming {}

#endif  // RUB_MIMO_FRAMING_H
