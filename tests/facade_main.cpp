// facade_main.cpp — the reference's main.cc flow (mimo/main.cc:1157-1469) on top of the
// framing.h-compatible facade, with the USRP replaced by fc32 files:
//   facade_main tx <dir>   build sctype / msequences / framegen, random data -> modem -> frame,
//                          fixed 2x2 channel + noise, write <dir>/rx%d.dat, tx_data%d.dat (no GPU)
//   facade_main rx <dir>   re-read the capture (main.cc:906-918), framesync::execute, demodulate,
//                          write rx_sig%d.dat / rx_data%d.dat, print the report (needs the GPU)
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>

#define NUM_SUBCARRIERS 64
#define CP_LENGTH 16
#define PID_MAX 200
#define NUM_ACCESS_CODES 20
#define MODEM_BITS_PER_SYMBOL 2
#include "rub_mimo/framing.h"

static std::vector<gr_complex *> rx_sig;
static unsigned int num_frames_detected = 0, num_valid_packets_received = 0;
static unsigned long received_sample_counter = 0;

// mimo/main.cc:104-133
void *callback(std::vector<gr_complex *> x, unsigned int occupied_carriers) {
  num_frames_detected++;
  if (num_valid_packets_received == PID_MAX) return NULL;
  num_valid_packets_received++;
  for (unsigned int stream = 0; stream < NUM_STREAMS; stream++)
    memmove(rx_sig[stream] + received_sample_counter, x[stream], sizeof(gr_complex) * occupied_carriers);
  received_sample_counter += occupied_carriers;
  return NULL;
}

static uint64_t lcg = 88172645463325252ull;
static double urand() { lcg ^= lcg << 13; lcg ^= lcg >> 7; lcg ^= lcg << 17; return ((lcg >> 11) + 0.5) / 9007199254740992.0; }

int main(int argc, char **argv) {
  if (argc < 3) { fprintf(stderr, "usage: %s tx|rx <dir>\n", argv[0]); return 2; }
  const std::string mode = argv[1], dir = argv[2];
  const unsigned int M = NUM_SUBCARRIERS, cp_len = CP_LENGTH, num_streams = NUM_STREAMS, L = M + cp_len;
  std::vector<unsigned char> p(M);
  unsigned int n_null, n_pilot, n_data;
  ofdmframe_init_default_sctype(p.data(), M);
  ofdmframe_validate_sctype(p.data(), M, &n_null, &n_pilot, &n_data);
  const unsigned int Mo = n_pilot + n_data, tx_sig_len = Mo * PID_MAX;
  msequence ms_S0 = msequence_create(LFSR_SMALL_LENGTH, LFSR_SMALL_0_GEN_POLY, 1);
  std::vector<msequence> ms_S1(num_streams);
  ms_S1[0] = msequence_create(LFSR_LARGE_LENGTH, LFSR_LARGE_0_GEN_POLY, 1);
  ms_S1[1] = msequence_create(LFSR_LARGE_LENGTH, LFSR_LARGE_1_GEN_POLY, 1);
  unsigned char *pp = p.data();
  try {
    if (mode == "tx") {
      rx_beamforming::framegen fg(M, cp_len, num_streams, NUM_ACCESS_CODES, pp, ms_S0, ms_S1);
      modem mod = modem_create(LIQUID_MODEM_QAM4);
      std::vector<std::vector<unsigned int> > tx_data(num_streams, std::vector<unsigned int>(tx_sig_len));
      std::vector<std::vector<gr_complex> > tx_sig(num_streams, std::vector<gr_complex>(tx_sig_len));
      srand(7);
      for (unsigned int c = 0; c < num_streams; c++)
        for (unsigned int s = 0; s < tx_sig_len; s++) { tx_data[c][s] = rand() % ARITY; modem_modulate(mod, tx_data[c][s], &tx_sig[c][s]); }
      const unsigned int sync_len = (NUM_ACCESS_CODES * num_streams + 1) * L;
      const size_t total = (size_t)sync_len * 3 + (size_t)PID_MAX * L;  // zeros | sync | payload | zeros
      std::vector<std::vector<gr_complex> > tx(num_streams, std::vector<gr_complex>(total, gr_complex(0, 0)));
      std::vector<gr_complex *> ptr(num_streams), in(num_streams);
      for (unsigned int c = 0; c < num_streams; c++) ptr[c] = tx[c].data() + sync_len;
      if (fg.write_sync_words(ptr) != sync_len) return 3;
      for (unsigned int pid = 0; pid < PID_MAX; pid++) {
        for (unsigned int c = 0; c < num_streams; c++) { ptr[c] = tx[c].data() + 2 * (size_t)sync_len + (size_t)pid * L; in[c] = tx_sig[c].data() + (size_t)pid * Mo; }
        fg.assemble_mimo_packet(ptr, in);
      }
      // BASEBAND_GAIN (main.cc:1049), channel H = [[1, .5], [.5i, 1]], noise at ~30 dB
      const gr_complex H[2][2] = {{gr_complex(1, 0), gr_complex(0.5f, 0)}, {gr_complex(0, 0.5f), gr_complex(1, 0)}};
      const double sigma = 0.25 * std::sqrt(1.25 / 1000.0 / 2.0);
      for (unsigned int r = 0; r < num_streams; r++) {
        std::vector<gr_complex> rx(total);
        for (size_t n = 0; n < total; n++) {
          gr_complex a(0, 0);
          for (unsigned int t = 0; t < num_streams; t++) a += H[r][t] * (tx[t][n] * (float)BASEBAND_GAIN);
          const double u1 = urand(), u2 = urand(), rr = std::sqrt(-2.0 * std::log(u1));
          rx[n] = a + gr_complex((float)(sigma * rr * std::cos(6.283185307179586 * u2)), (float)(sigma * rr * std::sin(6.283185307179586 * u2)));
        }
        rub_file_write_fc32((dir + "/rx" + std::to_string(r + 1) + ".dat").c_str(), reinterpret_cast<float *>(rx.data()), total);
        rub_file_write_u32((dir + "/tx_data" + std::to_string(r + 1) + ".dat").c_str(), tx_data[r].data(), tx_sig_len);
      }
      printf("{\"mode\": \"tx\", \"samples\": %zu, \"occupied\": %u}\n", total, Mo);
      return 0;
    }
    // ---- rx ----
    rx_beamforming::framesync fs(M, cp_len, num_streams, NUM_ACCESS_CODES, pp, ms_S0, ms_S1, callback);
    const size_t cap = (size_t)(NUM_ACCESS_CODES * num_streams + 1) * L * 3 + (size_t)PID_MAX * L;
    std::vector<std::vector<gr_complex> > rx(num_streams, std::vector<gr_complex>(cap));
    std::vector<gr_complex *> rx_buffer(num_streams);
    std::vector<std::vector<gr_complex> > sig(num_streams, std::vector<gr_complex>((size_t)PID_MAX * L));
    uint64_t n_read = 0;
    for (unsigned int r = 0; r < num_streams; r++) {
      if (rub_file_read_fc32((dir + "/rx" + std::to_string(r + 1) + ".dat").c_str(), reinterpret_cast<float *>(rx[r].data()), cap, &n_read)) return 4;
      rx_buffer[r] = rx[r].data();
      rx_sig.push_back(sig[r].data());
    }
    framesync_states_t st = fs.execute(rx_buffer, (unsigned int)n_read);
    modem dem = modem_create(LIQUID_MODEM_QAM4);
    unsigned long valid[2] = {0, 0};
    for (unsigned int c = 0; c < num_streams; c++) {
      std::vector<unsigned int> tx_data(tx_sig_len), rx_data(tx_sig_len);
      FILE *f = fopen((dir + "/tx_data" + std::to_string(c + 1) + ".dat").c_str(), "rb");
      if (!f || fread(tx_data.data(), 4, tx_sig_len, f) != tx_sig_len) return 5;
      fclose(f);
      for (unsigned int s = 0; s < tx_sig_len; s++) { modem_demodulate(dem, rx_sig[c][s], &rx_data[s]); valid[c] += rx_data[s] == tx_data[s]; }
      rub_file_write_fc32((dir + "/rx_sig" + std::to_string(c + 1) + ".dat").c_str(), reinterpret_cast<float *>(rx_sig[c]), tx_sig_len);
      rub_file_write_u32((dir + "/rx_data" + std::to_string(c + 1) + ".dat").c_str(), rx_data.data(), tx_sig_len);
    }
    std::vector<std::vector<std::vector<gr_complex> > > G = fs.get_G();
    printf("{\"mode\": \"rx\", \"state\": %d, \"sync_index\": %lu, \"plateau_start\": [%lu, %lu], \"plateau_end\": [%lu, %lu], "
           "\"frames_detected\": %u, \"valid_packets\": %u, \"valid_symbols\": [%lu, %lu], \"symbols\": %u, "
           "\"G00\": [%f, %f], \"G01\": [%f, %f]}\n",
           (int)st, fs.get_sync_index(), fs.get_plateau_start(0), fs.get_plateau_start(1), fs.get_plateau_end(0),
           fs.get_plateau_end(1), num_frames_detected, num_valid_packets_received, valid[0], valid[1], tx_sig_len,
           G[3][0][0].real(), G[3][0][0].imag(), G[3][0][1].real(), G[3][0][1].imag());
    return 0;
  } catch (const std::exception &e) {
    fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
}
