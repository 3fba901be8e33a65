/*
 * rub_oracle.h — CPU ORACLE for the RUB_MIMO receive hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain-C restatement of the reference's receive algorithm
 * (/root/reference/mimo/framing.cc, mimo/main.cc), generalised to NxN / MMSE / square-QAM /
 * max-log LLR in the reference's conventions.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product (rub_mimo_b200/, librubmimo_b200.so) never includes, links or calls it.
 *
 * PARITY STATUS: pinned to the reference's OWN framing code for the path it implements (2x2, ZF,
 * full-band LS): `make -C oracle ref` compiles /root/reference/mimo/framing.cc where it lies against
 * stand-in headers for the libraries that are not installed (oracle/shim/: FFTW3f plans run by this
 * file's FFT, VOLK generic kernels, liquid-dsp msequence / wdelay / window / firfilt, gr_complex,
 * boost::format) into oracle/_ref/, oracle/make_ref_fixtures.py runs it and commits its outputs as
 * tests/golden/ref_*.npz, and tests/test_ref_fixtures.py requires this oracle (and the CUDA path)
 * to reproduce them BIT FOR BIT: framegen waveform, Schmidl & Cox plateau / sync index / sample
 * count, timing search, LS estimate G (quirks Q1/Q2/Q4), invert() and 1000 decoded OFDM symbols.
 * What the stand-ins cannot pin is the rounding of the real FFTW / VOLK / liquid builds (none is
 * vendored or version-pinned upstream: mimo/makefile:8-13 only has -l flags), and upstream has no
 * tests, fixtures or golden vectors of its own.  Everything the reference has no
 * code for (N > 2, MMSE, QAM > 4, LLRs, comb pilots) stays pinned by
 *   (1) known-answer tests derived from the reference source (tests/test_oracle_*.py),
 *   (2) an independent float64 numpy model (oracle/oracle_f64.py: np.fft + np.linalg),
 *   (3) committed golden fixtures regenerated only by oracle/make_golden.py.
 *
 * Arithmetic contract ("mirror fp32"): every float operation below is IEEE-754 binary32,
 * evaluated in source order, no contraction except the explicit fmaf() calls
 * (-ffp-contract=off).  The CUDA path implements the same operation sequence, so hard
 * decisions and error counters can be compared bit-exactly.
 */
#ifndef RUB_ORACLE_H
#define RUB_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { float re, im; } ocf;

/* liquid OFDMFRAME_SCTYPE_* */
#define ORC_SC_NULL 0
#define ORC_SC_PILOT 1
#define ORC_SC_DATA 2

#define ORC_DET_ZF 0
#define ORC_DET_MMSE 1
#define ORC_EST_FULLBAND 0
#define ORC_EST_COMB 1
#define ORC_FLAG_Q1 0x1u
#define ORC_FLAG_UNBIASED 0x2u
#define ORC_FLAG_ZF_CHOLESKY 0x4u

typedef struct {
  uint32_t M, cp_len, N, nac, D, q, detector, estimator, P, flags;
  float noise_var;
  const uint8_t *sctype; /* NULL = all data */
} orc_config;

/* ---- msequence (liquid <= 1.3 semantics; mimo/main.cc:1268-1270) ---- */
typedef struct { uint32_t m, g, a, n, v, b; } orc_mseq;
void orc_mseq_init(orc_mseq *ms, uint32_t m, uint32_t g, uint32_t a);
void orc_mseq_reset(orc_mseq *ms);
uint32_t orc_mseq_advance(orc_mseq *ms);
uint32_t orc_mseq_symbol(orc_mseq *ms, uint32_t bps);

/* ---- subcarrier allocation (mimo/framing.cc:949-1030) ---- */
void orc_init_default_sctype(uint8_t *p, uint32_t M, int use_all, int add_null);
int orc_validate_sctype(const uint8_t *p, uint32_t M, uint32_t *Mn, uint32_t *Mp, uint32_t *Md);

/* ---- FFT (unnormalised; forward sign -1 as FFTW_FORWARD, mimo/framing.cc:368-372) ---- */
void orc_fft_twiddles(uint32_t M, ocf *tw);                 /* tw[i] = exp(-2*pi*i*i/M)  */
void orc_fft_forward(uint32_t M, const ocf *in, ocf *out);  /* M in {64..4096}           */
void orc_fft_backward(uint32_t M, const ocf *in, ocf *out); /* conj(fwd(conj(x)))        */

/* ---- preambles (mimo/framing.cc:1053-1111, :1214-1262) ---- */
int orc_init_S0(const uint8_t *p, uint32_t M, ocf *S0, ocf *s0, orc_mseq *ms);
int orc_init_S1(const uint8_t *p, uint32_t M, uint32_t nac, ocf *S1, ocf *s1, orc_mseq *ms);

/* ---- framegen (mimo/framing.cc:169-235) ---- */
uint32_t orc_write_sync_words(const orc_config *c, const ocf *s0, const ocf *s1 /*[N][nac][M]*/,
                              ocf *const *tx_buff);
uint32_t orc_write_comb_words(const orc_config *c, const ocf *S1 /*[N][nac][M]*/,
                              ocf *const *tx_buff);
uint32_t orc_assemble_mimo_packet(const orc_config *c, ocf *const *tx_buff,
                                  const ocf *const *in_buff);

/* ---- modem (liquid square QAM; mimo/main.cc:1237, :1405) ---- */
ocf orc_modulate(uint32_t q, uint32_t sym);
uint32_t orc_demodulate(uint32_t q, ocf x);
/* max-log LLRs of one equalised symbol; llr[q]; isig = 1/sigma_eff^2 */
void orc_llr(uint32_t q, ocf x, float isig, float *llr);

/* ---- weights (mimo/framing.cc:1344-1367 + generalisation) ---- */
float orc_invert_2x2(ocf W[4], const ocf G[4]);
/* G row-major [rx][tx]; W row-major [stream][rx]; gain[N], isig[N] */
void orc_weights(const orc_config *c, const ocf *G, ocf *W, float *gain, float *isig);

/* ---- receive chain on one pre-aligned (or timing-table) frame ----
 * rx[r] points at the rx row; timing (may be NULL) [r][T] FFT-window starts; payload_start
 * < 0 = aligned.  Outputs may be NULL.  Layouts as include/rub_mimo/rub_mimo.h.          */
typedef struct {
  ocf *eq; float *llr; uint8_t *bits; uint8_t *rx_data; ocf *G; ocf *W; float *gain;
  float *isig; uint64_t *counters; /* [N][4] accumulated */
} orc_frame_out;
int orc_rx_frame(const orc_config *c, const ocf *S1, const ocf *const *rx, uint64_t first_sample,
                 const int32_t *timing, int64_t payload_start, const uint8_t *tx_data,
                 orc_frame_out *out);
/* batch over dense [frame][rx][row] input; n_threads <= 1 => serial (the reference decodes on
 * one thread, mimo/main.cc:922); > 1 => OpenMP over frames.  Returns 0 on success.        */
int orc_rx_batch(const orc_config *c, const ocf *S1, const ocf *iq, uint64_t frame_stride,
                 uint64_t rx_stride, uint64_t first_sample, uint32_t n_frames,
                 const uint8_t *tx_data, ocf *eq, float *llr, uint8_t *bits, uint8_t *rx_data,
                 ocf *G, uint64_t *counters, int n_threads);

/* ---- faithful framesync state machine (mimo/framing.cc:471-506, :591-886) ---- */
typedef struct {
  int state;                 /* framesync_states_t value at return                       */
  uint64_t sync_index, num_samples_processed;
  uint64_t plateau_start[8], plateau_end[8];
  int32_t *corr_indices;     /* [N][nac*N] (caller allocates)                            */
  int32_t s0_corr_index[8];
  uint64_t window_start;     /* absolute sample index of window-buffer element 0          */
  int64_t payload_start;     /* buffer-relative                                           */
  uint32_t symbols_decoded;  /* callbacks fired (may exceed D, quirk Q14)                 */
} orc_sync_result;
/* in_buff[N] capture; eq receives the first D symbols per stream ([stream][sym][j]);
 * G [k][rx][tx] as the reference's get_G().  Returns 0, or 1 if no sync.                  */
int orc_framesync_execute(const orc_config *c, const ocf *S0, const ocf *S1,
                          const ocf *const *in_buff, uint64_t num_samples, float threshold,
                          orc_sync_result *res, ocf *eq, ocf *G_ref, ocf *W_ref, float *gain_ref);
/* Schmidl&Cox metric trace of one stream (mimo/framing.cc:626-637); y[num_samples]       */
void orc_sc_metric(uint32_t M, const ocf *x, uint64_t num_samples, float *y);

uint32_t orc_num_training(const orc_config *c);
uint32_t orc_num_occupied(const orc_config *c);
const char *orc_build_info(void);

#ifdef __cplusplus
}
#endif
#endif
