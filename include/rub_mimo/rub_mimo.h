/*
 * rub_mimo.h — C ABI of librubmimo_b200.so, the B200-native MIMO-OFDM receive path.
 *
 * The reference (yefeng22222/RUB_MIMO) has no FFI/plugin layer: its boundary is the C++
 * class API of mimo/framing.h.  This header is the C-ABI a maintainer binds instead of
 * linking mimo/framing.cc; every entry point cites the reference interface it replaces.
 * A framing.h-compatible C++ facade on top of this ABI is in include/rub_mimo/framing.h.
 *
 * Conventions
 *   - All complex samples are interleaved float32 (re, im) = the reference's gr_complex
 *     (mimo/framing.h:23, std::complex<float>).
 *   - No exceptions, no exit(): every call returns a rub_status (the reference uses
 *     assert/printf/exit(1), mimo/framing.cc:395-401, :498-499, :1021-1022).
 *   - Device entry points take raw CUDA device pointers and run asynchronously on the
 *     handle's CUDA stream.  There is NO CPU fallback: without a CUDA device the device
 *     entry points return RUB_ERR_NO_DEVICE.
 *   - One host thread per handle (the reference objects are single-threaded,
 *     mimo/main.cc:922).
 */
#ifndef RUB_MIMO_H
#define RUB_MIMO_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RUB_ABI_VERSION 1u

/* ---------------------------------------------------------------- status ----------- */
typedef enum rub_status {
  RUB_OK = 0,
  RUB_ERR_INVALID_ARG = 1,   /* bad pointer / size / enum                               */
  RUB_ERR_UNSUPPORTED = 2,   /* legal config this build has no kernel for               */
  RUB_ERR_NO_DEVICE = 3,     /* no CUDA device: the receive path has no CPU fallback    */
  RUB_ERR_CUDA = 4,          /* a CUDA runtime call failed (see rub_last_error)         */
  RUB_ERR_NOMEM = 5,
  RUB_ERR_NCCL = 6,
  RUB_ERR_IO = 7
} rub_status;

/* ------------------------------------------------------- subcarrier types ---------- */
/* liquid-dsp OFDMFRAME_SCTYPE_* values used by mimo/framing.cc:218, :312, :949-998.   */
#define RUB_SCTYPE_NULL 0
#define RUB_SCTYPE_PILOT 1
#define RUB_SCTYPE_DATA 2

/* ------------------------------------------------------------ enums ---------------- */
/* modulation: bits per symbol of a square Gray QAM in liquid-dsp's convention
 * (replaces MODEM_SCHEME / ARITY, mimo/config.h:107-108; GUI names MOD_QUAM4/16/64,
 * Interface/usrp_device.h:11-14).                                                     */
#define RUB_MOD_QPSK 2u
#define RUB_MOD_QAM16 4u
#define RUB_MOD_QAM64 6u
#define RUB_MOD_QAM256 8u

#define RUB_DET_ZF 0u   /* zero forcing (the reference's only detector, framing.cc:1344) */
#define RUB_DET_MMSE 1u /* W = (G^H G + noise_var I)^-1 G^H                              */

#define RUB_EST_LS_FULLBAND 0u    /* TDMA full-band LS, mimo/framing.cc:801-824          */
#define RUB_EST_LS_COMB_INTERP 1u /* comb pilots k = t (mod P) + linear interpolation    */

/* flags (quirk switches follow SURVEY.md appendix A numbering)                         */
#define RUB_FLAG_Q1_IDENTITY_INIT 0x1u /* G accumulates onto identity, framing.cc:302-319 */
#define RUB_FLAG_MMSE_UNBIASED 0x2u    /* scale MMSE rows by 1/mu_s                       */
#define RUB_FLAG_ZF_CHOLESKY 0x4u      /* N=2 ZF: use the general Cholesky form instead of
                                          the reference adjugate form (framing.cc:1352)  */

/* output mask bits for rub_rx_process_batch                                            */
#define RUB_OUT_EQ 0x1u      /* equalised symbols (what mimo_callback receives)          */
#define RUB_OUT_LLR 0x2u     /* max-log LLRs                                             */
#define RUB_OUT_BITS 0x4u    /* packed hard bits                                         */
#define RUB_OUT_RXDATA 0x8u  /* demodulated symbol indices (rx_data, main.cc:1405)       */
#define RUB_OUT_G 0x10u      /* channel estimate G                                       */

/* execution path selector (rub_rx_set_path)                                            */
#define RUB_PATH_AUTO 0u
#define RUB_PATH_STAGED 1u /* FFT -> HBM -> estimate -> weights -> detect (any config)   */
#define RUB_PATH_FUSED 2u  /* one persistent kernel per frame batch (eligible configs)   */

/* ------------------------------------------------------------ config --------------- */
/* Runtime mirror of the compile-time macros of mimo/config.h:65-108.                   */
typedef struct rub_config {
  uint32_t struct_size;       /* = sizeof(rub_config), ABI guard                         */
  uint32_t M;                 /* NUM_SUBCARRIERS (config.h:65), power of two 64..4096    */
  uint32_t cp_len;            /* CP_LENGTH (config.h:66), <= M                           */
  uint32_t num_streams;       /* NUM_STREAMS (config.h:106): N = N_tx = N_rx, 1..8       */
  uint32_t num_access_codes;  /* NUM_ACCESS_CODES (config.h:104)                         */
  uint32_t num_data_symbols;  /* PID_MAX (config.h:92): payload OFDM symbols per frame   */
  uint32_t modulation;        /* RUB_MOD_* (bits per symbol)                             */
  uint32_t detector;          /* RUB_DET_*                                               */
  uint32_t estimator;         /* RUB_EST_*                                               */
  uint32_t pilot_spacing;     /* comb spacing P (framing.cc:974 uses 8); 0 -> 8          */
  uint32_t flags;             /* RUB_FLAG_*                                              */
  float noise_var;            /* per-subcarrier noise variance in the units of Y         */
  const uint8_t *sctype;      /* p[M] RUB_SCTYPE_* or NULL = all DATA (USE_ALL_CARRIERS) */
} rub_config;

/* number of training OFDM symbols per frame: nac*N (full-band TDMA, framing.cc:191-204)
 * or nac (comb).                                                                       */
uint32_t rub_config_num_training_symbols(const rub_config *cfg);
/* M_occupied = M_pilot + M_data (framing.cc:329)                                        */
uint32_t rub_config_num_occupied(const rub_config *cfg);
/* validate a configuration; RUB_OK or the reason                                        */
rub_status rub_config_validate(const rub_config *cfg);

/* batch layout of the IQ input: iq[frame*frame_stride + rx*rx_stride + sample], in
 * complex samples.  0 = dense ([frame][rx][ (T+D)*(M+cp) ]).                            */
typedef struct rub_iq_layout {
  uint64_t frame_stride;
  uint64_t rx_stride;
  uint64_t first_sample; /* offset of the first training symbol's CP inside each rx row  */
} rub_iq_layout;

/* Per-frame outputs (any pointer may be NULL when its RUB_OUT_* bit is clear).
 *   eq      [frame][stream][sym][j]       complex64   j over occupied carriers, ascending
 *   llr     [frame][stream][sym][j][bit]  float32     bit 0 = MSB of the symbol index;
 *                                                     positive => bit 0
 *   bits    [frame][stream][sym][row]     uint8       row = ceil(Mo*q/8) bytes, MSB first
 *   rx_data [frame][stream][sym][j]       uint8       liquid-style symbol index
 *   G       [frame][rx][tx][k]            complex64   all M bins (0 on NULL carriers)
 *   counters[stream][4] uint64: bit_errors, bits, symbol_errors, symbols
 *           (accumulated; only touched when tx_data != NULL)
 * tx_data  [frame][stream][sym][j] uint8  transmitted symbol indices (main.cc:1236)     */
typedef struct rub_rx_io {
  const float *iq;
  rub_iq_layout layout;
  const uint8_t *tx_data;
  /* optional per-link FFT window starts (quirk Q2, framing.cc:805-808):
   * timing[frame][rx][T] int32 sample offsets relative to the rx row start; and
   * payload_start[frame] (quirk Q4, framing.cc:857).  NULL = pre-aligned frames.        */
  const int32_t *timing;
  const int32_t *payload_start;
  float *eq;
  float *llr;
  uint8_t *bits;
  uint8_t *rx_data;
  float *G;
  /* error counters [stream][bit errors, bits, symbol errors, symbols] (uint64), only with tx_data.
   *   rub_rx_process_batch:       a DEVICE pointer the kernels atomicAdd this batch's counts into, or NULL to
   *                               accumulate in the handle's own device counters (rub_rx_read_counters);
   *   rub_rx_process_batch_host,
   *   rub_rx_process_capture:     a HOST pointer that receives a copy of the handle's CUMULATIVE counters
   *                               after the call (not this batch's alone; rub_rx_reset_counters zeroes them).
   * Passing a host pointer to rub_rx_process_batch is an error the library cannot detect.                 */
  uint64_t *counters;
  uint32_t out_mask;
} rub_rx_io;

/* ------------------------------------------------------- receiver handle ----------- */
typedef struct rub_rx rub_rx;

/* Replaces rx_beamforming::framesync::framesync (mimo/framing.h:189-196): builds every
 * table the receive chain needs.  S1 is [tx][code][k] complex64 (the frequency-domain
 * access codes, framing.cc:1214-1262) or NULL to generate them from the default LFSRs
 * (rub_default_S1).  device < 0 = current CUDA device.  cuda_stream = a cudaStream_t
 * (NULL = the handle creates its own non-blocking stream).                              */
rub_status rub_rx_create(rub_rx **out, const rub_config *cfg, const float *S1, int device,
                         void *cuda_stream);
void rub_rx_destroy(rub_rx *h);

/* Replaces framesync::execute's decode work (mimo/framing.cc:535-589, :801-832, and the
 * demod/count loop mimo/main.cc:1403-1410) for a batch of pre-aligned frames whose IQ
 * samples are resident in device memory.  Asynchronous on the handle's stream.          */
rub_status rub_rx_process_batch(rub_rx *h, const rub_rx_io *io, uint32_t n_frames);

/* Same call with HOST buffers (pageable or pinned): copies inputs host->device, runs the
 * chain, copies the requested outputs back, in chunks pipelined over internal streams.
 * Synchronous.  This is what the framing.h facade and the offline IQ-file driver call. */
rub_status rub_rx_process_batch_host(rub_rx *h, const rub_rx_io *io, uint32_t n_frames);
/* frames per pipeline chunk of rub_rx_process_batch_host (0 = automatic, about 256 MB of traffic)  */
rub_status rub_rx_set_host_chunk(rub_rx *h, uint32_t frames);

rub_status rub_rx_sync(rub_rx *h);
rub_status rub_rx_set_path(rub_rx *h, uint32_t path);
uint32_t rub_rx_get_path(const rub_rx *h);               /* path the last batch used   */
rub_status rub_rx_reset_counters(rub_rx *h);
/* device pointer of the internal uint64[4*N] counters (used when io->counters==NULL)    */
uint64_t *rub_rx_device_counters(rub_rx *h);
rub_status rub_rx_read_counters(rub_rx *h, uint64_t *host_out /* [4*N] */);
/* kernels launched by this handle since creation (for bench.py's gpu_launches)          */
uint64_t rub_rx_launch_count(const rub_rx *h);
/* CUDA-event time of the last rub_rx_process_batch on the handle's stream, in ms;
 * requires rub_rx_sync first.  dominant_ms = the detect/fused kernel alone.             */
rub_status rub_rx_last_timing(rub_rx *h, float *total_ms, float *dominant_ms);
/* name of the dominant kernel of the last batch ("k_rx_ws", "k_rx_fused", "k_detect_lean", "k_detect") */
const char *rub_rx_last_kernel(const rub_rx *h);
/* algorithmic bytes of one rub_rx_process_batch call (SURVEY.md 8d formula)             */
uint64_t rub_rx_algorithmic_bytes(const rub_rx *h, uint32_t n_frames, uint32_t out_mask,
                                  int with_tx_data);

/* ------------------------------------------------ synchronisation (rows f1 / f2) ---- */
/* Schmidl & Cox timing metric of one rx stream, framesync::execute_sc_sync(x, stream)
 * (mimo/framing.cc:626-637): y[n] = |P[n]|^2 / R[n]^2 with P the M/2-lag autocorrelation over
 * M/2 samples and R = 0.5 * sum of the last M |x|^2 (samples before the first are zero).  One
 * GPU thread per output sample evaluates both sums oldest-to-newest, which is the summation
 * order of the oracle: the metric is bit-exact.  x and y are HOST buffers.                  */
rub_status rub_rx_sc_metric(rub_rx *h, const float *x, uint64_t num_samples, float *y);
/* The same metric in a selectable form.  RUB_SYNC_FIR: as above (the reference's two FIR dot products per
 * sample in liquid's order: O(M) per sample, bit-identical metric, so y crosses PLATEAU_THREASHOLD exactly
 * where the reference's does).  RUB_SYNC_SCAN: sliding sums P(n) = P(n-1) + c(n) - c(n-M/2), R likewise, one
 * block scan per tile of 2048 samples: O(1) per sample; the sums are added in another order, so y differs in
 * the last bits and a plateau start can move by a sample where y sits within rounding of the threshold.      */
#define RUB_SYNC_FIR 0u
#define RUB_SYNC_SCAN 1u
rub_status rub_rx_sc_metric_ex(rub_rx *h, const float *x, uint64_t num_samples, float *y, uint32_t mode);
/* metric form used by rub_rx_process_capture (default RUB_SYNC_SCAN)                                          */
rub_status rub_rx_set_sync_mode(rub_rx *h, uint32_t mode);
/* Debug sinks of the reference's DEBUG_LOG build (mimo/framing.cc:598-600, :676-680, :873-883; read by
 * mimo/apps/plot.py:27-40): when dir is set, rub_rx_process_capture writes dir/f_sc_<s>.dat (float32 metric of
 * every capture sample of rx stream s = 1..N) and, for the last burst found, dir/corr_<s>_<a>.dat (float32
 * [access_code_buffer_len - M] correlation powers |X . conj(S1)|^2 / M^2 at the candidate offsets of access code
 * a = 1..nac*N; a = 0 = the S0 preamble, only after rub_rx_set_S0).  NULL or "" switches them off.            */
rub_status rub_rx_set_debug_dir(rub_rx *h, const char *dir);
/* Access-code timing search of estimate_channel (mimo/framing.cc:702-744, USE_NEW_CHANNEL_EST)
 * over the window buffer (HOST, [N][window_len] complex64, oldest sample first): for every
 * candidate i in [0, M+cp) and every (rx, ac_id) it correlates the M samples at
 * i + (M+cp)*(ac_id+1) with the access code and keeps the first maximum.  The correlation is
 * evaluated in the time domain against s1 (sum_k X[k] conj(S1[k]) = sqrt(M) sum_n x[n]
 * conj(s1[n])), so no FFT is needed.  corr_indices [N][nac*N], s0_corr_index [N] (may be NULL;
 * correlation with the S0 preamble at offset i, framing.cc:711-721; needs rub_rx_set_S0).    */
rub_status rub_rx_timing_search(rub_rx *h, const float *window, uint64_t window_len,
                                int32_t *corr_indices, int32_t *s0_corr_index);
/* time-domain S0 preamble (M complex64) used by rub_rx_timing_search's optional S0 output  */
rub_status rub_rx_set_S0(rub_rx *h, const float *s0);

/* ----------------------------------------------------------- multi-GPU ------------- */
/* Frames are independent (framing.cc:653-886 recomputes all state per frame), so a batch
 * is sharded by frame range with no data-path collective; the only exchange is one
 * ncclAllReduce(sum, uint64) over the 4*N error counters.  NCCL is resolved with dlopen
 * at first use (no link-time dependency).                                               */
#define RUB_NCCL_UNIQUE_ID_BYTES 128
rub_status rub_comm_get_unique_id(uint8_t id[RUB_NCCL_UNIQUE_ID_BYTES]);
rub_status rub_comm_init(rub_rx *h, const uint8_t id[RUB_NCCL_UNIQUE_ID_BYTES], int rank,
                         int world_size);
/* Global counters = sum over ranks of each rank's cumulative local counters (the handle's own
 * counters are only read, so the call is idempotent and may follow every batch).  The local
 * counters are snapshotted on the handle's stream and reduced on a side stream: the next batch
 * does not wait for the collective.  With world_size 1 (no rub_comm_init) the result equals the
 * local counters.  rub_rx_read_counters_global waits for the latest reduction and copies it out. */
rub_status rub_allreduce_counters(rub_rx *h);
rub_status rub_rx_read_counters_global(rub_rx *h, uint64_t *host_out /* [4*N] */);
rub_status rub_comm_destroy(rub_rx *h);
/* frame range [begin, end) of `rank` when n_frames are sharded over world_size ranks    */
void rub_shard_range(uint64_t n_frames, int rank, int world_size, uint64_t *begin,
                     uint64_t *end);

/* ------------------------------------------ host-only setup (a8 / a9 rows) --------- */
/* liquid-dsp msequence stand-in (Fibonacci LFSR, liquid <= 1.3 semantics) used at
 * mimo/main.cc:1268-1270, mimo/framing.cc:1075, :1240.                                  */
typedef struct rub_msequence {
  uint32_t m, g, a, n, v, b;
} rub_msequence;
void rub_msequence_init(rub_msequence *ms, uint32_t m, uint32_t g, uint32_t a);
void rub_msequence_reset(rub_msequence *ms);
uint32_t rub_msequence_advance(rub_msequence *ms);
uint32_t rub_msequence_generate_symbol(rub_msequence *ms, uint32_t bps);

/* mimo/framing.cc:949-998; use_all_carriers/add_null_carriers mirror config.h:95-96     */
void rub_ofdmframe_init_default_sctype(uint8_t *p, uint32_t M, int use_all_carriers,
                                       int add_null_carriers);
/* mimo/framing.cc:1000-1030 (returns RUB_ERR_INVALID_ARG instead of exit(1))            */
rub_status rub_ofdmframe_validate_sctype(const uint8_t *p, uint32_t M, uint32_t *M_null,
                                         uint32_t *M_pilot, uint32_t *M_data);
/* mimo/framing.cc:1053-1111 (USE_NEW_INIT_S0)                                            */
rub_status rub_ofdmframe_init_S0(const uint8_t *p, uint32_t M, float *S0, float *s0,
                                 rub_msequence *ms);
/* mimo/framing.cc:1214-1262 (USE_NEW_INIT_S1, BPSK)                                      */
rub_status rub_ofdmframe_init_S1(const uint8_t *p, uint32_t M, uint32_t num_access_codes,
                                 float *S1, float *s1, rub_msequence *ms);
/* S1 for all N streams from the default generator polynomials (config.h:70-75 for
 * streams 0/1, further degree-13 primitive polynomials for streams 2..7).
 * S1 [N][nac][M] complex64, s1 (time domain, may be NULL) same shape.                   */
rub_status rub_default_S1(const rub_config *cfg, float *S1, float *s1);
rub_status rub_default_S0(const rub_config *cfg, float *S0, float *s0);
uint32_t rub_default_lfsr_poly(uint32_t stream);

/* mimo/framing.cc:1344-1367: 2x2 adjugate "inverse"; W, G row-major [2][2] complex64;
 * returns 1/|det|^2 in *gain.                                                            */
rub_status rub_invert_2x2(float *W, const float *G, float *gain);

/* liquid-dsp square-QAM modem in the convention of SURVEY.md 8c (replaces
 * modem_modulate / modem_demodulate, mimo/main.cc:1237, :1405).                         */
rub_status rub_modem_modulate(uint32_t bits_per_symbol, uint32_t sym, float out[2]);
rub_status rub_modem_demodulate(uint32_t bits_per_symbol, const float in[2], uint32_t *sym);

/* ------------------------------------------------------- frame generator ----------- */
/* Replaces rx_beamforming::framegen (mimo/framing.h:42-103).                            */
typedef struct rub_framegen rub_framegen;
rub_status rub_framegen_create(rub_framegen **out, const rub_config *cfg, const float *S0,
                               const float *s0, const float *s1 /* [N][nac][M] */);
void rub_framegen_destroy(rub_framegen *fg);
/* framegen::write_sync_words (framing.cc:169-208): tx_buff[stream] each
 * (nac*N+1)*(M+cp) complex; returns the sample count.                                   */
uint32_t rub_framegen_write_sync_words(rub_framegen *fg, float *const *tx_buff);
/* comb-pilot training symbols for RUB_EST_LS_COMB_INTERP: nac symbols, all tx active on
 * disjoint combs.  tx_buff[stream] each nac*(M+cp) complex.                             */
uint32_t rub_framegen_write_comb_words(rub_framegen *fg, float *const *tx_buff);
/* framegen::assemble_mimo_packet (framing.cc:210-235): in_buff[stream] has Mo symbols,
 * tx_buff[stream] receives M+cp samples; returns symbol_len.                            */
uint32_t rub_framegen_assemble_mimo_packet(rub_framegen *fg, float *const *tx_buff,
                                           const float *const *in_buff);

/* Batched transmit waveform on the GPU: the access codes of
 * framegen::write_sync_words (framing.cc:191-204; S0 is not written) followed by
 * framegen::assemble_mimo_packet (framing.cc:210-235) for every payload symbol, for whole
 * batches of frames: symbol indices -> modulate -> carrier mapping -> IFFT ->
 * dft_normalizer -> cyclic prefix -> * baseband_gain.  Bit-identical to rub_framegen_*.
 * `ctx` supplies the device, stream and tables (any receiver handle of the same config).
 * tx_data: DEVICE [n_frames][N][D][Mo] symbol indices.  out: DEVICE complex64, sample n of
 * (frame f, stream s) at out[2*(f*frame_stride + s*stream_stride + n)], strides in complex
 * samples, (T + D)*(M + cp) samples written per row.  Asynchronous on the handle's stream. */
rub_status rub_framegen_batch_device(rub_rx *ctx, const uint8_t *tx_data, uint32_t n_frames,
                                     float *out, uint64_t frame_stride, uint64_t stream_stride,
                                     float baseband_gain);

/* --------------------------------------------- synthetic / offline IQ source -------- */
/* Replaces the USRP stream (mimo/main.cc:872-898) with a synthetic source: for each
 * frame, random symbol indices -> modulate -> framegen -> BASEBAND_GAIN -> per-link
 * L-tap Rayleigh (or a fixed flat channel) -> AWGN, written pre-aligned in the batch
 * layout.  Counter-based RNG keyed by (seed, global frame index) so any frame can be
 * regenerated independently (results do not depend on how frames are sharded).          */
typedef struct rub_synth_params {
  uint64_t seed;
  uint64_t first_frame;   /* global index of frame 0 of this call                       */
  uint32_t n_taps;        /* channel taps per link, each CN(0, 1/n_taps); 0 = fixed_H   */
  float snr_db;           /* per-rx SNR (signal power measured through the channel)      */
  float baseband_gain;    /* BASEBAND_GAIN, config.h:59                                  */
  const float *fixed_H;   /* [rx][tx] complex64 flat channel when n_taps == 0           */
  uint32_t include_s0;    /* prepend zeros + S0 like the reference burst (appendix C)   */
  uint32_t lead_zeros;    /* leading zero samples when include_s0                       */
  uint32_t n_threads;     /* host threads (0 = all)                                     */
} rub_synth_params;
/* samples per rx row the generator writes for one frame                                 */
uint64_t rub_synth_row_samples(const rub_config *cfg, const rub_synth_params *sp);
/* noise variance per subcarrier (units of Y) the generator used for `snr_db`            */
rub_status rub_synth_frames(const rub_config *cfg, const rub_synth_params *sp,
                            const float *S0s0 /* NULL=default */, const float *S1,
                            const float *s1, uint32_t n_frames, float *iq /* host */,
                            uint8_t *tx_data /* host, may be NULL */,
                            float *noise_var_out /* may be NULL */);

/* on-disk formats of the reference (mimo/main.cc:831-833, :906-918, :1243-1262,
 * :1413-1419; documented by mimo/apps/plot.py:27-40): raw fc32 / uint32 / float32.      */
rub_status rub_file_read_fc32(const char *path, float *dst, uint64_t max_samples,
                              uint64_t *n_read);
rub_status rub_file_write_fc32(const char *path, const float *src, uint64_t n_samples);
rub_status rub_file_write_u32(const char *path, const uint32_t *src, uint64_t n);

/* Offline IQ-file driver: the seam of mimo/main.cc:906-918, where the
 * reference reads "/tmp/rx%d.dat" instead of the USRP, with the sinks of :1413-1419.
 * One fc32 capture per rx antenna; frame f of every file starts at complex sample
 * first_sample + f*frame_stride.  The files are read in chunks into two pinned staging slots by a
 * reader thread while the previous chunk is on the GPU (rub_rx_process_batch_host pipelines its
 * H2D / kernels / D2H inside a chunk), and the requested outputs are appended to the sinks.
 * Any path array / path may be NULL.  Sinks per stream s: eq_paths[s] fc32 equalised symbols
 * ("rx_sig%d.dat"), rx_data_paths[s] uint32 decisions ("rx_data%d.dat"); tx_data_paths[s] is a
 * uint32 input of transmitted symbol indices, D*Mo per frame ("tx_data%d.dat"), enabling the
 * error counters.  llr_path / bits_path receive the batch layout of rub_rx_io, frame-major.   */
typedef struct rub_file_job {
  uint32_t struct_size;
  uint32_t n_frames;
  uint64_t first_sample;
  uint64_t frame_stride;            /* 0 = (T+D)*(M+cp): frames back to back                 */
  const char *const *rx_paths;      /* [N]                                                    */
  const char *const *tx_data_paths; /* [N] or NULL                                            */
  const char *const *eq_paths;      /* [N] or NULL                                            */
  const char *const *rx_data_paths; /* [N] or NULL                                            */
  const char *llr_path;             /* or NULL                                                */
  const char *bits_path;            /* or NULL                                                */
  uint32_t chunk_frames;            /* frames per staging slot, 0 = about 64 MB of input      */
} rub_file_job;
/* frames_done (may be NULL) = frames fully processed; a capture shorter than n_frames is
 * not an error, the run stops at the last complete frame.                                    */
rub_status rub_rx_process_files(rub_rx *h, const rub_file_job *job, uint64_t *frames_done);

/* Multi-burst capture (rows f1/f2 put together: no pre-aligned frames): the reference's receive
 * loop — framesync::execute with Schmidl & Cox plateau search, access-code buffering, timing
 * search, LS estimate, invert and decode (framing.cc:471-506, :591-651, :653-886) — over a HOST
 * capture [N][n_samples] complex64 that may hold any number of bursts.  The metric, one batched
 * timing search over all bursts found and one batched decode run on the GPU; the decode reads the
 * capture in place through per-link timing tables (quirks Q2/Q4).  `out` supplies HOST buffers
 * sized for max_frames bursts in the layouts of rub_rx_io (iq / layout / timing / payload_start
 * are ignored; tx_data, if given, enables the counters; out->counters, if given, receives the
 * handle's counters).  sync_index[f] (may be NULL) = mean plateau start of burst f in capture
 * samples.  Bursts whose window would start before the capture or run past its end are skipped.
 * After a burst the search restarts behind its window (the reference stops at the first burst
 * and is reset by its caller).  Synchronous.                                               */
rub_status rub_rx_process_capture(rub_rx *h, const float *capture, uint64_t n_samples,
                                  float threshold, uint32_t max_frames, const rub_rx_io *out,
                                  uint32_t *n_found, uint64_t *sync_index);

/* ------------------------------------------- configuration front-end (row f4) ------ */
/* The fields of the reference's `options` struct (mimo/main.cc:129-156) that are not part of
 * rub_config: the offline path carries them so that a command line / GUI record written for the
 * reference parses unchanged.                                                               */
typedef struct rub_frontend_options {
  double cent_freq, samp_rate, txgain, rxgain;
  float dsp_gain;                 /* BASEBAND_GAIN, applied by the transmit side              */
  char rx_addr[64], tx_addr[64], rx_subdev[32], tx_subdev[32];
  int32_t verbose;                /* 1 after -v/--verbose, 0 after -q/--quite                 */
  int32_t help;                   /* 1 after -h/--help                                        */
  uint32_t num_nullcarriers;      /* GUI record only ("Number of Nullcarriers")               */
} rub_frontend_options;
/* read_options (mimo/main.cc:174-240): the same flag names — --freq/-f, --rate/-r, --dsp_gain,
 * --tx_gain, --rx_gain, --num_subcarriers, --cp_len, --rx_addr, --tx_addr, --tx_subdev,
 * --rx_subdev, --verbose/-v, --quite/-q, --help/-h — as "--flag value" or "--flag=value".
 * cfg and fe must be initialised by the caller (defaults); only the given flags are changed.
 * An unknown flag or a missing / malformed value is RUB_ERR_INVALID_ARG (boost throws).     */
rub_status rub_config_from_args(int argc, const char *const *argv, rub_config *cfg,
                                rub_frontend_options *fe);
/* One device record of the GUI's JSON configuration (Interface/usrp_device.cpp:13-29, keys
 * "Number of Subcarriers", "Number of Nullcarriers", "Prefix Length", "Training Sequences",
 * "TX Gain", "RX Gain", "Center Freq.", "Samp. Rate", "Adress", "Subdevice Specifications").
 * Flat object only; keys that are absent leave cfg / fe unchanged.                          */
rub_status rub_config_from_json(const char *json, rub_config *cfg, rub_frontend_options *fe);

/* ------------------------------------------------------------- misc ---------------- */
const char *rub_strerror(rub_status s);
const char *rub_last_error(void); /* thread-local detail string of the last failure     */
uint32_t rub_abi_version(void);
int rub_device_count(void);       /* 0 when no CUDA device is visible                   */

#ifdef __cplusplus
}
#endif
#endif /* RUB_MIMO_H */
