// rub_rx.cu — receiver handle, kernel dispatch and the device-side C ABI of
// librubmimo_b200.so (include/rub_mimo/rub_mimo.h).  There is no CPU execution path in this
// file: without a CUDA device every entry point returns RUB_ERR_NO_DEVICE.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>

#include "rub_internal.h"
#include "rub_launch.h"
#include "rub_kernels_staged.cuh"
#include "rub_kernels_sync.cuh"
#include "rub_kernels_tx.cuh"

using namespace rub;

#define CUDA_TRY(x)                                                                       \
  do {                                                                                    \
    cudaError_t e_ = (x);                                                                 \
    if (e_ != cudaSuccess) {                                                              \
      set_error("%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
      return RUB_ERR_CUDA;                                                                \
    }                                                                                     \
  } while (0)

// ---------------------------------------------------------------- NCCL via dlopen -----
typedef struct { char internal[128]; } nccl_uid;
typedef int (*fn_ncclGetUniqueId)(nccl_uid *);
typedef int (*fn_ncclCommInitRank)(void **, int, nccl_uid, int);
typedef int (*fn_ncclAllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef int (*fn_ncclCommDestroy)(void *);
typedef const char *(*fn_ncclGetErrorString)(int);
static struct {
  void *lib;
  fn_ncclGetUniqueId GetUniqueId;
  fn_ncclCommInitRank CommInitRank;
  fn_ncclAllReduce AllReduce;
  fn_ncclCommDestroy CommDestroy;
  fn_ncclGetErrorString GetErrorString;
} g_nccl;
static rub_status nccl_load() {
  if (g_nccl.lib) return RUB_OK;
  void *l = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!l) l = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!l) { set_error("cannot load libnccl.so.2: %s", dlerror()); return RUB_ERR_NCCL; }
  g_nccl.GetUniqueId = (fn_ncclGetUniqueId)dlsym(l, "ncclGetUniqueId");
  g_nccl.CommInitRank = (fn_ncclCommInitRank)dlsym(l, "ncclCommInitRank");
  g_nccl.AllReduce = (fn_ncclAllReduce)dlsym(l, "ncclAllReduce");
  g_nccl.CommDestroy = (fn_ncclCommDestroy)dlsym(l, "ncclCommDestroy");
  g_nccl.GetErrorString = (fn_ncclGetErrorString)dlsym(l, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy) {
    set_error("libnccl.so.2 lacks a required symbol");
    return RUB_ERR_NCCL;
  }
  g_nccl.lib = l;
  return RUB_OK;
}

// ---------------------------------------------------------------- handle --------------
struct rub_rx {
  HostCfg h;
  int device = 0;
  int num_sms = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  // tables
  cf *d_tw = nullptr;
  unsigned short *d_occ = nullptr;
  float *d_sgn = nullptr;
  cf *d_s1 = nullptr;  // time-domain access codes [tx][code][n] (timing search)
  cf *d_s0 = nullptr;  // time-domain S0 (optional)
  float s0_corr_scale = 0.f;  // M_S0 / M^2: turns the time-domain S0 correlation power into the reference's |X . conj(S0)|^2 / M^2
  void *d_sync = nullptr;  // grow-only scratch of the synchronisation calls
  size_t sync_bytes = 0;
  void *d_capt = nullptr;  // grow-only scratch of rub_rx_process_capture (capture, metric, flags, outputs)
  size_t capt_bytes = 0;
  uint32_t sync_mode = RUB_SYNC_SCAN;  // metric form of rub_rx_process_capture
  std::string debug_dir;               // f_sc_%d.dat / corr_%d_%d.dat sinks of rub_rx_process_capture ("" = off)
  unsigned char *d_null = nullptr;
  DemapConst lut;
  WeightMode wm;
  // staged scratch
  void *d_scratch = nullptr;
  size_t scratch_bytes = 0;
  // fused scratch
  cf *d_fW = nullptr;
  float *d_fG = nullptr;
  cf *d_fAcc = nullptr;             // warp-specialised kernel: LS estimate of the next frame, per CTA
  unsigned char *d_sgn8 = nullptr;  // warp-specialised kernel: packed access-code signs
  std::vector<float> sgn_host;      // access-code signs [tx][code][M]
  int fused_grid = 0;
  int fused_variant = 0;  // monolithic fused kernel: 1 = TMA task records (rub_launch.h)
  bool ws = false;  // the warp-specialised fused kernel (rub_kernels_ws.cuh) serves this configuration
  size_t fused_smem = 0;
  bool fused_ready = false;
  bool fused_unfit = false;  // no fused kernel fits this configuration (shared memory): staged path
  uint64_t *d_counters = nullptr;
  uint32_t path = RUB_PATH_AUTO, last_path = 0;
  const char *last_kernel = "";
  uint64_t launches = 0;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  bool timed = false;
  // host pipeline
  cudaStream_t s_in = nullptr, s_out = nullptr;
  void *d_pipe = nullptr;
  size_t pipe_bytes = 0;
  uint32_t host_chunk = 0;  // frames per pipeline chunk of the host entry point (0 = sized to ~256 MB)
  cudaEvent_t pev[12] = {};
  // comm: the counter all-reduce runs on its own stream against double-buffered snapshots
  void *comm = nullptr;
  int rank = 0, world = 1;
  cudaStream_t s_comm = nullptr;
  uint64_t *d_snap = nullptr;    // [2][32] snapshot of d_counters | [2][32] reduced result
  cudaEvent_t cev[4] = {};       // snapshot taken [2], reduction done [2]
  uint32_t n_reduce = 0;         // reductions issued so far
};

// fused kernels live in their own translation units (rub_fused.cu, rub_ws.cu)
static rub_status fused_prepare_dispatch(rub_rx *h, size_t *smem, int *grid) {
  const char *e = getenv("RUB_FUSED_WS");  // development switch: RUB_FUSED_WS=0 keeps the monolithic kernel
  h->ws = !(e && e[0] == '0') && ws_has_instance(h->h.log2M, h->h.N);
  int occ = 0;
  cudaError_t ce = cudaErrorInvalidValue;
  if (h->ws) {
    ce = ws_prepare(h->h.log2M, h->h.N, h->h.q, smem, &occ);
    if (ce != cudaSuccess || occ < 1) { cudaGetLastError(); h->ws = false; }  // e.g. 256-QAM staging does not fit: monolithic kernel
  }
  if (!h->ws) ce = fused_prepare(h->h.log2M, h->h.N, h->h.q, smem, &occ, &h->fused_variant);
  if (ce != cudaSuccess || occ < 1) {
    cudaGetLastError();
    set_error("no fused kernel fits this configuration (shared memory %zu B)", *smem);
    return RUB_ERR_UNSUPPORTED;
  }
  *grid = (h->ws ? 1 : occ) * h->num_sms;
  return RUB_OK;
}
static void fused_launch_dispatch(rub_rx *h, int grid, size_t smem, const FusedArgs &fa) {
  if (h->ws) ws_launch(h->h.log2M, h->h.N, grid, smem, h->stream, fa, h->lut);
  else fused_launch(h->h.log2M, h->h.N, grid, smem, h->stream, fa, h->lut, h->fused_variant);
}

// is the fused kernel applicable to this configuration + call?
static bool fused_eligible(const rub_rx *h, const rub_rx_io *io, uint64_t frame_stride, uint64_t rx_stride) {
  const HostCfg &c = h->h;
  if (!fused_has_instance(c.log2M, c.N)) return false;
  if (c.Mo != c.M) return false;                                   // ragged allocations -> staged
  if (c.c.estimator != RUB_EST_LS_FULLBAND) return false;          // comb -> staged
  if (io->timing || io->payload_start) return false;               // per-link windows -> staged
  // cp.async.bulk needs 16-byte aligned sources: even sample offsets everywhere
  if ((c.cp | c.L | io->layout.first_sample | frame_stride | rx_stride) & 1) return false;
  if (((uintptr_t)io->iq & 15) || ((uintptr_t)io->llr & 15) || ((uintptr_t)io->bits & 15) || ((uintptr_t)io->eq & 15)) return false;
  if (((uintptr_t)io->rx_data & 1) || ((uintptr_t)io->tx_data & 15)) return false;
  return true;
}

static void fill_weight_mode(rub_rx *h) {
  const HostCfg &c = h->h;
  h->wm.nv = c.c.noise_var;
  h->wm.mmse = (c.c.detector == RUB_DET_MMSE && c.c.noise_var > 0.f) ? 1 : 0;
  h->wm.unbiased = (c.c.flags & RUB_FLAG_MMSE_UNBIASED) ? 1 : 0;
  h->wm.zf2_adjugate = (c.N == 2 && c.c.detector == RUB_DET_ZF && !(c.c.flags & RUB_FLAG_ZF_CHOLESKY)) ? 1 : 0;
}

extern "C" int rub_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

extern "C" rub_status rub_rx_create(rub_rx **out, const rub_config *cfg, const float *S1, int device, void *cuda_stream) {
  if (!out) { set_error("out is NULL"); return RUB_ERR_INVALID_ARG; }
  *out = nullptr;
  rub_rx *h = new rub_rx();
  rub_status st = host_cfg_init(h->h, cfg);
  if (st) { delete h; return st; }
  if (rub_device_count() < 1) {
    delete h;
    set_error("no CUDA device visible: librubmimo_b200 has no CPU fallback for the receive path");
    return RUB_ERR_NO_DEVICE;
  }
  const HostCfg &c = h->h;
  if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
  h->device = device;
#define CT(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { set_error("%s failed: %s", #x, cudaGetErrorString(e_)); rub_rx_destroy(h); return RUB_ERR_CUDA; } } while (0)
  CT(cudaSetDevice(device));
  CT(cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, device));
  if (cuda_stream) h->stream = (cudaStream_t)cuda_stream;
  else { CT(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)); h->own_stream = true; }
  for (int i = 0; i < 4; i++) CT(cudaEventCreate(&h->ev[i]));
  // tables
  std::vector<cf> master, packed;
  build_twiddles(c.log2M, master, packed);
  CT(cudaMalloc(&h->d_tw, sizeof(cf) * packed.size()));
  CT(cudaMemcpy(h->d_tw, packed.data(), sizeof(cf) * packed.size(), cudaMemcpyHostToDevice));
  std::vector<unsigned short> occ;
  std::vector<unsigned char> nul(c.M);
  for (uint32_t k = 0; k < c.M; k++) { nul[k] = c.sctype[k] == RUB_SCTYPE_NULL; if (!nul[k]) occ.push_back((unsigned short)k); }
  CT(cudaMalloc(&h->d_occ, sizeof(unsigned short) * occ.size()));
  CT(cudaMemcpy(h->d_occ, occ.data(), sizeof(unsigned short) * occ.size(), cudaMemcpyHostToDevice));
  CT(cudaMalloc(&h->d_null, c.M));
  CT(cudaMemcpy(h->d_null, nul.data(), c.M, cudaMemcpyHostToDevice));
  // access-code signs: S1 is +-1 on occupied carriers (mimo/framing.cc:1240-1248), so X/S1 is a
  // sign flip; keep only the sign
  std::vector<float> S1v((size_t)c.N * c.nac * c.M * 2);
  if (S1) memcpy(S1v.data(), S1, S1v.size() * sizeof(float));
  else {
    rub_config c2 = c.c;
    c2.sctype = c.sctype.data();
    st = rub_default_S1(&c2, S1v.data(), nullptr);
    if (st) { rub_rx_destroy(h); return st; }
  }
  std::vector<float> sgn((size_t)c.N * c.nac * c.M);
  for (size_t i = 0; i < sgn.size(); i++) {
    const float re = S1v[2 * i], im = S1v[2 * i + 1];
    const bool isnull = nul[i % c.M];
    if (im != 0.f || (!isnull && re != 1.0f && re != -1.0f)) {
      set_error("S1 must be BPSK (+-1+0i) on occupied carriers (mimo/framing.cc:1246)");
      rub_rx_destroy(h);
      return RUB_ERR_UNSUPPORTED;
    }
    sgn[i] = isnull ? 0.f : re;
  }
  h->sgn_host = sgn;
  CT(cudaMalloc(&h->d_sgn, sizeof(float) * sgn.size()));
  CT(cudaMemcpy(h->d_sgn, sgn.data(), sizeof(float) * sgn.size(), cudaMemcpyHostToDevice));
  {
    // time-domain access codes s1 = IFFT(S1) * sqrt(1/M) (mimo/framing.cc:1228, :1253-1257)
    std::vector<cf> s1t((size_t)c.N * c.nac * c.M), blk(c.M);
    const float g1 = (float)sqrt(1.0 / (double)(float)c.M);
    for (size_t b = 0; b < (size_t)c.N * c.nac; b++) {
      for (uint32_t k = 0; k < c.M; k++) blk[k] = mk(nul[k] ? 0.f : S1v[2 * (b * c.M + k)], 0.f);
      host_fft_backward(c.log2M, blk.data(), s1t.data() + b * c.M, packed.data());
      for (uint32_t k = 0; k < c.M; k++) s1t[b * c.M + k] = cscale(s1t[b * c.M + k], g1);
    }
    CT(cudaMalloc(&h->d_s1, sizeof(cf) * s1t.size()));
    CT(cudaMemcpy(h->d_s1, s1t.data(), sizeof(cf) * s1t.size(), cudaMemcpyHostToDevice));
  }
  build_demap_const(c.q, h->lut);
  fill_weight_mode(h);
  CT(cudaMalloc(&h->d_counters, sizeof(uint64_t) * 4 * 8));
  CT(cudaMemset(h->d_counters, 0, sizeof(uint64_t) * 4 * 8));
#undef CT
  *out = h;
  return RUB_OK;
}

extern "C" void rub_rx_destroy(rub_rx *h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->s_comm) cudaStreamSynchronize(h->s_comm);
  if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
  cudaFree(h->d_tw); cudaFree(h->d_occ); cudaFree(h->d_sgn); cudaFree(h->d_null); cudaFree(h->d_s1); cudaFree(h->d_s0); cudaFree(h->d_sync); cudaFree(h->d_capt);
  cudaFree(h->d_scratch); cudaFree(h->d_fW); cudaFree(h->d_fG); cudaFree(h->d_fAcc); cudaFree(h->d_sgn8); cudaFree(h->d_counters); cudaFree(h->d_pipe);
  for (auto &e : h->ev) if (e) cudaEventDestroy(e);
  for (auto &e : h->pev) if (e) cudaEventDestroy(e);
  cudaFree(h->d_snap);
  for (auto &e : h->cev) if (e) cudaEventDestroy(e);
  if (h->s_comm) cudaStreamDestroy(h->s_comm);
  if (h->s_in) cudaStreamDestroy(h->s_in);
  if (h->s_out) cudaStreamDestroy(h->s_out);
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

extern "C" rub_status rub_rx_set_path(rub_rx *h, uint32_t path) {
  if (!h || path > RUB_PATH_FUSED) return RUB_ERR_INVALID_ARG;
  h->path = path;
  return RUB_OK;
}
extern "C" rub_status rub_rx_set_host_chunk(rub_rx *h, uint32_t frames) {
  if (!h) return RUB_ERR_INVALID_ARG;
  h->host_chunk = frames;
  return RUB_OK;
}
extern "C" uint32_t rub_rx_get_path(const rub_rx *h) { return h ? h->last_path : 0; }
extern "C" uint64_t rub_rx_launch_count(const rub_rx *h) { return h ? h->launches : 0; }
extern "C" const char *rub_rx_last_kernel(const rub_rx *h) { return h ? h->last_kernel : ""; }
extern "C" uint64_t *rub_rx_device_counters(rub_rx *h) { return h ? h->d_counters : nullptr; }
extern "C" rub_status rub_rx_reset_counters(rub_rx *h) {
  if (!h) return RUB_ERR_INVALID_ARG;
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaMemsetAsync(h->d_counters, 0, sizeof(uint64_t) * 4 * 8, h->stream));
  return RUB_OK;
}
extern "C" rub_status rub_rx_read_counters(rub_rx *h, uint64_t *host_out) {
  if (!h || !host_out) return RUB_ERR_INVALID_ARG;
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaMemcpyAsync(host_out, h->d_counters, sizeof(uint64_t) * 4 * h->h.N, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return RUB_OK;
}
extern "C" rub_status rub_rx_sync(rub_rx *h) {
  if (!h) return RUB_ERR_INVALID_ARG;
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return RUB_OK;
}
extern "C" rub_status rub_rx_last_timing(rub_rx *h, float *total_ms, float *dominant_ms) {
  if (!h || !h->timed) { set_error("no timed batch"); return RUB_ERR_INVALID_ARG; }
  if (total_ms) CUDA_TRY(cudaEventElapsedTime(total_ms, h->ev[0], h->ev[1]));
  if (dominant_ms) CUDA_TRY(cudaEventElapsedTime(dominant_ms, h->ev[2], h->ev[3]));
  return RUB_OK;
}

// Algorithmic (compulsory) HBM bytes of one batch, SURVEY.md 8d:
//   8*N*(M+cp)*(T+D)                       input samples incl. training symbols
// + D*N*Mo*(8[eq] + 4q[llr] + q/8[bits] + 1[rx_data] + 1[tx_data])   per-symbol outputs / ref
// + 8*M*N^2 [G]  + 32*N [counters]
extern "C" uint64_t rub_rx_algorithmic_bytes(const rub_rx *h, uint32_t n_frames, uint32_t out_mask, int with_tx) {
  if (!h) return 0;
  const HostCfg &c = h->h;
  uint64_t per = 8ull * c.N * c.L * (c.T + c.D);
  const uint64_t syms = (uint64_t)c.D * c.N * c.Mo;
  if (out_mask & RUB_OUT_EQ) per += syms * 8;
  if (out_mask & RUB_OUT_LLR) per += syms * 4 * c.q;
  if (out_mask & RUB_OUT_BITS) per += (uint64_t)c.D * c.N * c.row_bytes;
  if (out_mask & RUB_OUT_RXDATA) per += syms;
  if (with_tx) per += syms;
  if (out_mask & RUB_OUT_G) per += 8ull * c.M * c.N * c.N;
  return per * n_frames + (with_tx ? 32ull * c.N : 0);
}

// ---------------------------------------------------------------- staged dispatch -----
template <int LOG2M>
static void launch_fft(const ChainArgs &a, cudaStream_t st, cudaError_t *err) {
  constexpr int NT = FftPlan<LOG2M>::NT;
  constexpr int F = NT >= 128 ? 1 : 128 / NT;
  const size_t smem = (size_t)F * 2 * fft_padded_size(1 << LOG2M) * sizeof(cf);
  // the attribute is per device and per function: set once per device, safely across host threads
  static std::atomic<bool> attr_done[64];
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_done[dev & 63].load(std::memory_order_acquire)) {
    *err = cudaFuncSetAttribute(k_fft_staged<LOG2M, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (*err != cudaSuccess) return;
    attr_done[dev & 63].store(true, std::memory_order_release);
  }
  const long long total = (long long)a.n_frames * (a.T + a.D) * a.N;
  const unsigned grid = (unsigned)((total + F - 1) / F);
  k_fft_staged<LOG2M, F><<<grid, NT * F, smem, st>>>(a);
}
template <int N>
static void launch_weights_detect(const ChainArgs &a, const rub_rx *h, cudaStream_t st, cudaEvent_t e0, cudaEvent_t e1,
                                  int comb_fused, int write_G) {
  const long long tw = (long long)a.n_frames * a.M;
  if (comb_fused) {
    // comb LS + weights in one kernel (the pilots of a 128-carrier tile in shared memory)
    const size_t sm = (size_t)N * N * (128 / a.P + 2) * sizeof(cf);
    k_lscomb_weights<N><<<(unsigned)(a.n_frames * ((a.M + 127) / 128)), 128, sm, st>>>(a, h->wm, write_G);
  } else {
    k_weights<N><<<(unsigned)((tw + 127) / 128), 128, 0, st>>>(a, h->wm);
  }
  if (e0) cudaEventRecord(e0, st);
  if (detect_lean_launch(a, h->lut, st)) {
    if (e1) cudaEventRecord(e1, st);
    const_cast<rub_rx *>(h)->last_kernel = "k_detect_lean";
    return;
  }
  const_cast<rub_rx *>(h)->last_kernel = "k_detect";
  const int groups = (a.Mo + 7) / 8;
  dim3 grid((unsigned)((long long)a.n_frames * a.D * N), (unsigned)((groups + 127) / 128));
  k_detect<N><<<grid, 128, 0, st>>>(a, h->lut);
  if (e1) cudaEventRecord(e1, st);
}

static rub_status run_staged(rub_rx *h, ChainArgs a, const rub_rx_io *io, uint32_t n_frames, bool timed) {
  const HostCfg &c = h->h;
  // per-frame scratch: Y, (G), W, gain, isig
  const size_t yb = (size_t)(c.T + c.D) * c.N * c.M * sizeof(cf);
  const size_t gb = (size_t)c.N * c.N * c.M * sizeof(cf);
  const size_t fb = (size_t)c.N * c.M * sizeof(float);
  const bool user_G = (io->out_mask & RUB_OUT_G) && io->G;
  const size_t per_frame = yb + (user_G ? 0 : gb) + gb + 2 * fb;
  const size_t budget = (size_t)6 << 30;
  uint32_t chunk = (uint32_t)std::max<size_t>(1, std::min<size_t>(n_frames, budget / per_frame));
  const size_t need = per_frame * chunk;
  if (need > h->scratch_bytes) {
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    cudaFree(h->d_scratch);
    h->d_scratch = nullptr;
    h->scratch_bytes = 0;
    CUDA_TRY(cudaMalloc(&h->d_scratch, need));
    h->scratch_bytes = need;
  }
  const size_t per_sym = (size_t)c.N * c.D * c.Mo;
  for (uint32_t f0 = 0; f0 < n_frames; f0 += chunk) {
    const uint32_t nf = std::min(chunk, n_frames - f0);
    ChainArgs b = a;
    b.n_frames = (int)nf;
    b.iq = a.iq + (size_t)f0 * a.frame_stride;
    if (a.timing) b.timing = a.timing + (size_t)f0 * c.N * c.T;
    if (a.payload_start) b.payload_start = a.payload_start + f0;
    unsigned char *p = (unsigned char *)h->d_scratch;
    b.Y = (cf *)p; p += yb * nf;
    if (user_G) b.G = reinterpret_cast<cf *>(io->G) + (size_t)f0 * c.N * c.N * c.M;
    else { b.G = (cf *)p; p += gb * nf; }
    b.W = (cf *)p; p += gb * nf;
    b.gain = (float *)p; p += fb * nf;
    b.isig = (float *)p;
    if (a.eq) b.eq = a.eq + f0 * per_sym;
    if (a.llr) b.llr = a.llr + f0 * per_sym * c.q;
    if (a.bits) b.bits = a.bits + (size_t)f0 * c.N * c.D * c.row_bytes;
    if (a.rx_data) b.rx_data = a.rx_data + f0 * per_sym;
    if (a.tx_data) b.tx_data = a.tx_data + f0 * per_sym;
    b.wrec = detect_lean_records(b) ? 1 : 0;  // k_detect_lean reads W / gain / isig as task records
    cudaError_t ferr = cudaSuccess;
    switch (c.log2M) {
      case 6: launch_fft<6>(b, h->stream, &ferr); break;
      case 7: launch_fft<7>(b, h->stream, &ferr); break;
      case 8: launch_fft<8>(b, h->stream, &ferr); break;
      case 9: launch_fft<9>(b, h->stream, &ferr); break;
      case 10: launch_fft<10>(b, h->stream, &ferr); break;
      case 11: launch_fft<11>(b, h->stream, &ferr); break;
      case 12: launch_fft<12>(b, h->stream, &ferr); break;
    }
    CUDA_TRY(ferr);
    const long long tg = (long long)nf * c.N * c.N * c.M;
    // comb estimator: LS and weights share one kernel unless the pilot spacing is so tight that a tile's pilots
    // do not fit the static bound of its shared-memory table (P < N never happens for a valid comb)
    const int comb_fused = c.c.estimator == RUB_EST_LS_COMB_INTERP && c.P >= c.N && c.P <= 128;
    if (comb_fused) {
    } else if (c.c.estimator == RUB_EST_LS_COMB_INTERP) {
      const long long tp = tg / c.P;  // one thread per pilot
      k_ls_comb<<<(unsigned)((tp + 255) / 256), 256, 0, h->stream>>>(b);
    }
    else k_ls_fullband<<<(unsigned)((tg + 255) / 256), 256, 0, h->stream>>>(b);
    const bool last = f0 + nf >= n_frames;
    cudaEvent_t e0 = (timed && last) ? h->ev[2] : nullptr, e1 = (timed && last) ? h->ev[3] : nullptr;
    switch (c.N) {
      case 1: launch_weights_detect<1>(b, h, h->stream, e0, e1, comb_fused, user_G ? 1 : 0); break;
      case 2: launch_weights_detect<2>(b, h, h->stream, e0, e1, comb_fused, user_G ? 1 : 0); break;
      case 3: launch_weights_detect<3>(b, h, h->stream, e0, e1, comb_fused, user_G ? 1 : 0); break;
      case 4: launch_weights_detect<4>(b, h, h->stream, e0, e1, comb_fused, user_G ? 1 : 0); break;
      case 5: launch_weights_detect<5>(b, h, h->stream, e0, e1, comb_fused, user_G ? 1 : 0); break;
      case 6: launch_weights_detect<6>(b, h, h->stream, e0, e1, comb_fused, user_G ? 1 : 0); break;
      case 7: launch_weights_detect<7>(b, h, h->stream, e0, e1, comb_fused, user_G ? 1 : 0); break;
      case 8: launch_weights_detect<8>(b, h, h->stream, e0, e1, comb_fused, user_G ? 1 : 0); break;
    }
    h->launches += 4;
    CUDA_TRY(cudaGetLastError());
  }
  h->last_path = RUB_PATH_STAGED;
  return RUB_OK;
}

static rub_status run_fused(rub_rx *h, const ChainArgs &a, uint32_t n_frames, bool timed) {
  const HostCfg &c = h->h;
  if (!h->fused_ready) {
    int grid = 0;
    size_t smem = 0;
    rub_status st = fused_prepare_dispatch(h, &smem, &grid);
    if (st) { h->fused_unfit = st == RUB_ERR_UNSUPPORTED; return st; }
    h->fused_grid = grid;
    h->fused_smem = smem;
    // per-CTA scratch: G -> W in place, gain, isig (re-read D times per frame: L2 resident)
    // (the kernels that fetch W by TMA keep W, gain and isig together as task records in the first buffer)
    CUDA_TRY(cudaMalloc(&h->d_fW, (size_t)grid * (c.N * c.N * c.M * sizeof(cf) + 2 * c.N * c.M * sizeof(float))));
    CUDA_TRY(cudaMalloc(&h->d_fG, (size_t)grid * 2 * c.N * c.M * sizeof(float)));
    if (h->ws) {
      // the warp-specialised kernel estimates frame f+1 while frame f is detected
      CUDA_TRY(cudaMalloc(&h->d_fAcc, (size_t)grid * c.N * c.N * c.M * sizeof(cf)));
      std::vector<unsigned char> s8(ws_sign_bytes(c.log2M, c.N, c.nac));
      ws_pack_signs(c.log2M, c.N, c.nac, h->sgn_host.data(), s8.data());
      CUDA_TRY(cudaMalloc(&h->d_sgn8, s8.size()));
      CUDA_TRY(cudaMemcpy(h->d_sgn8, s8.data(), s8.size(), cudaMemcpyHostToDevice));
    }
    h->fused_ready = true;
  }
  const size_t smem = h->fused_smem;
  FusedArgs fa;
  fa.a = a;
  fa.a.n_frames = (int)n_frames;
  fa.scratchW = h->d_fW;
  fa.scratchG = h->d_fG;
  fa.scratchAcc = h->d_fAcc;
  fa.sgn8 = h->d_sgn8;
  fa.llr_stage_bytes = 256 * (int)c.q;
  fa.wm = h->wm;
  const int grid = (int)std::min<uint32_t>((uint32_t)h->fused_grid, n_frames);
  if (timed) cudaEventRecord(h->ev[2], h->stream);
  fused_launch_dispatch(h, grid, smem, fa);
  if (timed) cudaEventRecord(h->ev[3], h->stream);
  h->launches += 1;
  CUDA_TRY(cudaGetLastError());
  h->last_path = RUB_PATH_FUSED;
  h->last_kernel = h->ws ? "k_rx_ws" : "k_rx_fused";
  return RUB_OK;
}

static rub_status process_device(rub_rx *h, const rub_rx_io *io, uint32_t n_frames, bool timed, bool shared_rows = false) {
  if (!h || !io) { set_error("process_batch: NULL argument"); return RUB_ERR_INVALID_ARG; }
  if (n_frames == 0) return RUB_OK;  // an empty batch is a no-op (no launch, counters untouched)
  if (!io->iq) { set_error("process_batch: iq is NULL"); return RUB_ERR_INVALID_ARG; }
  const HostCfg &c = h->h;
  CUDA_TRY(cudaSetDevice(h->device));
  const uint64_t rx_stride = io->layout.rx_stride ? io->layout.rx_stride : (uint64_t)(c.T + c.D) * c.L + io->layout.first_sample;
  const uint64_t frame_stride = io->layout.frame_stride ? io->layout.frame_stride : rx_stride * c.N;
  ChainArgs a;
  memset(&a, 0, sizeof(a));
  a.iq = reinterpret_cast<const cf *>(io->iq);
  // shared_rows: every frame addresses the same capture rows, the timing tables hold absolute offsets
  a.frame_stride = shared_rows ? 0 : frame_stride; a.rx_stride = rx_stride; a.first_sample = io->layout.first_sample;
  a.timing = io->timing; a.payload_start = io->payload_start;
  a.M = c.M; a.cp = c.cp; a.L = c.L; a.N = c.N; a.nac = c.nac; a.D = c.D; a.T = c.T; a.Mo = c.Mo; a.q = c.q; a.P = c.P;
  a.n_frames = (int)n_frames; a.row_bytes = c.row_bytes;
  a.dn = c.dn; a.s_ls = c.s_ls; a.flags = c.c.flags; a.estimator = c.c.estimator;
  a.tw = h->d_tw; a.occ = h->d_occ; a.sgn = h->d_sgn; a.scnull = h->d_null;
  a.eq = (io->out_mask & RUB_OUT_EQ) ? reinterpret_cast<cf *>(io->eq) : nullptr;
  a.llr = (io->out_mask & RUB_OUT_LLR) ? io->llr : nullptr;
  a.bits = (io->out_mask & RUB_OUT_BITS) ? io->bits : nullptr;
  a.rx_data = (io->out_mask & RUB_OUT_RXDATA) ? io->rx_data : nullptr;
  a.G = (io->out_mask & RUB_OUT_G) ? reinterpret_cast<cf *>(io->G) : nullptr;
  a.tx_data = io->tx_data;
  a.counters = io->counters ? (unsigned long long *)io->counters : (unsigned long long *)h->d_counters;
  if (((io->out_mask & RUB_OUT_EQ) && !io->eq) || ((io->out_mask & RUB_OUT_LLR) && !io->llr) ||
      ((io->out_mask & RUB_OUT_BITS) && !io->bits) || ((io->out_mask & RUB_OUT_RXDATA) && !io->rx_data) ||
      ((io->out_mask & RUB_OUT_G) && !io->G)) {
    set_error("process_batch: out_mask selects an output whose pointer is NULL");
    return RUB_ERR_INVALID_ARG;
  }
  const bool can_fuse = fused_eligible(h, io, frame_stride, rx_stride);
  const bool req_fused = h->path == RUB_PATH_FUSED;
  if (req_fused && (!can_fuse || h->fused_unfit)) {
    set_error("fused path requested but the configuration / buffers are not eligible");
    return RUB_ERR_UNSUPPORTED;
  }
  const bool use_fused = req_fused || (h->path == RUB_PATH_AUTO && can_fuse);
  if (timed) cudaEventRecord(h->ev[0], h->stream);
  rub_status st = (use_fused && !h->fused_unfit) ? run_fused(h, a, n_frames, timed) : run_staged(h, a, io, n_frames, timed);
  // AUTO: a configuration whose fused kernel does not fit the SM (found at the first launch) is served staged
  if (st == RUB_ERR_UNSUPPORTED && h->fused_unfit && !req_fused) st = run_staged(h, a, io, n_frames, timed);
  if (timed) cudaEventRecord(h->ev[1], h->stream);
  h->timed = timed && st == RUB_OK;
  return st;
}

extern "C" rub_status rub_rx_process_batch(rub_rx *h, const rub_rx_io *io, uint32_t n_frames) {
  return process_device(h, io, n_frames, true);
}

// ---------------------------------------------------------------- host-buffer path ----
// chunked 3-stream pipeline: H2D (s_in) -> chain (handle stream) -> D2H (s_out), two slots
extern "C" rub_status rub_rx_process_batch_host(rub_rx *h, const rub_rx_io *io, uint32_t n_frames) {
  if (!h || !io || !io->iq) { set_error("process_batch_host: NULL argument"); return RUB_ERR_INVALID_ARG; }
  const HostCfg &c = h->h;
  CUDA_TRY(cudaSetDevice(h->device));
  if (!h->s_in) {
    CUDA_TRY(cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking));
    for (auto &e : h->pev) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  const uint64_t rx_stride = io->layout.rx_stride ? io->layout.rx_stride : (uint64_t)(c.T + c.D) * c.L + io->layout.first_sample;
  const uint64_t frame_stride = io->layout.frame_stride ? io->layout.frame_stride : rx_stride * c.N;
  const size_t per_sym = (size_t)c.N * c.D * c.Mo;
  const size_t in_b = (size_t)frame_stride * sizeof(cf);
  const size_t eq_b = (io->out_mask & RUB_OUT_EQ) ? per_sym * 8 : 0;
  const size_t llr_b = (io->out_mask & RUB_OUT_LLR) ? per_sym * 4 * c.q : 0;
  const size_t bits_b = (io->out_mask & RUB_OUT_BITS) ? (size_t)c.N * c.D * c.row_bytes : 0;
  const size_t rxd_b = (io->out_mask & RUB_OUT_RXDATA) ? per_sym : 0;
  const size_t g_b = (io->out_mask & RUB_OUT_G) ? (size_t)c.N * c.N * c.M * 8 : 0;
  const size_t tx_b = io->tx_data ? per_sym : 0;
  const size_t tim_b = io->timing ? (size_t)c.N * c.T * 4 : 0;
  const size_t ps_b = io->payload_start ? 4 : 0;
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  const size_t per_frame = in_b + eq_b + llr_b + bits_b + rxd_b + g_b + tx_b + tim_b + ps_b;
  // chunk sized to ~256 MB of traffic so copies and compute overlap
  uint32_t chunk = (uint32_t)std::max<size_t>(1, std::min<size_t>(n_frames, ((size_t)256 << 20) / std::max<size_t>(1, per_frame)));
  if (h->host_chunk) chunk = std::min<uint32_t>(h->host_chunk, std::max<uint32_t>(1, n_frames));
  const size_t slot_b = al(in_b * chunk) + al(eq_b * chunk) + al(llr_b * chunk) + al(bits_b * chunk) +
                        al(rxd_b * chunk) + al(g_b * chunk) + al(tx_b * chunk) + al(tim_b * chunk) + al(ps_b * chunk) + 4096;
  if (2 * slot_b > h->pipe_bytes) {
    CUDA_TRY(cudaDeviceSynchronize());
    cudaFree(h->d_pipe);
    h->d_pipe = nullptr;
    h->pipe_bytes = 0;
    CUDA_TRY(cudaMalloc(&h->d_pipe, 2 * slot_b));
    h->pipe_bytes = 2 * slot_b;
  }
  struct Slot { unsigned char *in, *eq, *llr, *bits, *rxd, *g, *tx, *tim, *ps; } sl[2];
  for (int s = 0; s < 2; s++) {
    unsigned char *p = (unsigned char *)h->d_pipe + (size_t)s * slot_b;
    sl[s].in = p; p += al(in_b * chunk);
    sl[s].eq = p; p += al(eq_b * chunk);
    sl[s].llr = p; p += al(llr_b * chunk);
    sl[s].bits = p; p += al(bits_b * chunk);
    sl[s].rxd = p; p += al(rxd_b * chunk);
    sl[s].g = p; p += al(g_b * chunk);
    sl[s].tx = p; p += al(tx_b * chunk);
    sl[s].tim = p; p += al(tim_b * chunk);
    sl[s].ps = p;
  }
  // pev[0..1] in-done, pev[2..3] compute-done, pev[4..5] out-done (per slot)
  uint32_t ci = 0;
  const uint32_t saved_launch_path = h->last_path;
  (void)saved_launch_path;
  for (uint32_t f0 = 0; f0 < n_frames; f0 += chunk, ci++) {
    const uint32_t nf = std::min(chunk, n_frames - f0);
    const int s = ci & 1;
    // slot input free once its previous compute is done
    if (ci >= 2) CUDA_TRY(cudaStreamWaitEvent(h->s_in, h->pev[2 + s], 0));
    CUDA_TRY(cudaMemcpyAsync(sl[s].in, reinterpret_cast<const cf *>(io->iq) + (size_t)f0 * frame_stride, in_b * nf, cudaMemcpyHostToDevice, h->s_in));
    if (tx_b) CUDA_TRY(cudaMemcpyAsync(sl[s].tx, io->tx_data + f0 * per_sym, tx_b * nf, cudaMemcpyHostToDevice, h->s_in));
    if (tim_b) CUDA_TRY(cudaMemcpyAsync(sl[s].tim, io->timing + (size_t)f0 * c.N * c.T, tim_b * nf, cudaMemcpyHostToDevice, h->s_in));
    if (ps_b) CUDA_TRY(cudaMemcpyAsync(sl[s].ps, io->payload_start + f0, ps_b * nf, cudaMemcpyHostToDevice, h->s_in));
    CUDA_TRY(cudaEventRecord(h->pev[0 + s], h->s_in));
    // compute: needs the input, and the slot's previous outputs copied out
    CUDA_TRY(cudaStreamWaitEvent(h->stream, h->pev[0 + s], 0));
    if (ci >= 2) CUDA_TRY(cudaStreamWaitEvent(h->stream, h->pev[4 + s], 0));
    rub_rx_io d = *io;
    d.iq = (const float *)sl[s].in;
    d.layout.frame_stride = frame_stride; d.layout.rx_stride = rx_stride;
    d.tx_data = tx_b ? sl[s].tx : nullptr;
    d.timing = tim_b ? (const int32_t *)sl[s].tim : nullptr;
    d.payload_start = ps_b ? (const int32_t *)sl[s].ps : nullptr;
    d.eq = (float *)sl[s].eq; d.llr = (float *)sl[s].llr; d.bits = sl[s].bits; d.rx_data = sl[s].rxd; d.G = (float *)sl[s].g;
    d.counters = nullptr;  // accumulate in the handle's device counters
    rub_status st = process_device(h, &d, nf, false);
    if (st) return st;
    CUDA_TRY(cudaEventRecord(h->pev[2 + s], h->stream));
    // copy out
    CUDA_TRY(cudaStreamWaitEvent(h->s_out, h->pev[2 + s], 0));
    if (eq_b) CUDA_TRY(cudaMemcpyAsync(io->eq + (size_t)f0 * per_sym * 2, sl[s].eq, eq_b * nf, cudaMemcpyDeviceToHost, h->s_out));
    if (llr_b) CUDA_TRY(cudaMemcpyAsync(io->llr + (size_t)f0 * per_sym * c.q, sl[s].llr, llr_b * nf, cudaMemcpyDeviceToHost, h->s_out));
    if (bits_b) CUDA_TRY(cudaMemcpyAsync(io->bits + (size_t)f0 * c.N * c.D * c.row_bytes, sl[s].bits, bits_b * nf, cudaMemcpyDeviceToHost, h->s_out));
    if (rxd_b) CUDA_TRY(cudaMemcpyAsync(io->rx_data + (size_t)f0 * per_sym, sl[s].rxd, rxd_b * nf, cudaMemcpyDeviceToHost, h->s_out));
    if (g_b) CUDA_TRY(cudaMemcpyAsync(io->G + (size_t)f0 * c.N * c.N * c.M * 2, sl[s].g, g_b * nf, cudaMemcpyDeviceToHost, h->s_out));
    CUDA_TRY(cudaEventRecord(h->pev[4 + s], h->s_out));
  }
  CUDA_TRY(cudaStreamSynchronize(h->s_out));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  if (io->counters) CUDA_TRY(cudaMemcpy(io->counters, h->d_counters, sizeof(uint64_t) * 4 * c.N, cudaMemcpyDeviceToHost));
  return RUB_OK;
}

// ---------------------------------------------------------------- multi-GPU ------------
extern "C" rub_status rub_comm_get_unique_id(uint8_t id[RUB_NCCL_UNIQUE_ID_BYTES]) {
  rub_status st = nccl_load();
  if (st) return st;
  nccl_uid u;
  int r = g_nccl.GetUniqueId(&u);
  if (r) { set_error("ncclGetUniqueId: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"); return RUB_ERR_NCCL; }
  memcpy(id, u.internal, RUB_NCCL_UNIQUE_ID_BYTES);
  return RUB_OK;
}
extern "C" rub_status rub_comm_init(rub_rx *h, const uint8_t id[RUB_NCCL_UNIQUE_ID_BYTES], int rank, int world) {
  if (!h || !id || world < 1 || rank < 0 || rank >= world) return RUB_ERR_INVALID_ARG;
  rub_status st = nccl_load();
  if (st) return st;
  CUDA_TRY(cudaSetDevice(h->device));
  nccl_uid u;
  memcpy(u.internal, id, RUB_NCCL_UNIQUE_ID_BYTES);
  int r = g_nccl.CommInitRank(&h->comm, world, u, rank);
  if (r) { set_error("ncclCommInitRank: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"); h->comm = nullptr; return RUB_ERR_NCCL; }
  h->rank = rank;
  h->world = world;
  return RUB_OK;
}
// Global error counters = sum over ranks of every rank's cumulative local counters.  Idempotent: the
// handle's own counters are only read.  A snapshot of them is taken on the handle's stream (a 256-byte
// device copy ordered after the batches issued so far), and ncclAllReduce(sum, uint64) turns it into
// the global counters on a side stream, so the next batch never waits for the collective.  Snapshot and
// result are double buffered; rub_rx_read_counters_global returns the latest reduction.
extern "C" rub_status rub_allreduce_counters(rub_rx *h) {
  if (!h) return RUB_ERR_INVALID_ARG;
  if (!h->comm && h->world != 1) { set_error("rub_comm_init not called"); return RUB_ERR_NCCL; }
  CUDA_TRY(cudaSetDevice(h->device));
  if (!h->s_comm) {
    CUDA_TRY(cudaStreamCreateWithFlags(&h->s_comm, cudaStreamNonBlocking));
    CUDA_TRY(cudaMalloc(&h->d_snap, sizeof(uint64_t) * 4 * 32));
    CUDA_TRY(cudaMemset(h->d_snap, 0, sizeof(uint64_t) * 4 * 32));
    for (auto &e : h->cev) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  const uint32_t i = h->n_reduce & 1;
  uint64_t *snap = h->d_snap + 32 * i, *glob = h->d_snap + 64 + 32 * i;
  // the reduction issued two calls ago read this snapshot slot
  if (h->n_reduce >= 2) CUDA_TRY(cudaStreamWaitEvent(h->stream, h->cev[2 + i], 0));
  CUDA_TRY(cudaMemcpyAsync(snap, h->d_counters, sizeof(uint64_t) * 4 * h->h.N, cudaMemcpyDeviceToDevice, h->stream));
  CUDA_TRY(cudaEventRecord(h->cev[i], h->stream));
  CUDA_TRY(cudaStreamWaitEvent(h->s_comm, h->cev[i], 0));
  if (h->comm) {
    int r = g_nccl.AllReduce(snap, glob, 4 * h->h.N, /*ncclUint64*/ 5, /*ncclSum*/ 0, h->comm, h->s_comm);
    if (r) { set_error("ncclAllReduce: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"); return RUB_ERR_NCCL; }
    h->launches += 1;
  } else {
    CUDA_TRY(cudaMemcpyAsync(glob, snap, sizeof(uint64_t) * 4 * h->h.N, cudaMemcpyDeviceToDevice, h->s_comm));
  }
  CUDA_TRY(cudaEventRecord(h->cev[2 + i], h->s_comm));
  h->n_reduce++;
  return RUB_OK;
}
extern "C" rub_status rub_rx_read_counters_global(rub_rx *h, uint64_t *host_out) {
  if (!h || !host_out) return RUB_ERR_INVALID_ARG;
  if (!h->n_reduce) { set_error("read_counters_global: rub_allreduce_counters has not been called"); return RUB_ERR_INVALID_ARG; }
  CUDA_TRY(cudaSetDevice(h->device));
  const uint32_t i = (h->n_reduce - 1) & 1;
  CUDA_TRY(cudaMemcpyAsync(host_out, h->d_snap + 64 + 32 * i, sizeof(uint64_t) * 4 * h->h.N, cudaMemcpyDeviceToHost, h->s_comm));
  CUDA_TRY(cudaStreamSynchronize(h->s_comm));
  return RUB_OK;
}
extern "C" rub_status rub_comm_destroy(rub_rx *h) {
  if (!h) return RUB_ERR_INVALID_ARG;
  if (h->s_comm) cudaStreamSynchronize(h->s_comm);
  if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
  h->comm = nullptr;
  return RUB_OK;
}

// ---------------------------------------------------------------- synchronisation -----
// grow-only device scratch shared by the synchronisation calls (they are synchronous)
static rub_status sync_scratch(rub_rx *h, size_t bytes, void **out) {
  if (bytes > h->sync_bytes) {
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    cudaFree(h->d_sync);
    h->d_sync = nullptr;
    h->sync_bytes = 0;
    const size_t want = std::max(bytes, (size_t)1 << 20);
    if (cudaMalloc(&h->d_sync, want) != cudaSuccess) { set_error("out of device memory (%zu B of sync scratch)", want); return RUB_ERR_NOMEM; }
    h->sync_bytes = want;
  }
  *out = h->d_sync;
  return RUB_OK;
}
extern "C" rub_status rub_rx_set_S0(rub_rx *h, const float *s0) {
  if (!h || !s0) return RUB_ERR_INVALID_ARG;
  CUDA_TRY(cudaSetDevice(h->device));
  if (!h->d_s0) CUDA_TRY(cudaMalloc(&h->d_s0, sizeof(cf) * h->h.M));
  CUDA_TRY(cudaMemcpy(h->d_s0, s0, sizeof(cf) * h->h.M, cudaMemcpyHostToDevice));
  {
    // number of carriers S0 occupies (every other enabled one, framing.cc:1075-1110): the bins of FFT(s0) that are not empty
    const HostCfg &c = h->h;
    std::vector<cf> master, packed, X(c.M);
    build_twiddles(c.log2M, master, packed);
    host_fft_forward(c.log2M, reinterpret_cast<const cf *>(s0), X.data(), packed.data());
    float mx = 0.f;
    for (uint32_t k = 0; k < c.M; k++) mx = std::max(mx, X[k].x * X[k].x + X[k].y * X[k].y);
    uint32_t m_s0 = 0;
    for (uint32_t k = 0; k < c.M; k++) m_s0 += (X[k].x * X[k].x + X[k].y * X[k].y) > 0.25f * mx;
    h->s0_corr_scale = (float)m_s0 / ((float)c.M * (float)c.M);
  }
  return RUB_OK;
}

// Schmidl & Cox metric of n samples of one stream already on the device (row stride for batched streams)
static cudaError_t launch_sc_metric(rub_rx *h, uint32_t mode, const cf *dx, uint64_t n, uint64_t row_stride, uint32_t streams, float *dy) {
  const int M = (int)h->h.M;
  if (mode == RUB_SYNC_SCAN) {
    constexpr int PER = 8, TILE = 256 * PER;
    const size_t smem = (size_t)(M + M / 2 + TILE) * sizeof(cf);
    cudaError_t e = cudaFuncSetAttribute(k_sc_metric_scan<PER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dim3 grid((unsigned)((n + TILE - 1) / TILE), streams);
    k_sc_metric_scan<PER><<<grid, 256, smem, h->stream>>>(dx, n, row_stride, M, dy);
    h->launches += 1;
  } else {
    const size_t smem = (size_t)(M + M / 2 + 256) * sizeof(cf);
    cudaError_t e = cudaFuncSetAttribute(k_sc_metric, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    for (uint32_t s = 0; s < streams; s++)
      k_sc_metric<<<(unsigned)((n + 255) / 256), 256, smem, h->stream>>>(dx + (size_t)s * row_stride, n, M, dy + (size_t)s * row_stride);
    h->launches += streams;
  }
  return cudaGetLastError();
}

// framesync::execute_sc_sync(x, stream), mimo/framing.cc:626-637.  mode RUB_SYNC_FIR: the reference's two FIR dot
// products per sample in liquid's order (bit-identical metric, O(M) per sample); RUB_SYNC_SCAN: sliding sums
// (O(1) per sample; last-bit differences)
extern "C" rub_status rub_rx_sc_metric_ex(rub_rx *h, const float *x, uint64_t n, float *y, uint32_t mode) {
  if (!h || !x || !y) { set_error("sc_metric: NULL argument"); return RUB_ERR_INVALID_ARG; }
  if (mode > RUB_SYNC_SCAN) { set_error("sc_metric: unknown mode"); return RUB_ERR_INVALID_ARG; }
  if (n == 0) return RUB_OK;
  CUDA_TRY(cudaSetDevice(h->device));
  void *scr = nullptr;
  const size_t xb = (sizeof(cf) * n + 255) & ~(size_t)255;
  rub_status st = sync_scratch(h, xb + sizeof(float) * n, &scr);
  if (st) return st;
  cf *dx = reinterpret_cast<cf *>(scr);
  float *dy = reinterpret_cast<float *>((unsigned char *)scr + xb);
  cudaError_t e = cudaMemcpyAsync(dx, x, sizeof(cf) * n, cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) e = launch_sc_metric(h, mode, dx, n, n, 1, dy);
  if (e == cudaSuccess) e = cudaMemcpyAsync(y, dy, sizeof(float) * n, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (e != cudaSuccess) { set_error("sc_metric: %s", cudaGetErrorString(e)); st = RUB_ERR_CUDA; }
  return st;
}
extern "C" rub_status rub_rx_sc_metric(rub_rx *h, const float *x, uint64_t n, float *y) {
  return rub_rx_sc_metric_ex(h, x, n, y, RUB_SYNC_FIR);
}
extern "C" rub_status rub_rx_set_sync_mode(rub_rx *h, uint32_t mode) {
  if (!h || mode > RUB_SYNC_SCAN) return RUB_ERR_INVALID_ARG;
  h->sync_mode = mode;
  return RUB_OK;
}
extern "C" rub_status rub_rx_set_debug_dir(rub_rx *h, const char *dir) {
  if (!h) return RUB_ERR_INVALID_ARG;
  h->debug_dir = dir ? dir : "";
  return RUB_OK;
}

// timing search of estimate_channel, mimo/framing.cc:702-744
extern "C" rub_status rub_rx_timing_search(rub_rx *h, const float *window, uint64_t wlen, int32_t *corr_indices,
                                           int32_t *s0_corr_index) {
  if (!h || !window || !corr_indices) { set_error("timing_search: NULL argument"); return RUB_ERR_INVALID_ARG; }
  const HostCfg &c = h->h;
  const uint32_t max_ac = c.nac * c.N;
  // the last candidate window ends at (L-1) + L*max_ac + M (framing.cc:724-727)
  if (wlen < (uint64_t)c.L * (max_ac + 1) + c.M) { set_error("timing_search: window shorter than the preamble"); return RUB_ERR_INVALID_ARG; }
  if (s0_corr_index && !h->d_s0) { set_error("timing_search: S0 index requested but rub_rx_set_S0 not called"); return RUB_ERR_INVALID_ARG; }
  CUDA_TRY(cudaSetDevice(h->device));
  void *scr = nullptr;
  const size_t wb = (sizeof(cf) * wlen * c.N + 255) & ~(size_t)255;
  const uint32_t nkeys = c.N * (max_ac + 1);
  rub_status st = sync_scratch(h, wb + sizeof(unsigned long long) * nkeys, &scr);
  if (st) return st;
  cf *dw = reinterpret_cast<cf *>(scr);
  unsigned long long *dk = reinterpret_cast<unsigned long long *>((unsigned char *)scr + wb);
  std::vector<unsigned long long> keys(nkeys);
  const size_t smem = (size_t)(2 * c.M + 256) * sizeof(cf);
  cudaError_t e = cudaFuncSetAttribute(k_timing_search, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dw, window, sizeof(cf) * wlen * c.N, cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(dk, 0, sizeof(unsigned long long) * nkeys, h->stream);
  if (e == cudaSuccess) {
    dim3 grid((c.L + 255) / 256, nkeys);
    k_timing_search<<<grid, 256, smem, h->stream>>>(dw, wlen, wlen, nullptr, h->d_s1, s0_corr_index ? h->d_s0 : nullptr, (int)c.M,
                                                    (int)c.L, (int)c.N, (int)c.nac, dk, nullptr, 0.f);
    h->launches += 1;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(keys.data(), dk, sizeof(unsigned long long) * nkeys, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (e != cudaSuccess) { set_error("timing_search: %s", cudaGetErrorString(e)); return RUB_ERR_CUDA; }
  // decode the winners: index 0 when no correlation rose above the reference's initial max = 0
  for (uint32_t r = 0; r < c.N; r++)
    for (uint32_t slot = 0; slot <= max_ac; slot++) {
      const unsigned long long k = keys[r * (max_ac + 1) + slot];
      const bool hit = (k >> 32) != 0;
      const uint32_t i = hit ? 0xffffffffu - (uint32_t)(k & 0xffffffffu) : 0u;
      if (slot == 0) { if (s0_corr_index) s0_corr_index[r] = (int32_t)i; }
      else corr_indices[r * max_ac + slot - 1] = hit ? (int32_t)(c.L * slot + i) : 0;
    }
  return RUB_OK;
}

// ---------------------------------------------------------------- multi-frame capture -
// The reference's receive loop (framesync::execute, mimo/framing.cc:471-506, :591-651, :653-886) over a
// capture that holds any number of bursts, device resident from the H2D copy of the capture to the D2H copy
// of the results: Schmidl & Cox metric (sliding sums or the bit-exact FIR form), the plateau rule per
// stream, the state machine restarted behind every burst, one batched timing search for all bursts, the
// per-link timing tables, and one batched LS / invert / decode launch sequence that addresses the capture rows
// in place.  The capture sits behind a lead-in of Wlen zero samples, the content of the reference's window
// buffer before the first sample, so a burst near the start of the capture is decoded like any other.  All
// device scratch is cached on the handle (grow only); the host waits once for the burst count and once for
// the results.
static rub_status write_floats(const std::string &path, const float *p, size_t n) {
  FILE *f = fopen(path.c_str(), "wb");
  if (!f) { set_error("cannot open %s", path.c_str()); return RUB_ERR_IO; }
  const bool ok = fwrite(p, sizeof(float), n, f) == n;
  fclose(f);
  if (!ok) { set_error("short write to %s", path.c_str()); return RUB_ERR_IO; }
  return RUB_OK;
}
extern "C" rub_status rub_rx_process_capture(rub_rx *h, const float *capture, uint64_t n_samples, float threshold,
                                             uint32_t max_frames, const rub_rx_io *out, uint32_t *n_found,
                                             uint64_t *sync_index) {
  if (n_found) *n_found = 0;
  if (!h || !capture || !out || !n_found) { set_error("process_capture: NULL argument"); return RUB_ERR_INVALID_ARG; }
  const HostCfg &c = h->h;
  const uint64_t L = c.L, acb_len = L * (c.nac * c.N + 4), tx_sig_len = (uint64_t)c.D * L, Wlen = acb_len + tx_sig_len;
  const uint32_t max_ac = c.nac * c.N, slots = max_ac + 1;
  if (c.c.estimator != RUB_EST_LS_FULLBAND) { set_error("process_capture: TDMA access codes only"); return RUB_ERR_UNSUPPORTED; }
  const uint64_t lead = Wlen, n = n_samples + lead;   // padded row length
  if (n >= 0x7fffffffull) { set_error("process_capture: capture longer than 2^31 samples"); return RUB_ERR_INVALID_ARG; }
  if (c.N > 8) { set_error("process_capture: at most 8 streams"); return RUB_ERR_UNSUPPORTED; }
  if (max_frames == 0 || n_samples < L) return RUB_OK;
  CUDA_TRY(cudaSetDevice(h->device));
  // --- device scratch, carved from one grow-only allocation
  const size_t pts = (size_t)c.N * c.D * c.Mo;
  const size_t eq_b = (out->out_mask & RUB_OUT_EQ) ? pts * sizeof(cf) : 0, llr_b = (out->out_mask & RUB_OUT_LLR) ? pts * c.q * sizeof(float) : 0,
               bits_b = (out->out_mask & RUB_OUT_BITS) ? (size_t)c.N * c.D * c.row_bytes : 0, rd_b = (out->out_mask & RUB_OUT_RXDATA) ? pts : 0,
               g_b = (out->out_mask & RUB_OUT_G) ? (size_t)c.N * c.N * c.M * sizeof(cf) : 0, tx_b = out->tx_data ? pts : 0;
  const bool sinks = !h->debug_dir.empty();
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  size_t need = 0;
  auto carve = [&](size_t bytes) { const size_t o = need; need += al(bytes); return o; };
  const size_t o_cap = carve(sizeof(cf) * n * c.N), o_y = carve(sizeof(float) * n * c.N), o_ok = carve(n * c.N),
               o_first = carve(sizeof(long long) * ((n + 1023) / 1024)),
               o_off = carve(sizeof(long long) * max_frames), o_syn = carve(sizeof(unsigned long long) * max_frames), o_cnt = carve(256),
               o_keys = carve(sizeof(unsigned long long) * (size_t)max_frames * c.N * slots),
               o_tim = carve(sizeof(int32_t) * (size_t)max_frames * c.N * c.T), o_pay = carve(sizeof(int32_t) * max_frames),
               o_eq = carve(eq_b * max_frames), o_llr = carve(llr_b * max_frames), o_bits = carve(bits_b * max_frames),
               o_rd = carve(rd_b * max_frames), o_g = carve(g_b * max_frames), o_tx = carve(tx_b * max_frames),
               o_corr = carve(sinks ? sizeof(float) * (size_t)c.N * slots * L : 0);
  if (need > h->capt_bytes) {
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    cudaFree(h->d_capt);
    h->d_capt = nullptr;
    h->capt_bytes = 0;
    if (cudaMalloc(&h->d_capt, need) != cudaSuccess) { cudaGetLastError(); set_error("out of device memory (%zu B of capture scratch)", need); return RUB_ERR_NOMEM; }
    h->capt_bytes = need;
  }
  unsigned char *base = (unsigned char *)h->d_capt;
  cf *d_cap = (cf *)(base + o_cap);
  float *d_y = (float *)(base + o_y);
  unsigned char *d_ok = base + o_ok;
  long long *d_off = (long long *)(base + o_off);
  unsigned long long *d_syn = (unsigned long long *)(base + o_syn), *d_keys = (unsigned long long *)(base + o_keys);
  unsigned *d_cnt = (unsigned *)(base + o_cnt);
  int32_t *d_tim = (int32_t *)(base + o_tim), *d_pay = (int32_t *)(base + o_pay);
  // --- capture behind its zero lead-in, metric, plateau flags, the walk over the bursts
  for (uint32_t s = 0; s < c.N; s++) {
    CUDA_TRY(cudaMemsetAsync(d_cap + (size_t)s * n, 0, sizeof(cf) * lead, h->stream));
    CUDA_TRY(cudaMemcpyAsync(d_cap + (size_t)s * n + lead, reinterpret_cast<const cf *>(capture) + (size_t)s * n_samples,
                             sizeof(cf) * n_samples, cudaMemcpyHostToDevice, h->stream));
  }
  CUDA_TRY(launch_sc_metric(h, h->sync_mode, d_cap, n, n, c.N, d_y));
  {
    dim3 grid((unsigned)((n + 1023) / 1024), c.N);
    k_plateau_ok<<<grid, 256, 0, h->stream>>>(d_y, n, n, (int)c.cp, threshold, d_ok);
    PlateauWalk w;
    w.n = (long long)n; w.L = (long long)L; w.acb_len = (long long)acb_len; w.tx_sig_len = (long long)tx_sig_len; w.Wlen = (long long)Wlen;
    w.N = (int)c.N; w.cp = (int)c.cp; w.threshold = threshold; w.max_frames = max_frames;
    long long *d_first = (long long *)(base + o_first);
    k_plateau_first<<<(unsigned)((n + 1023) / 1024), 1024, 0, h->stream>>>(d_ok, (long long)n, n, (int)c.N, d_first);
    k_plateau_walk<<<1, 1024, 0, h->stream>>>(d_ok, d_first, d_y, n, w, d_off, d_syn, d_cnt);
    h->launches += 3;
    CUDA_TRY(cudaGetLastError());
  }
  uint32_t F = 0;
  CUDA_TRY(cudaMemcpyAsync(&F, d_cnt, sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
  std::vector<unsigned long long> syncs(max_frames);
  CUDA_TRY(cudaMemcpyAsync(syncs.data(), d_syn, sizeof(unsigned long long) * max_frames, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));   // the burst count sizes the launches below
  if (sinks) {
    // f_sc_%d.dat (framing.cc:598-600): the metric of every capture sample of stream %d (1-based), float32
    std::vector<float> yh((size_t)n_samples);
    for (uint32_t s = 0; s < c.N; s++) {
      CUDA_TRY(cudaMemcpy(yh.data(), d_y + (size_t)s * n + lead, sizeof(float) * n_samples, cudaMemcpyDeviceToHost));
      rub_status ws = write_floats(h->debug_dir + "/f_sc_" + std::to_string(s + 1) + ".dat", yh.data(), yh.size());
      if (ws) return ws;
    }
  }
  if (F == 0) return RUB_OK;
  // --- batched timing search (framing.cc:702-744) straight from the capture rows, tables on the device
  CUDA_TRY(cudaMemsetAsync(d_keys, 0, sizeof(unsigned long long) * (size_t)F * c.N * slots, h->stream));
  {
    const size_t smem = (size_t)(2 * c.M + 256) * sizeof(cf);
    CUDA_TRY(cudaFuncSetAttribute(k_timing_search, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((c.L + 255) / 256, c.N * slots, F);
    k_timing_search<<<grid, 256, smem, h->stream>>>(d_cap, Wlen, n, d_off, h->d_s1, nullptr, (int)c.M, (int)c.L,
                                                    (int)c.N, (int)c.nac, d_keys, nullptr, 0.f);
    const int tot = (int)(F * c.N * max_ac);
    k_timing_tables<<<(tot + 255) / 256, 256, 0, h->stream>>>(d_keys, d_off, (int)F, (int)c.N, (int)max_ac, (int)c.L, (int)c.M, d_tim, d_pay);
    h->launches += 2;
    CUDA_TRY(cudaGetLastError());
  }
  if (sinks) {
    // corr_%d_%d.dat (framing.cc:676-680, :873-883): per rx stream (1-based) and access code (1-based; 0 = the S0
    // preamble when rub_rx_set_S0 was called) a float32 array of access_code_buffer_len - M correlation powers,
    // non-zero at the candidate offsets i + symbol_len * ac_id, of the LAST burst (the reference reopens the files
    // for every burst)
    float *d_corr = (float *)(base + o_corr);
    unsigned long long *d_k2 = d_keys + (size_t)F * c.N * slots - (size_t)c.N * slots;  // rewrite the last burst's keys: same values
    const size_t smem = (size_t)(2 * c.M + 256) * sizeof(cf);
    dim3 grid((c.L + 255) / 256, c.N * slots, 1);
    CUDA_TRY(cudaMemsetAsync(d_corr, 0, sizeof(float) * (size_t)c.N * slots * L, h->stream));
    k_timing_search<<<grid, 256, smem, h->stream>>>(d_cap, Wlen, n, d_off + (F - 1), h->d_s1, h->d_s0, (int)c.M, (int)c.L,
                                                    (int)c.N, (int)c.nac, d_k2, d_corr, h->s0_corr_scale);
    h->launches += 1;
    std::vector<float> ch((size_t)c.N * slots * L), file(acb_len - c.M);
    CUDA_TRY(cudaMemcpyAsync(ch.data(), d_corr, sizeof(float) * ch.size(), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    for (uint32_t r = 0; r < c.N; r++)
      for (uint32_t slot = 0; slot < slots; slot++) {
        if (slot == 0 && !h->d_s0) continue;
        std::fill(file.begin(), file.end(), 0.f);
        const size_t off = (size_t)L * slot;   // slot 0 (S0): offsets [0, L); access code a: i + L * (a + 1)
        for (uint32_t i = 0; i < L && off + i < file.size(); i++) file[off + i] = ch[((size_t)r * slots + slot) * L + i];
        rub_status ws = write_floats(h->debug_dir + "/corr_" + std::to_string(r + 1) + "_" + std::to_string(slot) + ".dat", file.data(), file.size());
        if (ws) return ws;
      }
  }
  // --- LS / invert / decode of all bursts in one batch, outputs gathered on the device
  rub_rx_io d;
  memset(&d, 0, sizeof(d));
  d.iq = reinterpret_cast<const float *>(d_cap);
  d.layout.rx_stride = n;
  d.timing = d_tim;
  d.payload_start = d_pay;
  d.out_mask = out->out_mask;
  if (eq_b) d.eq = (float *)(base + o_eq);
  if (llr_b) d.llr = (float *)(base + o_llr);
  if (bits_b) d.bits = base + o_bits;
  if (rd_b) d.rx_data = base + o_rd;
  if (g_b) d.G = (float *)(base + o_g);
  if (tx_b) {
    CUDA_TRY(cudaMemcpyAsync(base + o_tx, out->tx_data, tx_b * F, cudaMemcpyHostToDevice, h->stream));
    d.tx_data = base + o_tx;
  }
  rub_status st = process_device(h, &d, F, false, /*shared_rows=*/true);
  if (st) return st;
  if (eq_b) CUDA_TRY(cudaMemcpyAsync(out->eq, d.eq, eq_b * F, cudaMemcpyDeviceToHost, h->stream));
  if (llr_b) CUDA_TRY(cudaMemcpyAsync(out->llr, d.llr, llr_b * F, cudaMemcpyDeviceToHost, h->stream));
  if (bits_b) CUDA_TRY(cudaMemcpyAsync(out->bits, d.bits, bits_b * F, cudaMemcpyDeviceToHost, h->stream));
  if (rd_b) CUDA_TRY(cudaMemcpyAsync(out->rx_data, d.rx_data, rd_b * F, cudaMemcpyDeviceToHost, h->stream));
  if (g_b) CUDA_TRY(cudaMemcpyAsync(out->G, d.G, g_b * F, cudaMemcpyDeviceToHost, h->stream));
  if (out->counters) CUDA_TRY(cudaMemcpyAsync(out->counters, h->d_counters, sizeof(uint64_t) * 4 * c.N, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  if (sync_index) for (uint32_t f = 0; f < F; f++) sync_index[f] = syncs[f] - lead;   // capture coordinates
  *n_found = F;
  return RUB_OK;
}

// ---------------------------------------------------------------- transmit side -------
// framegen::write_sync_words access codes + assemble_mimo_packet, mimo/framing.cc:191-235
extern "C" rub_status rub_framegen_batch_device(rub_rx *h, const uint8_t *tx_data, uint32_t n_frames, float *out,
                                                uint64_t frame_stride, uint64_t stream_stride, float baseband_gain) {
  if (!h || !tx_data || !out) { set_error("framegen_batch: NULL argument"); return RUB_ERR_INVALID_ARG; }
  if (n_frames == 0) return RUB_OK;
  const HostCfg &c = h->h;
  const uint64_t row = (uint64_t)(c.T + c.D) * c.L;
  if (stream_stride < row || frame_stride < stream_stride * (c.N - 1) + row) {
    set_error("framegen_batch: strides smaller than a row of %llu samples", (unsigned long long)row);
    return RUB_ERR_INVALID_ARG;
  }
  CUDA_TRY(cudaSetDevice(h->device));
  TxArgs a;
  a.tx_data = tx_data;
  a.out = reinterpret_cast<cf *>(out);
  a.frame_stride = (long long)frame_stride;
  a.stream_stride = (long long)stream_stride;
  a.tw = h->d_tw; a.occ = h->d_occ; a.scnull = h->d_null; a.sgn = h->d_sgn; a.s1 = h->d_s1;
  a.n_frames = (int)n_frames; a.N = (int)c.N; a.nac = (int)c.nac; a.T = (int)c.T; a.D = (int)c.D;
  a.M = (int)c.M; a.Mo = (int)c.Mo; a.cp = (int)c.cp; a.L = (int)c.L; a.q = (int)c.q; a.P = (int)c.P;
  a.comb = c.c.estimator == RUB_EST_LS_COMB_INTERP;
  a.alpha = c.alpha; a.dn = c.dn; a.g1 = (float)sqrt(1.0 / (double)(float)c.M); a.gain = baseband_gain;
  const long long ctas = (long long)n_frames * (c.T + c.D) * c.N;
  if (ctas > 0x7fffffffLL) { set_error("framegen_batch: batch too large"); return RUB_ERR_INVALID_ARG; }
  switch (c.log2M) {
#define X(L2)                                                                                              \
  case L2: {                                                                                               \
    const size_t smem = (size_t)2 * fft_padded_size(1 << L2) * sizeof(cf);                                 \
    CUDA_TRY(cudaFuncSetAttribute(k_framegen<L2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    k_framegen<L2><<<(unsigned)ctas, FftPlan<L2>::NT, smem, h->stream>>>(a);                               \
  } break;
    X(6) X(7) X(8) X(9) X(10) X(11) X(12)
#undef X
    default: set_error("framegen_batch: unsupported M"); return RUB_ERR_UNSUPPORTED;
  }
  h->launches += 1;
  CUDA_TRY(cudaGetLastError());
  return RUB_OK;
}

// ---------------------------------------------------------------- offline file driver -
// mimo/main.cc:906-918 (read /tmp/rx%d.dat) and :1413-1419 (rx_sig%d.dat, rx_data%d.dat sinks)
namespace {
struct FileSet {
  std::vector<FILE *> f;
  ~FileSet() { for (FILE *p : f) if (p) fclose(p); }
  bool open(const char *const *paths, uint32_t n, const char *mode) {
    f.assign(n, nullptr);
    if (!paths) return true;
    for (uint32_t i = 0; i < n; i++) {
      f[i] = fopen(paths[i], mode);
      if (!f[i]) { set_error("cannot open %s", paths[i]); return false; }
    }
    return true;
  }
  bool any() const { for (FILE *p : f) if (p) return true; return false; }
};
struct Slot {  // one pinned staging slot
  unsigned char *base = nullptr;
  cf *iq = nullptr; uint8_t *tx = nullptr; cf *eq = nullptr; float *llr = nullptr; uint8_t *bits = nullptr, *rxd = nullptr;
  uint32_t frames = 0;  // complete frames the reader found
};
}  // namespace

extern "C" rub_status rub_rx_process_files(rub_rx *h, const rub_file_job *job, uint64_t *frames_done) {
  if (frames_done) *frames_done = 0;
  if (!h || !job || job->struct_size != sizeof(rub_file_job) || !job->rx_paths) {
    set_error("process_files: bad job");
    return RUB_ERR_INVALID_ARG;
  }
  const HostCfg &c = h->h;
  const uint64_t row = (uint64_t)(c.T + c.D) * c.L;
  const uint64_t fstride = job->frame_stride ? job->frame_stride : row;
  if (fstride < row) { set_error("process_files: frame_stride shorter than a frame"); return RUB_ERR_INVALID_ARG; }
  FileSet rx, txf, eqf, rdf, misc;
  if (!rx.open(job->rx_paths, c.N, "rb") || !txf.open(job->tx_data_paths, c.N, "rb") ||
      !eqf.open(job->eq_paths, c.N, "wb") || !rdf.open(job->rx_data_paths, c.N, "wb"))
    return RUB_ERR_IO;
  const char *mp[2] = {job->llr_path, job->bits_path};
  misc.f.assign(2, nullptr);
  for (int i = 0; i < 2; i++)
    if (mp[i] && !(misc.f[i] = fopen(mp[i], "wb"))) { set_error("cannot open %s", mp[i]); return RUB_ERR_IO; }
  const bool want_tx = txf.any(), want_eq = eqf.any(), want_rd = rdf.any(), want_llr = misc.f[0], want_bits = misc.f[1];
  const size_t pts = (size_t)c.D * c.Mo;  // symbols per (frame, stream)
  const size_t in_b = (size_t)c.N * row * sizeof(cf), tx_b = c.N * pts, eq_b = c.N * pts * sizeof(cf),
               llr_b = c.N * pts * c.q * sizeof(float), bits_b = (size_t)c.N * c.D * c.row_bytes, rd_b = c.N * pts;
  uint32_t chunk = job->chunk_frames ? job->chunk_frames : (uint32_t)std::max<size_t>(1, ((size_t)64 << 20) / in_b);
  chunk = std::min<uint32_t>(chunk, std::max<uint32_t>(1, job->n_frames));
  auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
  const size_t slot_b = al(in_b * chunk) + al(tx_b * chunk) + al(eq_b * chunk) + al(llr_b * chunk) + al(bits_b * chunk) + al(rd_b * chunk);
  CUDA_TRY(cudaSetDevice(h->device));
  Slot sl[2];
  unsigned char *pinned = nullptr;
  CUDA_TRY(cudaHostAlloc((void **)&pinned, 2 * slot_b, cudaHostAllocDefault));
  for (int i = 0; i < 2; i++) {
    unsigned char *p = pinned + (size_t)i * slot_b;
    sl[i].base = p;
    sl[i].iq = (cf *)p; p += al(in_b * chunk);
    sl[i].tx = p; p += al(tx_b * chunk);
    sl[i].eq = (cf *)p; p += al(eq_b * chunk);
    sl[i].llr = (float *)p; p += al(llr_b * chunk);
    sl[i].bits = p; p += al(bits_b * chunk);
    sl[i].rxd = p;
  }
  for (uint32_t r = 0; r < c.N; r++)
    if (job->first_sample && fseeko(rx.f[r], (off_t)(job->first_sample * sizeof(cf)), SEEK_SET)) {
      cudaFreeHost(pinned);
      set_error("process_files: cannot seek %s", job->rx_paths[r]);
      return RUB_ERR_IO;
    }
  // reader: frames [f0, f0+nf) of every antenna file -> slot (dense [frame][rx][row]); stops at a short read
  std::vector<uint32_t> tmp32(pts);
  auto read_chunk = [&](Slot &s, uint32_t nf) {
    s.frames = 0;
    for (uint32_t f = 0; f < nf; f++) {
      bool ok = true;
      for (uint32_t r = 0; r < c.N && ok; r++) {
        ok = fread(s.iq + ((size_t)f * c.N + r) * row, sizeof(cf), row, rx.f[r]) == row;
        if (ok && fstride > row) ok = fseeko(rx.f[r], (off_t)((fstride - row) * sizeof(cf)), SEEK_CUR) == 0;
      }
      for (uint32_t r = 0; r < c.N && ok && want_tx; r++) {
        ok = txf.f[r] && fread(tmp32.data(), sizeof(uint32_t), pts, txf.f[r]) == pts;
        uint8_t *d = s.tx + ((size_t)f * c.N + r) * pts;
        for (size_t i = 0; ok && i < pts; i++) d[i] = (uint8_t)tmp32[i];
      }
      if (!ok) break;
      s.frames++;
    }
  };
  rub_status st = RUB_OK;
  uint64_t done = 0;
  read_chunk(sl[0], std::min(chunk, job->n_frames));
  std::vector<uint32_t> out32(pts);
  for (int cur = 0; st == RUB_OK && sl[cur].frames > 0; cur ^= 1) {
    Slot &s = sl[cur];
    // prefetch the next chunk from disk while this one is on the GPU
    const bool more = s.frames == std::min(chunk, job->n_frames - (uint32_t)done) && done + s.frames < job->n_frames;
    const uint32_t next_n = more ? std::min<uint32_t>(chunk, job->n_frames - (uint32_t)(done + s.frames)) : 0;
    sl[cur ^ 1].frames = 0;
    std::thread reader;
    if (next_n) reader = std::thread(read_chunk, std::ref(sl[cur ^ 1]), next_n);
    rub_rx_io io;
    memset(&io, 0, sizeof(io));
    io.iq = (const float *)s.iq;
    io.tx_data = want_tx ? s.tx : nullptr;
    if (want_eq) { io.eq = (float *)s.eq; io.out_mask |= RUB_OUT_EQ; }
    if (want_llr) { io.llr = s.llr; io.out_mask |= RUB_OUT_LLR; }
    if (want_bits) { io.bits = s.bits; io.out_mask |= RUB_OUT_BITS; }
    if (want_rd) { io.rx_data = s.rxd; io.out_mask |= RUB_OUT_RXDATA; }
    st = rub_rx_process_batch_host(h, &io, s.frames);
    if (st == RUB_OK) {
      bool ok = true;
      for (uint32_t f = 0; f < s.frames && ok; f++)
        for (uint32_t r = 0; r < c.N && ok; r++) {
          const size_t o = ((size_t)f * c.N + r) * pts;
          if (eqf.f[r]) ok = fwrite(s.eq + o, sizeof(cf), pts, eqf.f[r]) == pts;
          if (ok && rdf.f[r]) {
            for (size_t i = 0; i < pts; i++) out32[i] = s.rxd[o + i];
            ok = fwrite(out32.data(), sizeof(uint32_t), pts, rdf.f[r]) == pts;
          }
        }
      if (ok && want_llr) ok = fwrite(s.llr, 1, llr_b * s.frames, misc.f[0]) == llr_b * s.frames;
      if (ok && want_bits) ok = fwrite(s.bits, 1, bits_b * s.frames, misc.f[1]) == bits_b * s.frames;
      if (!ok) { set_error("process_files: short write"); st = RUB_ERR_IO; }
      else done += s.frames;
    }
    if (reader.joinable()) reader.join();
  }
  cudaFreeHost(pinned);
  if (frames_done) *frames_done = done;
  return st;
}
