// rub_fused.cu — instances and host launchers of the monolithic fused kernel and of the
// block-mapped detect kernel (rub_kernels_fused.cuh).
#include "rub_kernels_fused.cuh"
#include "rub_launch.h"

namespace rub {

// the (log2 M, N) pairs the fused kernel is instantiated for
#define RUB_FUSED_LIST(X) X(9, 2) X(9, 4) X(10, 2) X(10, 4) X(11, 1) X(11, 2) X(11, 4) X(12, 1) X(12, 2)

bool fused_has_instance(uint32_t l2, uint32_t N) {
#define X(L, NN) if (l2 == L && N == NN) return true;
  RUB_FUSED_LIST(X)
#undef X
  return false;
}
template <int LOG2M, int N>
static cudaError_t prepare(uint32_t q, size_t *smem_out, int *occ, int *variant) {
  using TR = FusedTraits<LOG2M, N>;
  *variant = 0;
  if (TR::HAS_WTMA && !getenv("RUB_FUSED_NO_WTMA")) {
    // the TMA-record variant, if it keeps the occupancy the classic one is built for
    const size_t sm = TR::smem_bytes((int)q, true);
    int o = 0;
    if (cudaFuncSetAttribute(k_rx_fused<LOG2M, N, TR::HAS_WTMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm) == cudaSuccess &&
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, k_rx_fused<LOG2M, N, TR::HAS_WTMA>, TR::THREADS, sm) == cudaSuccess &&
        o >= TR::MIN_CTAS) {
      *smem_out = sm; *occ = o; *variant = 1;
      return cudaSuccess;
    }
    cudaGetLastError();
  }
  const size_t smem = TR::smem_bytes((int)q);
  cudaError_t e = cudaFuncSetAttribute(k_rx_fused<LOG2M, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  *smem_out = smem;
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, k_rx_fused<LOG2M, N>, TR::THREADS, smem);
}
cudaError_t fused_prepare(uint32_t l2, uint32_t N, uint32_t q, size_t *smem, int *occ, int *variant) {
#define X(L, NN) if (l2 == L && N == NN) return prepare<L, NN>(q, smem, occ, variant);
  RUB_FUSED_LIST(X)
#undef X
  return cudaErrorInvalidValue;
}
template <int LOG2M, int N>
static void launch_fused(int grid, size_t smem, cudaStream_t st, const FusedArgs &fa, const DemapConst &dc, int variant) {
  using TR = FusedTraits<LOG2M, N>;
  if (variant && TR::HAS_WTMA) k_rx_fused<LOG2M, N, TR::HAS_WTMA><<<grid, TR::THREADS, smem, st>>>(fa, dc);
  else k_rx_fused<LOG2M, N><<<grid, TR::THREADS, smem, st>>>(fa, dc);
}
void fused_launch(uint32_t l2, uint32_t N, int grid, size_t smem, cudaStream_t st, const FusedArgs &fa, const DemapConst &dc, int variant) {
#define X(L, NN) if (l2 == L && N == NN) { launch_fused<L, NN>(grid, smem, st, fa, dc, variant); return; }
  RUB_FUSED_LIST(X)
#undef X
}

// block-mapped detect kernel: all carriers occupied, 16-byte aligned outputs, N in {1,2,4,8}
template <int N, int MB>
static void launch_lean(const ChainArgs &a, const DemapConst &dc, cudaStream_t st) {
  const int llr_stage = 256 * 2 * MB;
  // 8 warps x (2 staging slots + 2 W records of N * 512 + 512 bytes)
  const size_t smem = (size_t)8 * 2 * (llr_stage + 64) + (DetectLeanTmaW<N>::value ? (size_t)8 * 2 * (N * 512 + 512 + 64) : 0);
  cudaFuncSetAttribute(k_detect_lean<N, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const long long nwork = (long long)a.n_frames * a.D * (a.M / 64);
  k_detect_lean<N, MB><<<(unsigned)((nwork + 7) / 8), 256, smem, st>>>(a, dc, llr_stage);
}
template <int N>
static void launch_lean_q(const ChainArgs &a, const DemapConst &dc, cudaStream_t st) {
  switch (a.q) {
    case 2: launch_lean<N, 1>(a, dc, st); break;
    case 4: launch_lean<N, 2>(a, dc, st); break;
    case 6: launch_lean<N, 3>(a, dc, st); break;
    default: launch_lean<N, 4>(a, dc, st); break;
  }
}
bool detect_lean_eligible(const ChainArgs &a) {
  if (a.Mo != a.M || (a.M % 64)) return false;
  if (((uintptr_t)a.llr & 15) || ((uintptr_t)a.bits & 15) || ((uintptr_t)a.eq & 15) || ((uintptr_t)a.W & 15)) return false;
  if (((uintptr_t)a.rx_data & 1) || ((uintptr_t)a.tx_data & 1)) return false;
  if (a.N >= 4 && ((uintptr_t)a.tx_data & 15)) return false;  // tx_data rides in the TMA ring slot
  return a.N == 1 || a.N == 2 || a.N == 4 || a.N == 8;
}
// do the weights kernels have to write task records for this call?
bool detect_lean_records(const ChainArgs &a) { return detect_lean_eligible(a) && a.N >= 4; }
bool detect_lean_launch(const ChainArgs &a, const DemapConst &dc, cudaStream_t st) {
  if (!detect_lean_eligible(a) || (a.wrec != 0) != detect_lean_records(a)) return false;
  switch (a.N) {
    case 1: launch_lean_q<1>(a, dc, st); return true;
    case 2: launch_lean_q<2>(a, dc, st); return true;
    case 4: launch_lean_q<4>(a, dc, st); return true;
    case 8: launch_lean_q<8>(a, dc, st); return true;
    default: return false;
  }
}

}  // namespace rub
