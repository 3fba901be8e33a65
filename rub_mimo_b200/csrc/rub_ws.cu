// rub_ws.cu — instances and host launchers of the warp-specialised fused kernel (rub_kernels_ws.cuh).
#include "rub_kernels_ws.cuh"
#include "rub_launch.h"

namespace rub {

#define RUB_WS_LIST(X) X(11, 4)

bool ws_has_instance(uint32_t l2, uint32_t N) {
#define X(L, NN) if (l2 == L && N == NN) return true;
  RUB_WS_LIST(X)
#undef X
  return false;
}
template <int LOG2M, int N, int MB>
static cudaError_t prepare_mb(size_t *smem_out, int *occ) {
  using TR = WsTraits<LOG2M, N>;
  const size_t smem = TR::smem_bytes(2 * MB);
  cudaError_t e = cudaFuncSetAttribute(k_rx_ws<LOG2M, N, MB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_rx_ws<LOG2M, N, MB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  *smem_out = smem;
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, k_rx_ws<LOG2M, N, MB, true>, TR::THREADS, smem);
}
template <int LOG2M, int N>
static cudaError_t prepare(uint32_t q, size_t *smem_out, int *occ) {
  switch (q) {
    case 2: return prepare_mb<LOG2M, N, 1>(smem_out, occ);
    case 4: return prepare_mb<LOG2M, N, 2>(smem_out, occ);
    case 6: return prepare_mb<LOG2M, N, 3>(smem_out, occ);
    default: return cudaErrorInvalidValue;  // 256-QAM: the LLR staging does not fit beside the rings (monolithic kernel)
  }
}
template <int LOG2M, int N, int MB>
static void launch_mb(int grid, size_t smem, cudaStream_t st, const FusedArgs &fa, const DemapConst &dc) {
  constexpr int T = WsTraits<LOG2M, N>::THREADS;
  const ChainArgs &a = fa.a;
  // the usual output set gets the instance without per-task null checks of the output pointers
  if (a.eq && a.llr && a.bits && a.tx_data && !a.rx_data) k_rx_ws<LOG2M, N, MB, true><<<grid, T, smem, st>>>(fa, dc);
  else k_rx_ws<LOG2M, N, MB, false><<<grid, T, smem, st>>>(fa, dc);
}
template <int LOG2M, int N>
static void launch(int grid, size_t smem, cudaStream_t st, const FusedArgs &fa, const DemapConst &dc) {
  switch (fa.a.q) {
    case 2: launch_mb<LOG2M, N, 1>(grid, smem, st, fa, dc); break;
    case 4: launch_mb<LOG2M, N, 2>(grid, smem, st, fa, dc); break;
    case 6: launch_mb<LOG2M, N, 3>(grid, smem, st, fa, dc); break;
    default: break;  // never prepared (see prepare)
  }
}
cudaError_t ws_prepare(uint32_t l2, uint32_t N, uint32_t q, size_t *smem, int *occ) {
#define X(L, NN) if (l2 == L && N == NN) return prepare<L, NN>(q, smem, occ);
  RUB_WS_LIST(X)
#undef X
  return cudaErrorInvalidValue;
}
void ws_pack_signs(uint32_t l2, uint32_t N, uint32_t nac, const float *sgn, unsigned char *out) {
#define X(L, NN) if (l2 == L && N == NN) { ws_pack_signs<L>((int)N, (int)nac, sgn, out); return; }
  RUB_WS_LIST(X)
#undef X
}
size_t ws_sign_bytes(uint32_t l2, uint32_t N, uint32_t nac) {
#define X(L, NN) if (l2 == L && N == NN) return (size_t)N * nac * (Fft<L>::M / (Fft<L>::S2::P / Fft<L>::S2::B));
  RUB_WS_LIST(X)
#undef X
  return 0;
}
void ws_launch(uint32_t l2, uint32_t N, int grid, size_t smem, cudaStream_t st, const FusedArgs &fa, const DemapConst &dc) {
#define X(L, NN) if (l2 == L && N == NN) { launch<L, NN>(grid, smem, st, fa, dc); return; }
  RUB_WS_LIST(X)
#undef X
}

}  // namespace rub
