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
template <int LOG2M, int N>
static cudaError_t prepare(uint32_t q, size_t *smem_out, int *occ) {
  using TR = WsTraits<LOG2M, N>;
  const size_t smem = TR::smem_bytes((int)q);
  cudaError_t e = cudaFuncSetAttribute(k_rx_ws<LOG2M, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  *smem_out = smem;
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, k_rx_ws<LOG2M, N>, TR::THREADS, smem);
}
cudaError_t ws_prepare(uint32_t l2, uint32_t N, uint32_t q, size_t *smem, int *occ) {
#define X(L, NN) if (l2 == L && N == NN) return prepare<L, NN>(q, smem, occ);
  RUB_WS_LIST(X)
#undef X
  return cudaErrorInvalidValue;
}
void ws_launch(uint32_t l2, uint32_t N, int grid, size_t smem, cudaStream_t st, const FusedArgs &fa, const DemapConst &dc) {
#define X(L, NN) if (l2 == L && N == NN) { k_rx_ws<L, NN><<<grid, WsTraits<L, NN>::THREADS, smem, st>>>(fa, dc); return; }
  RUB_WS_LIST(X)
#undef X
}

}  // namespace rub
