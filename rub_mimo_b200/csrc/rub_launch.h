// rub_launch.h — host-side launchers of the fused kernels.  Each kernel family is instantiated in
// its own translation unit (rub_fused.cu, rub_ws.cu) so the library builds in parallel.
#pragma once
#include <cuda_runtime.h>

#include "rub_kernels_args.cuh"

namespace rub {

// monolithic fused kernel (rub_kernels_fused.cuh)
bool fused_has_instance(uint32_t log2M, uint32_t N);
// variant 1 = W / gain / isig as TMA task records (two-stream instances, when the shared memory allows)
cudaError_t fused_prepare(uint32_t log2M, uint32_t N, uint32_t q, size_t *smem, int *ctas_per_sm, int *variant);
void fused_launch(uint32_t log2M, uint32_t N, int grid, size_t smem, cudaStream_t st, const FusedArgs &fa, const DemapConst &dc, int variant);
// warp-specialised fused kernel (rub_kernels_ws.cuh)
bool ws_has_instance(uint32_t log2M, uint32_t N);
cudaError_t ws_prepare(uint32_t log2M, uint32_t N, uint32_t q, size_t *smem, int *ctas_per_sm);
// sign bytes of the access codes in the kernel's last-stage register order (host side, sgn = [tx][code][M] of +-1)
size_t ws_sign_bytes(uint32_t log2M, uint32_t N, uint32_t nac);
void ws_pack_signs(uint32_t log2M, uint32_t N, uint32_t nac, const float *sgn, unsigned char *out);
void ws_launch(uint32_t log2M, uint32_t N, int grid, size_t smem, cudaStream_t st, const FusedArgs &fa, const DemapConst &dc);
// block-mapped detect kernel of the staged path (rub_kernels_fused.cuh: k_detect_lean)
// (its W / gain / isig come as task records: the weights kernels are told through ChainArgs::wrec)
bool detect_lean_eligible(const ChainArgs &a);
bool detect_lean_records(const ChainArgs &a);
bool detect_lean_launch(const ChainArgs &a, const DemapConst &dc, cudaStream_t st);

}  // namespace rub
