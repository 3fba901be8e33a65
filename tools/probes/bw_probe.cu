// HBM bandwidth probe: pure read, pure write (st.global and TMA bulk store), copy, 30/70 mix.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bw_probe bw_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void k_read(const float4 *p, size_t n, float *out) {
  float acc = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float4 v = __ldcs(p + i);
    acc += v.x + v.y + v.z + v.w;
  }
  if (acc == 123.456f) *out = acc;
}
__global__ void k_write(float4 *p, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    __stcs(p + i, make_float4(1.f, 2.f, 3.f, (float)i));
}
__global__ void k_copy(const float4 *s, float4 *d, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    __stcs(d + i, __ldcs(s + i));
}
// read nr float4 and write nw float4 per iteration-chunk (mix)
__global__ void k_mix(const float4 *s, float4 *d, size_t n_r, size_t n_w) {
  const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x, T = (size_t)gridDim.x * blockDim.x;
  float acc = 0.f;
  for (size_t i = t; i < n_r; i += T) { float4 v = __ldcs(s + i); acc += v.x; }
  for (size_t i = t; i < n_w; i += T) __stcs(d + i, make_float4(acc, 2.f, 3.f, 4.f));
}
// TMA bulk stores: each warp stages 1536 B in smem and bulk-stores it
__global__ void k_write_tma(unsigned char *p, size_t nbytes) {
  extern __shared__ __align__(128) unsigned char sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, W = blockDim.x >> 5;
  unsigned char *slot = sm + warp * 2 * 1536;
  const size_t chunks = nbytes / 1536;
  int it = 0;
  for (size_t c = blockIdx.x * (size_t)W + warp; c < chunks; c += (size_t)gridDim.x * W, it++) {
    unsigned char *sl = slot + (it & 1) * 1536;
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    __syncwarp();
    for (int j = 0; j < 3; j++) reinterpret_cast<float4 *>(sl)[lane + 32 * j] = make_float4(1.f, 2.f, (float)c, (float)j);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(p + c * 1536),
                   "r"((unsigned)__cvta_generic_to_shared(sl)), "r"(1536u) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
  const size_t bytes = (size_t)4 << 30;
  float4 *a, *b; float *o;
  CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes)); CK(cudaMalloc(&o, 4));
  CK(cudaMemset(a, 1, bytes)); CK(cudaMemset(b, 2, bytes));
  const size_t n = bytes / 16;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int grid = sms * 8, thr = 256;
  auto timeit = [&](const char *name, double gb, auto fn) {
    for (int i = 0; i < 2; i++) fn();
    cudaEventRecord(e0);
    for (int i = 0; i < 5; i++) fn();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    printf("%-28s %.3f ms  %.0f GB/s\n", name, ms, gb / (ms * 1e-3));
  };
  timeit("read 4 GiB", bytes / 1e9, [&] { k_read<<<grid, thr>>>(a, n, o); });
  timeit("write 4 GiB (st.cs)", bytes / 1e9, [&] { k_write<<<grid, thr>>>(b, n); });
  timeit("copy 4+4 GiB", 2 * bytes / 1e9, [&] { k_copy<<<grid, thr>>>(a, b, n); });
  timeit("mix read 1.8 / write 4.07 GB", (1.8e9 + 4.07e9) / 1e9, [&] { k_mix<<<grid, thr>>>(a, b, (size_t)(1.8e9 / 16), (size_t)(4.07e9 / 16)); });
  cudaFuncSetAttribute(k_write_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 2 * 1536);
  timeit("write 4 GiB (TMA 1536 B)", bytes / 1e9, [&] { k_write_tma<<<sms * 2, 512, 16 * 2 * 1536>>>((unsigned char *)b, bytes); });
  timeit("write 4 GiB (TMA, 1 CTA/SM)", bytes / 1e9, [&] { k_write_tma<<<sms, 512, 16 * 2 * 1536>>>((unsigned char *)b, bytes); });
  CK(cudaDeviceSynchronize());
  return 0;
}
