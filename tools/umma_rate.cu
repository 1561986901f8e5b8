// Microbenchmark: sustained issue rate of tcgen05.mma.cta_group::1.kind::f16 (bf16, M=128) from shared-memory operands.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate tools/umma_rate.cu && ./umma_rate
// Prints SM cycles per MMA for N = 256 / 128 / 64 with 1 or 2 independent accumulators, one CTA per SM on all SMs.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../nerf-and-dietnerf_b200/csrc/tc_common.cuh"
using namespace nerf::tc;

__global__ void __launch_bounds__(128, 1) rate_kernel(int n_mma, int N, int n_acc, int a_advance, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const uint32_t sbase = smem_u32(smem);
  for (int i = threadIdx.x; i < (64 * 1024 + 32 * 1024) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(smem_u32(&tmem_base_s), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(N);
    const uint32_t a_addr = sbase, b_addr = sbase + 64 * 1024;
    long long t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      const int k = i & 3, panel = (i >> 2) & 3;
      umma_bf16(tmem_base + (uint32_t)((i / 16) % n_acc) * 256u, make_desc_kmajor(a_addr + (a_advance ? panel * 16384 : 0) + k * 32),
                make_desc_kmajor(b_addr + k * 32), idesc, 1u);
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem_base, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  const int smem = 96 * 1024;
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int n_mma = 8192;
  for (int grid : {1, 148}) for (int N : {256, 128, 64}) for (int n_acc : {1, 2}) for (int adv : {0, 1}) {
    rate_kernel<<<grid, 128, smem>>>(n_mma, N, n_acc, adv, d);   // warm-up
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    rate_kernel<<<grid, 128, smem>>>(n_mma, N, n_acc, adv, d);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long cyc; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
    double flops = 2.0 * 128 * N * 16 * (double)n_mma * grid;
    printf("grid %3d N %3d acc %d a_adv %d: %7.1f cycles/MMA  (%.3f ms, %.1f TFLOP/s, clock %.2f GHz) %s\n", grid, N, n_acc, adv,
           (double)cyc / n_mma, ms, flops / ms / 1e9, cyc / (ms * 1e6), cudaGetErrorString(err));
  }
  return 0;
}
