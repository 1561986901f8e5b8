// Aggregate L2 -> SM bandwidth of bulk async copies (cp.async.bulk, TMA engine) with EVERY SM pulling at once -- the
// number that bounds the backward of the NeRF MLP ("the bandwidth wall").
//   mode 0: every CTA streams its OWN slice of a window that fits L2 (distinct data, L2 hits after the first pass)
//   mode 1: every CTA streams the SAME slice (what the weight-streaming forward does: requests de-duplicated in L2)
//   mode 2: every CTA streams its own slice of a window far larger than L2 (HBM)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/l2_bw_probe tools/l2_bw_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar) : "memory");
}

constexpr int kRing = 196608;      // default bytes in flight per SM, whatever the copy size

// kChunk: bytes per bulk copy; kLanes: lanes of the warp that each issue one copy per round (a 16 x 1 KB gather = 16 lanes)
template <int kChunk, int kLanes, int kRingBytes = kRing>
__global__ void __launch_bounds__(32, 1) probe(const uint8_t* __restrict__ base, size_t slice_bytes, size_t cta_stride, int n_chunks,
                                              long long* cycles) {
  constexpr int kStage = kChunk * kLanes, kStages = kRingBytes / kStage;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[kStages];
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(smem_u32(&full[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const int lane = threadIdx.x;
  const uint8_t* src = base + (size_t)blockIdx.x * cta_stride;
  const size_t per_slice = slice_bytes / kStage;
  const long long t0 = clock64();
  for (int i = 0; i < n_chunks + kStages; ++i) {
    const int s = i % kStages;
    if (i >= kStages) {                       // stage i - kStages landed?
      const uint32_t ph = ((i - kStages) / kStages) & 1u;
      if (lane == 0) while (!mbar_try_wait(smem_u32(&full[s]), ph)) {}
      __syncwarp();
    }
    if (i < n_chunks) {
      if (lane == 0) mbar_arrive_expect_tx(smem_u32(&full[s]), kStage);
      __syncwarp();
      if (lane < kLanes)
        bulk_g2s(smem_u32(smem) + s * kStage + lane * kChunk, src + (size_t)(i % per_slice) * kStage + (size_t)lane * kChunk, kChunk,
                 smem_u32(&full[s]));
    }
  }
  if (lane == 0) cycles[blockIdx.x] = clock64() - t0;
}

template <int kChunk, int kLanes, int kRingBytes = kRing>
static void run(const char* name, const uint8_t* buf, size_t slice, size_t stride, int sms, long long* cyc) {
  constexpr int kStage = kChunk * kLanes;
  slice = slice / kStage * kStage;
  if (stride) stride = slice;
  cudaFuncSetAttribute(probe<kChunk, kLanes, kRingBytes>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRingBytes);
  const int n = (int)(((size_t)128 << 20) / kStage);          // 128 MB per SM
  for (int rep = 0; rep < 2; ++rep) {                         // the first repetition warms L2
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<kChunk, kLanes, kRingBytes><<<sms, 32, kRingBytes>>>(buf, slice, stride, n, cyc);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep == 1)
      printf("%-46s %2d x %5d B per stage, %2d stages: %6.2f TB/s over %d SMs (%6.1f GB/s per SM)  %s\n", name, kLanes, kChunk,
             kRingBytes / kStage,
             (double)sms * n * kStage / (ms * 1e-3) / 1e12, sms, (double)n * kStage / (ms * 1e-3) / 1e9, cudaGetErrorString(err));
  }
}

int main(int argc, char** argv) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t big = (size_t)8 << 30;
  uint8_t* buf;
  if (cudaMalloc(&buf, big) != cudaSuccess) { printf("cudaMalloc failed\n"); return 1; }
  cudaMemset(buf, 1, big);
  long long* cyc;
  cudaMalloc(&cyc, sms * sizeof(long long));
  const size_t l2 = ((size_t)64 << 20) / sms, hbm = (size_t)50 << 20;
  printf("bytes in flight per SM: %d\n", kRing);
  run<32768, 1>("distinct data, 64 MB window (L2-resident)", buf, l2, 1, sms, cyc);
  run<16384, 1>("distinct data, 64 MB window (L2-resident)", buf, l2, 1, sms, cyc);
  run<1024, 16>("distinct data, 64 MB window (L2-resident)", buf, l2, 1, sms, cyc);
  run<1024, 1>("distinct data, 64 MB window (L2-resident)", buf, l2, 1, sms, cyc);
  run<32768, 1>("same 1 MB for every SM (de-duplicated in L2)", buf, (size_t)1 << 20, 0, sms, cyc);
  run<32768, 1>("distinct data, 7.4 GB window (HBM)", buf, hbm, 1, sms, cyc);
  run<16384, 1>("distinct data, 7.4 GB window (HBM)", buf, hbm, 1, sms, cyc);
  run<1024, 16>("distinct data, 7.4 GB window (HBM)", buf, hbm, 1, sms, cyc);
  // the weight ring of the forward / chain kernels: the same ~1.2 MB for every SM, 64 KB of ring
  run<16384, 1, 65536>("same 1 MB for every SM (weight ring)", buf, (size_t)1 << 20, 0, sms, cyc);
  run<32768, 1, 65536>("same 1 MB for every SM (weight ring)", buf, (size_t)1 << 20, 0, sms, cyc);
  run<16384, 1, 98304>("same 1 MB for every SM (weight ring)", buf, (size_t)1 << 20, 0, sms, cyc);
  run<32768, 1, 98304>("same 1 MB for every SM (weight ring)", buf, (size_t)1 << 20, 0, sms, cyc);
  run<16384, 2, 65536>("same 1 MB for every SM (weight ring)", buf, (size_t)1 << 20, 0, sms, cyc);
  run<8192, 2, 65536>("same 1 MB for every SM (weight ring)", buf, (size_t)1 << 20, 0, sms, cyc);
  run<32768, 1, 163840>("distinct data, 64 MB window (L2-resident)", buf, l2, 1, sms, cyc);
  run<32768, 1, 131072>("distinct data, 64 MB window (L2-resident)", buf, l2, 1, sms, cyc);
  run<32768, 2, 196608>("distinct data, 64 MB window (L2-resident)", buf, l2, 1, sms, cyc);
  run<32768, 1, 163840>("distinct data, 7.4 GB window (HBM)", buf, hbm, 1, sms, cyc);
  return 0;
}
