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

// Loads and stores at the same time: warp 0 streams 32 KB bulk copies (as above), `store_warps` more warps write
// 512-byte warp stores (st.global.v4, what the epilogues of the MLP kernels issue) into an L2-resident window of their own.
// Each role reports its own cycle count: do the SM's load requests and its store data share a path?
__global__ void __launch_bounds__(32 * 17, 1) probe_ls(const uint8_t* __restrict__ base, size_t slice_bytes, int n_chunks,
                                                       uint8_t* __restrict__ st_base, size_t st_slice, long long st_iters,
                                                       int store_warps, long long* cyc_load, long long* cyc_store, int st_kind) {
  constexpr int kChunk = 32768, kStages = 5;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[kStages];
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(smem_u32(&full[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long t0 = clock64();
  if (warp == 0) {
    const uint8_t* src = base + (size_t)blockIdx.x * slice_bytes;
    const size_t per_slice = slice_bytes / kChunk;
    for (int i = 0; i < n_chunks + kStages; ++i) {
      const int s = i % kStages;
      if (i >= kStages) {
        const uint32_t ph = ((i - kStages) / kStages) & 1u;
        if (lane == 0) while (!mbar_try_wait(smem_u32(&full[s]), ph)) {}
        __syncwarp();
      }
      if (i < n_chunks && lane == 0) {
        mbar_arrive_expect_tx(smem_u32(&full[s]), kChunk);
        bulk_g2s(smem_u32(smem) + s * kChunk, src + (size_t)(i % per_slice) * kChunk, kChunk, smem_u32(&full[s]));
      }
    }
    if (lane == 0) cyc_load[blockIdx.x] = clock64() - t0;
  } else if (warp <= store_warps) {
    uint8_t* dst = st_base + (size_t)blockIdx.x * st_slice + (size_t)(warp - 1) * 512 + lane * 16;
    const size_t stride = (size_t)store_warps * 512, wrap = st_slice / stride;
    if (st_kind == 0) {
      for (long long i = 0; i < st_iters; ++i) {
        const uint4 v = make_uint4((uint32_t)i, lane, warp, 7u);
        asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(dst + (size_t)(i % wrap) * stride), "r"(v.x), "r"(v.y), "r"(v.z),
                     "r"(v.w) : "memory");
      }
    } else if (st_kind == 1) {
      // 256-bit stores: a warp writes 1 KB per instruction (half as many instructions for the same bytes)
      uint8_t* d8 = st_base + (size_t)blockIdx.x * st_slice + (size_t)(warp - 1) * 1024 + lane * 32;
      const size_t stride8 = (size_t)store_warps * 1024, wrap8 = st_slice / stride8;
      for (long long i = 0; i < st_iters / 2; ++i) {
        const uint32_t a = (uint32_t)i;
        asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(d8 + (size_t)(i % wrap8) * stride8), "r"(a),
                     "r"(a), "r"(a), "r"(a), "r"(a), "r"(a), "r"(a), "r"(a) : "memory");
      }
    } else {
      // bulk shared -> global copies of 16 KB (TMA engine) out of the ring's memory, one lane per warp, 4 in flight per warp
      if (lane == 0) {
        uint8_t* db = st_base + (size_t)blockIdx.x * st_slice;
        const size_t pieces = st_slice / 16384;
        const long long n = st_iters * 512 / 16384;
        for (long long i = 0; i < n; ++i) {
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(db + (size_t)((i * store_warps + warp - 1) % pieces) * 16384),
                       "r"(smem_u32(smem) + (uint32_t)((warp - 1) & 7) * 16384u), "r"(16384) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      }
    }
    if (lane == 0 && warp == 1) cyc_store[blockIdx.x] = clock64() - t0;
  }
}

static void run_ls(const uint8_t* buf, uint8_t* stbuf, int sms, long long* cyc, size_t load_mb, size_t store_mb, int store_warps,
                   bool hbm_loads, int st_kind = 0) {
  cudaFuncSetAttribute(probe_ls, cudaFuncAttributeMaxDynamicSharedMemorySize, 5 * 32768);
  const size_t slice = hbm_loads ? ((size_t)40 << 20) : (((size_t)48 << 20) / sms / 32768 * 32768);
  const size_t st_slice = ((size_t)24 << 20) / sms / 8192 * 8192;
  const int n_chunks = (int)((load_mb << 20) / 32768);
  const long long st_iters = store_warps ? (long long)((store_mb << 20) / ((size_t)store_warps * 512)) : 0;
  long long h[2][160];
  for (int rep = 0; rep < 2; ++rep) {
    cudaMemset(cyc, 0, 2 * 160 * sizeof(long long));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe_ls<<<sms, 32 * 17, 5 * 32768>>>(buf, slice, n_chunks, stbuf, st_slice, st_iters, store_warps, cyc, cyc + 160, st_kind);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    if (rep == 1) {
      double cl = 0, cs = 0;
      for (int i = 0; i < sms; ++i) { cl += (double)h[0][i] / sms; cs += (double)h[1][i] / sms; }
      printf("[%s] loads %4zu MB/SM (%s) + stores %4zu MB/SM by %2d warps: kernel %7.3f ms | load role %9.0f cycles = %5.1f B/clk/SM | "
             "store role %9.0f cycles = %5.1f B/clk/SM  %s\n",
             st_kind == 0 ? "st.v4" : st_kind == 1 ? "st.v8" : "bulk S2G", n_chunks ? load_mb : 0, hbm_loads ? "HBM" : "L2", store_warps ? store_mb : 0, store_warps, ms, cl,
             cl > 0 ? (double)n_chunks * 32768 / cl : 0.0, cs, cs > 0 ? (double)st_iters * store_warps * 512 / cs : 0.0,
             cudaGetErrorString(err));
    }
  }
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
  cudaMalloc(&cyc, 2 * 160 * sizeof(long long));
  const size_t l2 = ((size_t)64 << 20) / sms, hbm = (size_t)50 << 20;
  if (argc > 1 && atoi(argv[1]) == 1) {
    // loads and stores together (see probe_ls)
    uint8_t* stbuf = buf + ((size_t)7 << 30);
    printf("-- bulk loads alone, stores alone, both (5 x 32 KB in flight; stores = 512-byte warp stores into an L2-resident window)\n");
    run_ls(buf, stbuf, sms, cyc, 96, 0, 0, false);
    run_ls(buf, stbuf, sms, cyc, 0, 32, 16, false);
    run_ls(buf, stbuf, sms, cyc, 0, 32, 4, false);
    run_ls(buf, stbuf, sms, cyc, 96, 32, 16, false);
    run_ls(buf, stbuf, sms, cyc, 96, 16, 16, false);
    run_ls(buf, stbuf, sms, cyc, 96, 8, 16, false);
    run_ls(buf, stbuf, sms, cyc, 96, 32, 4, false);
    run_ls(buf, stbuf, sms, cyc, 96, 0, 0, true);
    run_ls(buf, stbuf, sms, cyc, 96, 32, 16, true);
    for (int kind = 1; kind <= 2; ++kind) {
      run_ls(buf, stbuf, sms, cyc, 0, 32, 16, false, kind);
      run_ls(buf, stbuf, sms, cyc, 0, 32, 4, false, kind);
      run_ls(buf, stbuf, sms, cyc, 96, 32, 16, false, kind);
      run_ls(buf, stbuf, sms, cyc, 96, 32, 4, false, kind);
    }
    return 0;
  }
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
