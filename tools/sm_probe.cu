// Microbenchmarks behind the MLP kernel design:
//   (1) tcgen05.ld throughput per SM (TMEM -> registers) for 4 / 8 / 16 warps;
//   (2) tcgen05.mma issue rate, cta_group::1 vs cta_group::2 (M = 256 over a CTA pair), N = 256 / 128;
//   (3) the same MMA stream while a producer thread streams weight chunks from L2 into shared memory with bulk copies
//       and epilogue-like warps write shared memory with st.shared.v4 -- the shared-memory contention the forward
//       kernel sees; also reports the L2 -> SM bulk-copy bandwidth with all SMs pulling.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/sm_probe tools/sm_probe.cu && tools/sm_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../nerf-and-dietnerf_b200/csrc/tc_common.cuh"
using namespace nerf::tc;

__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
               : "memory");
}
__host__ __device__ constexpr uint32_t make_idesc_m(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// ---- (1) tcgen05.ld --------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512, 1) tmem_ld_kernel(int n_iter, int n_warps, long long* out, uint32_t* sink) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  uint32_t x = 0;
  long long t0 = 0, t1 = 0;
  __syncthreads();
  if (warp < n_warps) {
    const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) & 3) * 128u;
    t0 = clock64();
    for (int i = 0; i < n_iter; ++i) {
      uint32_t a[32], b[32];
      tmem_ld32(taddr + ((i & 1) * 64), a);
      tmem_ld32(taddr + ((i & 1) * 64) + 32, b);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) x ^= a[j] + b[j];
    }
    t1 = clock64();
  }
  if (x == 0x12345678u) sink[0] = x;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// ---- (2)+(3) MMA stream with optional contention ---------------------------------------------------------------------------
// smem: A 64 KB | B ring 2 x 32 KB | scratch 32 KB for st.shared
struct ProbeOut { long long mma_cycles, fill_cycles, fill_bytes; };

template <int kCta>
__global__ void __launch_bounds__(320, 1)
mma_kernel(int n_mma, int N, int fill, int n_fill, int n_store_warps, int n_ld_warps, const uint8_t* __restrict__ wsrc, uint32_t wbytes, ProbeOut* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_done, bar_fill[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ volatile int stop_flag;
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (160 * 1024) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (fill & 16) {
    // random bf16 operands in [-2, 2): tensor-core power depends on the data
    uint32_t x = 0x9E3779B9u * (threadIdx.x + 1) + blockIdx.x;
    for (int i = threadIdx.x; i < (160 * 1024) / 4; i += blockDim.x) {
      x = x * 1664525u + 1013904223u;
      const uint32_t lo = 0x3F00u | ((x >> 8) & 0x80FFu), hi = 0x3F00u | ((x >> 16) & 0x80FFu);
      reinterpret_cast<uint32_t*>(smem)[i] = lo | (hi << 16);
    }
  }
  fill &= 15;
  fence_proxy_async();
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar_done), 1);
    mbar_init(smem_u32(&bar_fill[0]), 1);
    mbar_init(smem_u32(&bar_fill[1]), 1);
    stop_flag = 0;
    fence_barrier_init();
  }
  if (kCta == 2) cluster_sync_all();
  if (warp == 0) { if (kCta == 2) tmem_alloc2(smem_u32(&tmem_base_s), 512); else tmem_alloc(smem_u32(&tmem_base_s), 512); }
  tc_fence_before();
  __syncthreads();
  if (kCta == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const bool leader = (kCta == 1) || cluster_ctarank() == 0;
  const int b_rows = (kCta == 2) ? N / 2 : N;                 // B rows held by this CTA
  if (warp == 0 && lane == 0) {
    long long t0 = clock64();
    if (leader) {
      const uint32_t idesc = make_idesc_m(128 * kCta, N);
      for (int i = 0; i < n_mma; ++i) {
        const int k = i & 3, panel = (i >> 2) & 3, st = (i >> 2) & 1;
        const uint64_t ad = make_desc_kmajor(sbase + panel * 16384 + k * 32);
        const uint64_t bd = make_desc_kmajor(sbase + 65536 + st * 32768 + k * 32);
        if (kCta == 2) umma2_bf16(tmem_base + (uint32_t)((i / 16) & 1) * 256u, ad, bd, idesc, 1u);
        else umma_bf16(tmem_base + (uint32_t)((i / 16) & 1) * 256u, ad, bd, idesc, 1u);
      }
      if (kCta == 2) umma2_commit_mc(smem_u32(&bar_done), 3); else umma_commit(smem_u32(&bar_done));
    }
    mbar_wait(smem_u32(&bar_done), 0);
    long long t1 = clock64();
    stop_flag = 1;
    if (blockIdx.x == 0) out->mma_cycles = t1 - t0;
  } else if (warp == 1 && lane == 0 && fill) {
    // producer: keep two bulk copies of b_rows*128 bytes in flight until the MMA stream ends (or n_fill copies when n_fill > 0)
    const uint32_t bytes = (uint32_t)b_rows * 128u;
    long long t0 = clock64();
    long long total = 0;
    uint32_t src_off = (blockIdx.x * 40960u) % wbytes;
    uint32_t g = 0;
    for (;; ++g) {
      const uint32_t st = g & 1u;
      if (g >= 2) mbar_wait(smem_u32(&bar_fill[st]), ((g >> 1) - 1) & 1u);
      const bool done = n_fill > 0 ? (int)g >= n_fill : stop_flag != 0;
      if (done) break;
      mbar_arrive_expect_tx(smem_u32(&bar_fill[st]), bytes);
      bulk_g2s(sbase + 65536 + st * 32768, wsrc + src_off, bytes, smem_u32(&bar_fill[st]));
      src_off += bytes; if (src_off + bytes > wbytes) src_off = 0;
      total += bytes;
    }
    if (g >= 1) mbar_wait(smem_u32(&bar_fill[(g - 1) & 1u]), ((g - 1) >> 1) & 1u);   // the copy still in flight
    long long t1 = clock64();
    if (blockIdx.x == 0) { out->fill_cycles = t1 - t0; out->fill_bytes = total; }
  } else if (warp >= 2 && warp < 2 + n_ld_warps) {
    // epilogue-like readers: tcgen05.ld of the accumulator the MMA stream is NOT writing would need exact phase
    // knowledge; read columns 256..511 lanes (warp % 4) while the MMAs alternate over both halves
    const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + 256u + (uint32_t)(((warp - 2) >> 2) & 1) * 128u;
    uint32_t x = 0;
    while (!stop_flag) {
      uint32_t a[32];
      tmem_ld32(taddr, a);
      tmem_ld32(taddr + 32, a);
      tmem_ld32(taddr + 64, a);
      tmem_ld32(taddr + 96, a);
      tmem_ld_wait();
      x ^= a[0];
    }
    if (x == 0x12345u) out->fill_bytes = x;
  } else if (warp >= 2 && warp < 2 + n_store_warps) {
    // epilogue-like writers: st.shared.v4 into the scratch region, 512 B per warp-instruction
    const uint32_t base = sbase + 131072 + (uint32_t)((warp - 2) & 7) * 4096 + lane * 16;
    uint32_t v = lane;
    while (!stop_flag) {
#pragma unroll
      for (int j = 0; j < 8; ++j) sts128(base + j * 512, make_uint4(v, v + 1, v + 2, v + 3));
      ++v;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (kCta == 2) cluster_sync_all();
  if (warp == 0) { if (kCta == 2) tmem_dealloc2(tmem_base, 512); else tmem_dealloc(tmem_base, 512); }
}

// ---- (4) operand major-ness: K-major vs MN-major (128-byte swizzle and un-swizzled), M = 128, N = 256 --------------------
__global__ void __launch_bounds__(128, 1) major_kernel(int n_mma, int a_mode, int b_mode, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const uint32_t sbase = smem_u32(smem);
  for (int i = threadIdx.x; i < (128 * 1024) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(smem_u32(&tmem_base_s), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (threadIdx.x == 0) {
    // modes: 0 = K-major SW128, 1 = MN-major SW128 (64-row slab: LBO 8192, SBO 1024), 2 = MN-major no swizzle
    // (chunk stride 1024, k-group stride 128), 3 = MN-major no swizzle (chunk stride 128, k-group stride 4096)
    const uint32_t idesc = make_idesc(256, a_mode ? 1 : 0, b_mode ? 1 : 0);
    const uint32_t a_addr = sbase, b_addr = sbase + 64 * 1024;
    long long t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      const int k = i & 3, st = (i >> 2) & 1;
      auto desc = [&](uint32_t base, int mode) -> uint64_t {
        if (mode == 0) return make_desc_kmajor(base + k * 32);
        if (mode == 1) return make_desc(base + k * 2048, 8192, 1024);
        uint64_t d = 0;
        const uint32_t addr = base + (mode == 2 ? k * 256 : k * 8192);
        const uint32_t lbo = mode == 2 ? 128 : 4096, sbo = mode == 2 ? 1024 : 128;
        d |= (uint64_t)((addr & 0x3FFFF) >> 4);
        d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
        d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
        d |= (uint64_t)1 << 46;
        return d;
      };
      umma_bf16(tmem_base + (uint32_t)((i / 16) & 1) * 256u, desc(a_addr + st * 32768, a_mode), desc(b_addr + st * 32768, b_mode), idesc, 1u);
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem_base, 512);
}

// ---- (5) the weight pipeline alone: producer (bulk G2S ring) -> MMA issuer (4 MMAs per stage, commit frees the stage) ------
__global__ void __launch_bounds__(128, 1)
pipe_kernel(int n_chunks, int n_stages, int stage_bytes, int N, int spin, int var, const uint8_t* __restrict__ wsrc, uint32_t wbytes, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[8], empty[8], done;
  __shared__ uint32_t tmem_base_s;
  const uint32_t sbase = smem_u32(smem);
  for (int i = threadIdx.x; i < (64 * 1024) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  if (threadIdx.x == 0) {
    for (int s = 0; s < 8; ++s) { mbar_init(smem_u32(&full[s]), 1); mbar_init(smem_u32(&empty[s]), 1); }
    mbar_init(smem_u32(&done), 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) tmem_alloc(smem_u32(&tmem_base_s), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t ring = sbase + 65536;
  if (threadIdx.x == 32) {           // producer
    uint32_t off = (blockIdx.x * 40960u) % wbytes;
    for (int g = 0; g < n_chunks; ++g) {
      const uint32_t s = g % n_stages, ph = (g / n_stages) & 1u;
      if (spin) mbar_wait_spin(smem_u32(&empty[s]), ph ^ 1u); else mbar_wait(smem_u32(&empty[s]), ph ^ 1u);
      if (var & 1) { mbar_arrive(smem_u32(&full[s])); continue; }
      mbar_arrive_expect_tx(smem_u32(&full[s]), stage_bytes);
      bulk_g2s(ring + s * stage_bytes, wsrc + off, stage_bytes, smem_u32(&full[s]));
      off += stage_bytes; if (off + stage_bytes > wbytes) off = 0;
    }
  } else if ((threadIdx.x >> 5) == 2 && (var & 32)) {
    // converged issuer warp (CUTLASS style): all 32 lanes run the loop and the waits, one elected lane issues
    const uint32_t idesc = make_idesc(N);
    long long t0 = clock64();
    const bool leader_lane = (threadIdx.x & 31) == 0;
    for (int g = 0; g < n_chunks; ++g) {
      const uint32_t s = g % n_stages, ph = (g / n_stages) & 1u;
      mbar_wait(smem_u32(&full[s]), ph);
      tc_fence_after();
      if (leader_lane) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_base + (uint32_t)((g >> 2) & 1) * 256u, make_desc_kmajor(sbase + ((g & 3) * 16384) + k * 32),
                    make_desc_kmajor(ring + s * stage_bytes + k * 32), idesc, 1u);
        umma_commit(smem_u32(&empty[s]));
      }
      __syncwarp();
    }
    if (leader_lane) {
      umma_commit(smem_u32(&done));
      mbar_wait(smem_u32(&done), 0);
      if (blockIdx.x == 0) out[0] = clock64() - t0;
    }
  } else if (threadIdx.x == 64 && !(var & 16)) {    // MMA issuer
    const uint32_t idesc = make_idesc(N);
    long long t0 = clock64();
    if (var & 8) {
      // software-pipelined: the wait for stage g+1 sits between the MMAs of stage g, the commit after its last MMA
      mbar_wait(smem_u32(&full[0]), 0);
      for (int g = 0; g < n_chunks; ++g) {
        const uint32_t s = g % n_stages;
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          umma_bf16(tmem_base + (uint32_t)((g >> 2) & 1) * 256u, make_desc_kmajor(sbase + ((g & 3) * 16384) + k * 32),
                    make_desc_kmajor(ring + s * stage_bytes + k * 32), idesc, 1u);
          if (k == 1 && g + 1 < n_chunks) mbar_wait(smem_u32(&full[(g + 1) % n_stages]), ((g + 1) / n_stages) & 1u);
        }
        umma_commit(smem_u32(&empty[s]));
      }
    } else {
      long long tw = 0, tm = 0, tc = 0;
      for (int g = 0; g < n_chunks; ++g) {
        const uint32_t s = g % n_stages, ph = (g / n_stages) & 1u;
        long long a = clock64();
        if (!(var & 2)) { if (spin) mbar_wait_spin(smem_u32(&full[s]), ph); else mbar_wait(smem_u32(&full[s]), ph); }
        tc_fence_after();
        long long b = clock64();
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_base + (uint32_t)((g >> 2) & 1) * 256u, make_desc_kmajor(sbase + ((g & 3) * 16384) + k * 32),
                    make_desc_kmajor(ring + s * stage_bytes + k * 32), idesc, 1u);
        long long c = clock64();
        umma_commit(smem_u32(&empty[s]));
        long long d = clock64();
        tw += b - a; tm += c - b; tc += d - c;
      }
      if (blockIdx.x == 0 && (var & 64))
        printf("   MMA thread per stage: wait full %lld, issue 4 MMAs %lld, commit %lld cycles\n", tw / n_chunks, tm / n_chunks, tc / n_chunks);
    }
    umma_commit(smem_u32(&done));
    mbar_wait(smem_u32(&done), 0);
    if (blockIdx.x == 0) out[0] = clock64() - t0;
  } else if ((threadIdx.x == 64 || threadIdx.x == 96) && (var & 16)) {
    // two issuer threads (different warps): even / odd stages, each with its own accumulator
    const int me = threadIdx.x == 96;
    const uint32_t idesc = make_idesc(N);
    long long t0 = clock64();
    for (int g = me; g < n_chunks; g += 2) {
      const uint32_t s = g % n_stages, ph = (g / n_stages) & 1u;
      mbar_wait(smem_u32(&full[s]), ph);
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16(tmem_base + (uint32_t)me * 256u, make_desc_kmajor(sbase + ((g & 3) * 16384) + k * 32),
                  make_desc_kmajor(ring + s * stage_bytes + k * 32), idesc, 1u);
      umma_commit(smem_u32(&empty[s]));
    }
    if (me == 0) {
      // the other thread's MMAs are not covered by this commit: give it time, then commit + wait
      umma_commit(smem_u32(&done));
      mbar_wait(smem_u32(&done), 0);
      if (blockIdx.x == 0) out[0] = clock64() - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem_base, 512);
}

// ---- (6) issue loop: what does a per-stage tcgen05.commit / tcgen05.fence cost? -------------------------------------------
__global__ void __launch_bounds__(128, 1) issue_kernel(int n_iter, int N, int mmas_per_iter, int do_fence, int do_commit, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t dummy[4], done;
  __shared__ uint32_t tmem_base_s;
  const uint32_t sbase = smem_u32(smem);
  for (int i = threadIdx.x; i < (128 * 1024) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  if (threadIdx.x == 0) {
    for (int s = 0; s < 4; ++s) mbar_init(smem_u32(&dummy[s]), 1);
    mbar_init(smem_u32(&done), 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) tmem_alloc(smem_u32(&tmem_base_s), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (threadIdx.x == 64) {
    const uint32_t idesc = make_idesc(N);
    long long t0 = clock64();
    for (int g = 0; g < n_iter; ++g) {
      if (do_fence) tc_fence_after();
      for (int k = 0; k < mmas_per_iter; ++k)
        umma_bf16(tmem_base + (uint32_t)((g >> 2) & 1) * 256u, make_desc_kmajor(sbase + ((g & 3) * 16384) + (k & 3) * 32),
                  make_desc_kmajor(sbase + 65536 + (g & 1) * 32768 + (k & 3) * 32), idesc, 1u);
      if (do_commit) umma_commit(smem_u32(&dummy[g & 3]));
    }
    umma_commit(smem_u32(&done));
    mbar_wait(smem_u32(&done), 0);
    if (blockIdx.x == 0) out[0] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem_base, 512);
}

int main(int argc, char** argv) {
  if (argc > 1 && argv[1][0] == 'i') {
    long long* d; cudaMalloc(&d, 64);
    const int smem = 132 * 1024;
    cudaFuncSetAttribute(issue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int N : {256, 128}) for (int per : {4, 8, 16}) for (int fence : {0, 1}) for (int commit : {0, 1}) {
      const int n_iter = 4096;
      issue_kernel<<<148, 128, smem>>>(n_iter, N, per, fence, commit, d);
      issue_kernel<<<148, 128, smem>>>(n_iter, N, per, fence, commit, d);
      cudaError_t err = cudaDeviceSynchronize();
      long long cyc; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
      printf("N=%3d %2d MMAs/iter fence %d commit %d: %7.1f cycles/iter = %6.1f cycles/MMA %s\n", N, per, fence, commit,
             (double)cyc / n_iter, (double)cyc / n_iter / per, cudaGetErrorString(err));
    }
    return 0;
  }
  if (argc > 1 && argv[1][0] == 'p') {
    long long* d; cudaMalloc(&d, 64);
    uint8_t* w; cudaMalloc(&w, 2u << 20); cudaMemset(w, 0, 2u << 20);
    const int smem = 196 * 1024;
    cudaFuncSetAttribute(pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int cfgs[][3] = {{2, 32768, 256}, {4, 32768, 256}, {4, 16384, 128}};
    // var: 1 = producer arrives without copying, 2 = MMA thread does not wait for full, 8 = software-pipelined waits,
    // 16 = two issuer threads
    for (auto& c : cfgs) for (int var : {64, 65, 66}) for (int grid : {148}) {
      const int n_chunks = 4096, spin = 0;
      pipe_kernel<<<grid, 128, smem>>>(n_chunks, c[0], c[1], c[2], spin, var, w, 2u << 20, d);
      pipe_kernel<<<grid, 128, smem>>>(n_chunks, c[0], c[1], c[2], spin, var, w, 2u << 20, d);
      cudaError_t err = cudaDeviceSynchronize();
      long long cyc; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
      printf("grid %3d ring %d x %5d B, N=%3d, var %d: %6.1f cycles per 4-MMA stage (ideal %d) %s\n", grid, c[0], c[1], c[2], var,
             (double)cyc / n_chunks, c[2] == 256 ? 512 : 267, cudaGetErrorString(err));
    }
    return 0;
  }
  if (argc > 1 && argv[1][0] == 't') {
    // MMA stream (N = 256, alternating accumulators) with 0 / 4 / 8 warps hammering tcgen05.ld
    ProbeOut* po; cudaMalloc(&po, sizeof(ProbeOut));
    uint8_t* w; cudaMalloc(&w, 2u << 20); cudaMemset(w, 0, 2u << 20);
    const int smem = 164 * 1024;
    cudaFuncSetAttribute(mma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int nld : {0, 4, 8}) for (int fill : {0, 1}) {
      ProbeOut h = {};
      for (int rep = 0; rep < 2; ++rep) mma_kernel<1><<<148, 320, smem>>>(8192, 256, fill, 0, 0, nld, w, 2u << 20, po);
      cudaError_t err = cudaDeviceSynchronize();
      cudaMemcpy(&h, po, sizeof(h), cudaMemcpyDeviceToHost);
      printf("MMA N=256 with %d tcgen05.ld warps, fill %d: %6.1f cycles/MMA %s\n", nld, fill, (double)h.mma_cycles / 8192, cudaGetErrorString(err));
    }
    return 0;
  }
  if (argc > 1 && argv[1][0] == 'r') {
    // zeros vs random operands: cycles per MMA and wall-clock rate (power throttling shows up in the clock)
    ProbeOut* po; cudaMalloc(&po, sizeof(ProbeOut));
    uint8_t* w; cudaMalloc(&w, 2u << 20); cudaMemset(w, 0, 2u << 20);
    const int smem = 164 * 1024;
    cudaFuncSetAttribute(mma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int rnd : {0, 16, 0, 16}) {
      ProbeOut h = {};
      const int n_mma = 1 << 18;
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      mma_kernel<1><<<148, 320, smem>>>(n_mma, 256, rnd, 0, 0, 0, w, 2u << 20, po);
      cudaEventRecord(e0);
      mma_kernel<1><<<148, 320, smem>>>(n_mma, 256, rnd, 0, 0, 0, w, 2u << 20, po);
      cudaEventRecord(e1);
      cudaError_t err = cudaDeviceSynchronize();
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      cudaMemcpy(&h, po, sizeof(h), cudaMemcpyDeviceToHost);
      printf("%s operands: %6.1f cycles/MMA, %.2f ms -> %.0f TFLOP/s, SM clock %.2f GHz %s\n", rnd ? "random" : "zero  ",
             (double)h.mma_cycles / n_mma, ms, 2.0 * 128 * 256 * 16 * (double)n_mma * 148 / ms / 1e9, h.mma_cycles / (ms * 1e6), cudaGetErrorString(err));
    }
    return 0;
  }
  if (argc > 1 && argv[1][0] == 'm') {
    long long* d; cudaMalloc(&d, 64);
    const int smem = 132 * 1024;
    cudaFuncSetAttribute(major_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const char* names[4] = {"K-major SW128", "MN-major SW128", "MN-major nosw(chunk 1024)", "MN-major nosw(chunk 128)"};
    for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b) {
      major_kernel<<<148, 128, smem>>>(4096, a, b, d);
      major_kernel<<<148, 128, smem>>>(4096, a, b, d);
      cudaError_t err = cudaDeviceSynchronize();
      long long cyc; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
      printf("A %-26s B %-26s: %6.1f cycles/MMA (M=128 N=256 K=16) %s\n", names[a], names[b], (double)cyc / 4096, cudaGetErrorString(err));
    }
    return 0;
  }
  long long* d; cudaMalloc(&d, 64);
  uint32_t* sink; cudaMalloc(&sink, 64);
  ProbeOut* po; cudaMalloc(&po, sizeof(ProbeOut));
  const uint32_t wbytes = 2u << 20;
  uint8_t* w; cudaMalloc(&w, wbytes); cudaMemset(w, 0, wbytes);
  // (1)
  for (int nw : {4, 8, 16}) {
    const int n_iter = 4096;
    tmem_ld_kernel<<<148, 512>>>(n_iter, nw, d, sink);
    tmem_ld_kernel<<<148, 512>>>(n_iter, nw, d, sink);
    cudaError_t err = cudaDeviceSynchronize();
    long long cyc; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
    double bytes = (double)n_iter * 2 * 32 * 32 * 4 * nw;
    printf("tcgen05.ld 32x32b.x32, %2d warps: %.1f B/clk/SM (%.0f cycles per 128 KB accumulator) %s\n", nw, bytes / cyc,
           131072.0 / (bytes / cyc), cudaGetErrorString(err));
  }
  // (2)+(3)
  const int smem = 164 * 1024;
  cudaFuncSetAttribute(mma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(mma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(mma_kernel<2>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  const int n_mma = 8192;
  for (int cta : {1, 2}) for (int N : {256, 128}) for (int fill : {0, 1}) for (int nsw : {0, 8}) {
    ProbeOut h = {};
    cudaMemset(po, 0, sizeof(ProbeOut));
    for (int rep = 0; rep < 2; ++rep) {
      if (cta == 1) {
        mma_kernel<1><<<148, 320, smem>>>(n_mma, N, fill, 0, nsw, 0, w, wbytes, po);
      } else {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(148); cfg.blockDim = dim3(320); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, mma_kernel<2>, n_mma, N, fill, 0, nsw, 0, (const uint8_t*)w, wbytes, po);
      }
    }
    cudaError_t err = cudaDeviceSynchronize();
    cudaMemcpy(&h, po, sizeof(h), cudaMemcpyDeviceToHost);
    double per_sm_flop = 2.0 * 128 * N * 16;   // per SM per MMA (cta_group::2: 256 rows over two SMs)
    printf("cta_group::%d N %3d fill %d store_warps %d: %6.1f cycles/MMA (%.0f FLOP/clk/SM)  fill %.1f B/clk/SM  %s\n", cta, N, fill, nsw,
           (double)h.mma_cycles / n_mma, per_sm_flop / ((double)h.mma_cycles / n_mma),
           h.fill_cycles ? (double)h.fill_bytes / h.fill_cycles : 0.0, cudaGetErrorString(err));
  }
  // L2 -> SM bulk-copy bandwidth alone (no MMA): n_fill copies
  for (int cta : {1}) for (int N : {256, 128}) {
    ProbeOut h = {};
    cudaMemset(po, 0, sizeof(ProbeOut));
    mma_kernel<1><<<148, 320, smem>>>(0, N, 1, 64, 0, 0, w, wbytes, po);
    cudaMemset(po, 0, sizeof(ProbeOut));
    mma_kernel<1><<<148, 320, smem>>>(0, N, 1, 4096, 0, 0, w, wbytes, po);
    cudaError_t err = cudaDeviceSynchronize();
    cudaMemcpy(&h, po, sizeof(h), cudaMemcpyDeviceToHost);
    printf("bulk G2S from L2, %5d-byte copies, 2 in flight, 148 SMs: %.1f B/clk/SM %s\n", N * 128,
           h.fill_cycles ? (double)h.fill_bytes / h.fill_cycles : 0.0, cudaGetErrorString(err));
  }
  return 0;
}
