// The gradient exchange of the sharded train step ("one all-reduce of the MLP gradients per step",
// src/NeRF.py:160-167 under a ray-sharded batch) as kernels of our own over NVLink peer memory instead of an NCCL call:
//
//   nerf_peer_barrier       every GPU of the box publishes "my gradient buffer is final" in the other GPUs' signal pads and
//                           waits for theirs (system-scope release / acquire, one thread per peer);
//   nerf_peer_reduce_adam   ONE-SHOT all-reduce fused with the optimizer: every GPU reads the slice of every peer's flat
//                           gradient buffer straight over NVLink (P2P loads through NVSwitch, 16 bytes per lane), adds the
//                           ranks in rank order -- the same order on every GPU, so the replicas stay bit-identical without
//                           any broadcast -- and applies the Keras-2.7 Adam update (nerf_adam_step's arithmetic) in the same
//                           pass.  2 MB per network: the exchange is latency-bound, and one kernel replaces ring all-reduce +
//                           Adam kernel and their two launch gaps on the critical path at the end of the step.
//
// The buffers are symmetric allocations (torch.distributed._symmetric_memory on the host side: same size on every GPU,
// peer-mapped); the caller passes DEVICE arrays of the per-rank base pointers.  Write-after-read safety is the caller's:
// the host package alternates between two gradient buffers, so a buffer is rewritten two barriers after its last remote read.
#include <math.h>

#include "common.cuh"

namespace nerf {

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer_f4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_peer_f1(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}

// pads[r]: rank r's signal pad (uint32 words).  Word slot * 64 + s of rank r's pad is written by rank s only.
// Epochs only grow (the caller passes its step counter), so the pads never need a reset.
__global__ void peer_barrier_kernel(uint32_t* const* __restrict__ pads, int rank, int world, uint32_t epoch, int slot) {
  const int peer = threadIdx.x;
  if (peer >= world) return;
  __threadfence_system();                                  // this GPU's earlier kernels' writes, system-wide
  st_release_sys(pads[peer] + slot * 64 + rank, epoch);
  const uint32_t* mine = pads[rank] + slot * 64 + peer;
  long long t0 = clock64();
  while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) {
    __nanosleep(200);
    if (clock64() - t0 > 40000000000LL) {                  // ~20 s: a rank died or the ranks disagree on the sequence
      printf("nerf_b200: peer barrier timeout (rank %d waits for rank %d, slot %d, epoch %u)\n", rank, peer, slot, epoch);
      __trap();
    }
  }
}

__device__ __forceinline__ void adam_update(float& p, float g, float& m, float& v, float lr_t, float omb1, float omb2, float eps) {
  const float mi = m + (g - m) * omb1;                     // same arithmetic as adam_kernel (optim.cu)
  const float vi = v + (g * g - v) * omb2;
  m = mi;
  v = vi;
  p = p - lr_t * mi / (sqrtf(vi) + eps);
}

// grads[r] + offset: the slice on rank r.  params == nullptr: reduction only (reduced must be given).
__global__ void __launch_bounds__(256)
peer_reduce_adam_kernel(float* __restrict__ params, const float* const* __restrict__ grads, int world, int64_t offset, int64_t n,
                        float* __restrict__ m, float* __restrict__ v, float lr_t, float omb1, float omb2, float eps,
                        float* __restrict__ reduced) {
  const int64_t i4 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) * 4;
  if (i4 >= n) return;
  if (i4 + 4 <= n) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < world; ++r) {                      // rank order: identical sums on every GPU
      const float4 g = ld_peer_f4(grads[r] + offset + i4);
      s.x += g.x; s.y += g.y; s.z += g.z; s.w += g.w;
    }
    if (reduced) *reinterpret_cast<float4*>(reduced + i4) = s;
    if (params) {
      float4 p = *reinterpret_cast<float4*>(params + i4), mm = *reinterpret_cast<float4*>(m + i4),
             vv = *reinterpret_cast<float4*>(v + i4);
      adam_update(p.x, s.x, mm.x, vv.x, lr_t, omb1, omb2, eps);
      adam_update(p.y, s.y, mm.y, vv.y, lr_t, omb1, omb2, eps);
      adam_update(p.z, s.z, mm.z, vv.z, lr_t, omb1, omb2, eps);
      adam_update(p.w, s.w, mm.w, vv.w, lr_t, omb1, omb2, eps);
      *reinterpret_cast<float4*>(params + i4) = p;
      *reinterpret_cast<float4*>(m + i4) = mm;
      *reinterpret_cast<float4*>(v + i4) = vv;
    }
  } else {
    for (int64_t i = i4; i < n; ++i) {
      float s = 0.f;
      for (int r = 0; r < world; ++r) s += ld_peer_f1(grads[r] + offset + i);
      if (reduced) reduced[i] = s;
      if (params) adam_update(params[i], s, m[i], v[i], lr_t, omb1, omb2, eps);
    }
  }
}

}  // namespace nerf

using namespace nerf;

extern "C" {

int nerf_peer_barrier(void* const* pads_dev, int32_t rank, int32_t world, uint32_t epoch, int32_t slot, void* stream) {
  NERF_CHECK_ARG(pads_dev, "null pointer");
  NERF_CHECK_ARG(world >= 1 && world <= 64 && rank >= 0 && rank < world && slot >= 0 && slot < 4, "bad rank / world / slot");
  peer_barrier_kernel<<<1, 64, 0, (cudaStream_t)stream>>>((uint32_t* const*)pads_dev, rank, world, epoch, slot);
  NERF_CHECK_LAUNCH();
  return NERF_OK;
}

int nerf_peer_reduce_adam(float* params_or_null, void* const* grads_dev, int32_t world, int64_t offset, int64_t n, float* m,
                          float* v, float lr, float beta1, float beta2, float eps, int64_t t, float* reduced_or_null,
                          void* stream) {
  NERF_CHECK_ARG(grads_dev && (params_or_null || reduced_or_null), "null pointer");
  NERF_CHECK_ARG(!params_or_null || (m && v && t >= 1), "Adam needs its moments and the 1-based step");
  NERF_CHECK_ARG(world >= 1 && world <= 64 && n >= 0 && offset >= 0 && (offset & 3) == 0, "bad world / size / offset (16-byte aligned slices)");
  NERF_CHECK_ARG(!params_or_null || ((((uintptr_t)params_or_null | (uintptr_t)m | (uintptr_t)v) & 15) == 0), "16-byte aligned vectors");
  if (n == 0) return NERF_OK;
  double lr_t = 0.0;
  if (params_or_null) lr_t = (double)lr * sqrt(1.0 - pow((double)beta2, (double)t)) / (1.0 - pow((double)beta1, (double)t));
  const int64_t threads = ceil_div(n, 4);
  peer_reduce_adam_kernel<<<(unsigned)ceil_div(threads, 256), 256, 0, (cudaStream_t)stream>>>(
      params_or_null, (const float* const*)grads_dev, world, offset, n, m, v, (float)lr_t, 1.0f - beta1, 1.0f - beta2, eps,
      reduced_or_null);
  NERF_CHECK_LAUNCH();
  return NERF_OK;
}

}  // extern "C"
