// Per-sample arithmetic of ray_marching (src/UtilsNeuralRadianceField.py:88-115) shared by the compositing kernels
// (composite.cu) and the hierarchical-sampling kernel that forms the coarse weights itself (sampler.cu): ONE definition,
// so the weights agree bit for bit wherever they are computed.
#pragma once
#include "common.cuh"

namespace nerf {

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ float warp_incl_scan_mul(float v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    float o = __shfl_up_sync(kFull, v, d);
    if (lane >= d) v *= o;
  }
  return v;
}

__device__ __forceinline__ float warp_incl_rscan_add(float v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    float o = __shfl_down_sync(kFull, v, d);
    if (lane + d < 32) v += o;
  }
  return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(kFull, v, d);
  return v;
}

// sigmoid and exp(-sigma*delta) on the SFU (ex2.approx / rcp.approx, ~2 ulp): the kernels are issue-bound, not
// HBM-bound, with libm's expf and an IEEE division per colour channel; the absolute error (< 2e-7) is below the
// 2e-6 parity bound of the compositing tests and far below the 1e-5 render bound.  The .ftz forms are used directly:
// __expf / __fdividef wrap every MUFU in a denormal-range test and two predicated scalings (three issue slots per
// call, twelve per 32-sample block), and a flushed denormal changes nothing here (1 + 1e-39 and 1 - 1e-39 are 1).
__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_ftz(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
constexpr float kLog2e = 1.4426950408889634f;
__device__ __forceinline__ float exp_neg(float x) { return ex2_ftz(x * -kLog2e); }          // exp(-x)
__device__ __forceinline__ float sigmoidf_(float x) { return rcp_ftz(1.0f + exp_neg(x)); }

// per-sample forward quantities
struct SampleFwd {
  float sigma, delta, alpha, x;  // x = 1 - alpha
};

__device__ __forceinline__ SampleFwd sample_fwd(float raw_sigma, float z_cur, float z_next, bool last) {
  SampleFwd r;
  r.sigma = fmaxf(raw_sigma, 0.f);
  r.delta = last ? 1e9f : z_next - z_cur;
  r.alpha = 1.0f - exp_neg(r.sigma * r.delta);
  r.x = 1.0f - r.alpha;
  return r;
}

}  // namespace nerf
