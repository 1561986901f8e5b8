// K3: alpha compositing (ray_marching, src/UtilsNeuralRadianceField.py:88-115) forward and backward.
//
// HBM-bound: 44 B/sample + 12 B/ray with every reference output, 24 B/sample lean.
// One warp per ray (= one scan segment).  Sample s of the ray lives in lane (s % 32), block (s / 32), so every
// load/store instruction of the warp touches one contiguous 128/512-byte span.  The exclusive cumprod of
// (1-alpha) is a Kogge-Stone warp-shuffle scan per 32-sample block with a running carry across blocks; the
// backward needs the mirrored reverse exclusive cumsum of g*w.
#include "composite.cuh"

namespace nerf {

// Fused loss (kLoss): keras MeanSquaredError against `target` right where the ray's colour is reduced
// (src/NeRF.py:151,157): d_rgb = grad_scale * (rgb - target), sq_err_sum += sum (rgb - target)^2 (one atomic per block).
struct LossArgs {
  const float* target;   // (N,3)
  float grad_scale;      // 2 * loss_weight / (3 * n_total_rays)
  float* sq_err_sum;     // scalar accumulator (caller zeroes)
  float* d_rgb;          // (N,3) out, may be null
};

// block-wide sum of one value per warp (lane 0 holds it) -> one atomicAdd; every thread of the block must call it
__device__ __forceinline__ void block_accumulate(float warp_value, int lane, float* target) {
  __shared__ float part[8];
  const int warp = threadIdx.x >> 5;
  if (lane == 0) part[warp] = warp_value;
  __syncthreads();
  if (threadIdx.x == 0 && target) {
    float sum = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) sum += part[w];
    atomicAdd(target, sum);
  }
}

// kExtras = false: cumprod, alpha_out and rgb_s are all null (the lean call of render / train), so the per-block pointer
// tests (two 64-bit compares each) are compiled out.  kExact: S == 32 C (64, 128, 192, 256: every train / render shape
// of the reference configs), so every `s < S` test, its selects and the predicate spills they cause are compiled out too.
#ifndef NERF_COMPOSITE_FWD_MIN_BLOCKS
#define NERF_COMPOSITE_FWD_MIN_BLOCKS 5   // 48 registers: 40 resident warps per SM instead of 32 (S = 192 lean: 8 bytes of spill)
#endif
template <int C, bool kLoss, bool kExtras, bool kExact>
__global__ void __launch_bounds__(256, (C <= 4 || (C <= 6 && !kExtras && !kLoss)) ? NERF_COMPOSITE_FWD_MIN_BLOCKS : 1)
composite_fwd_kernel(const float4* __restrict__ raw4, const float* __restrict__ z,
                                                            int64_t n_rays, int S, float* __restrict__ rgb,
                                                            float* __restrict__ weights, float* __restrict__ cumprod,
                                                            float* __restrict__ alpha_out, float* __restrict__ rgb_s,
                                                            float* __restrict__ depth, float* __restrict__ acc,
                                                            LossArgs loss) {
  const int lane = threadIdx.x & 31;
  int64_t ray = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  const bool active = ray < n_rays;
  if (!kLoss && !active) return;
  if (!active) ray = n_rays - 1;           // kLoss: keep the warp for the block-wide reduction; it stores nothing
  const float4* raw_r = raw4 + ray * S;
  const float* z_r = z + ray * S;

  float4 raw[C];
  float zc[C];
#pragma unroll
  for (int k = 0; k < C; ++k) {
    int s = k * 32 + lane;
    if (kExact || s < S) {
      raw[k] = __ldcs(raw_r + s);
      zc[k] = __ldcs(z_r + s);
    } else {
      raw[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      zc[k] = 0.f;
    }
  }
  float carry = 1.0f;
  float r_acc = 0.f, g_acc = 0.f, b_acc = 0.f, d_acc = 0.f, a_acc = 0.f;
#pragma unroll
  for (int k = 0; k < C; ++k) {
    int s = k * 32 + lane;
    if (!kExact && k * 32 >= S) break;
    // z of the next sample: lane+1 of this block, or lane 0 of the next block
    float z_next = __shfl_down_sync(kFull, zc[k], 1);
    float z_next_blk = (k + 1 < C) ? __shfl_sync(kFull, zc[(k + 1 < C) ? k + 1 : k], 0) : 0.f;
    if (lane == 31) z_next = z_next_blk;
    const bool valid = kExact || s < S;
    SampleFwd f = sample_fwd(raw[k].w, zc[k], z_next, kExact ? (k == C - 1 && lane == 31) : (s == S - 1));
    float x = valid ? f.x : 1.0f;
    float incl = warp_incl_scan_mul(x, lane);
    float excl = __shfl_up_sync(kFull, incl, 1);
    if (lane == 0) excl = 1.0f;
    float T = carry * excl;
    carry *= __shfl_sync(kFull, incl, 31);
    if (valid && active) {
      float w = f.alpha * T;
      float cr = sigmoidf_(raw[k].x), cg = sigmoidf_(raw[k].y), cb = sigmoidf_(raw[k].z);
      r_acc = __fmaf_rn(w, cr, r_acc);      // explicit contractions: every instantiation (fused loss or not, exact
      g_acc = __fmaf_rn(w, cg, g_acc);      // block count or not) must round identically - the fused-kernel tests
      b_acc = __fmaf_rn(w, cb, b_acc);      // compare them bit for bit
      d_acc = __fmaf_rn(w, zc[k], d_acc);
      a_acc += w;
      int64_t o = ray * S + s;
      if (weights) __stcs(weights + o, w);
      if (kExtras && cumprod) __stcs(cumprod + o, T);
      if (kExtras && alpha_out) __stcs(alpha_out + o, f.alpha);
      if (kExtras && rgb_s) {
        __stcs(rgb_s + o * 3 + 0, cr);
        __stcs(rgb_s + o * 3 + 1, cg);
        __stcs(rgb_s + o * 3 + 2, cb);
      }
    }
  }
  r_acc = warp_sum(r_acc);
  g_acc = warp_sum(g_acc);
  b_acc = warp_sum(b_acc);
  if (depth) d_acc = warp_sum(d_acc);
  if (acc) a_acc = warp_sum(a_acc);
  if (lane == 0 && active) {
    if (rgb) {
      rgb[ray * 3 + 0] = r_acc;
      rgb[ray * 3 + 1] = g_acc;
      rgb[ray * 3 + 2] = b_acc;
    }
    if (depth) depth[ray] = d_acc;
    if (acc) acc[ray] = a_acc;
  }
  if (kLoss) {
    float e = 0.f;
    if (lane == 0 && active) {
      const float er = r_acc - loss.target[ray * 3 + 0], eg = g_acc - loss.target[ray * 3 + 1],
                  eb = b_acc - loss.target[ray * 3 + 2];
      e = er * er + eg * eg + eb * eb;
      if (loss.d_rgb) {
        loss.d_rgb[ray * 3 + 0] = loss.grad_scale * er;
        loss.d_rgb[ray * 3 + 1] = loss.grad_scale * eg;
        loss.d_rgb[ray * 3 + 2] = loss.grad_scale * eb;
      }
    }
    block_accumulate(e, lane, loss.sq_err_sum);
  }
}

// kLoss: the ray's colour, its MSE against `target` and d_rgb are computed HERE from the forward scan the backward
// repeats anyway, so composite_fwd + mse + composite_bwd of the fine network are one launch (rgb_out optional).
#ifndef NERF_COMPOSITE_BWD6_MIN_BLOCKS
#define NERF_COMPOSITE_BWD6_MIN_BLOCKS 3
#endif
template <int C, bool kLoss, bool kExact>
__global__ void __launch_bounds__(256, (C <= 4 ? 4 : (C <= 6 ? NERF_COMPOSITE_BWD6_MIN_BLOCKS : 1)))
composite_bwd_kernel(const float4* __restrict__ raw4, const float* __restrict__ z,
                                                            const float* __restrict__ d_rgb,
                                                            const float* __restrict__ d_weights, int64_t n_rays, int S,
                                                            float4* __restrict__ d_raw4, float* __restrict__ d_z,
                                                            LossArgs loss, float* __restrict__ rgb_out) {
  const int lane = threadIdx.x & 31;
  int64_t ray = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  const bool active = ray < n_rays;
  if (!kLoss && !active) return;
  if (!active) ray = n_rays - 1;
  const float4* raw_r = raw4 + ray * S;
  const float* z_r = z + ray * S;
  float dr = 0.f, dg = 0.f, db = 0.f;
  if (!kLoss) { dr = __ldg(d_rgb + ray * 3 + 0); dg = __ldg(d_rgb + ray * 3 + 1); db = __ldg(d_rgb + ray * 3 + 2); }

  // Per 32-sample block only EIGHT values stay live between the phases (transmittance, weight, 1 - alpha, delta, sigma and
  // the three colours; each is overwritten by its successor - g T, g w, the colour gradients, d delta - as soon as it is
  // dead), so S = 192 fits 64 registers = 4 resident blocks per SM: the kernel had become latency-bound at 3 (ncu: 33 % of
  // the warp slots, 36 % of the issue slots, 59 % of DRAM).
  float4 raw[C];
  float zc[C], T[C], w[C], x[C], dl[C], sg[C], cr[C], cg[C], cb[C];
#pragma unroll
  for (int k = 0; k < C; ++k) {
    int s = k * 32 + lane;
    if (kExact || s < S) {
      raw[k] = __ldcs(raw_r + s);
      zc[k] = __ldcs(z_r + s);
    } else {
      raw[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      zc[k] = 0.f;
    }
  }
  // forward scan: T, w, colours (and, fused, the ray's colour -> loss -> d_rgb)
  float carry = 1.0f;
  float r_acc = 0.f, g_acc = 0.f, b_acc = 0.f;
#pragma unroll
  for (int k = 0; k < C; ++k) {
    int s = k * 32 + lane;
    float z_next = __shfl_down_sync(kFull, zc[k], 1);
    float z_next_blk = __shfl_sync(kFull, zc[(k + 1 < C) ? k + 1 : k], 0);
    if (lane == 31) z_next = z_next_blk;
    const bool valid = kExact || s < S;
    const SampleFwd f = sample_fwd(raw[k].w, zc[k], z_next, kExact ? (k == C - 1 && lane == 31) : (s == S - 1));
    x[k] = f.x;
    dl[k] = f.delta;
    sg[k] = f.sigma;
    float incl = warp_incl_scan_mul(valid ? f.x : 1.0f, lane);
    float excl = __shfl_up_sync(kFull, incl, 1);
    if (lane == 0) excl = 1.0f;
    T[k] = carry * excl;
    carry *= __shfl_sync(kFull, incl, 31);
    w[k] = f.alpha * T[k];
    cr[k] = sigmoidf_(raw[k].x);
    cg[k] = sigmoidf_(raw[k].y);
    cb[k] = sigmoidf_(raw[k].z);
    if (kLoss && valid) {
      r_acc = __fmaf_rn(w[k], cr[k], r_acc);
      g_acc = __fmaf_rn(w[k], cg[k], g_acc);
      b_acc = __fmaf_rn(w[k], cb[k], b_acc);
    }
  }
  if (kLoss) {
    r_acc = warp_sum(r_acc);
    g_acc = warp_sum(g_acc);
    b_acc = warp_sum(b_acc);
    const float er = r_acc - __ldg(loss.target + ray * 3 + 0), eg = g_acc - __ldg(loss.target + ray * 3 + 1),
                eb = b_acc - __ldg(loss.target + ray * 3 + 2);
    dr = loss.grad_scale * er;
    dg = loss.grad_scale * eg;
    db = loss.grad_scale * eb;
    if (lane == 0 && active) {
      if (rgb_out) { rgb_out[ray * 3 + 0] = r_acc; rgb_out[ray * 3 + 1] = g_acc; rgb_out[ray * 3 + 2] = b_acc; }
      if (loss.d_rgb) { loss.d_rgb[ray * 3 + 0] = dr; loss.d_rgb[ray * 3 + 1] = dg; loss.d_rgb[ray * 3 + 2] = db; }
    }
    block_accumulate((lane == 0 && active) ? er * er + eg * eg + eb * eb : 0.f, lane, loss.sq_err_sum);
  }
  // g = dL/dw per sample; from here on T holds g T, w holds g w and the colours hold the colour gradients
  // (every product written out: the loss-fused and the plain instantiations must round identically)
#pragma unroll
  for (int k = 0; k < C; ++k) {
    int s = k * 32 + lane;
    const bool valid = kExact || s < S;
    float gi = __fmaf_rn(db, cb[k], __fmaf_rn(dg, cg[k], __fmul_rn(dr, cr[k])));
    if (d_weights && valid) gi += __ldcs(d_weights + ray * S + s);
    if (!valid) gi = 0.f;
    cr[k] = __fmul_rn(__fmul_rn(__fmul_rn(w[k], dr), cr[k]), 1.f - cr[k]);
    cg[k] = __fmul_rn(__fmul_rn(__fmul_rn(w[k], dg), cg[k]), 1.f - cg[k]);
    cb[k] = __fmul_rn(__fmul_rn(__fmul_rn(w[k], db), cb[k]), 1.f - cb[k]);
    T[k] = __fmul_rn(gi, T[k]);
    w[k] = __fmul_rn(gi, w[k]);
  }
  // reverse exclusive cumsum of g*w, then the per-sample gradients (dl becomes d delta)
  float rcarry = 0.f;
#pragma unroll
  for (int k = C - 1; k >= 0; --k) {
    int s = k * 32 + lane;
    const bool valid = kExact || s < S;
    float incl = warp_incl_rscan_add(w[k], lane);
    float excl = __shfl_down_sync(kFull, incl, 1);
    if (lane == 31) excl = 0.f;
    float R = rcarry + excl;
    rcarry += __shfl_sync(kFull, incl, 0);
    // TF: d cumprod / d x = div_no_nan(R, x); d alpha = g*T - that.  x = 1 - alpha is 0 or >= 2^-24, never denormal
    const float dalpha = __fsub_rn(T[k], x[k] == 0.f ? 0.f : __fmul_rn(R, rcp_ftz(x[k])));
    const float dax = __fmul_rn(dalpha, x[k]);
    const float dsig = __fmul_rn(dax, dl[k]);
    const float ddel = (kExact ? (k == C - 1 && lane == 31) : (s == S - 1)) ? 0.f : __fmul_rn(dax, sg[k]);
    dl[k] = valid ? ddel : 0.f;
    if (valid && active) __stcs(d_raw4 + ray * S + s, make_float4(cr[k], cg[k], cb[k], sg[k] > 0.f ? dsig : 0.f));
  }
  float (&ddelta)[C] = dl;
  if (d_z) {
    // d z_s = d delta_{s-1} - d delta_s
#pragma unroll
    for (int k = 0; k < C; ++k) {
      int s = k * 32 + lane;
      float prev = __shfl_up_sync(kFull, ddelta[k], 1);
      float prev_blk = __shfl_sync(kFull, ddelta[(k > 0) ? k - 1 : 0], 31);
      if (lane == 0) prev = (k > 0) ? prev_blk : 0.f;
      if ((kExact || s < S) && active) __stcs(d_z + ray * S + s, prev - ddelta[k]);
    }
  }
}

template <int C, bool kExact>
static int launch_fwd_(const float* raw4, const float* z, int64_t n, int S, float* rgb, float* w, float* T, float* a,
                       float* rgb_s, float* depth, float* acc, const LossArgs* loss, cudaStream_t st) {
  const int warps = 8;
  const unsigned grid = (unsigned)ceil_div(n, warps);
  if (loss)
    composite_fwd_kernel<C, true, false, kExact><<<grid, warps * 32, 0, st>>>((const float4*)raw4, z, n, S, rgb, w, nullptr,
                                                                              nullptr, nullptr, depth, acc, *loss);
  else if (T || a || rgb_s)
    composite_fwd_kernel<C, false, true, kExact><<<grid, warps * 32, 0, st>>>((const float4*)raw4, z, n, S, rgb, w, T, a,
                                                                              rgb_s, depth, acc, LossArgs{});
  else
    composite_fwd_kernel<C, false, false, kExact><<<grid, warps * 32, 0, st>>>((const float4*)raw4, z, n, S, rgb, w, nullptr,
                                                                               nullptr, nullptr, depth, acc, LossArgs{});
  return 0;
}
template <int C>
static int launch_fwd(const float* raw4, const float* z, int64_t n, int S, float* rgb, float* w, float* T, float* a,
                      float* rgb_s, float* depth, float* acc, const LossArgs* loss, cudaStream_t st) {
  return S == 32 * C ? launch_fwd_<C, true>(raw4, z, n, S, rgb, w, T, a, rgb_s, depth, acc, loss, st)
                     : launch_fwd_<C, false>(raw4, z, n, S, rgb, w, T, a, rgb_s, depth, acc, loss, st);
}
template <int C, bool kExact>
static int launch_bwd_(const float* raw4, const float* z, const float* d_rgb, const float* d_w, int64_t n, int S,
                       float* d_raw4, float* d_z, const LossArgs* loss, float* rgb_out, cudaStream_t st) {
  const int warps = 8;
  const unsigned grid = (unsigned)ceil_div(n, warps);
  if (loss)
    composite_bwd_kernel<C, true, kExact><<<grid, warps * 32, 0, st>>>((const float4*)raw4, z, nullptr, d_w, n, S,
                                                                       (float4*)d_raw4, d_z, *loss, rgb_out);
  else
    composite_bwd_kernel<C, false, kExact><<<grid, warps * 32, 0, st>>>((const float4*)raw4, z, d_rgb, d_w, n, S,
                                                                        (float4*)d_raw4, d_z, LossArgs{}, nullptr);
  return 0;
}
template <int C>
static int launch_bwd(const float* raw4, const float* z, const float* d_rgb, const float* d_w, int64_t n, int S,
                      float* d_raw4, float* d_z, const LossArgs* loss, float* rgb_out, cudaStream_t st) {
  return S == 32 * C ? launch_bwd_<C, true>(raw4, z, d_rgb, d_w, n, S, d_raw4, d_z, loss, rgb_out, st)
                     : launch_bwd_<C, false>(raw4, z, d_rgb, d_w, n, S, d_raw4, d_z, loss, rgb_out, st);
}

}  // namespace nerf

using namespace nerf;

#define DISPATCH_C(S, CALL)                 \
  do {                                      \
    int c__ = ((S) + 31) / 32;              \
    if (c__ <= 1) { CALL(1); }              \
    else if (c__ <= 2) { CALL(2); }         \
    else if (c__ <= 4) { CALL(4); }         \
    else if (c__ <= 6) { CALL(6); }         \
    else if (c__ <= 8) { CALL(8); }         \
    else if (c__ <= 16) { CALL(16); }       \
    else { CALL(32); }                      \
  } while (0)

extern "C" {

int nerf_composite_fwd(const float* raw4, const float* z, int64_t n_rays, int32_t n_samples, float* rgb, float* weights,
                       float* cumprod, float* alpha, float* rgb_s, float* depth, float* acc, void* stream) {
  NERF_CHECK_ARG(raw4 && z, "null input");
  NERF_CHECK_ARG(n_rays >= 0 && n_samples > 0 && n_samples <= 1024, "n_samples must be in [1,1024]");
  if (n_rays == 0) return NERF_OK;
  cudaStream_t st = (cudaStream_t)stream;
#define CALL(C) launch_fwd<C>(raw4, z, n_rays, n_samples, rgb, weights, cumprod, alpha, rgb_s, depth, acc, nullptr, st)
  DISPATCH_C(n_samples, CALL);
#undef CALL
  NERF_CHECK_LAUNCH();
  return NERF_OK;
}

int nerf_composite_bwd(const float* raw4, const float* z, const float* d_rgb, const float* d_weights_or_null,
                       int64_t n_rays, int32_t n_samples, float* d_raw4, float* d_z_or_null, void* stream) {
  NERF_CHECK_ARG(raw4 && z && d_rgb && d_raw4, "null pointer");
  NERF_CHECK_ARG(n_rays >= 0 && n_samples > 0 && n_samples <= 1024, "n_samples must be in [1,1024]");
  if (n_rays == 0) return NERF_OK;
  cudaStream_t st = (cudaStream_t)stream;
#define CALL(C) launch_bwd<C>(raw4, z, d_rgb, d_weights_or_null, n_rays, n_samples, d_raw4, d_z_or_null, nullptr, nullptr, st)
  DISPATCH_C(n_samples, CALL);
#undef CALL
  NERF_CHECK_LAUNCH();
  return NERF_OK;
}

int nerf_composite_mse_fwd(const float* raw4, const float* z, const float* target, int64_t n_rays, int32_t n_samples,
                           int64_t n_total_rays, float loss_weight, float* rgb, float* weights, float* sq_err_sum,
                           float* d_rgb, void* stream) {
  NERF_CHECK_ARG(raw4 && z && target && sq_err_sum, "null pointer");
  NERF_CHECK_ARG(n_rays >= 0 && n_total_rays > 0 && n_samples > 0 && n_samples <= 1024, "bad shape");
  if (n_rays == 0) return NERF_OK;
  cudaStream_t st = (cudaStream_t)stream;
  LossArgs loss{target, (float)(2.0 * (double)loss_weight / (3.0 * (double)n_total_rays)), sq_err_sum, d_rgb};
#define CALL(C) launch_fwd<C>(raw4, z, n_rays, n_samples, rgb, weights, nullptr, nullptr, nullptr, nullptr, nullptr, &loss, st)
  DISPATCH_C(n_samples, CALL);
#undef CALL
  NERF_CHECK_LAUNCH();
  return NERF_OK;
}

int nerf_composite_mse_fwd_bwd(const float* raw4, const float* z, const float* target, int64_t n_rays, int32_t n_samples,
                               int64_t n_total_rays, float loss_weight, float* rgb_or_null, float* sq_err_sum,
                               float* d_raw4, float* d_z_or_null, void* stream) {
  NERF_CHECK_ARG(raw4 && z && target && sq_err_sum && d_raw4, "null pointer");
  NERF_CHECK_ARG(n_rays >= 0 && n_total_rays > 0 && n_samples > 0 && n_samples <= 1024, "bad shape");
  if (n_rays == 0) return NERF_OK;
  cudaStream_t st = (cudaStream_t)stream;
  LossArgs loss{target, (float)(2.0 * (double)loss_weight / (3.0 * (double)n_total_rays)), sq_err_sum, nullptr};
#define CALL(C) launch_bwd<C>(raw4, z, nullptr, nullptr, n_rays, n_samples, d_raw4, d_z_or_null, &loss, rgb_or_null, st)
  DISPATCH_C(n_samples, CALL);
#undef CALL
  NERF_CHECK_LAUNCH();
  return NERF_OK;
}

}  // extern "C"
