// Loss and optimizer kernels of the train step: keras.losses.MeanSquaredError (src/NeRF.py:151,157) and
// Keras-2.7 Adam (optimizer.apply_gradients, src/NeRF.py:164,167).  Elementwise, HBM-bound.
#include "common.cuh"

namespace nerf {

__global__ void mse_fwd_bwd_kernel(const float* __restrict__ rgb, const float* __restrict__ target, int64_t n_elems,
                                   float grad_scale, float* __restrict__ sq_err_sum, float* __restrict__ d_rgb) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  float e = 0.f;
  if (i < n_elems) {
    float diff = rgb[i] - target[i];
    e = diff * diff;
    if (d_rgb) d_rgb[i] = grad_scale * diff;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) e += __shfl_xor_sync(0xffffffffu, e, d);
  __shared__ float part[8];
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) part[warp] = e;
  __syncthreads();
  if (threadIdx.x == 0 && sq_err_sum) {
    float s = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += part[w];
    atomicAdd(sq_err_sum, s);
  }
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t n, float lr_t, float one_minus_b1, float one_minus_b2,
                            float eps) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float gi = g[i];
  float mi = m[i] + (gi - m[i]) * one_minus_b1;
  float vi = v[i] + (gi * gi - v[i]) * one_minus_b2;
  m[i] = mi;
  v[i] = vi;
  p[i] = p[i] - lr_t * mi / (sqrtf(vi) + eps);
}

// metrics of one train step from the two squared-error sums (src/NeRF.py:170-178, src/UtilsNeuralRadianceField.py:123-132):
// out = [loss, psnr_coarse, psnr_fine, loss_for_rays]; one thread -- it replaces ~10 elementwise launches per step.
__global__ void train_metrics_kernel(const float* __restrict__ sums, float inv_count, float coarse_weight, int has_fine,
                                     float* __restrict__ out) {
  const float mse_c = sums[0] * inv_count;
  const float mse_f = has_fine ? sums[1] * inv_count : 0.f;
  const float k = -10.0f / 2.302585092994046f;           // psnr = -10 ln(mse) / ln(10)
  out[0] = coarse_weight * mse_c + mse_f;
  out[1] = k * logf(mse_c);
  out[2] = has_fine ? k * logf(mse_f) : 0.f;
  out[3] = mse_c + mse_f;
}

}  // namespace nerf

using namespace nerf;

extern "C" {

int nerf_mse_fwd_bwd(const float* rgb, const float* target, int64_t n_rays, int64_t n_total_rays, float loss_weight,
                     float* sq_err_sum, float* d_rgb, void* stream) {
  NERF_CHECK_ARG(rgb && target, "null pointer");
  NERF_CHECK_ARG(n_rays >= 0 && n_total_rays > 0, "bad ray count");
  if (n_rays == 0) return NERF_OK;
  int64_t n = n_rays * 3;
  float scale = (float)(2.0 * (double)loss_weight / (3.0 * (double)n_total_rays));
  mse_fwd_bwd_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(rgb, target, n, scale, sq_err_sum, d_rgb);
  NERF_CHECK_LAUNCH();
  return NERF_OK;
}

int nerf_train_metrics(const float* sq_err_sums, int64_t n_total_rays, float coarse_loss_weight, int32_t has_fine,
                       float* out4, void* stream) {
  NERF_CHECK_ARG(sq_err_sums && out4, "null pointer");
  NERF_CHECK_ARG(n_total_rays > 0, "bad ray count");
  train_metrics_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(sq_err_sums, (float)(1.0 / (3.0 * (double)n_total_rays)),
                                                          coarse_loss_weight, has_fine, out4);
  NERF_CHECK_LAUNCH();
  return NERF_OK;
}

int nerf_adam_step(float* params, const float* grads, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                   float eps, int64_t t, void* stream) {
  NERF_CHECK_ARG(params && grads && m && v, "null pointer");
  NERF_CHECK_ARG(n >= 0 && t >= 1, "bad size or step");
  if (n == 0) return NERF_OK;
  double lr_t = (double)lr * sqrt(1.0 - pow((double)beta2, (double)t)) / (1.0 - pow((double)beta1, (double)t));
  adam_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(params, grads, m, v, n, (float)lr_t,
                                                                          1.0f - beta1, 1.0f - beta2, eps);
  NERF_CHECK_LAUNCH();
  return NERF_OK;
}

}  // extern "C"
