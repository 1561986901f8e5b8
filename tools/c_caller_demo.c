/* A caller of libnerf_b200.so that is not Python: plain C, cudaMalloc'ed buffers, the two whole-path entry points.
 * It fits the coarse + fine NeRF of the standard config to a synthetic 64x64 view for a few steps
 * (nerf_train_step_fused: NeRF.train_step, src/NeRF.py:136-178), renders the view back (nerf_render_fused_fwd:
 * NeRF.render, src/NeRF.py:109-134) and prints one JSON line.  This is the binding a TF custom-op shim would make.
 *
 *   nvcc --cudart shared -o tools/c_caller_demo tools/c_caller_demo.c -Lnerf-and-dietnerf_b200 -lnerf_b200 \
 *        -Xlinker -rpath -Xlinker '$ORIGIN/../nerf-and-dietnerf_b200'          (__graft_entry__.build() does this)
 *   tools/c_caller_demo [steps] [mode: 1 = bf16 (default), 0 = fp32] [anything: no side stream]
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "../include/nerf_b200.h"

#define CK(x)                                                             \
  do {                                                                    \
    cudaError_t e_ = (x);                                                 \
    if (e_ != cudaSuccess) {                                              \
      fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));            \
      return 2;                                                           \
    }                                                                     \
  } while (0)
#define NK(x)                                                             \
  do {                                                                    \
    int r_ = (x);                                                         \
    if (r_ != NERF_OK) {                                                  \
      fprintf(stderr, "%s -> %d: %s\n", #x, r_, nerf_last_error());       \
      return 3;                                                           \
    }                                                                     \
  } while (0)

static uint64_t g_state = 0x9E3779B97F4A7C15ull;
static float urand(void) { /* xorshift64: uniform in [0, 1) */
  g_state ^= g_state << 13;
  g_state ^= g_state >> 7;
  g_state ^= g_state << 17;
  return (float)((double)(g_state >> 40) * (1.0 / 16777216.0));
}

/* Keras Dense defaults (Glorot-uniform kernel, zero bias) in the layout nerf_param_count describes: [W (in,out), b] per
 * layer in creation order (src/NeRF.py:312-337) */
static void glorot_init(float* p, int64_t n_params) {
  static const int shape[11][2] = {{33, 256}, {256, 256}, {256, 256}, {256, 256}, {289, 256}, {256, 256},
                                   {256, 256}, {256, 256}, {280, 128}, {128, 3},  {280, 1}};
  int64_t off = 0;
  for (int l = 0; l < 11; ++l) {
    const int in = shape[l][0], out = shape[l][1];
    const float lim = sqrtf(6.0f / (float)(in + out));
    for (int64_t i = 0; i < (int64_t)in * out; ++i) p[off + i] = (2.0f * urand() - 1.0f) * lim;
    off += (int64_t)in * out;
    for (int i = 0; i < out; ++i) p[off + i] = 0.0f;
    off += out;
  }
  if (off != n_params) {
    fprintf(stderr, "layer table does not match nerf_param_count\n");
    exit(4);
  }
}

int main(int argc, char** argv) {
  const int steps = argc > 1 ? atoi(argv[1]) : 40;
  const int mode = argc > 2 ? atoi(argv[2]) : NERF_MODE_BF16;
  const int h = 64, w = 64;
  const int64_t n = (int64_t)h * w;
  nerf_net_cfg cfg = {5, 4, 2, 256, 128, 0.05f};
  nerf_render_cfg rc = {0.5f, 2.5f, 64, 128, mode};
  nerf_train_cfg tc = {1.0f, 0, 0, 5e-4f, 0.9f, 0.999f, 1e-7f};
  const int64_t np = nerf_param_count(&cfg);
  cudaStream_t st, side;
  CK(cudaStreamCreate(&st));
  CK(cudaStreamCreate(&side)); /* optional: the fine network's weight gradients overlap the coarse backward */

  /* parameters, Adam moments, gradients, bf16 packs */
  float* host_p = (float*)malloc((size_t)np * sizeof(float));
  float *params[2], *grads, *adam_m, *adam_v, *metrics;
  void* packed[2] = {NULL, NULL};
  const int64_t packed_bytes = nerf_packed_bytes(&cfg);
  for (int k = 0; k < 2; ++k) {
    glorot_init(host_p, np);
    CK(cudaMalloc((void**)&params[k], (size_t)np * sizeof(float)));
    CK(cudaMemcpy(params[k], host_p, (size_t)np * sizeof(float), cudaMemcpyHostToDevice));
    if (mode != NERF_MODE_FP32) {
      CK(cudaMalloc(&packed[k], (size_t)packed_bytes));
      NK(nerf_pack_weights(&cfg, params[k], packed[k], st));
    }
  }
  CK(cudaMalloc((void**)&grads, (size_t)(4 + 2 * np) * sizeof(float)));
  CK(cudaMalloc((void**)&adam_m, (size_t)(2 * np) * sizeof(float)));
  CK(cudaMalloc((void**)&adam_v, (size_t)(2 * np) * sizeof(float)));
  CK(cudaMemset(adam_m, 0, (size_t)(2 * np) * sizeof(float)));
  CK(cudaMemset(adam_v, 0, (size_t)(2 * np) * sizeof(float)));
  CK(cudaMalloc((void**)&metrics, 4 * sizeof(float)));

  /* one camera on the +z axis looking at the origin; the target image is a smooth pattern of the pixel position */
  const float c2w[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 1.2f, 0, 0, 0, 1};
  float *dirs, *origs, *target, *rgb;
  CK(cudaMalloc((void**)&dirs, (size_t)n * 4 * sizeof(float)));
  CK(cudaMalloc((void**)&origs, (size_t)n * 4 * sizeof(float)));
  CK(cudaMalloc((void**)&target, (size_t)n * 3 * sizeof(float)));
  CK(cudaMalloc((void**)&rgb, (size_t)n * 3 * sizeof(float)));
  NK(nerf_ray_directions(c2w, 0.69111f, h, w, 0, n, dirs, origs, st));
  float* host_t = (float*)malloc((size_t)n * 3 * sizeof(float));
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      float* t = host_t + 3 * ((int64_t)y * w + x);
      t[0] = 0.5f + 0.4f * sinf(0.2f * (float)x);
      t[1] = 0.5f + 0.4f * cosf(0.15f * (float)y);
      t[2] = 0.3f + 0.4f * (float)(x + y) / (float)(h + w);
    }
  CK(cudaMemcpy(target, host_t, (size_t)n * 3 * sizeof(float), cudaMemcpyHostToDevice));

  /* workspaces, sized by the library */
  const int64_t ws_train = nerf_train_workspace_bytes(&cfg, &rc, n), ws_render = nerf_render_workspace_bytes(&cfg, &rc, n);
  if (ws_train < 0 || ws_render < 0) {
    fprintf(stderr, "unsupported configuration\n");
    return 5;
  }
  void* ws;
  CK(cudaMalloc(&ws, (size_t)(ws_train > ws_render ? ws_train : ws_render)));

  float first[4] = {0, 0, 0, 0}, last[4] = {0, 0, 0, 0};
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int s = 0; s < steps; ++s) {
    nerf_rng_state rng = {7, 0, (uint32_t)s, 0};
    if (s == 1) CK(cudaEventRecord(e0, st));
    NK(nerf_train_step_fused(&cfg, &rc, &tc, params[0], packed[0], params[1], packed[1], origs, dirs, target, n, n, &rng,
                             grads, adam_m, adam_v, s + 1, metrics, ws, argc > 3 ? NULL : side, st));
    if (s == 0 || s == steps - 1) {
      CK(cudaMemcpyAsync(s == 0 ? first : last, metrics, 4 * sizeof(float), cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
    }
  }
  CK(cudaEventRecord(e1, st));
  CK(cudaStreamSynchronize(st));
  float ms = 0.f;
  if (steps > 2) CK(cudaEventElapsedTime(&ms, e0, e1));

  /* render the view with the trained weights and measure it against the target on the host */
  nerf_render_outs outs = {rgb, NULL, NULL, NULL, NULL, NULL, NULL, NULL};
  nerf_rng_state rng = {11, 0, 0, 0};
  NK(nerf_render_fused_fwd(&cfg, &rc, params[0], packed[0], params[1], packed[1], origs, dirs, n, &rng, &outs, ws, st));
  float* host_rgb = (float*)malloc((size_t)n * 3 * sizeof(float));
  CK(cudaMemcpyAsync(host_rgb, rgb, (size_t)n * 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  double se = 0.0;
  int finite = 1;
  for (int64_t i = 0; i < n * 3; ++i) {
    if (!isfinite(host_rgb[i])) finite = 0;
    se += ((double)host_rgb[i] - host_t[i]) * ((double)host_rgb[i] - host_t[i]);
  }
  const double psnr = -10.0 * log10(se / (double)(n * 3));
  printf("{\"caller\": \"c\", \"library\": \"%s\", \"mode\": %d, \"rays\": %lld, \"steps\": %d, \"loss_first\": %.6f, "
         "\"loss_last\": %.6f, \"psnr_fine_last\": %.3f, \"render_psnr\": %.3f, \"finite\": %d, \"ms_per_step\": %.3f, "
         "\"train_workspace_mb\": %.1f}\n",
         nerf_version(), mode, (long long)n, steps, first[0], last[0], last[2], psnr, finite,
         steps > 2 ? ms / (float)(steps - 1) : 0.f, (double)ws_train / 1048576.0);
  return finite && last[0] < first[0] ? 0 : 1;
}
