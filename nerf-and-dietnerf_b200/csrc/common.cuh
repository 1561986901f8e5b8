// Shared helpers for libnerf_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/nerf_b200.h"

namespace nerf {

void set_error(const char* fmt, ...);

#define NERF_CHECK_ARG(cond, msg)                         \
  do {                                                    \
    if (!(cond)) {                                        \
      nerf::set_error("%s: %s", __func__, msg);           \
      return NERF_E_ARG;                                  \
    }                                                     \
  } while (0)

#define NERF_CHECK_LAUNCH()                                                              \
  do {                                                                                   \
    cudaError_t e__ = cudaGetLastError();                                                \
    if (e__ != cudaSuccess) {                                                            \
      nerf::set_error("%s: CUDA error: %s", __func__, cudaGetErrorString(e__));          \
      return NERF_E_CUDA;                                                                \
    }                                                                                    \
  } while (0)

#define NERF_CUDA(call)                                                                  \
  do {                                                                                   \
    cudaError_t e__ = (call);                                                            \
    if (e__ != cudaSuccess) {                                                            \
      nerf::set_error("%s: %s failed: %s", __func__, #call, cudaGetErrorString(e__));    \
      return NERF_E_CUDA;                                                                \
    }                                                                                    \
  } while (0)

constexpr int kMaxSMs = 160;  // upper bound used for scratch sizing (B200: 148)

// SM count of the CURRENT device (cudaDeviceGetAttribute, cached per device; 148 when no device is reachable, so that the
// host-only size queries keep working on a CPU box).
int num_sms();
// true exactly once per (current device, slot): guards per-device one-time setup such as cudaFuncSetAttribute
// (thread-safe; the ABI is re-entrant per (device, stream)).  Slots: 0 = forward kernels, 1 = backward kernels.
bool device_first_use(int slot);

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// The rays behind the rows of a backward call (nerf_mlp_bwd_rays): with them the dX chain turns d(xyz encoding) into
// d z = d . PE'(o + d z) in the epilogue of its last step instead of writing the 33 floats per row out for
// encode_samples_bwd_z to read back.
struct BwdRays {
  const float4* origs;
  const float4* dirs;
  const float* z;          // (n_rays, n_samples) depths of the rows
  float* d_z;              // (n_rays, n_samples); null = no ray inputs
  int32_t n_samples;
  int32_t accumulate;      // 1: d_z += (the compositing backward wrote its share first)
};

// Pinhole ray of pixel index r = y w + x (get_rays_directions, src/UtilsCV.py:467-499): ONE definition for the stand-alone
// generator (rays.cu) and for the MLP prologue that generates its rays itself (nerf_mlp_fwd_camera), so the two agree bit
// for bit.  c2w: row-major 4 x 4.
__device__ __forceinline__ float4 pinhole_ray_dir(const float* c2w, float tan_half_fov, int h, int w, int64_t r) {
  const int y = (int)(r / w), x = (int)(r - (int64_t)y * w);
  const float xr = __fadd_rn((float)x, 0.5f), yr = __fadd_rn((float)y, 0.5f);
  const float x_ndc = __fdiv_rn(xr, (float)w), y_ndc = __fdiv_rn(yr, (float)h);
  const float xs = __fsub_rn(__fmul_rn(2.f, x_ndc), 1.f);
  const float ys = __fsub_rn(1.f, __fmul_rn(2.f, y_ndc));
  const float d0 = __fmul_rn(xs, tan_half_fov), d1 = __fmul_rn(ys, tan_half_fov), d2 = -1.f, d3 = 0.f;
  float o[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float acc = __fmul_rn(c2w[k * 4 + 0], d0);
    acc = __fadd_rn(acc, __fmul_rn(c2w[k * 4 + 1], d1));
    acc = __fadd_rn(acc, __fmul_rn(c2w[k * 4 + 2], d2));
    acc = __fadd_rn(acc, __fmul_rn(c2w[k * 4 + 3], d3));
    o[k] = acc;
  }
  return make_float4(o[0], o[1], o[2], o[3]);
}
// the camera behind the rays of a forward call (nerf_mlp_fwd_camera): ray i of the call is pixel ray_begin + i
struct CameraRays {
  float c2w[16];
  float tan_half_fov;
  int32_t h, w;
  int64_t ray_begin;
};

// coarse depths drawn inside the MLP prologue (nerf_mlp_fwd_rays_stratified)
struct StratifiedZ {
  float z_start, z_end;
  uint64_t seed, ray_offset;
  uint32_t step;
  float* z_out;
};

// ---- network geometry ---------------------------------------------------------------------------------
struct LayerDesc {
  int in, out;
  int64_t w_off, b_off;  // offsets (floats) into the flat parameter vector
};

struct NetGeom {
  int dx, dv, hidden, last_hidden, n_layers;
  bool view;
  LayerDesc layers[12];
  int64_t n_params;
};

inline bool make_geom(const nerf_net_cfg* cfg, NetGeom* g) {
  if (!cfg || cfg->n_pos_enc_xyz < 0 || cfg->n_pos_enc_xyz > 16 || cfg->n_angles < 0 || cfg->n_angles > 2 ||
      cfg->hidden <= 0 || cfg->last_hidden <= 0 || cfg->n_pos_enc_view < 0 || cfg->n_pos_enc_view > 16)
    return false;
  g->dx = 3 + 6 * cfg->n_pos_enc_xyz;
  g->view = cfg->n_angles > 0;
  g->dv = g->view ? 2 * cfg->n_pos_enc_view * (cfg->n_angles + 1) : 0;
  g->hidden = cfg->hidden;
  g->last_hidden = cfg->last_hidden;
  int H = cfg->hidden, HL = cfg->last_hidden, n = 0;
  int ins[12], outs[12];
  ins[n] = g->dx; outs[n++] = H;
  for (int i = 0; i < 3; ++i) { ins[n] = H; outs[n++] = H; }
  ins[n] = g->dx + H; outs[n++] = H;
  for (int i = 0; i < 3; ++i) { ins[n] = H; outs[n++] = H; }
  if (g->view) {
    ins[n] = H + g->dv; outs[n++] = HL;  // last hidden
    ins[n] = HL; outs[n++] = 3;          // rgb
    ins[n] = H + g->dv; outs[n++] = 1;   // sigma
  } else {
    ins[n] = H; outs[n++] = H;
    ins[n] = H; outs[n++] = HL;
    ins[n] = HL; outs[n++] = 3;
    ins[n] = H; outs[n++] = 1;
  }
  g->n_layers = n;
  int64_t off = 0;
  for (int i = 0; i < n; ++i) {
    g->layers[i].in = ins[i];
    g->layers[i].out = outs[i];
    g->layers[i].w_off = off;
    off += (int64_t)ins[i] * outs[i];
    g->layers[i].b_off = off;
    off += outs[i];
  }
  g->n_params = off;
  return true;
}

// ---- Philox4x32-10 (same stream as oracle/philox.py) -----------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

__device__ __forceinline__ float bits_to_uniform(uint32_t w) {
  return __uint_as_float((w & 0x7FFFFFu) | 0x3F800000u) - 1.0f;
}

// tf.linspace(start, stop, n)[i] in fp32 (src/UtilsCV.py:573): exact end points, start + delta * i in between
__device__ __forceinline__ float linspace_tf(float start, float stop, float delta, int i, int n) {
  if (i == 0) return start;
  if (i == n - 1) return stop;
  return __fadd_rn(start, __fmul_rn(delta, (float)i));
}

// get_z_values (src/UtilsCV.py:565-581) for sample s of a ray: linspace + jitter u * span / n.  ONE definition shared by
// the stand-alone kernel (nerf_stratified_z) and the MLP prologue that generates the coarse depths itself.
__device__ __forceinline__ float stratified_z_value(float z_start, float z_end, float span, int n_samples, int s, float u) {
  const float delta = n_samples > 1 ? __fdiv_rn(__fsub_rn(z_end, z_start), (float)(n_samples - 1)) : 0.f;
  const float lin = linspace_tf(z_start, z_end, delta, s, n_samples);
  return __fadd_rn(lin, __fdiv_rn(__fmul_rn(u, span), (float)n_samples));
}

// four uniforms: draws 4*block .. 4*block+3 of (ray, stream, step)
__device__ __forceinline__ float4 philox_uniform4(uint64_t seed, uint32_t ray, uint32_t block, uint32_t stream_id,
                                                  uint32_t step) {
  uint4 r = philox4x32_10(make_uint4(ray, block, stream_id, step),
                          make_uint2((uint32_t)(seed & 0xFFFFFFFFu), (uint32_t)(seed >> 32)));
  return make_float4(bits_to_uniform(r.x), bits_to_uniform(r.y), bits_to_uniform(r.z), bits_to_uniform(r.w));
}
#endif

}  // namespace nerf
