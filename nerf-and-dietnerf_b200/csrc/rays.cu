// K1: ray generation, stratified depths, sample positions, view inputs and sin/cos positional encodings.
// Reference functions replaced: src/UtilsCV.py:467-499 (get_rays_directions), :565-581 (get_z_values),
// :584-599 (sample_along_rays), :124-143 (get_view_directions); src/UtilsNeuralRadianceField.py:52-85 (PE).
// All of it is HBM/SFU-bound elementwise work: one thread per output element, coalesced stores,
// grids sized from the element count.  Arithmetic that the oracle pins bit-for-bit uses __fmul_rn/__fadd_rn
// so nvcc cannot contract it into FMAs.
#include <math.h>

#include "common.cuh"

namespace nerf {

constexpr float kPi = 3.14159265358979323846f;

// ---- ray directions --------------------------------------------------------------------------------------
struct C2W {
  float m[16];
};

__global__ void ray_directions_kernel(C2W c2w, float tan_half_fov, int h, int w, int64_t ray_begin, int64_t n_rays,
                                      float* __restrict__ dirs4, float* __restrict__ origs4) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n_rays) return;
  const float4 out = pinhole_ray_dir(c2w.m, tan_half_fov, h, w, ray_begin + i);
  reinterpret_cast<float4*>(dirs4)[i] = out;
  if (origs4) reinterpret_cast<float4*>(origs4)[i] = make_float4(c2w.m[3], c2w.m[7], c2w.m[11], c2w.m[15]);
}

// ---- stratified z ------------------------------------------------------------------------------------------

__global__ void stratified_z_kernel(float z_start, float z_end, float span, int64_t n_rays, int n_samples,
                                    const float* __restrict__ jitter, uint64_t seed, uint32_t step,
                                    uint64_t ray_offset, float* __restrict__ z) {
  int blocks_per_ray = (n_samples + 3) / 4;
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= n_rays * blocks_per_ray) return;
  int64_t ray = t / blocks_per_ray;
  int blk = (int)(t % blocks_per_ray);
  float u[4];
  if (jitter) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int s = blk * 4 + k;
      u[k] = s < n_samples ? jitter[ray * n_samples + s] : 0.f;
    }
  } else {
    float4 r = philox_uniform4(seed, (uint32_t)(ray + ray_offset), (uint32_t)blk, 0u, step);
    u[0] = r.x; u[1] = r.y; u[2] = r.z; u[3] = r.w;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    int s = blk * 4 + k;
    if (s < n_samples) z[ray * n_samples + s] = stratified_z_value(z_start, z_end, span, n_samples, s, u[k]);
  }
}

// ---- sample_along_rays / view dirs ----------------------------------------------------------------------------
__global__ void sample_along_rays_kernel(const float4* __restrict__ origs, const float4* __restrict__ dirs,
                                         const float* __restrict__ z, int64_t total, int n_samples,
                                         float4* __restrict__ out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= total) return;
  int64_t ray = i / n_samples;
  float4 o = __ldg(origs + ray), d = __ldg(dirs + ray);
  float zz = z[i];
  out[i] = make_float4(__fadd_rn(o.x, __fmul_rn(d.x, zz)), __fadd_rn(o.y, __fmul_rn(d.y, zz)),
                       __fadd_rn(o.z, __fmul_rn(d.z, zz)), __fadd_rn(o.w, __fmul_rn(d.w, zz)));
}

__global__ void view_directions_kernel(const float4* __restrict__ dirs, int64_t total, int n_samples, int n_comp,
                                       float* __restrict__ out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;  // over (N*S*n_comp)
  if (i >= total) return;
  int c = (int)(i % n_comp);
  int64_t ray = (i / n_comp) / n_samples;
  float4 d = __ldg(dirs + ray);
  float v = (n_comp == 3) ? (c == 0 ? d.x : (c == 1 ? d.y : d.z)) : (c == 0 ? d.x : d.z);
  out[i] = v;
}

// ---- positional encodings ----------------------------------------------------------------------------------
// theta = (2^k * pi_f32) * v ; 2^k * pi is exact scaling so the association matches the reference expression.
__device__ __forceinline__ float pe_theta(float v, int k) { return __fmul_rn(ldexpf(kPi, k), v); }

// element j of the xyz encoding row of width 3*(1+2L): per coordinate [c, s0, c0, ..., s_{L-1}, c_{L-1}]
__device__ __forceinline__ float pe_xyz_elem(float v, int within, int L) {
  if (L == 0 || within == 0) return v;
  int k = (within - 1) >> 1;
  float th = pe_theta(v, k);
  return ((within - 1) & 1) ? cosf(th) : sinf(th);
}

__global__ void posenc_xyz_kernel(const float* __restrict__ xyz, int64_t m, int L, float* __restrict__ out) {
  int per = 1 + 2 * L, width = 3 * per;
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= m * width) return;
  int64_t row = i / width;
  int j = (int)(i % width), c = j / per, within = j % per;
  out[i] = pe_xyz_elem(__ldg(xyz + row * 3 + c), within, L);
}

__global__ void posenc_views_kernel(const float* __restrict__ x, int64_t m, int ncomp, int L, float* __restrict__ out) {
  int per = 2 * L, width = ncomp * per;
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= m * width) return;
  int64_t row = i / width;
  int j = (int)(i % width), c = j / per, within = j % per;
  float th = pe_theta(__ldg(x + row * ncomp + c), within >> 1);
  out[i] = (within & 1) ? cosf(th) : sinf(th);
}

__device__ __forceinline__ float pe_xyz_grad_coord(float v, const float* __restrict__ d_row, int L) {
  // d_row points at this coordinate's (1+2L) slice of the upstream gradient
  float g = d_row[0];
  if (L == 0) return g;
  for (int k = 0; k < L; ++k) {
    float scale = ldexpf(kPi, k);
    float th = __fmul_rn(scale, v);
    float s, c;
    sincosf(th, &s, &c);
    g += (d_row[1 + 2 * k] * c - d_row[2 + 2 * k] * s) * scale;
  }
  return g;
}

__global__ void posenc_xyz_bwd_kernel(const float* __restrict__ xyz, const float* __restrict__ d_out, int64_t m, int L,
                                      float* __restrict__ d_xyz) {
  int per = 1 + 2 * L;
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;  // over m*3
  if (i >= m * 3) return;
  int64_t row = i / 3;
  int c = (int)(i % 3);
  d_xyz[i] = pe_xyz_grad_coord(xyz[i], d_out + row * 3 * per + c * per, L);
}

// ---- fused encode: (o, d, z) -> xyz_enc, view_enc ----------------------------------------------------------------
__global__ void encode_samples_kernel(const float4* __restrict__ origs, const float4* __restrict__ dirs,
                                      const float* __restrict__ z, int64_t n_samples_total, int n_samples, int Lx,
                                      int Lv, int n_angles, float* __restrict__ xyz_enc, float* __restrict__ view_enc) {
  int perx = 1 + 2 * Lx, dx = 3 * perx;
  int ncomp = n_angles > 0 ? n_angles + 1 : 0;
  int perv = 2 * Lv, dv = ncomp * perv;
  int width = dx + dv;
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n_samples_total * width) return;
  int64_t row = i / width;
  int j = (int)(i % width);
  int64_t ray = row / n_samples;
  float4 d = __ldg(dirs + ray);
  if (j < dx) {
    float4 o = __ldg(origs + ray);
    float zz = __ldg(z + row);
    int c = j / perx, within = j % perx;
    float oc = c == 0 ? o.x : (c == 1 ? o.y : o.z);
    float dc = c == 0 ? d.x : (c == 1 ? d.y : d.z);
    float p = __fadd_rn(oc, __fmul_rn(dc, zz));
    xyz_enc[row * dx + j] = pe_xyz_elem(p, within, Lx);
  } else {
    int jj = j - dx, c = jj / perv, within = jj % perv;
    float v = (ncomp == 3) ? (c == 0 ? d.x : (c == 1 ? d.y : d.z)) : (c == 0 ? d.x : d.z);
    float th = pe_theta(v, within >> 1);
    view_enc[row * dv + jj] = (within & 1) ? cosf(th) : sinf(th);
  }
}

// One thread per sample row.  The rows of a block (256 x 3 (1 + 2 Lx) floats, 33 KB at Lx = 5) are staged in shared
// memory with coalesced loads first: read straight from global memory every thread walked its own 132-byte row, 32
// scattered sectors per warp instruction (36 us per 524 k rows where the bytes take 11).  The odd row stride keeps the
// shared-memory reads conflict-free.
__global__ void __launch_bounds__(256)
encode_samples_bwd_z_kernel(const float4* __restrict__ origs, const float4* __restrict__ dirs,
                            const float* __restrict__ z, const float* __restrict__ d_xyz_enc,
                            int64_t total, int n_samples, int Lx, float* __restrict__ d_z,
                            int accumulate) {
  extern __shared__ float s_rows[];
  const int perx = 1 + 2 * Lx, width = 3 * perx;
  const int64_t row0 = blockIdx.x * (int64_t)blockDim.x;
  const int64_t n_here = min((int64_t)blockDim.x, total - row0);
  const float* src = d_xyz_enc + row0 * width;
  for (int64_t i = threadIdx.x; i < n_here * width; i += blockDim.x) s_rows[i] = __ldcs(src + i);
  __syncthreads();
  const int64_t row = row0 + threadIdx.x;
  if (row >= total) return;
  int64_t ray = row / n_samples;
  float4 o = __ldg(origs + ray), d = __ldg(dirs + ray);
  float zz = z[row];
  const float* g = s_rows + (size_t)threadIdx.x * width;
  float px = __fadd_rn(o.x, __fmul_rn(d.x, zz)), py = __fadd_rn(o.y, __fmul_rn(d.y, zz)),
        pz = __fadd_rn(o.z, __fmul_rn(d.z, zz));
  float gz = pe_xyz_grad_coord(px, g, Lx) * d.x + pe_xyz_grad_coord(py, g + perx, Lx) * d.y +
             pe_xyz_grad_coord(pz, g + 2 * perx, Lx) * d.z;
  d_z[row] = accumulate ? d_z[row] + gz : gz;
}

static inline unsigned grid_for(int64_t n, int block) { return (unsigned)ceil_div(n, block); }

}  // namespace nerf

using namespace nerf;

extern "C" {

int nerf_ray_directions(const float* c2w_host, float fov, int32_t h, int32_t w, int64_t ray_begin, int64_t n_rays,
                        float* dirs4, float* origs4, void* stream) {
  NERF_CHECK_ARG(c2w_host && dirs4, "null pointer");
  NERF_CHECK_ARG(h > 0 && w > 0 && ray_begin >= 0 && n_rays >= 0 && ray_begin + n_rays <= (int64_t)h * w,
                 "ray range outside the image");
  if (n_rays == 0) return NERF_OK;
  C2W m;
  for (int i = 0; i < 16; ++i) m.m[i] = c2w_host[i];
  float tan_half = tanf(fov / 2.0f);
  ray_directions_kernel<<<grid_for(n_rays, 256), 256, 0, (cudaStream_t)stream>>>(m, tan_half, h, w, ray_begin, n_rays,
                                                                                 dirs4, origs4);
  NERF_CHECK_LAUNCH();
  return NERF_OK;
}

int nerf_stratified_z(float z_start, float z_end, int64_t n_rays, int32_t n_samples, const float* jitter_or_null,
                      uint64_t seed, uint32_t step, uint64_t ray_offset, float* z, void* stream) {
  NERF_CHECK_ARG(z, "null pointer");
  NERF_CHECK_ARG(n_rays >= 0 && n_samples > 0, "bad shape");
  if (n_rays == 0) return NERF_OK;
  // (z_end - z_start) is formed in double from the Python floats and then cast (src/UtilsCV.py:580)
  float span = (float)((double)z_end - (double)z_start);
  int64_t threads = n_rays * ((n_samples + 3) / 4);
  stratified_z_kernel<<<grid_for(threads, 256), 256, 0, (cudaStream_t)stream>>>(z_start, z_end, span, n_rays, n_samples,
                                                                                jitter_or_null, seed, step, ray_offset, z);
  NERF_CHECK_LAUNCH();
  return NERF_OK;
}

int nerf_sample_along_rays(const float* origs4, const float* dirs4, const float* z, int64_t n_rays, int32_t n_samples,
                           float* coords4, void* stream) {
  NERF_CHECK_ARG(origs4 && dirs4 && z && coords4, "null pointer");
  NERF_CHECK_ARG(n_rays >= 0 && n_samples > 0, "bad shape");
  int64_t total = n_rays * n_samples;
  if (total == 0) return NERF_OK;
  sample_along_rays_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      (const float4*)origs4, (const float4*)dirs4, z, total, n_samples, (float4*)coords4);
  NERF_CHECK_LAUNCH();
  return NERF_OK;
}

int nerf_view_directions(const float* dirs4, int64_t n_rays, int32_t n_samples, int32_t n_angles, float* view_dirs,
                         void* stream) {
  NERF_CHECK_ARG(dirs4 && view_dirs, "null pointer");
  NERF_CHECK_ARG(n_angles == 1 || n_angles == 2, "n_angles_for_model should be 1 or 2.");
  int64_t total = n_rays * n_samples * (n_angles + 1);
  if (total == 0) return NERF_OK;
  view_directions_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const float4*)dirs4, total, n_samples,
                                                                                 n_angles + 1, view_dirs);
  NERF_CHECK_LAUNCH();
  return NERF_OK;
}

int nerf_posenc_xyz(const float* xyz, int64_t m, int32_t L, float* out, void* stream) {
  NERF_CHECK_ARG(xyz && out, "null pointer");
  NERF_CHECK_ARG(m >= 0 && L >= 0 && L <= 16, "bad shape");
  int64_t total = m * 3 * (1 + 2 * L);
  if (total == 0) return NERF_OK;
  posenc_xyz_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(xyz, m, L, out);
  NERF_CHECK_LAUNCH();
  return NERF_OK;
}

int nerf_posenc_views(const float* x, int64_t m, int32_t c, int32_t L, float* out, void* stream) {
  NERF_CHECK_ARG(x && out, "null pointer");
  NERF_CHECK_ARG(m >= 0 && c > 0 && L > 0 && L <= 16, "bad shape");
  int64_t total = m * c * 2 * L;
  if (total == 0) return NERF_OK;
  posenc_views_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(x, m, c, L, out);
  NERF_CHECK_LAUNCH();
  return NERF_OK;
}

int nerf_posenc_xyz_bwd(const float* xyz, const float* d_out, int64_t m, int32_t L, float* d_xyz, void* stream) {
  NERF_CHECK_ARG(xyz && d_out && d_xyz, "null pointer");
  NERF_CHECK_ARG(m >= 0 && L >= 0 && L <= 16, "bad shape");
  if (m == 0) return NERF_OK;
  posenc_xyz_bwd_kernel<<<grid_for(m * 3, 256), 256, 0, (cudaStream_t)stream>>>(xyz, d_out, m, L, d_xyz);
  NERF_CHECK_LAUNCH();
  return NERF_OK;
}

int nerf_encode_samples(const nerf_net_cfg* cfg, const float* origs4, const float* dirs4, const float* z, int64_t n_rays,
                        int32_t n_samples, float* xyz_enc, float* view_enc, void* stream) {
  NetGeom g;
  NERF_CHECK_ARG(make_geom(cfg, &g), "bad net config");
  NERF_CHECK_ARG(origs4 && dirs4 && z && xyz_enc && (view_enc || !g.view), "null pointer");
  int64_t rows = n_rays * n_samples;
  if (rows == 0) return NERF_OK;
  int64_t total = rows * (g.dx + g.dv);
  encode_samples_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      (const float4*)origs4, (const float4*)dirs4, z, rows, n_samples, cfg->n_pos_enc_xyz, cfg->n_pos_enc_view,
      cfg->n_angles, xyz_enc, view_enc);
  NERF_CHECK_LAUNCH();
  return NERF_OK;
}

int nerf_encode_samples_bwd_z(const nerf_net_cfg* cfg, const float* origs4, const float* dirs4, const float* z,
                              const float* d_xyz_enc, int64_t n_rays, int32_t n_samples, float* d_z, int32_t accumulate,
                              void* stream) {
  NetGeom g;
  NERF_CHECK_ARG(make_geom(cfg, &g), "bad net config");
  NERF_CHECK_ARG(origs4 && dirs4 && z && d_xyz_enc && d_z, "null pointer");
  int64_t rows = n_rays * n_samples;
  if (rows == 0) return NERF_OK;
  int block = 256;                                   // rows per block: as many as fit 96 KB of staged gradient rows
  const size_t row_bytes = (size_t)3 * (1 + 2 * cfg->n_pos_enc_xyz) * sizeof(float);
  while (block > 32 && block * row_bytes > 96 * 1024) block >>= 1;
  const size_t smem = block * row_bytes;
  NERF_CHECK_ARG(smem <= 200 * 1024, "n_pos_enc_xyz too large");
  if (smem > 48 * 1024)
    NERF_CUDA(cudaFuncSetAttribute(encode_samples_bwd_z_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  encode_samples_bwd_z_kernel<<<grid_for(rows, block), block, smem, (cudaStream_t)stream>>>(
      (const float4*)origs4, (const float4*)dirs4, z, d_xyz_enc, rows, n_samples, cfg->n_pos_enc_xyz, d_z, accumulate);
  NERF_CHECK_LAUNCH();
  return NERF_OK;
}

}  // extern "C"
