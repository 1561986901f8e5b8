// Hierarchical (inverse-CDF) sampling: get_z_vals_from_prob_dist_func, src/UtilsCV.py:502-539, forward and the
// gradient w.r.t. `weights` that TF autodiff produces inside NeRF.train_step (the reference does not detach
// z_from_dist, src/NeRF.py:155), plus the per-ray sorted merge of src/NeRF.py:132.
//
// Bit-exact indices: the searchsorted result depends on the fp32 summation order of reduce_sum and cumsum.
// The canonical order (shared with oracle/nerf_oracle.py) is sequential left-to-right fp32 with IEEE division,
// so one lane walks the S<=1024 entries of a ray that the whole warp staged in shared memory with coalesced
// loads; everything else (pdf, search, interpolation, rank sort) is lane-parallel.  1 KB/ray of HBM traffic.
#include "common.cuh"

namespace nerf {

constexpr unsigned kFullMask = 0xffffffffu;
constexpr int kWarpsPerBlock = 4;

struct RayCdf {
  float* w;    // [S] weights, then pdf
  float* cdf;  // [S]
  float* z;    // [S]
};

// Stage one ray and build pdf/cdf in the canonical order. Returns (sum + eps).
__device__ __forceinline__ float build_cdf(const float* __restrict__ weights, const float* __restrict__ z, int S,
                                           int lane, float* sw, float* scdf, float* sz) {
  for (int i = lane; i < S; i += 32) {
    sw[i] = __ldcs(weights + i);
    sz[i] = __ldcs(z + i);
  }
  __syncwarp();
  // Sequential left-to-right fp32 sum and cumsum, computed REDUNDANTLY by every lane from broadcast shared-memory reads:
  // the loads do not depend on the running value, so they pipeline and the dependent chain is one FADD per element
  // (a single lane doing load -> add -> store through the same array paid a shared-memory round trip per element).
  float total = 0.f;
#pragma unroll 8
  for (int i = 0; i < S; ++i) total = __fadd_rn(total, sw[i]);
  const float denom = __fadd_rn(total, 1e-7f);  // EPS, src/UtilsCV.py:30
  __syncwarp();
  for (int i = lane; i < S; i += 32) sw[i] = __fdiv_rn(sw[i], denom);   // pdf (the weights are not needed after this)
  __syncwarp();
  float run = 0.f;
#pragma unroll 8
  for (int i = 0; i < S; ++i) {
    run = __fadd_rn(run, sw[i]);
    if ((i & 31) == lane) scdf[i] = run;
  }
  __syncwarp();
  return denom;
}

struct Draw {
  int b, t;
  float lo, hi, zlo, zhi, den;
  bool floored;
};

__device__ __forceinline__ Draw locate(const float* scdf, const float* sz, int S, float u, int* idx_out) {
  // idx = #{i : cdf_i < u}  (tf.searchsorted side='left')
  int lo_i = 0, hi_i = S;
  while (lo_i < hi_i) {
    int mid = (lo_i + hi_i) >> 1;
    if (scdf[mid] < u) lo_i = mid + 1; else hi_i = mid;
  }
  int idx = lo_i;
  if (idx_out) *idx_out = idx;
  Draw d;
  d.b = max(0, idx - 1);
  d.t = min(S - 1, idx);
  d.lo = scdf[d.b];
  d.hi = scdf[d.t];
  int zb = min(max(d.b, 0), S - 2), zt = min(max(d.t, 0), S - 2);
  d.zlo = __fmul_rn(0.5f, __fadd_rn(sz[zb + 1], sz[zb]));
  d.zhi = __fmul_rn(0.5f, __fadd_rn(sz[zt + 1], sz[zt]));
  float den = __fsub_rn(d.hi, d.lo);
  d.floored = den < 1e-5f;
  d.den = d.floored ? 1e-5f : den;
  return d;
}

// Stable ascending sort of the Nf <= 32 E new samples of a ray by a warp-wide bitonic network on 64-bit keys
// (order-preserving bits of z in the high word, the draw index in the low word: ties keep draw order, exactly the stable
// sort the oracle performs, and no two keys are equal).  Element e = lane E + i lives in register i of lane `lane`, so the
// strides below E are register swaps and the others one 64-bit shuffle per element: log^2 steps (28 for 128 samples)
// instead of the Nf^2 / 32 pair tests per lane of the rank sort.
template <int E>
__device__ __forceinline__ void bitonic_sort_rows(const float* szs, int Nf, int lane, float* __restrict__ z_out,
                                                  int* __restrict__ perm_out) {
  unsigned long long k[E];
#pragma unroll
  for (int i = 0; i < E; ++i) {
    const int j = lane * E + i;
    if (j < Nf) {
      uint32_t u = __float_as_uint(szs[j]);
      if ((u << 1) == 0u) u = 0u;                                  // -0 sorts with +0 (a float compare calls them equal)
      u ^= (u >> 31) ? 0xffffffffu : 0x80000000u;
      k[i] = ((unsigned long long)u << 32) | (uint32_t)j;
    } else {
      k[i] = ~0ull;                                                // padding sorts to the end
    }
  }
#pragma unroll
  for (int size = 2; size <= 32 * E; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride >= 1; stride >>= 1) {
      if (stride >= E) {
        const int lane_stride = stride / E;
        const bool lower = (lane & lane_stride) == 0;
#pragma unroll
        for (int i = 0; i < E; ++i) {
          const unsigned long long o = __shfl_xor_sync(kFullMask, k[i], lane_stride);
          const bool asc = (((lane * E + i) & size) == 0) || size == 32 * E;
          const bool take_min = lower == asc;
          k[i] = ((o < k[i]) == take_min) ? o : k[i];                 // min or max with ONE compare (equal keys: either)
        }
      } else {
#pragma unroll
        for (int i = 0; i < E; ++i) {
          if ((i & stride) == 0) {
            const int p = i | stride;
            const bool asc = (((lane * E + i) & size) == 0) || size == 32 * E;
            const unsigned long long a = k[i], b = k[p];
            const bool swap = asc ? (a > b) : (a < b);
            k[i] = swap ? b : a;
            k[p] = swap ? a : b;
          }
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < E; ++i) {
    const int e = lane * E + i;
    if (e < Nf) {
      const int j = (int)(uint32_t)k[i];
      z_out[e] = szs[j];
      if (perm_out) perm_out[e] = j;
    }
  }
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
sample_pdf_fwd_kernel(const float* __restrict__ weights, const float* __restrict__ z, int64_t n_rays, int S, int Nf,
                      const float* __restrict__ u_in, uint64_t seed, uint32_t step, uint64_t ray_offset,
                      float* __restrict__ z_new, int* __restrict__ idx_out, int* __restrict__ perm_out,
                      float* __restrict__ u_out) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ray = blockIdx.x * (int64_t)kWarpsPerBlock + warp;
  if (ray >= n_rays) return;
  // per-warp region [w | cdf | z | new samples]; the last block starts on a 16-byte boundary (float4 loads in the sort)
  const int s3 = (3 * S + 3) & ~3, nf4 = (Nf + 3) & ~3;
  float* base = smem + (size_t)warp * (s3 + nf4);
  float *sw = base, *scdf = base + S, *sz = base + 2 * S, *szs = base + s3;
  build_cdf(weights + ray * S, z + ray * S, S, lane, sw, scdf, sz);

  const int n_blocks = (Nf + 3) / 4;
  for (int blk = lane; blk < n_blocks; blk += 32) {
    float u4[4];
    if (u_in) {
#pragma unroll
      for (int k = 0; k < 4; ++k) u4[k] = (blk * 4 + k < Nf) ? __ldcs(u_in + ray * Nf + blk * 4 + k) : 0.f;
    } else {
      float4 r = philox_uniform4(seed, (uint32_t)(ray + ray_offset), (uint32_t)blk, 1u, step);
      u4[0] = r.x; u4[1] = r.y; u4[2] = r.z; u4[3] = r.w;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int j = blk * 4 + k;
      if (j >= Nf) break;
      int idx;
      Draw d = locate(scdf, sz, S, u4[k], &idx);
      float t = __fdiv_rn(__fsub_rn(u4[k], d.lo), d.den);
      szs[j] = __fadd_rn(d.zlo, __fmul_rn(t, __fsub_rn(d.zhi, d.zlo)));
      if (idx_out) idx_out[ray * Nf + j] = idx;
      if (u_out) u_out[ray * Nf + j] = u4[k];
    }
  }
  __syncwarp();
  // Up to 256 new samples (every reference config: 64 ... 192): bitonic network, see bitonic_sort_rows.  More: the stable
  // rank sort below (tf.sort ascending; ties keep draw order like the oracle's stable sort):
  //   rank_j = #{k < j : z_k <= z_j} + #{k > j : z_k < z_j}.
  // The kernel is issue-bound on this O(Nf^2) loop, so it is arranged to cost one compare + one add per pair: in pass m
  // the warp ranks j = 32 m + lane, so every k below 32 m is "before" and every k from 32 (m + 1) on is "after" for ALL
  // lanes (warp-uniform bounds, 16-byte shared-memory loads); only the 32 k of the diagonal block need the index test.
  if (Nf <= 256) {
    float* zo = z_new + ray * Nf;
    int* po = perm_out ? perm_out + ray * Nf : nullptr;
    if (Nf <= 32) bitonic_sort_rows<1>(szs, Nf, lane, zo, po);
    else if (Nf <= 64) bitonic_sort_rows<2>(szs, Nf, lane, zo, po);
    else if (Nf <= 128) bitonic_sort_rows<4>(szs, Nf, lane, zo, po);
    else bitonic_sort_rows<8>(szs, Nf, lane, zo, po);
    return;
  }
  for (int j0 = 0; j0 < Nf; j0 += 32) {
    const int j = j0 + lane;
    const float v = j < Nf ? szs[j] : 0.f;
    int rank = 0;
    int k = 0;
    for (; k + 4 <= j0; k += 4) {                       // k < j for every lane
      const float4 o = *reinterpret_cast<const float4*>(szs + k);
      rank += (o.x <= v) + (o.y <= v) + (o.z <= v) + (o.w <= v);
    }
    const int diag_end = min(Nf, j0 + 32);
    for (; k < diag_end; ++k) {                         // diagonal block (j0 is a multiple of 4: k == j0 here)
      const float o = szs[k];
      rank += (o < v) || (o == v && k < j);
    }
    for (; k < Nf && (k & 3); ++k) rank += szs[k] < v;  // (diag_end is a multiple of 4 unless it is Nf)
    for (; k + 4 <= Nf; k += 4) {                       // k > j for every lane
      const float4 o = *reinterpret_cast<const float4*>(szs + k);
      rank += (o.x < v) + (o.y < v) + (o.z < v) + (o.w < v);
    }
    for (; k < Nf; ++k) rank += szs[k] < v;
    if (j < Nf) {
      z_new[ray * Nf + rank] = v;
      if (perm_out) perm_out[ray * Nf + rank] = j;
    }
  }
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
sample_pdf_bwd_kernel(const float* __restrict__ weights, const float* __restrict__ z, const float* __restrict__ u,
                      const int* __restrict__ perm, const float* __restrict__ d_z_new, int64_t n_rays, int S, int Nf,
                      float* __restrict__ d_weights) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ray = blockIdx.x * (int64_t)kWarpsPerBlock + warp;
  if (ray >= n_rays) return;
  // layout per warp: w[S] cdf[S] z[S] dcdf[S] | dzs[Nf] dlo[Nf] dhi[Nf] bt[Nf] (ints)
  float* base = smem + (size_t)warp * (4 * S + 4 * Nf);
  float *sw = base, *scdf = base + S, *sz = base + 2 * S, *sdc = base + 3 * S;
  float *sdz = base + 4 * S, *sdlo = sdz + Nf, *sdhi = sdlo + Nf;
  int* sbt = reinterpret_cast<int*>(sdhi + Nf);
  const float denom = build_cdf(weights + ray * S, z + ray * S, S, lane, sw, scdf, sz);

  for (int k = lane; k < Nf; k += 32) sdz[perm[ray * Nf + k]] = __ldcs(d_z_new + ray * Nf + k);
  for (int i = lane; i < S; i += 32) sdc[i] = 0.f;
  __syncwarp();
  for (int j = lane; j < Nf; j += 32) {
    float uj = __ldcs(u + ray * Nf + j);
    Draw d = locate(scdf, sz, S, uj, nullptr);
    float dt = sdz[j] * (d.zhi - d.zlo);
    float num = uj - d.lo;
    float dlo = -dt / d.den, dhi = 0.f;
    if (!d.floored) {
      float q = dt * num / (d.den * d.den);
      dlo += q;
      dhi = -q;
    }
    sdlo[j] = dlo;
    sdhi[j] = dhi;
    sbt[j] = d.b | (d.t << 16);
  }
  __syncwarp();
  // Per-draw cdf gradients -> d cdf: every cdf entry GATHERS its contributions in draw order (lower end before upper
  // end of the same draw), i.e. the same sequence of additions a sequential scatter would perform, lane-parallel over
  // the entries.  Then the reverse cumsum -> d pdf, redundantly in every lane (pipelined broadcast reads, one dependent
  // FADD per element), 32 entries at a time so no lane overwrites an entry another lane still has to read.
  for (int i = lane; i < S; i += 32) {
    float acc = 0.f;
    for (int j = 0; j < Nf; ++j) {
      const int bt = sbt[j];
      if ((bt & 0xffff) == i) acc += sdlo[j];
      if ((bt >> 16) == i) acc += sdhi[j];
    }
    sdc[i] = acc;
  }
  __syncwarp();
  float run = 0.f;
  for (int k0 = ((S - 1) >> 5) << 5; k0 >= 0; k0 -= 32) {
    float mine = 0.f;
    for (int i = min(S, k0 + 32) - 1; i >= k0; --i) {
      run += sdc[i];
      if ((i & 31) == lane) mine = run;
    }
    __syncwarp();
    if (k0 + lane < S) sdc[k0 + lane] = mine;
  }
  __syncwarp();
  // pdf = w / denom, denom = sum(w) + eps :  d w_k = dpdf_k/denom - sum_j dpdf_j w_j / denom^2
  //                                                  = dpdf_k/denom - (sum_j dpdf_j pdf_j) / denom   (sw holds the pdf)
  float part = 0.f;
  for (int i = lane; i < S; i += 32) part += sdc[i] * sw[i];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(kFullMask, part, d);
  const float corr = part / denom;
  for (int i = lane; i < S; i += 32) d_weights[ray * S + i] = sdc[i] / denom - corr;
}

__global__ void merge_sorted_kernel(const float* __restrict__ a, int sa, const float* __restrict__ b, int sb,
                                    int64_t n_rays, float* __restrict__ out, int32_t* __restrict__ rank_a) {
  const int st = sa + sb;
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n_rays * st) return;
  int64_t ray = i / st;
  int j = (int)(i % st);
  const float* ar = a + ray * sa;
  const float* br = b + ray * sb;
  float v;
  int rank;
  if (j < sa) {  // element of a: rank = j + #{b < v}
    v = ar[j];
    int lo = 0, hi = sb;
    while (lo < hi) { int m = (lo + hi) >> 1; if (br[m] < v) lo = m + 1; else hi = m; }
    rank = j + lo;
    if (rank_a) rank_a[ray * sa + j] = rank;
  } else {       // element of b: rank = jb + #{a <= v}
    int jb = j - sa;
    v = br[jb];
    int lo = 0, hi = sa;
    while (lo < hi) { int m = (lo + hi) >> 1; if (ar[m] <= v) lo = m + 1; else hi = m; }
    rank = jb + lo;
  }
  out[ray * st + rank] = v;
}

// d_a[ray][j] = d_out[ray][rank_a[ray][j]]: what the sort(concat) of src/NeRF.py:132 back-propagates to its first
// operand (a sort's gradient is a gather through its permutation).
__global__ void merge_sorted_bwd_kernel(const float* __restrict__ d_out, const int32_t* __restrict__ rank_a, int sa,
                                        int st, int64_t n_rays, float* __restrict__ d_a) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n_rays * sa) return;
  int64_t ray = i / sa;
  d_a[i] = d_out[ray * st + rank_a[i]];
}

}  // namespace nerf

using namespace nerf;

extern "C" {

int nerf_sample_pdf_fwd(const float* weights, const float* z, int64_t n_rays, int32_t n_samples, int32_t n_new,
                        const float* u_or_null, uint64_t seed, uint32_t step, uint64_t ray_offset, float* z_new,
                        int32_t* idx_or_null, int32_t* perm_or_null, float* u_out_or_null, void* stream) {
  NERF_CHECK_ARG(weights && z && z_new, "null pointer");
  NERF_CHECK_ARG(n_rays >= 0 && n_samples >= 2 && n_samples <= 1024 && n_new > 0 && n_new <= 1024,
                 "need 2 <= n_samples <= 1024 and 1 <= n_new <= 1024");
  if (n_rays == 0) return NERF_OK;
  size_t smem = (size_t)kWarpsPerBlock * (((3 * n_samples + 3) & ~3) + ((n_new + 3) & ~3)) * sizeof(float);
  if (smem > 48 * 1024)
    NERF_CUDA(cudaFuncSetAttribute(sample_pdf_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  sample_pdf_fwd_kernel<<<(unsigned)ceil_div(n_rays, kWarpsPerBlock), kWarpsPerBlock * 32, smem, (cudaStream_t)stream>>>(
      weights, z, n_rays, n_samples, n_new, u_or_null, seed, step, ray_offset, z_new, idx_or_null, perm_or_null,
      u_out_or_null);
  NERF_CHECK_LAUNCH();
  return NERF_OK;
}

int nerf_sample_pdf_bwd(const float* weights, const float* z, const float* u, const int32_t* perm, const float* d_z_new,
                        int64_t n_rays, int32_t n_samples, int32_t n_new, float* d_weights, void* stream) {
  NERF_CHECK_ARG(weights && z && u && perm && d_z_new && d_weights, "null pointer");
  NERF_CHECK_ARG(n_rays >= 0 && n_samples >= 2 && n_samples <= 1024 && n_new > 0 && n_new <= 1024,
                 "need 2 <= n_samples <= 1024 and 1 <= n_new <= 1024");
  if (n_rays == 0) return NERF_OK;
  size_t smem = (size_t)kWarpsPerBlock * (4 * n_samples + 4 * n_new) * sizeof(float);
  if (smem > 48 * 1024)
    NERF_CUDA(cudaFuncSetAttribute(sample_pdf_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  sample_pdf_bwd_kernel<<<(unsigned)ceil_div(n_rays, kWarpsPerBlock), kWarpsPerBlock * 32, smem, (cudaStream_t)stream>>>(
      weights, z, u, perm, d_z_new, n_rays, n_samples, n_new, d_weights);
  NERF_CHECK_LAUNCH();
  return NERF_OK;
}

int nerf_merge_sorted(const float* z_a, int32_t sa, const float* z_b, int32_t sb, int64_t n_rays, float* out,
                      void* stream) {
  NERF_CHECK_ARG(z_a && z_b && out, "null pointer");
  NERF_CHECK_ARG(sa > 0 && sb > 0 && n_rays >= 0, "bad shape");
  if (n_rays == 0) return NERF_OK;
  int64_t total = n_rays * (sa + sb);
  merge_sorted_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(z_a, sa, z_b, sb, n_rays, out,
                                                                                         nullptr);
  NERF_CHECK_LAUNCH();
  return NERF_OK;
}

int nerf_merge_sorted_rank(const float* z_a, int32_t sa, const float* z_b, int32_t sb, int64_t n_rays, float* out,
                           int32_t* rank_a, void* stream) {
  NERF_CHECK_ARG(z_a && z_b && out && rank_a, "null pointer");
  NERF_CHECK_ARG(sa > 0 && sb > 0 && n_rays >= 0, "bad shape");
  if (n_rays == 0) return NERF_OK;
  int64_t total = n_rays * (sa + sb);
  merge_sorted_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(z_a, sa, z_b, sb, n_rays, out,
                                                                                         rank_a);
  NERF_CHECK_LAUNCH();
  return NERF_OK;
}

int nerf_merge_sorted_bwd(const float* d_out, const int32_t* rank_a, int32_t sa, int32_t sb, int64_t n_rays,
                          float* d_a, void* stream) {
  NERF_CHECK_ARG(d_out && rank_a && d_a, "null pointer");
  NERF_CHECK_ARG(sa > 0 && sb > 0 && n_rays >= 0, "bad shape");
  if (n_rays == 0) return NERF_OK;
  merge_sorted_bwd_kernel<<<(unsigned)ceil_div(n_rays * sa, 256), 256, 0, (cudaStream_t)stream>>>(d_out, rank_a, sa,
                                                                                                  sa + sb, n_rays, d_a);
  NERF_CHECK_LAUNCH();
  return NERF_OK;
}

}  // extern "C"
