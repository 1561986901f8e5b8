// Hierarchical (inverse-CDF) sampling: get_z_vals_from_prob_dist_func, src/UtilsCV.py:502-539, forward and the
// gradient w.r.t. `weights` that TF autodiff produces inside NeRF.train_step (the reference does not detach
// z_from_dist, src/NeRF.py:155), plus the per-ray sorted merge of src/NeRF.py:132.
//
// Bit-exact indices: the searchsorted result depends on the fp32 summation order of reduce_sum and cumsum.
// The canonical order (shared with oracle/nerf_oracle.py) is sequential left-to-right fp32 with IEEE division,
// so every lane walks the S<=1024 entries of a ray that the whole warp staged in shared memory with coalesced
// loads (redundantly: the dependent chain is one FADD per entry); everything else (pdf, search, interpolation, sort) is
// lane-parallel.  1 KB/ray of HBM traffic.  nerf_hierarchical_sample runs the same kernel with the coarse ray_marching
// weights formed in place of the load and the merge with the coarse depths in place of the store (one launch between the
// two networks of a render).
#include "composite.cuh"

namespace nerf {

constexpr unsigned kFullMask = 0xffffffffu;
constexpr int kWarpsPerBlock = 4;

// The forward kernel is instruction-bound (ncu, profiles/r02_ay_*: 0.9 TB/s of algorithmic traffic, issue slots busy), so
// the hot shapes (64 coarse samples, 128 / 192 draws: every train / render step of the bench configs) are compile-time
// template arguments kS / kNf: every loop unrolls, the searches read shared memory at immediate offsets and the row
// moves 16 bytes at a time.  kS = kNf = 0 is the same code with run-time extents (any 2 <= S, Nf <= 1024).

// Stage one ray and build pdf/cdf in the canonical order. Returns (sum + eps).  `sz` receives the MID-POINTS
// 0.5 (z[i+1] + z[i]), i < S-1 (the only way the depths enter, src/UtilsCV.py:519), so a draw reads two values, not four.
// Rows that are SEARCHED (cdf, mid-points, sorted keys) are stored with one word of padding per 32 entries in the
// compile-time instantiations (kPad = 1): the probes of a binary search sit a power of two apart, i.e. in the same
// bank, and the kernel was bound by shared-memory wavefronts (2-way conflicts on the 64-entry cdf, 4-way on 128 keys).
template <int kPad>
__device__ __forceinline__ int padded(int i) { return i + kPad * (i >> 5); }

// scdf: linear cumsum row (lane 0 writes it 16 bytes at a time); scdfp: its padded copy (== scdf when kS = 0)
// raw4 != null (nerf_hierarchical_sample): the weights are not read but FORMED here from the coarse network's raw output -
// ray_marching's alpha / transmittance / weight with the arithmetic of composite_fwd_kernel (composite.cuh), bit for bit -
// and `szc` keeps the raw depths (padded row) for the merge at the end.
template <int kS, bool kFromRaw>
__device__ __forceinline__ float build_cdf(const float* __restrict__ weights, const float4* __restrict__ raw4,
                                           const float* __restrict__ z, int S_rt, int lane, float* sw, float* scdf,
                                           float* sz, float* scdfp, float* szc) {
  constexpr int kPad = kS ? 1 : 0;
  const int S = kS ? kS : S_rt;
  float carry = 1.0f;
#pragma unroll(kS ? (kS + 31) / 32 : 1)
  for (int i0 = 0; i0 < S; i0 += 32) {
    const int i = i0 + lane;
    const bool valid = i < S;
    float zi = 0.f, zn = 0.f;
    if (valid) {
      zi = __ldg(z + i);
      zn = __ldg(z + min(i + 1, S - 1));
      sz[padded<kPad>(i)] = __fmul_rn(0.5f, __fadd_rn(zn, zi));
      if (!kFromRaw) sw[i] = __ldcs(weights + i);
    }
    if (kFromRaw) {                                     // every lane: warp scan
      const SampleFwd f = sample_fwd(valid ? __ldg(&raw4[i].w) : 0.f, zi, zn, i == S - 1);
      const float incl = warp_incl_scan_mul(valid ? f.x : 1.0f, lane);
      float excl = __shfl_up_sync(kFullMask, incl, 1);
      if (lane == 0) excl = 1.0f;
      const float T = carry * excl;
      carry *= __shfl_sync(kFullMask, incl, 31);
      if (valid) {
        sw[i] = f.alpha * T;
        szc[padded<kPad>(i)] = zi;
      }
    }
  }
  __syncwarp();
  // Sequential left-to-right fp32 sum and cumsum, computed REDUNDANTLY by every lane from broadcast shared-memory reads:
  // the loads do not depend on the running value, so they pipeline and the dependent chain is one FADD per element
  // (a single lane doing load -> add -> store through the same array paid a shared-memory round trip per element).
  // 16-byte accesses when S is a multiple of four: 16 + 64 issue slots for the sum of 64 weights, 16 + 64 + 16 for the
  // cumsum (lane 0 stores it).
  const bool vec = (S & 3) == 0;
  float total = 0.f;
  if (vec) {
    const float4* sw4 = reinterpret_cast<const float4*>(sw);
#pragma unroll(kS ? kS / 4 : 4)
    for (int i = 0; i < (S >> 2); ++i) {
      const float4 v = sw4[i];
      total = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(total, v.x), v.y), v.z), v.w);
    }
  } else {
#pragma unroll 8
    for (int i = 0; i < S; ++i) total = __fadd_rn(total, sw[i]);
  }
  const float denom = __fadd_rn(total, 1e-7f);  // EPS, src/UtilsCV.py:30
  __syncwarp();
#pragma unroll(kS ? (kS + 31) / 32 : 1)
  for (int i0 = 0; i0 < S; i0 += 32) {            // pdf (the weights are not needed after this)
    const int i = i0 + lane;
    if (i < S) sw[i] = __fdiv_rn(sw[i], denom);
  }
  __syncwarp();
  float run = 0.f;
  if (vec) {
    const float4* sw4 = reinterpret_cast<const float4*>(sw);
    float4* sc4 = reinterpret_cast<float4*>(scdf);
#pragma unroll(kS ? kS / 4 : 4)
    for (int i = 0; i < (S >> 2); ++i) {
      const float4 v = sw4[i];
      float4 o;
      o.x = run = __fadd_rn(run, v.x);
      o.y = run = __fadd_rn(run, v.y);
      o.z = run = __fadd_rn(run, v.z);
      o.w = run = __fadd_rn(run, v.w);
      if (lane == 0) sc4[i] = o;
    }
  } else {
#pragma unroll 8
    for (int i = 0; i < S; ++i) {
      run = __fadd_rn(run, sw[i]);
      if ((i & 31) == lane) scdf[i] = run;
    }
  }
  __syncwarp();
  if (kS) {
#pragma unroll(kS ? (kS + 31) / 32 : 1)
    for (int i0 = 0; i0 < S; i0 += 32) {
      const int i = i0 + lane;
      if (i < S) scdfp[padded<kPad>(i)] = scdf[i];
    }
    __syncwarp();
  }
  return denom;
}

struct Draw {
  int b, t;
  float lo, hi, zlo, zhi, den;
  bool floored;
};

// #{i < n : a_i < v} for a sorted row in shared memory (tf.searchsorted side='left'): branch-free halving.  With a
// compile-time extent the probes are written in PTX so that each level is exactly LDS [addr + imm] / SETP / predicated
// ADD (the C++ form compiled to ~7 issue slots per level: pointer selects and re-derived addresses) and the kE
// independent chains of a lane are interleaved level by level.  ("memory": the rows were written by other lanes before a
// __syncwarp; the probes must stay behind it.)
template <typename T> struct SharedProbe;
template <> struct SharedProbe<float> {
  template <int kLoad, int kStep>
  static __device__ __forceinline__ void step(uint32_t& addr, float v) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .f32 x;\n\tld.shared.f32 x, [%0+%2];\n\tsetp.lt.f32 p, x, %1;\n\t@p add.u32 %0, %0, %3;\n\t}"
                 : "+r"(addr) : "f"(v), "n"(kLoad), "n"(kStep) : "memory");
  }
  static __device__ __forceinline__ uint32_t last(uint32_t addr, float v) {
    uint32_t inc;
    asm volatile("{\n\t.reg .pred p;\n\t.reg .f32 x;\n\tld.shared.f32 x, [%1];\n\tsetp.lt.f32 p, x, %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(inc) : "r"(addr), "f"(v) : "memory");
    return inc;
  }
};
template <> struct SharedProbe<uint32_t> {
  template <int kLoad, int kStep>
  static __device__ __forceinline__ void step(uint32_t& addr, uint32_t v) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .u32 x;\n\tld.shared.u32 x, [%0+%2];\n\tsetp.lt.u32 p, x, %1;\n\t@p add.u32 %0, %0, %3;\n\t}"
                 : "+r"(addr) : "r"(v), "n"(kLoad), "n"(kStep) : "memory");
  }
  static __device__ __forceinline__ uint32_t last(uint32_t addr, uint32_t v) {
    uint32_t inc;
    asm volatile("{\n\t.reg .pred p;\n\t.reg .u32 x;\n\tld.shared.u32 x, [%1];\n\tsetp.lt.u32 p, x, %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(inc) : "r"(addr), "r"(v) : "memory");
    return inc;
  }
};
// (N a power of two >= 32 in a padded row: the base of a level is a multiple of 2 half, so padded(base + x) - padded(base)
// is the same for every base)
template <int N, int kE, typename T>
__device__ __forceinline__ void lower_bound_levels(uint32_t (&addr)[kE], const T (&v)[kE]) {
  if constexpr (N > 1) {
    constexpr int half = N >> 1;
    constexpr int load = (half - 1) + ((half - 1) >> 5), step = half + (half >> 5);
#pragma unroll
    for (int e = 0; e < kE; ++e) SharedProbe<T>::template step<load * 4, step * 4>(addr[e], v[e]);
    lower_bound_levels<N - half, kE, T>(addr, v);
  }
}
// kE searches of one lane over the same row; kN > 0: compile-time extent (a power of two), PADDED row
template <int kN, int kE, typename T>
__device__ __forceinline__ void lower_bound(const T* a, int n_rt, const T (&v)[kE], int (&idx)[kE]) {
  if constexpr (kN > 0) {
    static_assert((kN & (kN - 1)) == 0 && kN >= 32, "padded search: power-of-two extent");
    const uint32_t a0 = (uint32_t)__cvta_generic_to_shared(a);
    uint32_t addr[kE];
#pragma unroll
    for (int e = 0; e < kE; ++e) addr[e] = a0;
    lower_bound_levels<kN, kE, T>(addr, v);
#pragma unroll
    for (int e = 0; e < kE; ++e) {
      const uint32_t pos = (addr[e] - a0) >> 2;                      // padded position: 33 words per 32 entries
      idx[e] = (int)(pos - pos / 33u + SharedProbe<T>::last(addr[e], v[e]));
    }
  } else {
    const T* p[kE];
#pragma unroll
    for (int e = 0; e < kE; ++e) p[e] = a;
    int n = n_rt;
    while (n > 1) {
      const int half = n >> 1;
#pragma unroll
      for (int e = 0; e < kE; ++e)
        if (p[e][half - 1] < v[e]) p[e] += half;
      n -= half;
    }
#pragma unroll
    for (int e = 0; e < kE; ++e) idx[e] = (int)(p[e] - a) + ((p[e][0] < v[e]) ? 1 : 0);
  }
}

// `szm`: the mid-points build_cdf left in shared memory; both rows padded<kPad>
template <int kPad>
__device__ __forceinline__ Draw locate(const float* scdf, const float* szm, int S, int idx) {
  Draw d;
  d.b = max(0, idx - 1);
  d.t = min(S - 1, idx);
  d.lo = scdf[padded<kPad>(d.b)];
  d.hi = scdf[padded<kPad>(d.t)];
  d.zlo = szm[padded<kPad>(min(d.b, S - 2))];
  d.zhi = szm[padded<kPad>(min(d.t, S - 2))];
  float den = __fsub_rn(d.hi, d.lo);
  d.floored = den < 1e-5f;
  d.den = d.floored ? 1e-5f : den;
  return d;
}

__device__ __forceinline__ uint32_t order_key(float v) {   // unsigned order == float order; -0 sorts with (and leaves as) +0
  uint32_t u = __float_as_uint(v);
  if ((u << 1) == 0u) u = 0u;
  return u ^ ((u >> 31) ? 0xffffffffu : 0x80000000u);
}
__device__ __forceinline__ float key_value(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k ^ 0x80000000u) : ~k);
}

// Stable ascending sort of the Nf <= 32 E new samples of a ray (tf.sort of src/NeRF.py:132; ties keep draw order, like
// the oracle's stable sort) and the permutation the backward needs.
//  * Values: a warp-wide bitonic network on bare 32-bit order-preserving keys, element e = lane E + i in register i of
//    lane `lane`.  All-ascending form: a merge of width `size` starts with a FLIP (e against e ^ (size - 1)) and continues
//    with half-cleaners (e against e ^ stride); the smaller index always keeps the minimum, so no compare-exchange needs
//    a direction select.  Partners across lanes cost one shuffle + a predicated min/max, partners inside a lane are
//    register pairs with compile-time indices (min + max).  28 steps for 128 samples.
//  * Permutation: every draw looks its own key up in the sorted row (lower bound in shared memory) = its rank when the
//    key is unique.  Runs of EQUAL keys (draws below the first / above the last cdf entry collapse onto one mid-point,
//    src/UtilsCV.py:519-527) are resolved one distinct value at a time: rank = first occurrence + the number of equal
//    draws with a smaller draw index (ballots over the E registers) - the stable order.
//  * Rays whose samples are all equal (empty rays) are already sorted: identity permutation.
// Replaces the 64-bit (z, draw index) network of the previous version: same results bit for bit
// (test_sample_pdf_bit_exact), about a third of its issue slots.
//  * kMerge (nerf_hierarchical_sample): instead of z_new the kernel writes z_all = sort(concat(z_new, z_c)) of
//    src/NeRF.py:132 - a sorted new sample lands at e + #{z_c < v}, a coarse depth at j + #{z_new <= v} (new samples precede
//    equal coarse ones: the stable order of the concatenation, the rule of merge_sorted_kernel), both counts by the same
//    searches over the padded rows.
template <int E, int kNf, int kS, bool kMerge>
__device__ __forceinline__ void sort_new_samples(const float* szs, uint32_t* ssort, int Nf_rt, int lane, bool vec_out,
                                                 float* __restrict__ z_out, int* __restrict__ perm_out,
                                                 const float* szc, int S_rt, float* __restrict__ z_all) {
  const int Nf = kNf ? kNf : Nf_rt;
  uint32_t k[E], orig[E];
  if (kNf == 32 * E && E == 4) {
    const float4 v = *reinterpret_cast<const float4*>(szs + lane * 4);
    k[0] = order_key(v.x); k[1] = order_key(v.y); k[2] = order_key(v.z); k[3] = order_key(v.w);
  } else {
#pragma unroll
    for (int i = 0; i < E; ++i) {
      const int j = lane * E + i;
      k[i] = j < Nf ? order_key(szs[j]) : 0xffffffffu;               // padding sorts to the end
    }
  }
#pragma unroll
  for (int i = 0; i < E; ++i) orig[i] = k[i];
  const uint32_t k0 = __shfl_sync(kFullMask, k[0], 0);
  bool same = true;
#pragma unroll
  for (int i = 0; i < E; ++i) same = same && (k[i] == k0 || lane * E + i >= Nf);
  if (!kMerge && __all_sync(kFullMask, same)) {
#pragma unroll
    for (int i = 0; i < E; ++i) {
      const int e = lane * E + i;
      if (e < Nf) {
        z_out[e] = szs[e];
        if (perm_out) perm_out[e] = e;
      }
    }
    return;
  }
#pragma unroll
  for (int size = 2; size <= 32 * E; size <<= 1) {
    if (size > E) {                                  // flip across lanes: lane ^ (size / E - 1), register E - 1 - i
      const bool lower = (lane & (size / (2 * E))) == 0;
      uint32_t o[E];
#pragma unroll
      for (int i = 0; i < E; ++i) o[i] = __shfl_xor_sync(kFullMask, k[E - 1 - i], size / E - 1);
#pragma unroll
      for (int i = 0; i < E; ++i) k[i] = lower ? min(k[i], o[i]) : max(k[i], o[i]);
    } else {                                         // flip inside the lane
#pragma unroll
      for (int i = 0; i < E; ++i) {
        if ((i & (size >> 1)) == 0) {
          const int p = i ^ (size - 1);
          const uint32_t a = k[i], b = k[p];
          k[i] = min(a, b);
          k[p] = max(a, b);
        }
      }
    }
#pragma unroll
    for (int stride = size >> 2; stride >= 1; stride >>= 1) {
      if (stride >= E) {
        const bool lower = (lane & (stride / E)) == 0;
#pragma unroll
        for (int i = 0; i < E; ++i) {
          const uint32_t o = __shfl_xor_sync(kFullMask, k[i], stride / E);
          k[i] = lower ? min(k[i], o) : max(k[i], o);
        }
      } else {
#pragma unroll
        for (int i = 0; i < E; ++i) {
          if ((i & stride) == 0) {
            const uint32_t a = k[i], b = k[i | stride];
            k[i] = min(a, b);
            k[i | stride] = max(a, b);
          }
        }
      }
    }
  }
  constexpr int kPad = kNf ? 1 : 0;                  // compile-time shapes: all 32 E keys (padding included) in a padded row
  if (kMerge) {
    const int S = kS ? kS : S_rt;
#pragma unroll
    for (int i = 0; i < E; ++i)
      if (kNf || lane * E + i < Nf) ssort[padded<kPad>(lane * E + i)] = k[i];
    __syncwarp();
    float v[E];
    int below[E];
#pragma unroll
    for (int i = 0; i < E; ++i) v[i] = key_value(k[i]);
    lower_bound<kS, E, float>(szc, S, v, below);                      // #{z_c < v}
#pragma unroll
    for (int i = 0; i < E; ++i)
      if (lane * E + i < Nf) z_all[lane * E + i + below[i]] = v[i];
#pragma unroll(kS ? (kS + 31) / 32 : 1)
    for (int j0 = 0; j0 < S; j0 += 32) {
      const int j = min(j0 + lane, S - 1);
      const float zc[1] = {szc[padded<kPad>(j)]};
      const uint32_t next_key[1] = {order_key(zc[0]) + 1u};           // #{key <= key(v)} = #{key < key(v) + 1}
      int not_above[1];
      lower_bound<(kNf ? 32 * E : 0), 1, uint32_t>(ssort, Nf, next_key, not_above);
      if (j0 + lane < S) z_all[j + not_above[0]] = zc[0];
    }
    return;
  }
  if (vec_out && kNf == 32 * E && (E == 4 || E == 8)) {
#pragma unroll
    for (int q = 0; q < E / 4; ++q)
      *reinterpret_cast<float4*>(z_out + lane * E + 4 * q) =
          make_float4(key_value(k[4 * q]), key_value(k[4 * q + 1]), key_value(k[4 * q + 2]), key_value(k[4 * q + 3]));
  } else {
#pragma unroll
    for (int i = 0; i < E; ++i)
      if (lane * E + i < Nf) z_out[lane * E + i] = key_value(k[i]);
  }
  if (!perm_out) return;
#pragma unroll
  for (int i = 0; i < E; ++i)
    if (kNf || lane * E + i < Nf) ssort[padded<kPad>(lane * E + i)] = k[i];
  __syncwarp();
  int rank[E];
  unsigned unresolved = 0u;
  lower_bound<(kNf ? 32 * E : 0), E, uint32_t>(ssort, Nf, orig, rank);
#pragma unroll
  for (int i = 0; i < E; ++i)
    if (lane * E + i < Nf && rank[i] + 1 < Nf && ssort[padded<kPad>(rank[i] + 1)] == orig[i]) unresolved |= 1u << i;
  const unsigned lanes_below = (1u << lane) - 1u;
  for (;;) {
    const unsigned pending = __ballot_sync(kFullMask, unresolved != 0u);
    if (!pending) break;
    uint32_t mine = 0u;
#pragma unroll
    for (int i = E - 1; i >= 0; --i)
      if ((unresolved >> i) & 1u) mine = orig[i];
    const uint32_t v = __shfl_sync(kFullMask, mine, __ffs(pending) - 1);
    int before = 0;
    bool eq[E];
#pragma unroll
    for (int i = 0; i < E; ++i) {
      eq[i] = orig[i] == v && lane * E + i < Nf;
      before += __popc(__ballot_sync(kFullMask, eq[i]) & lanes_below);
    }
#pragma unroll
    for (int i = 0; i < E; ++i) {
      if (eq[i]) {
        rank[i] += before;
        ++before;
        unresolved &= ~(1u << i);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < E; ++i)
    if (lane * E + i < Nf) perm_out[rank[i]] = lane * E + i;
}

// Shared-memory layout of one warp of the forward kernel, in floats; `padded_rows`: the compile-time instantiations
struct SamplerSmem {
  int cdf, zm, cdfp, szs, keys, zc, per_warp;
  __host__ __device__ SamplerSmem(int S, int Nf, bool padded_rows, bool merge = false) {
    const int pad_s = padded_rows ? ((S + 31) >> 5) : 0;
    cdf = S;
    zm = 2 * S;
    cdfp = padded_rows ? zm + S + pad_s : cdf;                       // run-time extents: the searches read the linear row
    szs = ((padded_rows ? cdfp + S + pad_s : 3 * S) + 3) & ~3;      // 16-byte boundary (float4 accesses)
    keys = szs + ((Nf + 3) & ~3);
    const int n_keys = Nf <= 32 ? 32 : (Nf <= 64 ? 64 : (Nf <= 128 ? 128 : 256));   // all 32 E keys of the network
    zc = (keys + (padded_rows ? n_keys + (n_keys >> 5) : ((Nf + 3) & ~3)) + 3) & ~3;       // raw depths (merge only)
    per_warp = (zc + (merge ? S + pad_s : 0) + 3) & ~3;
  }
};

// vec_io: z_new, u_out and idx_out rows start on 16-byte boundaries (checked by the launcher)
// kMerge: `weights` is the coarse network's raw output (N,S,4) and `z_new` receives z_all (N,S+Nf): the coarse weights, the
// draws, their sort and the merge with the coarse depths in ONE launch (nerf_hierarchical_sample).
template <int kS, int kNf, bool kMerge>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
sample_pdf_fwd_kernel(const float* __restrict__ weights, const float* __restrict__ z, int64_t n_rays, int S_rt, int Nf_rt,
                      const float* __restrict__ u_in, uint64_t seed, uint32_t step, uint64_t ray_offset,
                      float* __restrict__ z_new, int* __restrict__ idx_out, int* __restrict__ perm_out,
                      float* __restrict__ u_out, bool vec_io) {
  extern __shared__ float smem[];
  const int S = kS ? kS : S_rt, Nf = kNf ? kNf : Nf_rt;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ray = blockIdx.x * (int64_t)kWarpsPerBlock + warp;
  if (ray >= n_rays) return;
  // per-warp region (SamplerSmem): [w | cdf | z mid-points | padded cdf | new samples | sorted keys]
  constexpr int kPad = kS ? 1 : 0;
  const SamplerSmem L(S, Nf, kS != 0, kMerge);
  float* base = smem + (size_t)warp * L.per_warp;
  float *sw = base, *scdf = base + L.cdf, *sz = base + L.zm, *scdfp = base + L.cdfp, *szs = base + L.szs;
  float* szc = base + L.zc;
  uint32_t* ssort = reinterpret_cast<uint32_t*>(base + L.keys);
  build_cdf<kS, kMerge>(kMerge ? nullptr : weights + ray * S,
                        kMerge ? reinterpret_cast<const float4*>(weights) + ray * S : nullptr, z + ray * S, S, lane, sw,
                        scdf, sz, scdfp, szc);

  const int n_blocks = (Nf + 3) / 4;
  const bool vec = vec_io && (Nf & 3) == 0;
#pragma unroll(kNf ? (kNf + 127) / 128 : 1)
  for (int blk0 = 0; blk0 < n_blocks; blk0 += 32) {
    const int blk = blk0 + lane;
    if (blk >= n_blocks) break;
    float u4[4];
    if (u_in) {
#pragma unroll
      for (int k = 0; k < 4; ++k) u4[k] = (blk * 4 + k < Nf) ? __ldcs(u_in + ray * Nf + blk * 4 + k) : 0.f;
    } else {
      float4 r = philox_uniform4(seed, (uint32_t)(ray + ray_offset), (uint32_t)blk, 1u, step);
      u4[0] = r.x; u4[1] = r.y; u4[2] = r.z; u4[3] = r.w;
    }
    int idx4[4];
    float zn[4];
    lower_bound<kS, 4, float>(scdfp, S, u4, idx4);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const Draw d = locate<kPad>(scdfp, sz, S, idx4[k]);
      const float t = __fdiv_rn(__fsub_rn(u4[k], d.lo), d.den);
      zn[k] = __fadd_rn(d.zlo, __fmul_rn(t, __fsub_rn(d.zhi, d.zlo)));
    }
    if (vec) {
      *reinterpret_cast<float4*>(szs + 4 * blk) = make_float4(zn[0], zn[1], zn[2], zn[3]);
      if (idx_out) *reinterpret_cast<int4*>(idx_out + ray * Nf + 4 * blk) = make_int4(idx4[0], idx4[1], idx4[2], idx4[3]);
      if (u_out) *reinterpret_cast<float4*>(u_out + ray * Nf + 4 * blk) = make_float4(u4[0], u4[1], u4[2], u4[3]);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int j = blk * 4 + k;
        if (j < Nf) {
          szs[j] = zn[k];
          if (idx_out) idx_out[ray * Nf + j] = idx4[k];
          if (u_out) u_out[ray * Nf + j] = u4[k];
        }
      }
    }
  }
  __syncwarp();
  // Up to 256 new samples (every reference config: 64 ... 192): bitonic network, see sort_new_samples.  More: the stable
  // rank sort below (tf.sort ascending; ties keep draw order like the oracle's stable sort):
  //   rank_j = #{k < j : z_k <= z_j} + #{k > j : z_k < z_j},
  // one compare + one add per pair: in pass m the warp ranks j = 32 m + lane, so every k below 32 m is "before" and every
  // k from 32 (m + 1) on is "after" for ALL lanes (warp-uniform bounds, 16-byte shared-memory loads); only the 32 k of the
  // diagonal block need the index test.
  if (Nf <= 256) {
    float* zo = kMerge ? nullptr : z_new + ray * Nf;
    float* za = kMerge ? z_new + ray * (S + Nf) : nullptr;
    int* po = (perm_out && !kMerge) ? perm_out + ray * Nf : nullptr;
    if (kNf) {
      constexpr int E = kNf <= 32 ? 1 : (kNf <= 64 ? 2 : (kNf <= 128 ? 4 : 8));
      sort_new_samples<E, kNf, kS, kMerge>(szs, ssort, Nf, lane, vec, zo, po, szc, S, za);
    } else if (Nf <= 32) sort_new_samples<1, 0, 0, kMerge>(szs, ssort, Nf, lane, false, zo, po, szc, S, za);
    else if (Nf <= 64) sort_new_samples<2, 0, 0, kMerge>(szs, ssort, Nf, lane, false, zo, po, szc, S, za);
    else if (Nf <= 128) sort_new_samples<4, 0, 0, kMerge>(szs, ssort, Nf, lane, false, zo, po, szc, S, za);
    else sort_new_samples<8, 0, 0, kMerge>(szs, ssort, Nf, lane, false, zo, po, szc, S, za);
    return;
  }
  if constexpr (kNf == 0 && !kMerge) {                 // (the compile-time shapes and the merge never get here: <= 256 draws)
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
}

// Shared-memory layout of one warp of the backward kernel, in floats
struct SamplerBwdSmem {
  int cdf, zm, cdfp, dc, dz, lo, hi, bt, llo, lhi, cnt, start, per_warp;
  __host__ __device__ SamplerBwdSmem(int S, int Nf, bool padded_rows) {
    const int pad_s = padded_rows ? ((S + 31) >> 5) : 0;
    cdf = S;
    zm = 2 * S;
    cdfp = padded_rows ? zm + S + pad_s : cdf;
    dc = ((padded_rows ? cdfp + S + pad_s : 3 * S) + 3) & ~3;        // 16-byte boundary (float4 reverse cumsum)
    dz = dc + ((S + 3) & ~3);
    lo = dz + Nf; hi = lo + Nf; bt = hi + Nf; llo = bt + Nf; lhi = llo + Nf;
    cnt = lhi + Nf;
    start = cnt + S + 2;
    per_warp = (start + S + 3 + 3) & ~3;
  }
};

// Gradient w.r.t. the weights.  Per draw: d z_new -> d cdf[b], d cdf[t] (b = max(0, idx - 1), t = min(S - 1, idx)).  The
// per-draw contributions are then summed into d cdf DETERMINISTICALLY without the O(S Nf) gather of the previous version
// (every entry scanning every draw: 2 000 of the kernel's 3 000 issue slots per ray): a counting sort by idx.  Draws get a
// ticket inside their bin in a fixed order (register slot, then lane: __match_any_sync + the lanes below), a warp scan of
// the 65 bin counts gives the bin starts, the values are placed into two bin-ordered lists and every cdf entry adds up
// ITS contiguous list ranges in a fixed order (upper ends of bin i, lower ends of bin i + 1, the two clamped edge bins).
// Same inputs -> same additions in the same order on every run and every device.
template <int kS, int kNf>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
sample_pdf_bwd_kernel(const float* __restrict__ weights, const float* __restrict__ z, const float* __restrict__ u,
                      const int* __restrict__ perm, const float* __restrict__ d_z_new, int64_t n_rays, int S_rt, int Nf_rt,
                      float* __restrict__ d_weights) {
  extern __shared__ float smem[];
  constexpr int kPad = kS ? 1 : 0;
  const int S = kS ? kS : S_rt, Nf = kNf ? kNf : Nf_rt;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ray = blockIdx.x * (int64_t)kWarpsPerBlock + warp;
  if (ray >= n_rays) return;
  const SamplerBwdSmem L(S, Nf, kS != 0);
  float* base = smem + (size_t)warp * L.per_warp;
  float *sw = base, *scdf = base + L.cdf, *sz = base + L.zm, *scdfp = base + L.cdfp, *sdc = base + L.dc;
  float *sdz = base + L.dz, *slo = base + L.lo, *shi = base + L.hi, *llo = base + L.llo, *lhi = base + L.lhi;
  int *sbt = reinterpret_cast<int*>(base + L.bt), *cnt = reinterpret_cast<int*>(base + L.cnt),
      *start = reinterpret_cast<int*>(base + L.start);
  const float denom = build_cdf<kS, false>(weights + ray * S, nullptr, z + ray * S, S, lane, sw, scdf, sz, scdfp, nullptr);

#pragma unroll(kNf ? (kNf + 31) / 32 : 1)
  for (int k0 = 0; k0 < Nf; k0 += 32)
    if (k0 + lane < Nf) sdz[perm[ray * Nf + k0 + lane]] = __ldcs(d_z_new + ray * Nf + k0 + lane);
#pragma unroll(kS ? (kS + 33) / 32 : 1)
  for (int i0 = 0; i0 <= S; i0 += 32)
    if (i0 + lane <= S) cnt[i0 + lane] = 0;
  __syncwarp();
  const unsigned lanes_below = (1u << lane) - 1u;
  const int n_blocks = (Nf + 3) / 4;
#pragma unroll(kNf ? (kNf + 127) / 128 : 1)
  for (int blk0 = 0; blk0 < n_blocks; blk0 += 32) {          // every lane stays in the loop: __match_any_sync below
    const int blk = blk0 + lane;
    float u4[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) u4[k] = (blk < n_blocks && blk * 4 + k < Nf) ? __ldcs(u + ray * Nf + blk * 4 + k) : 0.f;
    int idx4[4];
    lower_bound<kS, 4, float>(scdfp, S, u4, idx4);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int j = blk * 4 + k;
      const bool valid = blk < n_blocks && j < Nf;
      if (valid) {
        const Draw d = locate<kPad>(scdfp, sz, S, idx4[k]);
        const float dt = sdz[j] * (d.zhi - d.zlo);
        const float num = u4[k] - d.lo;
        float dlo = -dt / d.den, dhi = 0.f;
        if (!d.floored) {
          const float q = dt * num / (d.den * d.den);
          dlo += q;
          dhi = -q;
        }
        slo[j] = dlo;
        shi[j] = dhi;
      }
      const int key = valid ? idx4[k] : -1 - lane;              // idle lanes match nobody
      const unsigned peers = __match_any_sync(kFullMask, key);
      const int taken = valid ? cnt[idx4[k]] : 0;
      if (valid) sbt[j] = idx4[k] | ((taken + __popc(peers & lanes_below)) << 16);
      __syncwarp();
      if (valid && (peers & lanes_below) == 0u) cnt[idx4[k]] = taken + __popc(peers);
      __syncwarp();
    }
  }
  // bin starts: exclusive scan of the S + 1 counts
  int carry = 0;
#pragma unroll(kS ? (kS + 32) / 32 : 1)
  for (int c0 = 0; c0 <= S; c0 += 32) {
    const int i = c0 + lane;
    const int v = i <= S ? cnt[i] : 0;
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(kFullMask, incl, d);
      if (lane >= d) incl += o;
    }
    if (i <= S) start[i] = carry + incl - v;
    carry += __shfl_sync(kFullMask, incl, 31);
  }
  if (lane == 0) start[S + 1] = carry;
  __syncwarp();
#pragma unroll(kNf ? (kNf + 31) / 32 : 1)
  for (int j0 = 0; j0 < Nf; j0 += 32) {
    const int j = j0 + lane;
    if (j < Nf) {
      const int bt = sbt[j];
      const int pos = start[bt & 0xffff] + (bt >> 16);
      llo[pos] = slo[j];
      lhi[pos] = shi[j];
    }
  }
  __syncwarp();
#pragma unroll(kS ? (kS + 31) / 32 : 1)
  for (int i0 = 0; i0 < S; i0 += 32) {
    const int i = i0 + lane;
    if (i < S) {
      const int a = start[i], b = start[i + 1], c = start[i + 2];
      float acc = 0.f;
      for (int p = a; p < b; ++p) acc += lhi[p];               // t == i: draws of bin i
      for (int p = b; p < c; ++p) acc += llo[p];               // b == i: draws of bin i + 1
      if (i == 0)
        for (int p = a; p < b; ++p) acc += llo[p];             // idx == 0: b == t == 0
      if (i == S - 1)
        for (int p = b; p < c; ++p) acc += lhi[p];             // idx == S: b == t == S - 1
      sdc[i] = acc;
    }
  }
  __syncwarp();
  // reverse cumsum -> d pdf, redundantly in every lane (pipelined broadcast reads, one dependent FADD per element)
  float run = 0.f;
  if ((S & 3) == 0) {
    float4* dc4 = reinterpret_cast<float4*>(sdc);
#pragma unroll(kS ? kS / 4 : 4)
    for (int i = (S >> 2) - 1; i >= 0; --i) {
      const float4 v = dc4[i];
      float4 o;
      o.w = run = run + v.w;
      o.z = run = run + v.z;
      o.y = run = run + v.y;
      o.x = run = run + v.x;
      if (lane == 0) dc4[i] = o;
    }
  } else {
    for (int k0 = ((S - 1) >> 5) << 5; k0 >= 0; k0 -= 32) {   // 32 entries at a time: no lane overwrites an entry another
      float mine = 0.f;                                        // lane still has to read
      for (int i = min(S, k0 + 32) - 1; i >= k0; --i) {
        run += sdc[i];
        if ((i & 31) == lane) mine = run;
      }
      __syncwarp();
      if (k0 + lane < S) sdc[k0 + lane] = mine;
    }
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
  const bool fixed = n_samples == 64 && (n_new == 128 || n_new == 192);
  size_t smem = (size_t)kWarpsPerBlock * SamplerSmem(n_samples, n_new, fixed).per_warp * sizeof(float);
  const bool vec_io = (((uintptr_t)z_new | (uintptr_t)idx_or_null | (uintptr_t)u_out_or_null) & 15) == 0;
  const unsigned grid = (unsigned)ceil_div(n_rays, kWarpsPerBlock);
#define NERF_SAMPLER_LAUNCH(KS, KNF)                                                                                   \
  do {                                                                                                                 \
    if (smem > 48 * 1024)                                                                                              \
      NERF_CUDA(cudaFuncSetAttribute(sample_pdf_fwd_kernel<KS, KNF, false>,                                            \
                                     cudaFuncAttributeMaxDynamicSharedMemorySize,                                      \
                                     (int)smem));                                                                      \
    sample_pdf_fwd_kernel<KS, KNF, false><<<grid, kWarpsPerBlock * 32, smem, (cudaStream_t)stream>>>(                  \
        weights, z, n_rays, n_samples, n_new, u_or_null, seed, step, ray_offset, z_new, idx_or_null, perm_or_null,     \
        u_out_or_null, vec_io);                                                                                        \
  } while (0)
  if (n_samples == 64 && n_new == 128) NERF_SAMPLER_LAUNCH(64, 128);
  else if (n_samples == 64 && n_new == 192) NERF_SAMPLER_LAUNCH(64, 192);
  else NERF_SAMPLER_LAUNCH(0, 0);
#undef NERF_SAMPLER_LAUNCH
  NERF_CHECK_LAUNCH();
  return NERF_OK;
}

int nerf_hierarchical_sample(const float* raw4_coarse, const float* z_coarse, int64_t n_rays, int32_t n_samples,
                             int32_t n_new, uint64_t seed, uint32_t step, uint64_t ray_offset, float* z_all, void* stream) {
  NERF_CHECK_ARG(raw4_coarse && z_coarse && z_all, "null pointer");
  NERF_CHECK_ARG(n_rays >= 0 && n_samples >= 2 && n_samples <= 1024 && n_new > 0 && n_new <= 256,
                 "need 2 <= n_samples <= 1024 and 1 <= n_new <= 256");
  NERF_CHECK_ARG(((uintptr_t)raw4_coarse & 15) == 0, "raw4_coarse must be 16-byte aligned");
  if (n_rays == 0) return NERF_OK;
  const bool fixed = n_samples == 64 && (n_new == 128 || n_new == 192);   // 128: the render setting of every reference YAML
  size_t smem = (size_t)kWarpsPerBlock * SamplerSmem(n_samples, n_new, fixed, true).per_warp * sizeof(float);
  const unsigned grid = (unsigned)ceil_div(n_rays, kWarpsPerBlock);
#define NERF_HSAMPLE_LAUNCH(KS, KNF)                                                                                   \
  do {                                                                                                                 \
    if (smem > 48 * 1024)                                                                                              \
      NERF_CUDA(cudaFuncSetAttribute(sample_pdf_fwd_kernel<KS, KNF, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                     (int)smem));                                                                      \
    sample_pdf_fwd_kernel<KS, KNF, true><<<grid, kWarpsPerBlock * 32, smem, (cudaStream_t)stream>>>(                   \
        raw4_coarse, z_coarse, n_rays, n_samples, n_new, nullptr, seed, step, ray_offset, z_all, nullptr, nullptr,     \
        nullptr, true);                                                                                                \
  } while (0)
  if (fixed && n_new == 128) NERF_HSAMPLE_LAUNCH(64, 128);
  else if (fixed) NERF_HSAMPLE_LAUNCH(64, 192);
  else NERF_HSAMPLE_LAUNCH(0, 0);
#undef NERF_HSAMPLE_LAUNCH
  NERF_CHECK_LAUNCH();
  return NERF_OK;
}

int nerf_sample_pdf_bwd(const float* weights, const float* z, const float* u, const int32_t* perm, const float* d_z_new,
                        int64_t n_rays, int32_t n_samples, int32_t n_new, float* d_weights, void* stream) {
  NERF_CHECK_ARG(weights && z && u && perm && d_z_new && d_weights, "null pointer");
  NERF_CHECK_ARG(n_rays >= 0 && n_samples >= 2 && n_samples <= 1024 && n_new > 0 && n_new <= 1024,
                 "need 2 <= n_samples <= 1024 and 1 <= n_new <= 1024");
  if (n_rays == 0) return NERF_OK;
  const bool fixed = n_samples == 64 && n_new == 128;
  size_t smem = (size_t)kWarpsPerBlock * SamplerBwdSmem(n_samples, n_new, fixed).per_warp * sizeof(float);
  const unsigned grid = (unsigned)ceil_div(n_rays, kWarpsPerBlock);
#define NERF_SAMPLER_BWD_LAUNCH(KS, KNF)                                                                               \
  do {                                                                                                                 \
    if (smem > 48 * 1024)                                                                                              \
      NERF_CUDA(cudaFuncSetAttribute(sample_pdf_bwd_kernel<KS, KNF>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                     (int)smem));                                                                      \
    sample_pdf_bwd_kernel<KS, KNF><<<grid, kWarpsPerBlock * 32, smem, (cudaStream_t)stream>>>(                         \
        weights, z, u, perm, d_z_new, n_rays, n_samples, n_new, d_weights);                                            \
  } while (0)
  if (fixed) NERF_SAMPLER_BWD_LAUNCH(64, 128);
  else NERF_SAMPLER_BWD_LAUNCH(0, 0);
#undef NERF_SAMPLER_BWD_LAUNCH
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
