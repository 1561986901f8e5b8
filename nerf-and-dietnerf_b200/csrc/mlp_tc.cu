// K2 in NERF_MODE_BF16: the coarse/fine NeRF MLP (src/NeRF.py:290-340, evaluated by model_predict,
// src/UtilsNeuralRadianceField.py:214-234) as ONE persistent tcgen05/TMEM kernel per call.
//
// Design (DESIGN.md "K2"):
//   * CTA pairs (cluster of 2 on one TPC, tcgen05 cta_group::2), one CTA per SM, 576 threads: warps 0-7 = epilogue of
//     row-tile A, warps 8-15 = epilogue of row-tile B, warp 16 = weight producer (1-D bulk async copies, TMA engine),
//     warp 17 = tcgen05.mma issuer in the leader CTA (one elected lane) / weight-arrival relay in the peer CTA;
//   * a pair owns TWO 256-sample super-tiles at a time (128 rows of each per CTA).  One MMA (M = 256, N = 256, K = 16)
//     covers a whole super-tile; each CTA streams only its half of every weight chunk ([N/2][64] bf16 = 16 KB,
//     128-byte swizzle, pre-swizzled in HBM by nerf_pack_weights) through a 4-stage ring;
//   * activations never leave the SM: the fp32 accumulator (128 x 256) of each tile lives in TMEM (2 x 256 columns),
//     the epilogue warps read it with tcgen05.ld, add bias, apply LeakyReLU, round to bf16 and write the next layer's
//     A operand back into the SAME swizzled shared-memory panels (all MMAs of the layer have completed by then);
//   * the 33-d xyz encoding and the 24-d view encoding share one 64-column input panel (cols 0..32 xyz, 40..63 view)
//     that feeds layer 0, the skip connection of layer 4 and the view branch of layer 8 (zero weight rows elsewhere);
//   * layer 8 is N = 144: 128 last-hidden units + the sigma head (column 128); the 128->3 rgb head is evaluated on
//     CUDA cores from the fp32 registers of the last epilogue;
//   * training mode also stores each layer's bf16 activations to HBM straight from the epilogue registers (coalesced
//     512-byte warp stores) in the row-block chunk-major layout (mlp_tc.cuh) the dW kernel bulk-loads as an un-swizzled
//     MN-major operand; only the input panel travels as a bulk S2G copy, once per tile.
#include <stdlib.h>

#include "mlp_tc.cuh"

namespace nerf {

// ---- weight packing -----------------------------------------------------------------------------------------------------
// One thread per bf16 element of every chunk, plus the fp32 tail (biases, rgb head).
// hi/lo split of an fp32 bias into two 16-bit values (bf16 or fp16) whose sum carries ~16 mantissa bits
template <typename T16>
__device__ __forceinline__ void split_bias(float b, T16* hi, T16* lo);
template <>
__device__ __forceinline__ void split_bias<__nv_bfloat16>(float b, __nv_bfloat16* hi, __nv_bfloat16* lo) {
  *hi = __float2bfloat16_rn(b);
  *lo = __float2bfloat16_rn(b - __bfloat162float(*hi));
}
template <>
__device__ __forceinline__ void split_bias<__half>(float b, __half* hi, __half* lo) {
  *hi = __float2half_rn(b);
  *lo = __float2half_rn(b - __half2float(*hi));
}

__device__ __forceinline__ float layer_bias(const NetGeom& g, const float* P, int layer, int n) {
  if (layer < 8 || !g.view) return n < g.layers[layer].out ? P[g.layers[layer].b_off + n] : 0.f;   // xyz-only: Dense l = layer l
  if (n < 128) return P[g.layers[8].b_off + n];
  if (n == 128) return P[g.layers[10].b_off];       // sigma head shares the [h8 ; view] input
  return 0.f;
}

template <typename T16>
__device__ __forceinline__ T16 to16(float v);
template <> __device__ __forceinline__ __nv_bfloat16 to16<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half to16<__half>(float v) { return __float2half_rn(v); }

template <typename T16>
__global__ void pack_weights_kernel(const __grid_constant__ TcPlan plan, NetGeom g, const float* __restrict__ P,
                                    uint8_t* __restrict__ packed) {
  const int chunk = blockIdx.y;
  if (chunk < plan.n_chunks) {
    // find the layer of this chunk
    int layer = 0;
    while (!(chunk >= plan.layer_first[layer] && chunk < plan.layer_first[layer] + plan.layer_nchunks[layer])) ++layer;
    const int n_rows = plan.layer_n[layer];
    const int e = blockIdx.x * blockDim.x + threadIdx.x;  // element index within [n_rows][64]
    const int src = plan.a_src[chunk];
    if (src == 5) {
      // bias slab [256][16], un-swizzled K-major: core matrices of 8 rows x 16 B; k = input-panel column - 32
      if (e >= 256 * 16) return;
      const int n = e >> 4, k = e & 15;
      T16 hi, lo;
      split_bias<T16>(n < n_rows ? layer_bias(g, P, layer, n) : 0.f, &hi, &lo);
      const T16 v = (k == kInpOneCol - 32) ? hi : (k == kInpOneCol + 1 - 32) ? lo : to16<T16>(0.f);
      *reinterpret_cast<T16*>(packed + plan.chunk_off[chunk] + (n >> 3) * 256 + (k >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2) = v;
      return;
    }
    if (e >= n_rows * 64) return;
    const int n = e >> 6, k = e & 63;
    // source Dense layer and output column
    int dense = layer, col = n;
    bool valid = n < g.layers[layer].out || (g.view && layer == 8);
    if (g.view && layer == 8) {
      if (n < 128) { dense = 8; col = n; }
      else if (n == 128) { dense = 10; col = 0; }   // sigma head shares the [h8 ; view] input
      else valid = false;
    }
    // source row of the Dense kernel (in, out)
    int row = -1;
    if (src < 4) {
      int feat = src * 64 + k;                        // activation feature
      row = (layer == 4) ? g.dx + feat : feat;         // layer 4 kernel rows: [xyz (dx) ; h4 (256)]
    } else {
      if (layer == 0 || layer == 4) { if (k < g.dx) row = k; }                       // xyz columns
      else if (layer == 8 && g.view) { if (k >= kInpViewCol && k < kInpViewCol + g.dv) row = g.hidden + (k - kInpViewCol); }
    }
    T16 v = to16<T16>(0.f);
    if (valid && row >= 0) {
      const LayerDesc& L = g.layers[dense];
      v = to16<T16>(P[L.w_off + (int64_t)row * L.out + col]);
    }
    if (src == 4 && (k == kInpOneCol || k == kInpOneCol + 1)) {   // the constant-1 columns meet the bias
      T16 hi, lo;
      split_bias<T16>(layer_bias(g, P, layer, n), &hi, &lo);
      v = (k == kInpOneCol) ? hi : lo;
    }
    *reinterpret_cast<T16*>(packed + plan.chunk_off[chunk] + panel_offset(n, k)) = v;
  } else {
    // tail: fp32 rgb head; xyz-only network: the sigma head (Dense 11, input h8) in the operand format, fp32 bias
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const LayerDesc& L = g.layers[g.view ? 9 : 10];
    if (e < 128)
      reinterpret_cast<float4*>(packed + plan.w_rgb_off)[e] =
          make_float4(P[L.w_off + e * 3 + 0], P[L.w_off + e * 3 + 1], P[L.w_off + e * 3 + 2], 0.f);
    if (e == 128) reinterpret_cast<float4*>(packed + plan.w_rgb_off)[128] = make_float4(P[L.b_off], P[L.b_off + 1], P[L.b_off + 2], 0.f);
    if (e < 256) reinterpret_cast<T16*>(packed + plan.w_sig_off)[e] = to16<T16>(g.view ? 0.f : P[g.layers[11].w_off + e]);
    if (e == 256) *reinterpret_cast<float*>(packed + plan.w_sig_off + 512) = g.view ? 0.f : P[g.layers[11].b_off];
  }
}

// ---- forward kernel --------------------------------------------------------------------------------------------------------
// NERF_TC_DEBUG & kDbgTiming: CTA 0 records the clock of its first handshake events and prints the timeline at exit
__device__ long long g_trace[3][160];
__device__ __forceinline__ void trace(bool on, int who, int& n, int code) {
  if (on && n < 160) { g_trace[who][n] = (clock64() << 8) | (long long)(code & 255); ++n; }
}

struct FwdBars {
  uint64_t full[kStages], empty[kStages], act_ready[2], acc_full[2];
  uint64_t panel_full[2], panel_free[2];   // train mode: epilogue warps <-> the tile's store warp
  uint32_t tmem_base;
};
static_assert(sizeof(FwdBars) <= 192, "barrier block overflows its shared-memory slot");

// Where the MLP input comes from: pre-encoded rows (xyz_enc / view_enc, the standalone model_predict entry point) or,
// when xyz_enc == nullptr, rays + depths: then the sample position o + d z and both sin/cos encodings are computed in
// the prologue and written straight into the shared-memory input panel (K1 fused into K2; nothing touches HBM).
// With gen_z the depths are not read either: the prologue draws the stratified coarse depths itself (get_z_values,
// src/UtilsCV.py:565-581 -- the same Philox stream and arithmetic as nerf_stratified_z, bit for bit) and writes them to
// z_out for the compositing / sampler kernels that follow.
struct FwdInput {
  const float* xyz_enc;
  const float* view_enc;
  const float4* origs;
  const float4* dirs;
  const float* z;
  int32_t n_samples, Lx, Lv, ncomp, dx, dv;
  int32_t gen_z;
  float z_start, z_end, span;
  uint32_t step;
  uint64_t seed, ray_offset;
  float* z_out;
  int32_t cam;             // 1: the rays are generated here from `camera` (origs / dirs unused)
  CameraRays camera;
};

// origin / direction of ray `ray` of the call: read, or generated from the camera (pinhole_ray_dir: the arithmetic of
// nerf_ray_directions, bit for bit).  A real call for the same reason as draw_stratified_z.
__device__ __noinline__ float4 camera_dir(const float* c2w, float tan_half_fov, int h, int w, int64_t r) {
  return pinhole_ray_dir(c2w, tan_half_fov, h, w, r);
}
__device__ __forceinline__ float4 ray_dir(const FwdInput& in, int64_t ray) {
  if (!in.cam) return __ldg(in.dirs + ray);
  return camera_dir(in.camera.c2w, in.camera.tan_half_fov, in.camera.h, in.camera.w, in.camera.ray_begin + ray);
}
__device__ __forceinline__ float4 ray_orig(const FwdInput& in, int64_t ray) {
  if (!in.cam) return __ldg(in.origs + ray);
  return make_float4(in.camera.c2w[3], in.camera.c2w[7], in.camera.c2w[11], in.camera.c2w[15]);
}

// A real call on purpose: inlined, the ten Philox rounds cost the kernel's hot loops their last spare registers (96 cap).
__device__ __noinline__ float draw_stratified_z(float z_start, float z_end, float span, int n_samples, int s, uint64_t seed,
                                                uint32_t ray, uint32_t step) {
  const float4 u4 = philox_uniform4(seed, ray, (uint32_t)(s >> 2), 0u, step);
  const float u = (s & 3) == 0 ? u4.x : (s & 3) == 1 ? u4.y : (s & 3) == 2 ? u4.z : u4.w;
  return stratified_z_value(z_start, z_end, span, n_samples, s, u);
}

// depth of MLP row `row` (ray-major): read, or drawn here (and published) when in.gen_z
__device__ __forceinline__ float row_depth(const FwdInput& in, int64_t row, int64_t ray) {
  if (!in.gen_z) return __ldg(in.z + row);
  const int s = (int)(row - ray * in.n_samples);
  const float zz = draw_stratified_z(in.z_start, in.z_end, in.span, in.n_samples, s, in.seed,
                                     (uint32_t)(ray + (int64_t)in.ray_offset), in.step);
  in.z_out[row] = zz;
  return zz;
}

template <bool kHalf>
__device__ __forceinline__ void sts_16(uint32_t panel_row_addr, int r, int col, float v) {
  const uint32_t both = pack_16x2<kHalf>(v, 0.f);
  const uint32_t addr = panel_row_addr + ((((uint32_t)col >> 3) ^ ((uint32_t)r & 7u)) << 4) + (((uint32_t)col & 7u) << 1);
  asm volatile("st.shared.b16 [%0], %1;" ::"r"(addr), "h"((uint16_t)(both & 0xffffu)) : "memory");
}

// Sign bits of 32 packed 16-bit values (8 words u[k] = high bytes of elements 4k..4k+3, from PRMT):
// element e lands in bit 8 (e & 3) + 7 - (e >> 2); a SET bit means NEGATIVE (LeakyReLU' = alpha).
__device__ __forceinline__ uint32_t hi_bytes(uint32_t a, uint32_t b) { return __byte_perm(a, b, 0x7531); }
__device__ __forceinline__ uint32_t neg_mask32(const uint32_t (&u)[8]) {
  uint32_t m = u[0] & 0x80808080u;
#pragma unroll
  for (int k = 1; k < 8; ++k) m |= (u[k] & 0x80808080u) >> k;
  return m;
}

// Epilogue of one 32-column group: acc (bias already inside, it rode in the MMA) -> LeakyReLU -> 16-bit -> swizzled
// panel row.  kMask: returns the negative-sign mask of the 32 values.
// pbx = (panel row address) | ((row & 7) << 4): a 16-byte chunk address is one XOR with an immediate.
template <bool kHalf>
__device__ __forceinline__ float2 unpack_16x2(uint32_t v) {
  if (kHalf) return __half22float2(*reinterpret_cast<const __half2*>(&v));
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v));
}

// kSig (xyz-only network, layer 7): also accumulates this thread's part of the sigma head, the dot product of the 16-bit
// activations just formed with the 16-bit sigma kernel (wsig_u32 = shared-memory address of this group's 32 weights).
template <bool kMask, bool kHalf, bool kSig = false>
__device__ __forceinline__ uint32_t epi_group32(const uint32_t (&acc)[32], float alpha, uint32_t pbx, int chunk_base,
                                                uint32_t wsig_u32 = 0u, float* sig = nullptr) {
  uint32_t u[8];
  const uint64_t alpha2 = pack_f32x2(alpha, alpha);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint32_t pk[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint64_t x = pack_f32x2(__uint_as_float(acc[8 * j + 2 * i]), __uint_as_float(acc[8 * j + 2 * i + 1]));
      const uint64_t lo = mul_f32x2(x, alpha2);
      float x0, x1, l0, l1;
      unpack_f32x2(x, x0, x1);
      unpack_f32x2(lo, l0, l1);
      pk[i] = pack_16x2<kHalf>(fmaxf(x0, l0), fmaxf(x1, l1)); // LeakyReLU for 0 <= alpha <= 1
      if (kSig) {
        const float2 a = unpack_16x2<kHalf>(pk[i]), w = unpack_16x2<kHalf>(lds32u(wsig_u32 + (uint32_t)(8 * j + 2 * i) * 2u));
        *sig = fmaf(a.x, w.x, fmaf(a.y, w.y, *sig));
      }
    }
    if (kMask) { u[2 * j] = hi_bytes(pk[0], pk[1]); u[2 * j + 1] = hi_bytes(pk[2], pk[3]); }
    sts128(pbx ^ (uint32_t)((chunk_base + j) << 4), make_uint4(pk[0], pk[1], pk[2], pk[3]));
  }
  return kMask ? neg_mask32(u) : 0u;
}

// sin / cos (pi t) for the fused prologue: exact range reduction r = t - 2 rint(t / 2) in [-1, 1] (exact in fp32), then
// the SFU on pi r in [-pi, pi] (absolute error ~1e-6, far below the 2^-9 / 2^-12 rounding of the bf16 / fp16 operand it
// becomes).  ~6 instructions instead of the ~50 of sincospif: the prologue sits on the tile boundary, where the tensor
// pipe has nothing else to do.
__device__ __forceinline__ void fast_sincospi(float t, float* sn, float* cs) {
  const float q = rintf(0.5f * t);
  const float a = fmaf(-2.f, q, t) * 3.14159265358979f;
  *sn = __sinf(a);
  *cs = __cosf(a);
}

template <bool kHalf>
__device__ __forceinline__ uint4 pack8_16(const float* v) {
  return make_uint4(pack_16x2<kHalf>(v[0], v[1]), pack_16x2<kHalf>(v[2], v[3]), pack_16x2<kHalf>(v[4], v[5]),
                    pack_16x2<kHalf>(v[6], v[7]));
}

// eight packed fp16 values -> eight packed bf16 values (round to nearest even; fp16's range fits bf16's)
__device__ __forceinline__ uint32_t half2_to_bf16x2(uint32_t h2) {
  const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&h2));
  return pack_bf16x2(f.x, f.y);
}
__device__ __forceinline__ uint4 half8_to_bf16(uint4 v) {
  return make_uint4(half2_to_bf16x2(v.x), half2_to_bf16x2(v.y), half2_to_bf16x2(v.z), half2_to_bf16x2(v.w));
}

// kXyz: the xyz-only network (src/NeRF.py:248-288): ten layers, no view columns, sigma head = a dot product with h8 in the
// epilogue of layer 7 (partial sums ride in a register to the last layer).
template <bool kSave, bool kHalf, bool kXyz>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(fwd_threads(kSave, kHalf), 1)
mlp_tc_fwd_kernel(const __grid_constant__ TcPlan plan, const uint8_t* __restrict__ packed,
                  const __grid_constant__ FwdInput in, int64_t M, float* __restrict__ out4, uint8_t* __restrict__ saved,
                  float alpha, uint32_t dbg) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  if ((sbase & 1023u) != 0u) __trap();               // swizzled panels need the 1024-byte alignment
  FwdBars* bars = reinterpret_cast<FwdBars*>(smem + kSmemBar);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // a "quad" = the four 128-row tiles a CTA pair works on at a time: tile = 4 quad + 2 t + rank (t = super-tile 0/1)
  const uint32_t rank = cluster_ctarank();
  const int64_t n_tiles = (M + kTileM - 1) / kTileM;
  const int64_t n_quads = (n_tiles + 3) / 4;
  const int64_t quad0 = cluster_id_x(), quad_step = num_clusters_x();

  if (threadIdx.x == 0) {
    // full[s] of the LEADER also counts the peer's relay arrive: both halves of the chunk have landed
    for (int s = 0; s < kStages; ++s) { mbar_init(smem_u32(&bars->full[s]), (rank == 0 && !(dbg & kDbgNoRelay)) ? 2 : 1); mbar_init(smem_u32(&bars->empty[s]), 1); }
    for (int t = 0; t < 2; ++t) {
      mbar_init(smem_u32(&bars->act_ready[t]), 2 * 8);   // leader's copy: one arrive per epilogue warp of BOTH CTAs
      mbar_init(smem_u32(&bars->acc_full[t]), 1);
      mbar_init(smem_u32(&bars->panel_full[t]), 8);      // one arrive per epilogue warp of the tile
      mbar_init(smem_u32(&bars->panel_free[t]), fwd_store_warps(kSave, kHalf));   // every store warp
    }
    fence_barrier_init();
  }
  cluster_sync_all();                                  // barriers of both CTAs exist before any remote arrive
  if (warp == kWarpMma) tmem_alloc_pair(smem_u32(&bars->tmem_base), 512);
  // L1 is ~3 KB next to 225 KB of shared memory: the fp32 rgb head lives in shared memory
  for (int i = threadIdx.x; i < 129; i += blockDim.x)
    reinterpret_cast<float4*>(smem + kSmemWrgb)[i] = __ldg(reinterpret_cast<const float4*>(packed + plan.w_rgb_off) + i);
  if (kXyz)
    for (int i = threadIdx.x; i < (256 * 2 + 16) / 4; i += blockDim.x)
      reinterpret_cast<uint32_t*>(smem + kSmemWsig)[i] = __ldg(reinterpret_cast<const uint32_t*>(packed + plan.w_sig_off) + i);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == kWarpProducer) {
    // ===== producer: weight chunks (ring) =====
    // kProdLanes lanes issue the copies of consecutive ring stages side by side.  Measured (tools/l2_bw_probe.cu,
    // profiles/r02_g_l2_bw_probe.log): ONE thread gets ~3.5 M bulk copies per second out of the TMA path whatever their
    // size (a 16 KB copy every ~550 cycles -- the four MMAs it feeds take 512), two lanes twice that.  With a single
    // producer lane the weight stream, not the tensor pipe or shared memory, set the pace of this kernel.
    if (lane < kProdLanes) {
      // Every producer lane walks the SAME static sequence (the chunk order of make_plan: pointer adds only -- no table
      // look-ups or divisions on the path from "stage released" to "stage full") and issues the items whose index is its
      // own modulo kProdLanes.
      const uint32_t full0 = smem_u32(&bars->full[0]), empty0 = smem_u32(&bars->empty[0]);
      const bool no_copy = (dbg & kDbgNoWeightCopy) != 0;
      uint32_t s = 0, ph = 1, k = 0;                     // ring stage, parity of the "stage is free" phase, item index
      // this CTA's half of a chunk: rows [rank N/2, (rank + 1) N/2) of [N][64] (or of the [256][16] bias slab)
      auto load = [&](const uint8_t* src, uint32_t half_bytes) {
        if ((k & (uint32_t)(kProdLanes - 1)) == (uint32_t)lane) {
          mbar_wait_spin(empty0 + 8u * s, ph);
          if (no_copy) {
            mbar_arrive(full0 + 8u * s);
          } else {
            mbar_arrive_expect_tx(full0 + 8u * s, half_bytes);
            bulk_g2s(sbase + kSmemStage + s * kStageBytes, src + rank * half_bytes, half_bytes, full0 + 8u * s);
          }
        }
        ++k;
        if (++s == kStages) { s = 0; ph ^= 1u; }
      };
      for (int64_t quad = quad0; quad < n_quads; quad += quad_step) {
        // staggered schedule: super-tile A runs layer l while super-tile B's previous accumulator is drained, so each
        // layer's chunks are streamed once per super-tile (they come from L2)
        const uint8_t* layer_base = packed;
#pragma unroll 1
        for (int l = 0; l < tc_n_layers(kXyz); ++l) {
          const uint32_t chunk_bytes = tc_layer_rows(kXyz, l) * 128u;
          const bool slab = tc_layer_slab(kXyz, l);
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            const uint8_t* p = layer_base;
            if (l > 0) {
#pragma unroll
              for (int ci = 0; ci < 4; ++ci, p += chunk_bytes) load(p, chunk_bytes >> 1);
            }
            // bias slab [256][16]: 32 bytes per output row, this CTA's N/2 rows (N = 128 in the xyz-only network's last layer)
            load(p, slab ? (chunk_bytes >> 3) : (chunk_bytes >> 1));
          }
          layer_base += (l == 0 ? chunk_bytes : 4u * chunk_bytes + (slab ? (uint32_t)kBiasSlabBytes : chunk_bytes));
        }
      }
    }
  } else if (warp == kWarpMma) {
    // ===== MMA issuer (leader CTA) / weight-arrival relay (peer CTA) =====
    if (lane == 0 && rank != 0 && (dbg & kDbgNoRelay)) {
    } else if (lane == 0 && rank != 0) {
      // peer CTA: forward "my half of the chunk has landed" to the leader's full barrier
      const uint32_t full0 = smem_u32(&bars->full[0]), full0_leader = mapa_shared(full0, 0);
      uint32_t s = 0, ph = 0;
      for (int64_t quad = quad0; quad < n_quads; quad += quad_step)
        for (int i = 0; i < 2 * plan.n_chunks; ++i) {
          mbar_wait_spin(full0 + 8u * s, ph);
          mbar_arrive_cluster(full0_leader + 8u * s);
          if (++s == kStages) { s = 0; ph ^= 1u; }
        }
    } else if (lane == 0) {
      // The issuing thread is latency-bound, not the tensor pipe: a generic loop (plan look-ups, descriptor rebuilds,
      // modulo ring indices) cost ~750 cycles per 4-MMA chunk (tools/sm_probe.cu "p"), 1.5x the 512 cycles the MMAs take.
      // So the schedule is spelled out statically (the chunk order of make_plan) and every descriptor is a register add.
      const bool timing = kTcTrace && (dbg & kDbgTiming) && blockIdx.x == 0;
      long long t_begin = clock64(), t_act = 0, t_full = 0, t_issue = 0;
      int n_tr = 0;
      const uint32_t fmt = kHalf ? 0 : 1;
      const uint32_t idesc256 = make_idesc(256, 0, 0, fmt, 256);
      const uint32_t idesc_last = make_idesc((int)tc_layer_rows(kXyz, tc_n_layers(kXyz) - 1), 0, 0, fmt, 256);
      const uint64_t b_sw = make_desc_kmajor(sbase + kSmemStage), b_slab = make_desc_k_nosw(sbase + kSmemStage, 128, 256);
      const uint32_t full0 = smem_u32(&bars->full[0]), empty0 = smem_u32(&bars->empty[0]);
      const bool no_mma = (dbg & kDbgNoMma) != 0;
      uint32_t s = 0, ph = 0, act_ph = 0;               // ring stage / phase, act_ready phase
      // one weight chunk: wait for both halves, 4 MMAs of K = 16 (or the single bias-slab MMA), release the stage
      auto chunk = [&](uint32_t d_tmem, uint64_t a_desc, bool slab, uint32_t idesc, uint32_t first_acc) {
        long long tw = timing ? clock64() : 0;
        mbar_wait_spin(full0 + 8u * s, ph);
        if (timing) { const long long now = clock64(); t_full += now - tw; tw = now; }
        tc_fence_after();
        const uint64_t soff = (uint64_t)(s * (uint32_t)(kStageBytes >> 4));
        if (!no_mma) {
          if (slab) {
            umma_pair(d_tmem, a_desc + 4, b_slab + soff, idesc, 1u);   // input-panel columns 32..47
          } else {
            const uint64_t b = b_sw + soff;
            umma_pair(d_tmem, a_desc, b, idesc, first_acc);
            umma_pair(d_tmem, a_desc + 2, b + 2, idesc, 1u);
            umma_pair(d_tmem, a_desc + 4, b + 4, idesc, 1u);
            umma_pair(d_tmem, a_desc + 6, b + 6, idesc, 1u);
          }
        }
        umma_commit_pair(empty0 + 8u * s);             // releases this stage in both CTAs
        if (timing) t_issue += clock64() - tw;
        if (++s == kStages) { s = 0; ph ^= 1u; }
      };
      for (int64_t quad = quad0; quad < n_quads; quad += quad_step) {
#pragma unroll 1
        for (int l = 0; l < tc_n_layers(kXyz); ++l) {
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            long long tw = timing ? clock64() : 0;
            mbar_wait_spin(smem_u32(&bars->act_ready[t]), act_ph);
            if (timing) t_act += clock64() - tw;
            trace(timing, 2, n_tr, l * 4 + t * 2);
            const uint32_t d_tmem = tmem_base + (uint32_t)t * 256u;
            const uint64_t a_act = make_desc_kmajor(sbase + kSmemAct + t * kActPanels * kPanelBytes);
            const uint64_t a_inp = make_desc_kmajor(sbase + kSmemInp + t * kPanelBytes);
            if (l == 0) {
              chunk(d_tmem, a_inp, false, idesc256, 0u);
            } else {
              const uint32_t idesc = l == tc_n_layers(kXyz) - 1 ? idesc_last : idesc256;
              chunk(d_tmem, a_act, false, idesc, 0u);
              chunk(d_tmem, a_act + 1 * (kPanelBytes >> 4), false, idesc, 1u);
              chunk(d_tmem, a_act + 2 * (kPanelBytes >> 4), false, idesc, 1u);
              chunk(d_tmem, a_act + 3 * (kPanelBytes >> 4), false, idesc, 1u);
              chunk(d_tmem, a_inp, tc_layer_slab(kXyz, l), idesc, 1u);   // input-panel chunk (skip / view layer) or bias slab
            }
            umma_commit_pair(smem_u32(&bars->acc_full[t]));
            trace(timing, 2, n_tr, l * 4 + t * 2 + 1);
          }
          act_ph ^= 1u;
        }
      }
      if (timing) {
        printf("fwd MMA thread: total %lld  wait act_ready %lld  wait full %lld  issue %lld (cycles)\n", clock64() - t_begin, t_act,
               t_full, t_issue);
        for (int i = 0; i < n_tr; ++i)
          printf("TRACE mma %lld l%d t%d %s\n", (g_trace[2][i] >> 8) - t_begin, (int)(g_trace[2][i] & 255) >> 2,
                 (int)(g_trace[2][i] & 2) >> 1, (g_trace[2][i] & 1) ? "issued" : "act_ready");
      }
    }
  } else if (warp >= kWarpStore) {
    // ===== store warps (train mode): tile t's freshly written activation panels -> HBM, row-block chunk-major =====
    // Global stores stall the issuing warp whenever the 32 B/clk SM->L2 port is busy (64 KB per tile-layer = 2048
    // cycles, as long as the MMAs of that tile-layer).  Issued from the epilogue warps they delayed the next epilogue;
    // a dedicated warp per tile absorbs the stalls while the epilogue warps go back to waiting for accumulators.
    // Both store warps work on EVERY copy (one half of the column chunks each): the two tiles' copies alternate in
    // time, and one warp alone does not saturate the port.
    if (kSave) {
      const int hw = warp - kWarpStore;                  // which half of the 32 column chunks
      const bool do_store = !(dbg & kDbgNoStore);
      uint32_t ph = 0;
      for (int64_t quad = quad0; quad < n_quads; quad += quad_step) {
        for (int l = 0; l < tc_n_layers(kXyz) - 1; ++l, ph ^= 1u) {
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            const uint32_t act_u32 = sbase + kSmemAct + t * kActPanels * kPanelBytes;
            const int64_t tile = quad * 4 + t * 2 + rank;
            const int64_t tile_dst = (dbg & kDbgStoreL2) ? (tile & 63) : tile;
            uint8_t* gblock = saved + (size_t)tile_dst * kSavedTileBytes + (size_t)saved_panel_h(l + 1) * kPanelBytes;
            mbar_wait(smem_u32(&bars->panel_full[t]), ph);
            if (do_store) {
#pragma unroll 8
              constexpr int kSW = fwd_store_warps(kSave, kHalf);
              for (int it = 0; it < 128 / kSW; ++it) {
                const int j = hw * (32 / kSW) + (it >> 2), r = (it & 3) * 32 + lane;     // 16-byte column chunk, row
                const float4 v = lds128f(act_u32 + (j >> 3) * kPanelBytes + r * 128 + (((j & 7) ^ (r & 7)) << 4));
                uint4 w = make_uint4(__float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w));
                if (kHalf) w = half8_to_bf16(w);          // the backward reads bf16
                stg128(gblock + rbcm_offset(r, j, 32), w);
              }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bars->panel_free[t]));
          }
        }
      }
    }
  } else {
    // ===== 16 epilogue warps: tile t = warp / 8, TMEM lane quarter q = warp % 4, column half = (warp / 4) % 2 =====
    const int t = warp >> 3, q = warp & 3, half = (warp >> 2) & 1;
    const int r = q * 32 + lane;                       // row inside the tile == TMEM lane
    const int gtid = threadIdx.x & (kEpiThreadsPerTile - 1);
    const int bar_id = 1 + t;
    const uint32_t act_u32 = sbase + kSmemAct + t * kActPanels * kPanelBytes;
    const uint32_t inp_u32 = sbase + kSmemInp + t * kPanelBytes;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)t * 256u;
    const uint32_t wrgb_u32 = sbase + kSmemWrgb;
    uint32_t acc_cnt = 0, copy_ph = 0;
    bool copy_pending = false;
    const bool timing = kTcTrace && (dbg & kDbgTiming) && blockIdx.x == 0 && gtid == 0;
    long long t_begin = clock64(), t_pro = 0, t_acc = 0, t_epi = 0, t_last = 0;
    int n_tr = 0;
    const uint32_t act_ready_leader = mapa_shared(smem_u32(&bars->act_ready[t]), 0);
    for (int64_t quad = quad0; quad < n_quads; quad += quad_step) {
      const int64_t tile = quad * 4 + t * 2 + rank;
      const int64_t row = tile * kTileM + r;
      const bool row_ok = row < M;
      long long tw = timing ? clock64() : 0;
      // ---- prologue: this thread's half of the input panel row: half 0 = xyz columns (16-byte chunks 0..4),
      //      half 1 = view columns (chunks 5..7).  Zero first, then scatter the encoded values as bf16.
      if (in.xyz_enc == nullptr && in.Lx == 5 && in.Lv == 4) {
        // fast path of the reference's network (46 of its 47 configs): the row is assembled in registers with compile-time
        // column indices and leaves as 16-byte chunks
        const uint32_t prow = inp_u32 + r * 128;
        const int64_t ray = row_ok ? row / in.n_samples : 0;
        const float4 d = row_ok ? ray_dir(in, ray) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (half == 0) {
          const float4 o = row_ok ? ray_orig(in, ray) : make_float4(0.f, 0.f, 0.f, 0.f);
          const float zz = row_ok ? row_depth(in, row, ray) : 0.f;
          float v[40];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float oc = c == 0 ? o.x : (c == 1 ? o.y : o.z), dc = c == 0 ? d.x : (c == 1 ? d.y : d.z);
            const float pc = __fadd_rn(oc, __fmul_rn(dc, zz));         // sample_along_rays, src/UtilsCV.py:598
            v[c * 11] = pc;
            float t = pc;
#pragma unroll
            for (int k = 0; k < 5; ++k, t *= 2.f) fast_sincospi(t, &v[c * 11 + 1 + 2 * k], &v[c * 11 + 2 + 2 * k]);
          }
#pragma unroll
          for (int i = 33; i < 40; ++i) v[i] = (i >= kInpOneCol) ? 1.f : 0.f;   // the constant-1 columns (bias rows / slabs)
#pragma unroll
          for (int j = 0; j < 5; ++j) sts128(prow + ((j ^ (r & 7)) << 4), pack8_16<kHalf>(v + 8 * j));
        } else {
          float v[24];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float vc = (in.ncomp == 3) ? (c == 0 ? d.x : (c == 1 ? d.y : d.z)) : (c == 0 ? d.x : d.z);
            float t = vc;
#pragma unroll
            for (int k = 0; k < 4; ++k, t *= 2.f) {
              float sn, cs;
              fast_sincospi(t, &sn, &cs);
              v[c * 8 + 2 * k] = (c < in.ncomp && row_ok) ? sn : 0.f;
              v[c * 8 + 2 * k + 1] = (c < in.ncomp && row_ok) ? cs : 0.f;
            }
          }
#pragma unroll
          for (int j = 0; j < 3; ++j) sts128(prow + (((5 + j) ^ (r & 7)) << 4), pack8_16<kHalf>(v + 8 * j));
        }
      } else {
        const uint32_t prow = inp_u32 + r * 128;
        const int cb = half ? 5 : 0, ce = half ? 8 : 5;
        for (int j = cb; j < ce; ++j) sts128(prow + ((j ^ (r & 7)) << 4), make_uint4(0u, 0u, 0u, 0u));
        if (half == 0) {                                  // the constant-1 columns the bias rows / slabs multiply
          sts_16<kHalf>(prow, r, kInpOneCol, 1.f);
          sts_16<kHalf>(prow, r, kInpOneCol + 1, 1.f);
        }
        if (row_ok) {
          if (in.xyz_enc != nullptr) {
            if (half == 0) {
              const float* xr = in.xyz_enc + row * in.dx;
              for (int i = 0; i < in.dx; ++i) sts_16<kHalf>(prow, r, i, __ldg(xr + i));
            } else {
              const float* vr = in.view_enc + row * in.dv;
              for (int i = 0; i < in.dv; ++i) sts_16<kHalf>(prow, r, kInpViewCol + i, __ldg(vr + i));
            }
          } else {
            const int64_t ray = row / in.n_samples;
            const float4 d = ray_dir(in, ray);
            if (half == 0) {
              const float4 o = ray_orig(in, ray);
              const float zz = row_depth(in, row, ray);
              const int per = 1 + 2 * in.Lx;
#pragma unroll
              for (int c = 0; c < 3; ++c) {
                const float oc = c == 0 ? o.x : (c == 1 ? o.y : o.z), dc = c == 0 ? d.x : (c == 1 ? d.y : d.z);
                const float pc = __fadd_rn(oc, __fmul_rn(dc, zz));       // sample_along_rays, src/UtilsCV.py:598
                sts_16<kHalf>(prow, r, c * per, pc);
                for (int k = 0; k < in.Lx; ++k) {
                  float sn, cs;
                  sincospif(ldexpf(pc, k), &sn, &cs);                      // sin/cos(2^k pi p)
                  sts_16<kHalf>(prow, r, c * per + 1 + 2 * k, sn);
                  sts_16<kHalf>(prow, r, c * per + 2 + 2 * k, cs);
                }
              }
            } else {
              for (int c = 0; c < in.ncomp; ++c) {
                const float vc = (in.ncomp == 3) ? (c == 0 ? d.x : (c == 1 ? d.y : d.z)) : (c == 0 ? d.x : d.z);
                for (int k = 0; k < in.Lv; ++k) {
                  float sn, cs;
                  sincospif(ldexpf(vc, k), &sn, &cs);
                  sts_16<kHalf>(prow, r, kInpViewCol + c * 2 * in.Lv + 2 * k, sn);
                  sts_16<kHalf>(prow, r, kInpViewCol + c * 2 * in.Lv + 2 * k + 1, cs);
                }
              }
            }
          }
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(act_ready_leader);
      if (timing) t_pro += clock64() - tw;
      uint8_t* saved_tile = kSave ? saved + (size_t)tile * kSavedTileBytes : nullptr;
      uint32_t* saved_mask = reinterpret_cast<uint32_t*>(saved_tile + (size_t)kSavedPanels * kPanelBytes);
      const bool do_store = kSave && !(dbg & kDbgNoStore);
      if (kSave) {
        named_bar_sync(bar_id, kEpiThreadsPerTile);
        if (kHalf) {
          // fp16 panel -> bf16 saved copy (same swizzled layout): this thread's half row, four 16-byte chunks
          if (do_store) {
#pragma unroll
            for (int j = half * 4; j < half * 4 + 4; ++j) {
              const uint32_t off = (uint32_t)r * 128u + (uint32_t)((j ^ (r & 7)) << 4);
              const float4 v = lds128f(inp_u32 + off);
              stg128(saved_tile + off, half8_to_bf16(make_uint4(__float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z),
                                                                __float_as_uint(v.w))));
            }
          }
        } else if (gtid == 0 && do_store) {
          bulk_s2g(saved_tile, inp_u32, kPanelBytes);
          bulk_commit();
        }
      }

      float sig_part = 0.f;                              // xyz-only: this thread's half of the sigma head's dot product
      for (int l = 0; l < tc_n_layers(kXyz); ++l) {
        tw = timing ? clock64() : 0;
        mbar_wait(smem_u32(&bars->acc_full[t]), acc_cnt & 1u);
        ++acc_cnt;
        tc_fence_after();
        if (timing) { const long long now = clock64(); t_acc += now - tw; tw = now; }
        trace(timing, t, n_tr, l * 2);
        if (kSave && copy_pending) {
          // the store warp must have read the panels this epilogue overwrites
          mbar_wait(smem_u32(&bars->panel_free[t]), copy_ph);
          copy_ph ^= 1u;
          copy_pending = false;
        }
        if (l < tc_n_layers(kXyz) - 1) {
          // TMEM loads are double-buffered: group cc+1 is in flight while group cc is processed
          uint32_t acc[kTmemBuffers][32];
          uint32_t mw[4] = {0u, 0u, 0u, 0u};
          if (kTmemBuffers == 2 && !(dbg & kDbgNoEpi)) tmem_ld32(taddr + half * 128, acc[0]);
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            if (dbg & kDbgNoEpi) break;
            const int c0 = half * 128 + cc * 32;
            if (kTmemBuffers == 1) tmem_ld32(taddr + c0, acc[0]);
            tmem_ld_wait();
            if (kTmemBuffers == 2 && cc + 1 < 4) tmem_ld32(taddr + c0 + 32, acc[(cc + 1) & 1]);
            const uint32_t pbx = (act_u32 + (c0 >> 6) * kPanelBytes + r * 128) | ((uint32_t)(r & 7) << 4);
            if (kXyz && l == 7)
              mw[cc] = epi_group32<kSave, kHalf, true>(acc[cc & (kTmemBuffers - 1)], alpha, pbx, (c0 & 63) >> 3,
                                                       sbase + kSmemWsig + (uint32_t)c0 * 2u, &sig_part);
            else
              mw[cc] = epi_group32<kSave, kHalf>(acc[cc & (kTmemBuffers - 1)], alpha, pbx, (c0 & 63) >> 3);
          }
          tc_fence_before();
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(act_ready_leader);
          if (timing) t_epi += clock64() - tw;
          trace(timing, t, n_tr, l * 2 + 1);
          if (kSave) {
            // hand the panels to the tile's store warp; the masks (4 words) leave from here
            if (lane == 0) mbar_arrive(smem_u32(&bars->panel_full[t]));
            copy_pending = true;
            if (do_store && !(dbg & kDbgNoEpi)) {
#pragma unroll
              for (int cc = 0; cc < 4; ++cc) saved_mask[(l * 8 + half * 4 + cc) * 128 + r] = mw[cc];
            }
          }
        } else {
          // last layer: cols 0..127 = last hidden (LeakyReLU), col 128 = sigma (linear); rgb head on CUDA cores.
          // half 0 owns cols 0..63, half 1 owns cols 64..127 and sigma; partial rgb sums meet in the (dead) input panel.
          float rr = 0.f, gg = 0.f, bb = 0.f;
          uint8_t* grow = do_store ? saved_tile + (size_t)kSavedPanelHL * kPanelBytes + rbcm_offset(r, 0, 16) : nullptr;
#pragma unroll 1
          for (int cc = 0; cc < 2; ++cc) {
            const int c0 = half * 64 + cc * 32;
            uint32_t acc[32];
            tmem_ld32(taddr + c0, acc);
            tmem_ld_wait();
            uint32_t u[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint32_t pk[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int e = 8 * j + 2 * i;
                float x0 = __uint_as_float(acc[e]), x1 = __uint_as_float(acc[e + 1]);     // bias rode in the MMA
                x0 = fmaxf(x0, alpha * x0);
                x1 = fmaxf(x1, alpha * x1);
                const float4 w0 = lds128f(wrgb_u32 + (c0 + e) * 16), w1 = lds128f(wrgb_u32 + (c0 + e + 1) * 16);
                rr = fmaf(x0, w0.x, fmaf(x1, w1.x, rr));
                gg = fmaf(x0, w0.y, fmaf(x1, w1.y, gg));
                bb = fmaf(x0, w0.z, fmaf(x1, w1.z, bb));
                pk[i] = pack_16x2<false>(x0, x1);          // only leaves for the backward (bf16) and the sign mask
              }
              if (kSave) { u[2 * j] = hi_bytes(pk[0], pk[1]); u[2 * j + 1] = hi_bytes(pk[2], pk[3]); }
              if (kSave && grow) stg128(grow + ((c0 >> 3) + j) * 1024, make_uint4(pk[0], pk[1], pk[2], pk[3]));
            }
            if (do_store) saved_mask[(kMaskRowHL * 8 + (c0 >> 5)) * 128 + r] = neg_mask32(u);
          }
          const uint32_t xch = act_u32 + r * 16;                  // activation panels are dead after this layer's MMAs
          if (half == 1) {
            float sigma = sig_part;                          // xyz-only: this half of the dot product with h8
            if (!kXyz) {
              uint32_t sg[16];
              tmem_ld16(taddr + 128, sg);
              tmem_ld_wait();
              sigma = __uint_as_float(sg[0]);
            }
            sts128(xch, make_uint4(__float_as_uint(rr), __float_as_uint(gg), __float_as_uint(bb), __float_as_uint(sigma)));
          }
          tc_fence_before();
          named_bar_sync(bar_id, kEpiThreadsPerTile);
          if (half == 0 && row_ok) {
            const float4 o = lds128f(xch), br = lds128f(wrgb_u32 + 128 * 16);
            const float sigma = kXyz ? o.w + sig_part + lds32f(sbase + kSmemWsig + 512) : o.w;
            reinterpret_cast<float4*>(out4)[row] = make_float4(rr + o.x + br.x, gg + o.y + br.y, bb + o.z + br.z, sigma);
          }
          // the next quad's first epilogue overwrites the exchange rows, its prologue the input panel (whose saved copy,
          // a bulk S2G issued after this pair's prologue, must have been read out by now)
          if (kSave && gtid == 0) bulk_wait_read0();
          named_bar_sync(bar_id, kEpiThreadsPerTile);
          if (timing) t_last += clock64() - tw;
          trace(timing, t, n_tr, l * 2 + 1);
        }
      }
    }
    if (timing)
      for (int i = 0; i < n_tr; ++i)
        printf("TRACE epi%d %lld l%d %s\n", t, (g_trace[t][i] >> 8) - t_begin, (int)(g_trace[t][i] & 255) >> 1,
               (g_trace[t][i] & 1) ? "done" : "acc_full");
    if (timing)
      printf("fwd epilogue tile %d: total %lld  prologue %lld  wait acc_full %lld  hidden epilogues %lld  last layer %lld (cycles)\n", t,
             clock64() - t_begin, t_pro, t_acc, t_epi, t_last);
    if (kSave && gtid == 0) bulk_wait0();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                  // the peer's shared memory / TMEM stay valid until both are done
  if (warp == kWarpMma) tmem_dealloc_pair(tmem_base, 512);
}

// ---- host side ----------------------------------------------------------------------------------------------------------------
uint32_t tc_debug_flags() {
  static const uint32_t flags = [] {
    const char* e = getenv("NERF_TC_DEBUG");
    return e ? (uint32_t)strtoul(e, nullptr, 0) : 0u;
  }();
  return flags;
}

int64_t mlp_tc_saved_bytes(const NetGeom& g, int64_t m) {
  int64_t tiles = (m + kTileM - 1) / kTileM;
  tiles = (tiles + 3) / 4 * 4;
  return tiles * (int64_t)kSavedTileBytes;
}

int64_t mlp_tc_workspace_bytes(const NetGeom& g, int64_t m, int backward) {
  (void)g;
  if (!backward) return 256;
  int64_t tiles = (m + kTileM - 1) / kTileM;
  tiles = (tiles + 3) / 4 * 4;
  return tiles * (int64_t)kDzTileBytes + 1024 + kDwScratchBytes + mlp_tc_bwd_flag_bytes(m);
}

// byte offset of the fp16 copy of the forward weight pack inside the packed buffer (layout: mlp_tc.cuh)
static uint32_t half_region_offset(const TcPlan& plan) { return pack_off_fwd_half(plan); }

static int launch_fwd(const nerf_net_cfg* cfg, const NetGeom& g, const void* packed, const FwdInput& in, int64_t m,
                      float* out4, void* saved, cudaStream_t st, bool half = false) {
  TcPlan plan;
  if (!make_plan(g, &plan) || cfg->leaky_alpha < 0.f || cfg->leaky_alpha > 1.f) {
    set_error("the tensor-core modes support hidden=256, last_hidden=128, xyz width <= 38, view width <= 24, "
              "0 <= leaky_relu_alpha <= 1");
    return NERF_E_UNSUPPORTED;
  }
  if (device_first_use(0)) {               // per device: the attribute lives in the device's copy of the function
    NERF_CUDA(cudaFuncSetAttribute(mlp_tc_fwd_kernel<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemAlloc));
    NERF_CUDA(cudaFuncSetAttribute(mlp_tc_fwd_kernel<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemAlloc));
    NERF_CUDA(cudaFuncSetAttribute(mlp_tc_fwd_kernel<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemAlloc));
    NERF_CUDA(cudaFuncSetAttribute(mlp_tc_fwd_kernel<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemAlloc));
    NERF_CUDA(cudaFuncSetAttribute(mlp_tc_fwd_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemAlloc));
    NERF_CUDA(cudaFuncSetAttribute(mlp_tc_fwd_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemAlloc));
    NERF_CUDA(cudaFuncSetAttribute(mlp_tc_fwd_kernel<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemAlloc));
    NERF_CUDA(cudaFuncSetAttribute(mlp_tc_fwd_kernel<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemAlloc));
  }
  int64_t n_quads = ((m + kTileM - 1) / kTileM + 3) / 4;
  const int n_pairs = num_sms() / 2;
  int grid = 2 * (int)(n_quads < n_pairs ? n_quads : n_pairs);             // CTA pairs
  // fp16 operands in training too; what is SAVED for the backward is converted to bf16 on its way out (store warps,
  // last-layer epilogue, input panel), because the backward's MMAs pair it with bf16 gradients
  const uint8_t* pk = (const uint8_t*)packed + (half ? half_region_offset(plan) : 0u);
  const uint32_t dbg = tc_debug_flags();
#define NERF_LAUNCH_FWD(SAVE, HALF, XYZ)                                                                                  \
  mlp_tc_fwd_kernel<SAVE, HALF, XYZ><<<grid, fwd_threads(SAVE, HALF), kSmemAlloc, st>>>(plan, pk, in, m, out4, (uint8_t*)saved,       \
                                                                            cfg->leaky_alpha, dbg)
  const int variant = (saved ? 1 : 0) | (half ? 2 : 0) | (plan.xyz_only ? 4 : 0);
  switch (variant) {
    case 0: NERF_LAUNCH_FWD(false, false, false); break;
    case 1: NERF_LAUNCH_FWD(true, false, false); break;
    case 2: NERF_LAUNCH_FWD(false, true, false); break;
    case 3: NERF_LAUNCH_FWD(true, true, false); break;
    case 4: NERF_LAUNCH_FWD(false, false, true); break;
    case 5: NERF_LAUNCH_FWD(true, false, true); break;
    case 6: NERF_LAUNCH_FWD(false, true, true); break;
    default: NERF_LAUNCH_FWD(true, true, true); break;
  }
#undef NERF_LAUNCH_FWD
  NERF_CHECK_LAUNCH();
  return NERF_OK;
}

int mlp_tc_fwd(const nerf_net_cfg* cfg, const NetGeom& g, const float* params, const void* packed, const float* xyz_enc,
               const float* view_enc, int64_t m, float* out4, void* saved, void* workspace, cudaStream_t st, bool half) {
  (void)params; (void)workspace;
  FwdInput in = {};
  in.xyz_enc = xyz_enc; in.view_enc = view_enc; in.dx = g.dx; in.dv = g.dv;
  in.Lx = cfg->n_pos_enc_xyz; in.Lv = cfg->n_pos_enc_view; in.ncomp = g.view ? cfg->n_angles + 1 : 0; in.n_samples = 1;
  return launch_fwd(cfg, g, packed, in, m, out4, saved, st, half);
}

int mlp_tc_fwd_rays(const nerf_net_cfg* cfg, const NetGeom& g, const void* packed, const float* origs4, const float* dirs4,
                    const float* z, int64_t n_rays, int n_samples, float* out4, void* saved, cudaStream_t st, bool half,
                    const StratifiedZ* gen, const CameraRays* cam) {
  FwdInput in = {};
  in.origs = (const float4*)origs4; in.dirs = (const float4*)dirs4; in.z = z; in.n_samples = n_samples;
  if (cam) { in.cam = 1; in.camera = *cam; }
  if (gen) {
    in.gen_z = 1; in.z_start = gen->z_start; in.z_end = gen->z_end;
    // (z_end - z_start) is formed in double from the Python floats and then cast (src/UtilsCV.py:580), as in nerf_stratified_z
    in.span = (float)((double)gen->z_end - (double)gen->z_start);
    in.seed = gen->seed; in.step = gen->step; in.ray_offset = gen->ray_offset; in.z_out = gen->z_out;
  }
  in.dx = g.dx; in.dv = g.dv; in.Lx = cfg->n_pos_enc_xyz; in.Lv = cfg->n_pos_enc_view; in.ncomp = g.view ? cfg->n_angles + 1 : 0;
  return launch_fwd(cfg, g, packed, in, n_rays * n_samples, out4, saved, st, half);
}

}  // namespace nerf

using namespace nerf;

extern "C" {

int64_t nerf_packed_bytes(const nerf_net_cfg* cfg) {
  NetGeom g;
  TcPlan plan;
  if (!make_geom(cfg, &g)) { set_error("nerf_packed_bytes: bad net config"); return NERF_E_ARG; }
  if (!make_plan(g, &plan)) { set_error("nerf_packed_bytes: config not supported by NERF_MODE_BF16"); return NERF_E_UNSUPPORTED; }
  return (int64_t)pack_total_bytes(plan);
}

int nerf_pack_weights(const nerf_net_cfg* cfg, const float* params, void* packed, void* stream) {
  NetGeom g;
  TcPlan plan;
  NERF_CHECK_ARG(make_geom(cfg, &g), "bad net config");
  NERF_CHECK_ARG(params && packed, "null pointer");
  if (!make_plan(g, &plan)) { set_error("nerf_pack_weights: config not supported by NERF_MODE_BF16"); return NERF_E_UNSUPPORTED; }
  dim3 grid((256 * 64 + 255) / 256, plan.n_chunks + 1);
  pack_weights_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(plan, g, params, (uint8_t*)packed);
  NERF_CHECK_LAUNCH();
  return bwd_pack_weights(g, params, (uint8_t*)packed + pack_off_bwd(plan), (cudaStream_t)stream);
}

int nerf_pack_weights_fp16(const nerf_net_cfg* cfg, const float* params, void* packed, void* stream) {
  NetGeom g;
  TcPlan plan;
  NERF_CHECK_ARG(make_geom(cfg, &g), "bad net config");
  NERF_CHECK_ARG(params && packed, "null pointer");
  if (!make_plan(g, &plan)) { set_error("nerf_pack_weights_fp16: config not supported by the tensor-core path"); return NERF_E_UNSUPPORTED; }
  dim3 grid((256 * 64 + 255) / 256, plan.n_chunks + 1);
  pack_weights_kernel<__half><<<grid, 256, 0, (cudaStream_t)stream>>>(plan, g, params, (uint8_t*)packed + half_region_offset(plan));
  NERF_CHECK_LAUNCH();
  // the backward of the fp16 mode is the bf16 backward: keep ITS pack current as well
  return bwd_pack_weights(g, params, (uint8_t*)packed + pack_off_bwd(plan), (cudaStream_t)stream);
}

}  // extern "C"
