// K4 in NERF_MODE_BF16: the gradients TF autodiff produces for the NeRF MLP inside NeRF.train_step
// (src/NeRF.py:149-167), as two tcgen05/TMEM kernels per network:
//
//  (1) mlp_tc_bwd_chain_kernel -- dX chain, the mirror image of the forward kernel.  Per 128-row tile the upstream
//      gradient d_out4 is turned into dZ_L (rgb head transposed on CUDA cores, LeakyReLU' from the saved sign masks),
//      then dH_l = dZ_{l+1} W_l^T runs layer by layer on the tensor cores with the fp32 accumulator in TMEM; the
//      epilogue applies the LeakyReLU' mask, rounds to bf16 and writes dZ_l back into the same swizzled panels (next A
//      operand) and, straight from its registers with coalesced 512-byte warp stores, to HBM in the row-block
//      chunk-major layout (mlp_tc.cuh) for the weight-gradient pass.  For the fine network two extra
//      N=64 steps accumulate d(xyz encoding) = dZ_5 W_4[0:33]^T + dZ_1 W_0^T (the reference does not detach the
//      importance samples, so this gradient flows on to the coarse network).
//  (2) mlp_tc_bwd_dw_kernel -- dW_l = A_l^T dZ_{l+1}: the contraction runs over ROWS, so both operands are read as
//      MN-major UMMA operands: a 64-row slab of a saved RBCM block is one contiguous bulk copy and IS the un-swizzled
//      MN-major operand (the saved input panel keeps the 128-byte-swizzled form).  The 148 CTAs are split over (layer, row-range) units;
//      each CTA keeps its whole 256x256 fp32 accumulator in TMEM (512 columns) across its row range, the idle epilogue
//      warps sum the bias gradients from the dZ stages in shared memory, and each CTA writes its partial result to a
//      split-K scratch block; mlp_tc_bwd_dw_reduce_kernel adds the partials of a unit in a FIXED order into the flat
//      gradient vector (deterministic, and 4x cheaper than the 9.7 M fp32 atomics it replaces).  HBM-bound, see DESIGN.md.
//
//  Overlapped mode (a second stream is given): the two kernels run CONCURRENTLY on disjoint SMs -- the chain on 2 P
//  CTA pairs' SMs of `stream`, the dW kernel on the remaining SMs of `side` -- and hand every dZ block over through a
//  per-(tile, block) ready counter in global memory (release by the chain's store warps, acquire by the dW kernel's
//  producer lane) so that the dW kernel reads dZ from L2 right after it was written instead of from HBM a kernel later.
//  The dependency is one-directional (the chain never waits for the dW kernel), so any serialisation of the two
//  launches (a profiler, a busy GPU) is still correct.  dW CTAs of a unit take tiles round-robin (tile = split + i n),
//  i.e. in the order the chain produces them.
#include <type_traits>

#include "mlp_tc.cuh"
#include "mlp_tc_bwd_pipe.cuh"

namespace nerf {

// ---- dZ hand-over flags (overlapped mode) ---------------------------------------------------------------------------
// flags[tile * kFlagsPerTile + b]: b = l - 1 for dZ_l (l = 1..9; two store warps -> target 2), b = 9 for the blocks the
// chain prologue writes (dZ_L', dOut; eight epilogue warps -> target 8)
constexpr int kFlagsPerTile = kHiddenSlots + 1;
constexpr uint32_t kFlagTargetStore = kStoreWarps, kFlagTargetPrologue = 8;
constexpr int kBwdMaxChunks = 48;
constexpr int kBwdMaxSteps = 12;
enum : int { STEP_MASK = 0, STEP_XSTASH = 1, STEP_XFINAL = 2 };

struct BwdPlan {
  uint32_t chunk_off[kBwdMaxChunks];
  uint32_t chunk_bytes[kBwdMaxChunks];
  int8_t step_first[kBwdMaxSteps], step_nch[kBwdMaxSteps], step_kind[kBwdMaxSteps], step_layer[kBwdMaxSteps];
  int16_t step_n[kBwdMaxSteps];
  int32_t n_steps, n_chunks;
  int32_t n_hidden;      // 256-wide hidden layers whose dZ the chain writes: 8 (view network) or 9 (xyz-only network)
  uint32_t w_sigma_off;  // fp32 [256]: sigma-head kernel rows 0..255
  uint32_t w_rgb_off;    // fp32 float4 [128]: rgb-head kernel rows
  uint32_t total_bytes;
};

// xyz: the xyz-only network (src/NeRF.py:248-288): one more step in front (dH9 = dZ_L W9^T), W8 is 256 x 256, and the
// sigma head's rank-1 term still enters at the step that produces dH8 (the chain kernel keys it on layer 8).
static void make_bwd_plan(BwdPlan* p, bool xyz = false) {
  memset(p, 0, sizeof(*p));
  p->n_hidden = xyz ? 9 : 8;
  int c = 0, s = 0;
  uint32_t off = 0;
  auto step = [&](int kind, int layer, int n, int nch) {
    p->step_first[s] = (int8_t)c;
    p->step_nch[s] = (int8_t)nch;
    p->step_kind[s] = (int8_t)kind;
    p->step_layer[s] = (int8_t)layer;
    p->step_n[s] = (int16_t)n;
    for (int i = 0; i < nch; ++i, ++c) {
      p->chunk_off[c] = off;
      p->chunk_bytes[c] = (uint32_t)n * 128u;
      off += (uint32_t)n * 128u;
    }
    ++s;
  };
  if (xyz) {
    step(STEP_MASK, 9, 256, 2);                     // dH9 = dZ_L W9^T
    step(STEP_MASK, 8, 256, 4);                     // dH8 = dZ_9 W8^T  (+ sigma term in the epilogue)
  } else {
    step(STEP_MASK, 8, 256, 2);                     // dH8 = dZ_L W8[0:256]^T  (+ sigma term in the epilogue)
  }
  for (int l = 7; l >= 5; --l) step(STEP_MASK, l, 256, 4);
  step(STEP_XSTASH, 4, 64, 4);                      // d xyz  = dZ5 W4[0:dx]^T
  step(STEP_MASK, 4, 256, 4);                       // dH4 = dZ5 W4[dx:dx+256]^T
  for (int l = 3; l >= 1; --l) step(STEP_MASK, l, 256, 4);
  step(STEP_XFINAL, 0, 64, 4);                      // d xyz += dZ1 W0^T
  p->n_steps = s;
  p->n_chunks = c;
  p->w_sigma_off = off;  off += 256 * 4;
  p->w_rgb_off = off;    off += 128 * 16;
  p->total_bytes = off;
}

// byte offset, inside the backward weight pack, of the first K chunk of W_l^T (the STEP_MASK step of layer l)
uint32_t bwd_pack_layer_offset(int layer) {
  BwdPlan p;
  make_bwd_plan(&p);          // (view network: the only one the pipe prototype handles)
  for (int s = 0; s < p.n_steps; ++s)
    if (p.step_kind[s] == STEP_MASK && p.step_layer[s] == layer) return p.chunk_off[p.step_first[s]];
  return 0xffffffffu;
}

uint32_t bwd_pack_bytes() {
  BwdPlan p;
  make_bwd_plan(&p, true);    // the larger of the two networks' packs: one size for both
  return p.total_bytes;
}

template <typename T16>
__global__ void pack_bwd_kernel(const __grid_constant__ BwdPlan plan, NetGeom g, const float* __restrict__ P,
                                uint8_t* __restrict__ packed) {
  const int chunk = blockIdx.y;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (chunk < plan.n_chunks) {
    int s = 0;
    while (!(chunk >= plan.step_first[s] && chunk < plan.step_first[s] + plan.step_nch[s])) ++s;
    const int rows = plan.step_n[s];
    if (e >= rows * 64) return;
    const int j = e >> 6, kk = e & 63;
    const int n = (chunk - plan.step_first[s]) * 64 + kk;   // output unit of the forward layer (contraction index)
    const int kind = plan.step_kind[s], layer = plan.step_layer[s];
    float v = 0.f;
    if (kind == STEP_MASK) {
      const LayerDesc& L = g.layers[layer];
      int row = (layer == 4) ? g.dx + j : j;
      if (n < L.out) v = P[L.w_off + (int64_t)row * L.out + n];
    } else {
      const LayerDesc& L = g.layers[kind == STEP_XSTASH ? 4 : 0];
      if (j < g.dx) v = P[L.w_off + (int64_t)j * L.out + n];
    }
    if constexpr (std::is_same<T16, __half>::value)
      *reinterpret_cast<__half*>(packed + plan.chunk_off[chunk] + panel_offset(j, kk)) = __float2half_rn(v);
    else
      *reinterpret_cast<__nv_bfloat16*>(packed + plan.chunk_off[chunk] + panel_offset(j, kk)) = __float2bfloat16_rn(v);
  } else {
    if (e < 256) reinterpret_cast<float*>(packed + plan.w_sigma_off)[e] = P[g.layers[g.view ? 10 : 11].w_off + e];
    if (e < 128) {
      const LayerDesc& L = g.layers[g.view ? 9 : 10];
      reinterpret_cast<float4*>(packed + plan.w_rgb_off)[e] =
          make_float4(P[L.w_off + e * 3 + 0], P[L.w_off + e * 3 + 1], P[L.w_off + e * 3 + 2], 0.f);
    }
  }
}

int bwd_pack_weights(const NetGeom& g, const float* params, uint8_t* packed_bwd, cudaStream_t st, bool half) {
  BwdPlan plan;
  make_bwd_plan(&plan, !g.view);
  dim3 grid((256 * 64 + 255) / 256, plan.n_chunks + 1);
  if (half) pack_bwd_kernel<__half><<<grid, 256, 0, st>>>(plan, g, params, packed_bwd);
  else pack_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(plan, g, params, packed_bwd);
  NERF_CHECK_LAUNCH();
  return NERF_OK;
}

// =====================================================================================================================
// (1) dX chain
// =====================================================================================================================
// CTA pairs like the forward kernel (mlp_tc.cu): this CTA's half of a weight chunk is <= 16 KB, six stages fit
constexpr int kCStageBytes = 16384, kCStages = 6;
constexpr int kSmemCAct = 0;                                          // [2 tiles][4 panels]
constexpr int kSmemCStage = 2 * kActPanels * kPanelBytes;             // [kCStages] x 16 KB
constexpr int kSmemCBar = kSmemCStage + kCStages * kCStageBytes;
constexpr int kSmemCConst = kSmemCBar + 256;          // fp32 [256] sigma-head kernel, then fp32 [128][3] rgb-head kernel
constexpr int kSmemCAlloc = kSmemCConst + 1024 + 1536;
static_assert(kSmemCAlloc <= 232448, "chain kernel exceeds the 227 KB shared-memory limit");

struct ChainBars {
  uint64_t full[kCStages], empty[kCStages], act_ready[2], acc_full[2];
  uint64_t panel_full[2], panel_free[2];   // epilogue warps <-> the tile's store warp (see the forward kernel)
  uint32_t tmem_base;
};

// sign-mask layout written by the forward epilogue (neg_mask32, mlp_tc.cu): element e of a 32-column group -> bit
// 8 (e & 3) + 7 - (e >> 2), SET when the activation is negative; mask_bit() is true for the positive side
__device__ __forceinline__ bool mask_bit(uint32_t mw, int e) { return !((mw >> (8 * (e & 3) + 7 - (e >> 2))) & 1u); }

__device__ __forceinline__ void store_chunk16(uint32_t panel_row_addr, int r, int chunk16, uint4 v) {
  sts128(panel_row_addr + ((chunk16 ^ (r & 7)) << 4), v);
}

// Epilogue of one dX step for this thread's 128 columns: accumulator (+ the sigma-head rank-1 term of the first step)
// -> LeakyReLU' from the sign mask -> bf16 -> swizzled panel row.  pbx = (panel row address) | ((row & 7) << 4) of the
// first of the thread's two panels, so a 16-byte chunk address is one XOR with an immediate.
template <bool kSigma>
__device__ __forceinline__ void chain_mask_epilogue(uint32_t taddr_half, const uint32_t (&mw4)[4], uint32_t pbx, float alpha,
                                                    float dsig, uint32_t w_sigma_half_u32) {
#pragma unroll
  for (int cc = 0; cc < 4; ++cc) {
    uint32_t acc[32];
    tmem_ld32(taddr_half + cc * 32, acc);
    tmem_ld_wait();
    const uint32_t mw = mw4[cc];
    const uint32_t pb = pbx + (cc >> 1) * kPanelBytes;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float w[8];
      if (kSigma) {
        const float4 w0 = lds128f(w_sigma_half_u32 + (cc * 32 + 8 * j) * 4), w1 = lds128f(w_sigma_half_u32 + (cc * 32 + 8 * j + 4) * 4);
        w[0] = w0.x; w[1] = w0.y; w[2] = w0.z; w[3] = w0.w; w[4] = w1.x; w[5] = w1.y; w[6] = w1.z; w[7] = w1.w;
      }
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float a = __uint_as_float(acc[8 * j + i]);
        if (kSigma) a = fmaf(dsig, w[i], a);
        v[i] = mask_bit(mw, 8 * j + i) ? a : alpha * a;
      }
      sts128(pb ^ (uint32_t)((((cc & 1) * 4 + j)) << 4),
             make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7])));
    }
  }
}

// d z of one row from its 33 gradients w.r.t. the xyz encoding (L = 5): d z = sum_c d_c (g_c0 + sum_k 2^k pi (g_sin cos - g_cos
// sin)(2^k pi p_c)), p = o + d z -- positional_encoding_for_xyz and sample_along_rays differentiated
// (src/UtilsNeuralRadianceField.py:52-69, src/UtilsCV.py:598), with the sin / cos of the forward prologue
// (exact range reduction, then the SFU).  Inlined with compile-time indices: as a real call the 33 gradients went through
// local memory (L1 is ~3 KB next to 225 KB of shared memory) and cost the chain 3 us per tile.
__device__ __forceinline__ float dz_from_enc_grad(const float (&g)[33], float4 o, float4 d, float zz) {
  float gz = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float oc = c == 0 ? o.x : (c == 1 ? o.y : o.z), dc = c == 0 ? d.x : (c == 1 ? d.y : d.z);
    const float pc = __fadd_rn(oc, __fmul_rn(dc, zz));
    float acc = g[11 * c], t = pc, scale = 3.14159265358979f;
#pragma unroll
    for (int k = 0; k < 5; ++k, t *= 2.f, scale *= 2.f) {
      const float q = rintf(0.5f * t);
      const float a = fmaf(-2.f, q, t) * 3.14159265358979f;
      acc += (g[11 * c + 1 + 2 * k] * __cosf(a) - g[11 * c + 2 + 2 * k] * __sinf(a)) * scale;
    }
    gz = fmaf(acc, dc, gz);
  }
  return gz;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreadsFwd, 1)
mlp_tc_bwd_chain_kernel(const __grid_constant__ BwdPlan plan, const uint8_t* __restrict__ packed,
                        const uint8_t* __restrict__ saved, const float* __restrict__ d_out4, int64_t M,
                        uint8_t* __restrict__ dz_ws, float* __restrict__ d_xyz_enc, int dx, float alpha, uint32_t dbg,
                        uint32_t* __restrict__ flags, uint32_t stagger_ns, const BwdRays rays) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  if ((sbase & 1023u) != 0u) __trap();
  if (stagger_ns) {
    // overlapped mode: spread the pairs over the nine layer phases so that every dW unit sees a steady stream of blocks
    // instead of one burst per round (keeps the dZ in flight small enough to stay in L2)
    for (uint32_t ns = (cluster_id_x() % 9u) * stagger_ns; ns;) {
      const uint32_t d = ns < 20000u ? ns : 20000u;
      __nanosleep(d);
      ns -= d;
    }
  }
  ChainBars* bars = reinterpret_cast<ChainBars*>(smem + kSmemCBar);
  // L1 is ~3 KB next to 200 KB of shared memory: the head kernels the epilogue needs live in shared memory
  for (int i = threadIdx.x; i < 256 + 384; i += blockDim.x) {
    const int j = i - 256;                              // rgb-head kernel is packed as float4 (w_r, w_g, w_b, 0) per unit
    reinterpret_cast<float*>(smem + kSmemCConst)[i] =
        i < 256 ? reinterpret_cast<const float*>(packed + plan.w_sigma_off)[i]
                : reinterpret_cast<const float*>(packed + plan.w_rgb_off)[(j / 3) * 4 + (j % 3)];
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool need_dx = d_xyz_enc != nullptr || rays.d_z != nullptr;
  // a "quad" = the four 128-row tiles a CTA pair works on at a time: tile = 4 quad + 2 t + rank (t = super-tile 0/1)
  const uint32_t rank = cluster_ctarank();
  const int64_t n_tiles = (M + kTileM - 1) / kTileM;
  const int64_t n_quads = (n_tiles + 3) / 4;
  const int64_t quad0 = cluster_id_x(), quad_step = num_clusters_x();
  int last_step = 0;
  for (int s = 0; s < plan.n_steps; ++s)
    if (need_dx || plan.step_kind[s] == STEP_MASK) last_step = s;

  if (threadIdx.x == 0) {
    // full[s] of the LEADER also counts the peer's relay arrive: both halves of the chunk have landed
    for (int s = 0; s < kCStages; ++s) { mbar_init(smem_u32(&bars->full[s]), rank == 0 ? 2 : 1); mbar_init(smem_u32(&bars->empty[s]), 1); }
    for (int t = 0; t < 2; ++t) {
      mbar_init(smem_u32(&bars->act_ready[t]), 2 * 8);    // leader's copy: one arrive per epilogue warp of BOTH CTAs
      mbar_init(smem_u32(&bars->acc_full[t]), 1);
      mbar_init(smem_u32(&bars->panel_full[t]), 8);
      mbar_init(smem_u32(&bars->panel_free[t]), kStoreWarps);   // every store warp
    }
    fence_barrier_init();
  }
  cluster_sync_all();
  if (warp == kWarpMma) tmem_alloc_pair(smem_u32(&bars->tmem_base), 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == kWarpProducer) {
    // kProdLanes lanes walk the same sequence and issue alternate items: one thread gets a bulk copy out every ~550 cycles,
    // the four MMAs a chunk feeds take 512 (mlp_tc.cu, tools/l2_bw_probe.cu)
    if (lane < kProdLanes) {
      const uint32_t full0 = smem_u32(&bars->full[0]), empty0 = smem_u32(&bars->empty[0]);
      const bool no_copy = (dbg & kDbgNoWeightCopy) != 0;
      uint32_t st = 0, ph = 1, k = 0;                    // ring stage, parity of the "stage is free" phase, item index
      for (int64_t quad = quad0; quad < n_quads; quad += quad_step) {
        for (int s = 0; s < plan.n_steps; ++s) {
          if (!need_dx && plan.step_kind[s] != STEP_MASK) continue;
          const int first = plan.step_first[s], nch = plan.step_nch[s];
          const uint32_t half_bytes = plan.chunk_bytes[first] >> 1;     // rows [rank N/2, (rank + 1) N/2) of [N][64]
          const uint8_t* src = packed + plan.chunk_off[first] + rank * half_bytes;
          // A step's chunks (at most four) are streamed ONCE per quad: the six-stage ring holds them for both super-tiles
          // (the MMA issuer releases a stage after its second use) and still has two stages to run ahead into the next step
          for (int ci = 0; ci < nch; ++ci, ++k) {
            if ((k & (uint32_t)(kProdLanes - 1)) == (uint32_t)lane) {
              mbar_wait_spin(empty0 + 8u * st, ph);
              if (no_copy) {
                mbar_arrive(full0 + 8u * st);
              } else {
                mbar_arrive_expect_tx(full0 + 8u * st, half_bytes);
                bulk_g2s(sbase + kSmemCStage + st * kCStageBytes, src + (size_t)ci * 2u * half_bytes, half_bytes, full0 + 8u * st);
              }
            }
            if (++st == kCStages) { st = 0; ph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == kWarpMma) {
    if (lane == 0 && rank != 0) {
      // peer CTA: forward "my half of the chunk has landed" to the leader's full barrier
      const uint32_t full0 = smem_u32(&bars->full[0]), full0_leader = mapa_shared(full0, 0);
      int per_quad = 0;
      for (int s = 0; s < plan.n_steps; ++s)
        if (need_dx || plan.step_kind[s] == STEP_MASK) per_quad += plan.step_nch[s];      // one fill per chunk and quad
      uint32_t st = 0, ph = 0;
      for (int64_t quad = quad0; quad < n_quads; quad += quad_step)
        for (int i = 0; i < per_quad; ++i) {
          mbar_wait_spin(full0 + 8u * st, ph);
          mbar_arrive_cluster(full0_leader + 8u * st);
          if (++st == kCStages) { st = 0; ph ^= 1u; }
        }
    } else if (lane == 0) {
      // leader: one thread issues every MMA of the pair; descriptors are register adds (see mlp_tc.cu)
      const uint32_t full0 = smem_u32(&bars->full[0]), empty0 = smem_u32(&bars->empty[0]);
      const uint64_t b0 = make_desc_kmajor(sbase + kSmemCStage);
      const bool no_mma = (dbg & kDbgNoMma) != 0;
      uint32_t st = 0, ph = 0, act_ph = 0;
      for (int64_t quad = quad0; quad < n_quads; quad += quad_step) {
#pragma unroll 1
        for (int s = 0; s < plan.n_steps; ++s) {
          if (!need_dx && plan.step_kind[s] != STEP_MASK) continue;
          const int nch = plan.step_nch[s];
          const uint32_t idesc = make_idesc(plan.step_n[s], 0, 0, 1, 256);
          const uint32_t st0 = st, ph0 = ph;               // the step's first stage: super-tile 1 walks the same stages again
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            mbar_wait_spin(smem_u32(&bars->act_ready[t]), act_ph);
            const uint32_t d_tmem = tmem_base + (uint32_t)t * 256u;
            uint64_t a = make_desc_kmajor(sbase + kSmemCAct + t * kActPanels * kPanelBytes);
            st = st0; ph = ph0;
            for (int ci = 0; ci < nch; ++ci, a += (kPanelBytes >> 4)) {
              if (t == 0) {                                // super-tile 0 waits for the chunk; it stays for super-tile 1
                mbar_wait_spin(full0 + 8u * st, ph);
                tc_fence_after();
              }
              const uint64_t b = b0 + (uint64_t)(st * (uint32_t)(kCStageBytes >> 4));
              if (!no_mma) {
                umma_pair(d_tmem, a, b, idesc, ci > 0 ? 1u : 0u);
                umma_pair(d_tmem, a + 2, b + 2, idesc, 1u);
                umma_pair(d_tmem, a + 4, b + 4, idesc, 1u);
                umma_pair(d_tmem, a + 6, b + 6, idesc, 1u);
              }
              if (t == 1) umma_commit_pair(empty0 + 8u * st);      // second use done: releases this stage in both CTAs
              if (++st == kCStages) { st = 0; ph ^= 1u; }
            }
            umma_commit_pair(smem_u32(&bars->acc_full[t]));
          }
          act_ph ^= 1u;
        }
      }
    }
  } else if (warp >= kWarpStore) {
    // store warps: tile t's freshly written dZ_l panels -> HBM (RBCM block of the dZ workspace); see mlp_tc.cu
    // both store warps work on every copy (one half of the column chunks each), in the order the epilogues finish
    const int hw = warp - kWarpStore;
    const bool do_store = !(dbg & kDbgNoStore);
    uint32_t ph = 0;
    for (int64_t quad = quad0; quad < n_quads; quad += quad_step) {
      for (int l = plan.n_hidden; l >= 1; --l, ph ^= 1u) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const uint32_t act_u32 = sbase + kSmemCAct + t * kActPanels * kPanelBytes;
          const int64_t tile = quad * 4 + t * 2 + rank;
          uint8_t* gblock = dz_ws + (size_t)tile * kDzTileBytes + (size_t)dz_panel(l) * kPanelBytes;
          mbar_wait(smem_u32(&bars->panel_full[t]), ph);
          if (do_store) {
#pragma unroll 8
            for (int it = 0; it < 128 / kStoreWarps; ++it) {
              const int j = hw * (32 / kStoreWarps) + (it >> 2), r = (it & 3) * 32 + lane;     // 16-byte column chunk, row
              const float4 v = lds128f(act_u32 + (j >> 3) * kPanelBytes + r * 128 + (((j & 7) ^ (r & 7)) << 4));
              stg128(gblock + rbcm_offset(r, j, 32),
                     make_uint4(__float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w)));
            }
          }
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(smem_u32(&bars->panel_free[t]));
            if (flags) flag_signal(flags + (size_t)tile * kFlagsPerTile + (l - 1));   // dZ_l of this tile: my half is out
          }
        }
      }
    }
  } else {
    // 16 epilogue warps: tile t = warp / 8, TMEM lane quarter q = warp % 4, column half = (warp / 4) % 2
    const int t = warp >> 3, q = warp & 3, half = (warp >> 2) & 1;
    const int r = q * 32 + lane;
    const int gtid = threadIdx.x & (kEpiThreadsPerTile - 1);
    const int bar_id = 1 + t;
    const uint32_t act_u32 = sbase + kSmemCAct + t * kActPanels * kPanelBytes;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)t * 256u;
    const uint32_t w_sigma_u32 = sbase + kSmemCConst, w_rgb_u32 = sbase + kSmemCConst + 1024;
    uint32_t acc_cnt = 0, copy_ph = 0;
    bool copy_pending = false;
    const uint32_t act_ready_leader = mapa_shared(smem_u32(&bars->act_ready[t]), 0);
    for (int64_t quad = quad0; quad < n_quads; quad += quad_step) {
      const int64_t tile = quad * 4 + t * 2 + rank;
      const int64_t row = tile * kTileM + r;
      const bool row_ok = row < M;
      const uint8_t* saved_tile = saved + (size_t)tile * kSavedTileBytes;
      const uint32_t* saved_mask = reinterpret_cast<const uint32_t*>(saved_tile + (size_t)kSavedPanels * kPanelBytes);
      uint8_t* dz_tile = dz_ws + (size_t)tile * kDzTileBytes;
      // ---- prologue: half 0 -> dZ_L cols 0..63 (panel 0) + sigma-gradient panel (2);
      //                half 1 -> dZ_L cols 64..127 (panel 1) + d_out panel (3)
      float4 d4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row_ok) d4 = __ldg(reinterpret_cast<const float4*>(d_out4) + row);
      uint32_t mwl[2];
#pragma unroll
      for (int w = 0; w < 2; ++w) mwl[w] = __ldg(saved_mask + (kMaskRowHL * 8 + half * 2 + w) * 128 + r);
      // the previous tile's last epilogue wrote these panel rows from other threads, and its store warp may still read them
      if (copy_pending) {
        mbar_wait(smem_u32(&bars->panel_free[t]), copy_ph);
        copy_ph ^= 1u;
        copy_pending = false;
      }
      named_bar_sync(bar_id, kEpiThreadsPerTile);
      const bool do_store = !(dbg & kDbgNoStore);
      {
        uint8_t* gL = dz_tile + (size_t)kDzPanelL * kPanelBytes;
#pragma unroll
        for (int jg = 0; jg < 8; ++jg) {
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int jl = jg * 8 + i;                   // column inside this half
            const uint32_t wa = w_rgb_u32 + (half * 64 + jl) * 12;
            float dh = d4.x * lds32f(wa) + d4.y * lds32f(wa + 4) + d4.z * lds32f(wa + 8);
            v[i] = mask_bit(mwl[jl >> 5], jl & 31) ? dh : alpha * dh;
          }
          uint4 pk = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                                pack_bf16x2(v[6], v[7]));
          store_chunk16(act_u32 + half * kPanelBytes + r * 128, r, jg, pk);
          if (do_store) stg128(gL + rbcm_offset(r, half * 8 + jg, kDzChunksL), pk);
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint4 v;
          if (half == 0) v = make_uint4(c == 0 ? pack_bf16x2(d4.w, 0.f) : 0u, 0u, 0u, 0u);
          else v = make_uint4(c == 0 ? pack_bf16x2(d4.x, d4.y) : 0u, c == 0 ? pack_bf16x2(d4.z, d4.w) : 0u, 0u, 0u);
          store_chunk16(act_u32 + (2 + half) * kPanelBytes + r * 128, r, c, v);
          if (do_store && c < 2) {                       // the dW pass reads chunks 16, 17 of dZ_L' and 0, 1 of dOut
            if (half == 0) stg128(gL + rbcm_offset(r, 16 + c, kDzChunksL), v);
            else stg128(dz_tile + (size_t)kDzPanelOut * kPanelBytes + rbcm_offset(r, c, kDzChunksOut), v);
          }
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_cluster(act_ready_leader);
        if (flags) flag_signal(flags + (size_t)tile * kFlagsPerTile + kHiddenSlots);   // this warp's part of dZ_L' / dOut is out
      }

      float xs[33];                                      // d(xyz encoding): half 0 -> cols 0..31 (+ col 32 when it forms d z
                                                         // itself), half 1 -> cols 32..39
#pragma unroll
      for (int i = 0; i < 33; ++i) xs[i] = 0.f;
      for (int s = 0; s < plan.n_steps; ++s) {
        const int kind = plan.step_kind[s];
        if (!need_dx && kind != STEP_MASK) continue;
        const int l = plan.step_layer[s];
        // the LeakyReLU' masks do not depend on the accumulator: fetch them (L2 latency) before waiting for the MMAs
        uint32_t mw4[4] = {0u, 0u, 0u, 0u};
        if (kind == STEP_MASK) {
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) mw4[cc] = __ldg(saved_mask + ((l - 1) * 8 + half * 4 + cc) * 128 + r);
        }
        mbar_wait(smem_u32(&bars->acc_full[t]), acc_cnt & 1u);
        ++acc_cnt;
        tc_fence_after();
        if (kind == STEP_MASK) {
          if (copy_pending) {                            // the store warp must have read the panels this epilogue overwrites
            mbar_wait(smem_u32(&bars->panel_free[t]), copy_ph);
            copy_ph ^= 1u;
            copy_pending = false;
          }
          if (!(dbg & kDbgNoEpi)) {
            const uint32_t pbx = (act_u32 + (half * 2) * kPanelBytes + r * 128) | ((uint32_t)(r & 7) << 4);
            if (l == 8) chain_mask_epilogue<true>(taddr + half * 128, mw4, pbx, alpha, d4.w, w_sigma_u32 + half * 512);
            else chain_mask_epilogue<false>(taddr + half * 128, mw4, pbx, alpha, 0.f, 0u);
          }
          tc_fence_before();
          fence_proxy_async();
          __syncwarp();
          if (s != last_step && lane == 0) mbar_arrive_cluster(act_ready_leader);
          if (lane == 0) mbar_arrive(smem_u32(&bars->panel_full[t]));   // hand dZ_l to the tile's store warp
          copy_pending = true;
        } else {
          if (half == 0) {
            uint32_t a0[32], a32 = 0u;
            tmem_ld32(taddr, a0);
            if (rays.d_z) tmem_ld1(taddr + 32, a32);      // column 32 = the last cosine of the z coordinate
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) xs[i] += __uint_as_float(a0[i]);
            xs[32] += __uint_as_float(a32);
          } else {
            uint32_t a1[16];
            tmem_ld16(taddr + 32, a1);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 8; ++i) xs[i] += __uint_as_float(a1[i]);
          }
          tc_fence_before();
          if (kind == STEP_XSTASH) {
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(act_ready_leader);
          } else if (row_ok) {
            if (rays.d_z && half == 0) {
              const int64_t ray = row / rays.n_samples;
              const float gz = dz_from_enc_grad(xs, __ldg(rays.origs + ray), __ldg(rays.dirs + ray), __ldg(rays.z + row));
              rays.d_z[row] = rays.accumulate ? rays.d_z[row] + gz : gz;
            }
            if (d_xyz_enc) {
              float* dst = d_xyz_enc + row * dx + half * 32;
              const int lim = dx - half * 32;
#pragma unroll
              for (int i = 0; i < 32; ++i) if (i < lim && (half == 0 || i < 8)) dst[i] = xs[i];
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                  // the peer's shared memory / TMEM stay valid until both are done
  if (warp == kWarpMma) tmem_dealloc_pair(tmem_base, 512);
}

// =====================================================================================================================
// (2) dW
// =====================================================================================================================
enum : int { OUT_MAIN = 0, OUT_D8A = 1, OUT_INP_XYZ = 2, OUT_INP_VIEW = 3, OUT_RGB = 4, OUT_SIGMA_X = 5 };

constexpr int kDwMaxStages = 10;
constexpr int kDwRingBytes = 196608;
constexpr int kDwOperandBytes = 32768;           // 64 rows x 256 columns bf16: the largest operand slab
constexpr int kDwHalf = 8192;                    // 64 rows of the swizzled input panel
struct DwUnit {
  int32_t a_off;          // byte offset of the A block inside a saved tile
  int32_t b_off;          // byte offset of the B block inside a dZ-workspace tile
  int16_t a_inp;          // 1: A = saved input panel (128-byte-swizzled [128][64]); 0: RBCM block
  int16_t a_chunks;       // column chunks of the A block (RBCM): 32 (h_l) or 16 (h_L)
  int16_t b_chunks;       // column chunks of the B block (its half stride)
  int16_t b_load;         // chunks actually loaded = ceil(n / 8)
  int16_t n, m_blocks, out_kind, dense, has_bias, first_cta, n_ctas;
  int16_t stages;         // ring depth: as many (A slab | B slab) stages as fit in the 192 KB ring, at most kDwMaxStages
  int32_t a_region;       // bytes reserved for the A slab inside a stage (the B slab follows)
  int32_t stage_bytes;
  float cost;             // modelled time per 64-row stage, the unit of the CTA allocation
  int16_t flag_idx;       // overlapped mode: which ready counter of a tile guards the B block, and its final value
  int16_t flag_target;
};
struct DwPlan {
  DwUnit u[16];
  int32_t n_units;
};

static void make_dw_plan(DwPlan* p, int n_ctas_total, bool xyz = false) {
  memset(p, 0, sizeof(*p));
  int n = 0;
  auto add = [&](int a_panel, int a_chunks, int b_panel, int b_chunks, int nn, int out_kind, int dense, int has_bias) {
    DwUnit& u = p->u[n++];
    u.a_off = a_panel * kPanelBytes; u.a_inp = (int16_t)(a_chunks == 0); u.a_chunks = (int16_t)a_chunks;
    u.b_off = b_panel * kPanelBytes; u.b_chunks = (int16_t)b_chunks; u.n = (int16_t)nn; u.b_load = (int16_t)((nn + 7) / 8);
    u.m_blocks = (int16_t)(a_chunks == 32 ? 2 : 1);
    u.out_kind = (int16_t)out_kind; u.dense = (int16_t)dense; u.has_bias = (int16_t)has_bias;
    // input-panel units: 8 KB of data + 8 KB that stay zero (the second 64-column MN block of the M = 128 operand)
    u.a_region = a_chunks == 0 ? 2 * kDwHalf : a_chunks * 1024;
    u.stage_bytes = u.a_region + u.b_load * 1024;
    int st = kDwRingBytes / u.stage_bytes;
    u.stages = (int16_t)(st > kDwMaxStages ? kDwMaxStages : st);
    // measured per-CTA cycle counts (NERF_TC_DEBUG=320, all 148 CTAs streaming): a 64-row stage costs ~26 cycles per KB
    // plus ~700 cycles that do not shrink with the ring depth -> a 27 KB-equivalent overhead per stage
    u.cost = (float)((a_chunks == 0 ? 8 : a_chunks) + u.b_load) + 27.f;
    // the view network's units, measured one by one (tools/dw_balance_probe.py, slowest CTA of the unit, cycles per tile
    // relative to a 256 x 256 unit = 91; profiles/r02_ah_dw_balance.log): the short-N units cost more than their bytes say
    // (Dense 8 + sigma head) and the input-panel / rgb-head units less
    if (!xyz) {
      if (out_kind == OUT_D8A) u.cost = 81.5f;
      else if (out_kind == OUT_INP_XYZ) u.cost = 60.2f;
      else if (out_kind == OUT_INP_VIEW) u.cost = 50.4f;
      else if (out_kind == OUT_RGB) u.cost = 40.1f;
    }
    const bool prologue_block = b_panel >= kDzPanelL;
    u.flag_idx = (int16_t)(prologue_block ? kHiddenSlots : b_panel / kActPanels);   // dz_panel(l) = (l - 1) kActPanels
    u.flag_target = (int16_t)(prologue_block ? kFlagTargetPrologue : kFlagTargetStore);
  };
  // Dense l (input h_l, or the input panel) with the gradient of its pre-activation output dZ_{l+1}
  add(0, 0, dz_panel(1), 32, 256, OUT_INP_XYZ, 0, 1);                                  // Dense 0
  for (int l = 1; l <= 7; ++l) add(saved_panel_h(l), 32, dz_panel(l + 1), 32, 256, OUT_MAIN, l, 1);  // Dense 1..7 (4: h4 rows)
  add(0, 0, dz_panel(5), 32, 256, OUT_INP_XYZ, 4, 0);                                  // Dense 4, xyz rows
  if (xyz) {
    add(saved_panel_h(8), 32, dz_panel(9), 32, 256, OUT_MAIN, 8, 1);                   // Dense 8: h8 -> h9
    add(saved_panel_h(9), 32, kDzPanelL, kDzChunksL, 128, OUT_MAIN, 9, 1);             // Dense 9: h9 -> last hidden
    add(kSavedPanelHL, 16, kDzPanelOut, kDzChunksOut, 16, OUT_RGB, 10, 1);             // rgb head (Dense 10)
    add(saved_panel_h(8), 32, kDzPanelOut, kDzChunksOut, 16, OUT_SIGMA_X, 11, 1);      // sigma head (Dense 11): h8^T d sigma
  } else {
    add(saved_panel_h(8), 32, kDzPanelL, kDzChunksL, 144, OUT_D8A, 8, 1);              // Dense 8 + sigma head, h8 rows
    add(0, 0, kDzPanelL, kDzChunksL, 144, OUT_INP_VIEW, 8, 0);                         // Dense 8 + sigma head, view rows
    add(kSavedPanelHL, 16, kDzPanelOut, kDzChunksOut, 16, OUT_RGB, 9, 1);              // rgb head (Dense 9)
  }
  p->n_units = n;
  // CTAs by modelled cost: the floor of the proportional share (at least one each), then every CTA that is left goes to
  // the unit with the largest cost per CTA -- the kernel ends with its slowest unit, so the split minimises the maximum
  float total = 0.f;
  for (int i = 0; i < n; ++i) total += p->u[i].cost;
  int used = 0;
  for (int i = 0; i < n; ++i) {
    int c = (int)((float)n_ctas_total * p->u[i].cost / total);
    if (c < 1) c = 1;
    p->u[i].n_ctas = (int16_t)c;
    used += c;
  }
  auto load = [&](int i) { return p->u[i].cost / (float)p->u[i].n_ctas; };
  while (used < n_ctas_total) {
    int best = 0;
    for (int i = 1; i < n; ++i) if (load(i) > load(best)) best = i;
    p->u[best].n_ctas++; ++used;
  }
  while (used > n_ctas_total) {               // fewer CTAs than units x 1 would need: take from the lightest loads
    int best = -1;
    for (int i = 0; i < n; ++i) if (p->u[i].n_ctas > 1 && (best < 0 || load(i) < load(best))) best = i;
    p->u[best].n_ctas--; --used;
  }
  int first = 0;
  for (int i = 0; i < n; ++i) { p->u[i].first_cta = (int16_t)first; first += p->u[i].n_ctas; }
}

constexpr int kSmemDwBar = kDwRingBytes;
constexpr int kSmemDwAlloc = kSmemDwBar + 256;
static_assert(kSmemDwAlloc <= 232448, "dW kernel exceeds the 227 KB shared-memory limit");
constexpr int kThreadsDw = 192;

struct DwBars {
  uint64_t full[kDwMaxStages], empty[kDwMaxStages], acc_full;
  uint32_t tmem_base;
};

__device__ __forceinline__ float* dw_target(const DwUnit& u, const NetGeom& g, float* G, int k, int n) {
  switch (u.out_kind) {
    case OUT_MAIN: {
      const LayerDesc& L = g.layers[u.dense];
      int row = (u.dense == 4 ? g.dx : 0) + k;
      return G + L.w_off + (int64_t)row * L.out + n;
    }
    case OUT_D8A:
      if (n < 128) return G + g.layers[8].w_off + (int64_t)k * 128 + n;
      if (n == 128) return G + g.layers[10].w_off + k;
      return nullptr;
    case OUT_INP_XYZ:
      if (k < g.dx) return G + g.layers[u.dense].w_off + (int64_t)k * 256 + n;
      return nullptr;
    case OUT_INP_VIEW: {
      if (k < kInpViewCol || k >= kInpViewCol + g.dv) return nullptr;
      int row = g.hidden + (k - kInpViewCol);
      if (n < 128) return G + g.layers[8].w_off + (int64_t)row * 128 + n;
      if (n == 128) return G + g.layers[10].w_off + row;
      return nullptr;
    }
    case OUT_RGB:
      if (n < 3) return G + g.layers[u.dense].w_off + (int64_t)k * 3 + n;
      return nullptr;
    case OUT_SIGMA_X:                      // dOut block = [d r, d g, d b, d sigma, 0 ...]: column 3 is the sigma head's
      if (n == 3) return G + g.layers[u.dense].w_off + k;
      return nullptr;
  }
  return nullptr;
}

__device__ __forceinline__ float* db_target(const DwUnit& u, const NetGeom& g, float* G, int n) {
  switch (u.out_kind) {
    case OUT_MAIN:
    case OUT_INP_XYZ: return G + g.layers[u.dense].b_off + n;
    case OUT_D8A:
      if (n < 128) return G + g.layers[8].b_off + n;
      if (n == 128) return G + g.layers[10].b_off;
      return nullptr;
    case OUT_RGB: return n < 3 ? G + g.layers[u.dense].b_off + n : nullptr;
    case OUT_SIGMA_X: return n == 3 ? G + g.layers[u.dense].b_off : nullptr;
  }
  return nullptr;
}

// Clusters of 2 only so that the CTAs fill whole TPCs and leave whole TPCs to the chain kernel's CTA pairs when the two
// kernels share the GPU (overlapped mode); the two CTAs of a cluster never talk to each other.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreadsDw, 1)
mlp_tc_bwd_dw_kernel(const __grid_constant__ DwPlan plan, const __grid_constant__ NetGeom g,
                     const uint8_t* __restrict__ saved, const uint8_t* __restrict__ dz_ws, int64_t M,
                     float* __restrict__ scratch, uint32_t dbg, const uint32_t* __restrict__ flags) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  if ((sbase & 1023u) != 0u) __trap();
  DwBars* bars = reinterpret_cast<DwBars*>(smem + kSmemDwBar);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  int ui = 0;
  while (ui + 1 < plan.n_units && (int)blockIdx.x >= plan.u[ui].first_cta + plan.u[ui].n_ctas) ++ui;
  const DwUnit& u = plan.u[ui];
  const int split = blockIdx.x - u.first_cta;
  const int64_t n_tiles = (M + kTileM - 1) / kTileM;
  // this split's tiles: split, split + n_ctas, split + 2 n_ctas, ... (the order the chain kernel produces them in)
  if (split >= n_tiles) return;               // whole CTA leaves together: nothing to do for this split
  const int64_t my_tiles = (n_tiles - split + u.n_ctas - 1) / u.n_ctas;
  const uint32_t n_stages_total = (uint32_t)my_tiles * 2u;
  const long long t_start = clock64();

  // zero the operand slots once: the unused second A block of the input-panel units must read as zeros
  for (int i = threadIdx.x; i < kDwRingBytes / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async();
  if (threadIdx.x == 0) {
    for (int s = 0; s < kDwMaxStages; ++s) { mbar_init(smem_u32(&bars->full[s]), 1); mbar_init(smem_u32(&bars->empty[s]), 5); }
    mbar_init(smem_u32(&bars->acc_full), 1);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc(smem_u32(&bars->tmem_base), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  const uint32_t a_bytes = u.a_inp ? (uint32_t)kDwHalf : (uint32_t)u.a_chunks * 1024u;   // per 64-row slab
  const uint32_t b_bytes = (uint32_t)u.b_load * 1024u, b_half_stride = (uint32_t)u.b_chunks * 1024u;
  const uint32_t n_ring = (uint32_t)u.stages, stage_bytes = (uint32_t)u.stage_bytes, a_region = (uint32_t)u.a_region;

  if (warp == 4) {
    if (lane == 0) {
      uint32_t gi = 0;
      for (int64_t tile = split; tile < n_tiles; tile += u.n_ctas) {
        const uint8_t* a_src = saved + (size_t)tile * kSavedTileBytes + (size_t)u.a_off;
        const uint8_t* b_src = dz_ws + (size_t)tile * kDzTileBytes + (size_t)u.b_off;
        if (flags) {
          // the chain kernel is writing this block right now: wait for its ready counter, then order the bulk copies
          // (async proxy) after the acquire
          flag_wait(flags + (size_t)tile * kFlagsPerTile + u.flag_idx, (uint32_t)u.flag_target);
          asm volatile("fence.proxy.async;" ::: "memory");
        }
        for (int half = 0; half < 2; ++half, ++gi) {
          const uint32_t st = gi % n_ring, ph = (gi / n_ring) & 1u;
          mbar_wait(smem_u32(&bars->empty[st]), ph ^ 1u);
          const uint32_t fb = smem_u32(&bars->full[st]);
          mbar_arrive_expect_tx(fb, a_bytes + b_bytes);
          const uint32_t dst = sbase + st * stage_bytes;
          bulk_g2s(dst, a_src + (size_t)half * a_bytes, a_bytes, fb);
          bulk_g2s(dst + a_region, b_src + (size_t)half * b_half_stride, b_bytes, fb);
        }
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc(u.n, 1, 1);
      for (uint32_t gi = 0; gi < n_stages_total; ++gi) {
        const uint32_t st = gi % n_ring, ph = (gi / n_ring) & 1u;
        mbar_wait(smem_u32(&bars->full[st]), ph);
        tc_fence_after();
        const uint32_t base = sbase + st * stage_bytes;
#pragma unroll
        for (int k = 0; k < 4; ++k) {                       // K = 16 rows per MMA
          if (dbg & kDbgNoMma) break;
          const uint64_t b_desc = make_desc_mn_nosw(base + a_region + k * 256, 128, 1024);
          for (int mb = 0; mb < u.m_blocks; ++mb) {
            const uint64_t a_desc = u.a_inp ? make_desc(base + k * 2048, kDwHalf, 1024)
                                            : make_desc_mn_nosw(base + mb * 16 * 1024 + k * 256, 128, 1024);
            umma_bf16(tmem_base + (uint32_t)mb * 256u, a_desc, b_desc, idesc, (gi > 0 || k > 0) ? 1u : 0u);
          }
        }
        umma_commit(smem_u32(&bars->empty[st]));
      }
      umma_commit(smem_u32(&bars->acc_full));
    }
  } else {
    // warps 0..3: bias-gradient column sums from the dZ slabs, then the accumulator drain
    const int tid = threadIdx.x;                       // 0..127, owns columns 2*tid, 2*tid+1 = chunk tid/4, word tid%4
    const int c = 2 * tid;
    const bool col_ok = u.has_bias && c < u.n;
    float s0 = 0.f, s1 = 0.f;
    for (uint32_t gi = 0; gi < n_stages_total; ++gi) {
      const uint32_t st = gi % n_ring, ph = (gi / n_ring) & 1u;
      mbar_wait(smem_u32(&bars->full[st]), ph);
      if (col_ok && !(dbg & kDbgNoBiasSum)) {
        // slab layout [chunk][row][16 B]: rows are visited rotated by (chunk & 7) so that the 8 chunks of a warp hit
        // 8 different bank groups
        const uint32_t pb = sbase + st * stage_bytes + a_region + (uint32_t)(tid >> 2) * 1024u + (uint32_t)(tid & 3) * 4u;
        const uint32_t rot = (uint32_t)(tid >> 2) & 7u;
        float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
#pragma unroll 8
        for (uint32_t i = 0; i < 64; i += 2) {
          const uint32_t w0 = lds32u(pb + (((i + rot) & 63u) << 4));
          const uint32_t w1 = lds32u(pb + (((i + 1 + rot) & 63u) << 4));
          a0 += __uint_as_float(w0 << 16);
          a1 += __uint_as_float(w0 & 0xffff0000u);
          b0 += __uint_as_float(w1 << 16);
          b1 += __uint_as_float(w1 & 0xffff0000u);
        }
        s0 += a0 + b0;
        s1 += a1 + b1;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bars->empty[st]));
    }
    // this CTA's split-K partial: [256 rows (k)][256 columns (n)] fp32, then the 256 bias column sums
    float* part = scratch + (size_t)blockIdx.x * kDwPartialFloats;
    part[256 * 256 + c] = s0;
    part[256 * 256 + c + 1] = s1;
    mbar_wait(smem_u32(&bars->acc_full), 0);
    tc_fence_after();
    const int q = warp;  // TMEM lane quarter
    for (int mb = 0; mb < u.m_blocks; ++mb) {
      if (dbg & kDbgNoDrain) break;
      const int k = mb * 128 + q * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)mb * 256u;
      for (int c0 = 0; c0 < u.n; c0 += 16) {
        uint32_t acc[16];
        tmem_ld16(taddr + c0, acc);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; i += 4)
          stg128(part + k * 256 + c0 + i, make_uint4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]));
      }
    }
    tc_fence_before();
    if ((dbg & kDbgTiming) && threadIdx.x == 0) {
      const long long t_acc = clock64();
      printf("dw cta %3d unit %2d (dense %d kind %d) tiles %4d  cycles %8lld\n", blockIdx.x, ui, u.dense, u.out_kind, (int)my_tiles,
             t_acc - t_start);
    }
  }
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem_base, 512);
}

// Adds the split-K partials of every unit into the flat gradient vector, in CTA order (deterministic).
// grid = (n_units, 64), 256 threads: thread = one float4 of the unit's [256][256] block (+ the bias row, block y == 0).
__global__ void __launch_bounds__(256)
mlp_tc_bwd_dw_reduce_kernel(const __grid_constant__ DwPlan plan, const __grid_constant__ NetGeom g,
                            const float* __restrict__ scratch, int64_t M, float* __restrict__ G) {
  const DwUnit& u = plan.u[blockIdx.x];
  const int64_t n_tiles = (M + kTileM - 1) / kTileM;
  const int live = (int)(n_tiles < u.n_ctas ? n_tiles : u.n_ctas);   // splits that had tiles (the others left without writing)
  const float* base = scratch + (size_t)u.first_cta * kDwPartialFloats;
  const int e4 = blockIdx.y * 256 + threadIdx.x;         // float4 index inside [256][64 float4]
  const int k = e4 >> 6, n0 = (e4 & 63) * 4;
  if (k < u.m_blocks * 128 && n0 < u.n) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8                                           // eight independent 16-byte loads in flight per thread; same order of additions
    for (int s = 0; s < live; ++s) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(base + (size_t)s * kDwPartialFloats + k * 256 + n0));
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    const float vals[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float* tgt = (n0 + i < u.n) ? dw_target(u, g, G, k, n0 + i) : nullptr;
      if (tgt) *tgt += vals[i];
    }
  }
  if (blockIdx.y == 0 && u.has_bias && threadIdx.x < u.n) {
    float acc = 0.f;
    for (int s = 0; s < live; ++s) acc += __ldg(base + (size_t)s * kDwPartialFloats + 256 * 256 + threadIdx.x);
    float* tgt = db_target(u, g, G, threadIdx.x);
    if (tgt) *tgt += acc;
  }
}

// ---- host ---------------------------------------------------------------------------------------------------------------
// env NERF_BWD_CHAIN_PAIRS: CTA pairs of the chain kernel in overlapped mode (the dW kernel gets the other SMs);
// env NERF_BWD_STAGGER_NS: start offset between consecutive layer phases of the chain pairs (0 = none)
static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

int64_t mlp_tc_bwd_flag_bytes(int64_t m) {
  int64_t tiles4 = ((m + kTileM - 1) / kTileM + 3) / 4 * 4;
  return tiles4 * kFlagsPerTile * (int64_t)sizeof(uint32_t);
}

// side != nullptr (and parts == 3): overlapped mode -- the chain on `st`, the dW kernel + reduction on `side` at the same
// time (see the top of this file); the caller joins `side` before it reads `grads`.
int mlp_tc_bwd(const nerf_net_cfg* cfg, const NetGeom& g, const float* params, const void* packed, const float* xyz_enc,
               const float* view_enc, const void* saved, const float* d_out4, int64_t m, float* grads, float* d_xyz_enc,
               void* workspace, cudaStream_t st, int parts, cudaStream_t side, bool half, const BwdRays* rays) {
  (void)params; (void)xyz_enc; (void)view_enc;
  BwdRays br;
  memset(&br, 0, sizeof(br));
  if (rays) {
    if (g.dx != 33 || !g.view) {
      set_error("nerf_mlp_bwd_rays: d z in the chain epilogue is built for n_pos_enc_dim_xyz = 5 view networks");
      return NERF_E_UNSUPPORTED;
    }
    br = *rays;
  }
  TcPlan fplan;
  if (!make_plan(g, &fplan)) {
    set_error("the tensor-core modes support hidden=256, last_hidden=128, xyz width <= 38, view width <= 24");
    return NERF_E_UNSUPPORTED;
  }
  if (device_first_use(1)) {
    NERF_CUDA(cudaFuncSetAttribute(mlp_tc_bwd_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemCAlloc));
    NERF_CUDA(cudaFuncSetAttribute(mlp_tc_bwd_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemDwAlloc));
  }
  BwdPlan bplan;
  make_bwd_plan(&bplan, !g.view);
  // Both 16-bit modes share this backward: tcgen05 kind::f16 wants ONE operand format per MMA (an fp16 A with a bf16 B is
  // an illegal instruction on sm_100a -- tried), and the gradients need bf16's range, so the forward of the fp16 mode
  // saves its activations converted to bf16 and the chain reads the bf16 W^T.
  (void)half;
  const uint8_t* packed_bwd = (const uint8_t*)packed + pack_off_bwd(fplan);
  uint8_t* dz_ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023);
  const int64_t tiles4 = ((m + kTileM - 1) / kTileM + 3) / 4 * 4;
  const int64_t n_quads = tiles4 / 4;
  float* scratch = reinterpret_cast<float*>(dz_ws + tiles4 * (int64_t)kDzTileBytes);
  uint32_t* flag_buf = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(scratch) + kDwScratchBytes);
  const uint32_t dbg = tc_debug_flags();
  const int sms = num_sms() & ~1;
  // Measured (profiles/r02_a_overlap_probe.log): the dW kernel is bound by what ONE SM can pull in (~35 GB/s per SM whether
  // dZ comes from L2 or HBM), so giving it fewer SMs costs more than the L2 hand-over saves (1.04 vs 0.76 ms at 262 k
  // rows).  The overlapped mode therefore stays opt-in (NERF_BWD_OVERLAP=1) as the experiment it was.
  const bool overlap = side != nullptr && side != st && parts == 3 && !(dbg & (kDbgNoChain | kDbgNoDw)) && sms >= 8 &&
                       env_int("NERF_BWD_OVERLAP", 0) != 0;
  int chain_pairs = sms / 2, dw_ctas = sms;
  uint32_t* flags = nullptr;
  uint32_t stagger_ns = 0;
  if (overlap) {
    // SM split: the chain costs ~2x the SM time of the dW kernel when the latter reads dZ from L2
    const int env_pairs = env_int("NERF_BWD_CHAIN_PAIRS", 0);
    const int env_stagger = env_int("NERF_BWD_STAGGER_NS", -1);
    chain_pairs = env_pairs > 0 ? env_pairs : (sms / 2) * 48 / 74;
    if (chain_pairs > sms / 2 - 2) chain_pairs = sms / 2 - 2;
    if (chain_pairs < 1) chain_pairs = 1;
    dw_ctas = sms - 2 * chain_pairs;
    flags = flag_buf;
    stagger_ns = n_quads >= 4 * (int64_t)chain_pairs ? (uint32_t)(env_stagger >= 0 ? env_stagger : 2500) : 0u;
    NERF_CUDA(cudaMemsetAsync(flags, 0, (size_t)mlp_tc_bwd_flag_bytes(m), st));
    cudaEvent_t ev;
    NERF_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    cudaError_t e = cudaEventRecord(ev, st);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(side, ev, 0);
    cudaEventDestroy(ev);
    if (e != cudaSuccess) { set_error("mlp_tc_bwd: stream hand-over failed: %s", cudaGetErrorString(e)); return NERF_E_CUDA; }
  }
  const int grid = 2 * (int)(n_quads < chain_pairs ? n_quads : chain_pairs);     // CTA pairs
  if (!(dbg & kDbgNoChain) && (parts & 1)) {
    mlp_tc_bwd_chain_kernel<<<grid, kThreadsFwd, kSmemCAlloc, st>>>(bplan, packed_bwd, (const uint8_t*)saved, d_out4, m,
                                                                   dz_ws, d_xyz_enc, g.dx, cfg->leaky_alpha, dbg, flags,
                                                                   stagger_ns, br);
    NERF_CHECK_LAUNCH();
  }
  if (!(dbg & kDbgNoDw) && (parts & 2)) {
    DwPlan dplan;
    make_dw_plan(&dplan, dw_ctas, !g.view);
    cudaStream_t dst = st;
    if (side != nullptr && side != st && parts == 3) {
      // the weight-gradient kernel goes to the caller's side stream: after the chain when the two run one after the
      // other (it then overlaps whatever the caller enqueues next on `st`), next to it in the overlapped mode
      dst = side;
      if (!overlap) {
        cudaEvent_t ev;
        NERF_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        cudaError_t e = cudaEventRecord(ev, st);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(side, ev, 0);
        cudaEventDestroy(ev);
        if (e != cudaSuccess) { set_error("mlp_tc_bwd: stream hand-over failed: %s", cudaGetErrorString(e)); return NERF_E_CUDA; }
      }
    }
    mlp_tc_bwd_dw_kernel<<<dw_ctas, kThreadsDw, kSmemDwAlloc, dst>>>(dplan, g, (const uint8_t*)saved, dz_ws, m, scratch, dbg,
                                                                     flags);
    NERF_CHECK_LAUNCH();
    mlp_tc_bwd_dw_reduce_kernel<<<dim3(dplan.n_units, 64), 256, 0, dst>>>(dplan, g, scratch, m, grads);
    NERF_CHECK_LAUNCH();
  }
  return NERF_OK;
}

}  // namespace nerf
