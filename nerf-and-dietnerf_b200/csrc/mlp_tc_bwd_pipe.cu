// K4 in NERF_MODE_BF16, layer-pipelined form: TF autodiff of the NeRF MLP inside NeRF.train_step (src/NeRF.py:149-167)
// with the input-gradient chain and the weight gradients of a layer computed BY THE SAME CTA PAIR from the same dZ, and
// dZ handed from layer to layer through L2 instead of HBM.
//
//   * The CTA pairs (cluster of 2, tcgen05 cta_group::2) are split into GROUPS, one per layer.  A group is
//     layer-STATIONARY: its pairs keep W_l^T resident in shared memory (64 KB per CTA: each CTA holds the 128 output rows
//     of every K chunk it contributes to the pair's MMA) and the whole 256 x 256 fp32 accumulator of dW_l in TMEM
//     (128 lanes x 256 columns per CTA, cta_group::2 M = 256) for the entire launch -- one split-K partial per PAIR.
//   * Per 256-row super-tile s (tiles 2s, 2s+1; CTA r owns the rows of tile 2s + r in the chain) a pair of group l
//       - waits for the ready counter of dZ_{l+1}[s] (written by a pair of group l+1 moments ago, so it sits in L2),
//       - chain:  dH_l = dZ_{l+1} W_l^T   (A = its own 128 rows, K-major; accumulator D1 = TMEM columns 0..255), epilogue =
//                 LeakyReLU' from the saved sign mask, bf16, straight from registers to the dZ_l block in global memory
//                 (512-byte warp stores), then the ready counter of dZ_l[s];
//       - dW:     dW_l += A_l^T dZ_{l+1}  (both operands MN-major, the pair splits the FEATURES: CTA r streams features
//                 128 r .. 128 r + 127 of the saved activations and of dZ_{l+1} for all 256 rows; accumulator D2 = TMEM
//                 columns 256..511), bias gradient = column sums of the same dZ pieces.
//   * Everything a CTA needs arrives as SIX 32 KB bulk copies per super-tile (blocks are tile chunk-major, TCM, mlp_tc.cuh):
//       c01, c23   its own tile's dZ, feature chunks 0..15 / 16..31: the chain's A operand, two 64-feature K panels each
//       B0, A0     tile 2s:    dZ and saved activations, this CTA's feature half -> dW, 8 K-steps of 16 rows
//       B1, A1     tile 2s+1:  the same
//     through ONE in-order ring of five 32 KB slots; the MMA groups C01 C23 D0 D1 take 1024 tensor cycles each, so every
//     copy is issued >= 2048 cycles before its data is needed and the chain accumulator drains (16 epilogue warps) under
//     D0 / D1.  Copy size and issue rate are what the first prototype of this kernel got wrong (16 x 1 KB gathers and 16 KB
//     pieces from one lane: 6.8 TB/s aggregate where the same ring moves 17-20 TB/s in 32 KB copies -- a thread gets one
//     bulk copy out every ~290 ns whatever its size, tools/l2_bw_probe.cu): here two producer lanes issue alternate items.
#include <stdlib.h>

#include "mlp_tc.cuh"
#include "mlp_tc_bwd_pipe.cuh"

namespace nerf {

// ---- shared-memory map ------------------------------------------------------------------------------------------------
constexpr int kPW = 0;                                   // resident W_l^T half: 4 K-chunks x [128 rows][64] swizzled
constexpr int kPWBytes = 4 * 16384;
constexpr int kPSlots = 5, kPSlotBytes = 32768;          // the ring
constexpr int kPRing = kPW + kPWBytes;
constexpr int kPBar = kPRing + kPSlots * kPSlotBytes;
constexpr int kPAlloc = kPBar + 256;
static_assert(kPAlloc <= 232448, "pipe kernel exceeds the 227 KB shared-memory limit");
constexpr int kPipeProdLanes = 2;

struct PipeBars {
  uint64_t wfull, full[kPSlots], empty[kPSlots];
  uint64_t acc_full, acc_empty, fin;
  uint32_t tmem_base;
};
static_assert(sizeof(PipeBars) <= 256, "barrier block overflows its slot");

// items of a super-tile in ring order (PIPE_ROLE_LAYER: all six; PIPE_ROLE_DW_ONLY: the last four)
enum : int { IT_C01 = 0, IT_C23 = 1, IT_B0 = 2, IT_A0 = 3, IT_B1 = 4, IT_A1 = 5 };

struct RingPos {
  uint32_t s, ph;
  __device__ __forceinline__ void next() { if (++s == (uint32_t)kPSlots) { s = 0; ph ^= 1u; } }
};

// ---- producer (warp 16 of both CTAs, kPipeProdLanes lanes) ----------------------------------------------------------------
template <bool kChain>
__device__ __forceinline__ void pipe_producer(const PipeGroup& G, PipeBars* bars, uint32_t sbase, uint32_t rank, int lane,
                                              const uint8_t* __restrict__ packed, const uint8_t* __restrict__ saved,
                                              const uint8_t* __restrict__ dz_ws, const uint32_t* flags, int64_t s_first,
                                              int64_t s_step, int64_t n_super, uint32_t dbg) {
  if (kChain && lane == 0) {
    const uint32_t wf = smem_u32(&bars->wfull);
    mbar_arrive_expect_tx(wf, kPWBytes);
#pragma unroll
    for (int kp = 0; kp < 4; ++kp)
      bulk_g2s(sbase + kPW + kp * 16384, packed + G.w_off + (size_t)kp * 32768 + rank * 16384u, 16384, wf);
  }
  RingPos p{0, 1};
  uint32_t k = 0;
  const uint32_t b_bytes = (uint32_t)G.b_chunks * 2048u;
  for (int64_t s = s_first; s < n_super; s += s_step) {
    if (G.wait_flag >= 0) {
      if (lane == 0) flag_wait(flags + (size_t)s * kPipeFlagsPerSuper + G.wait_flag, kPipeFlagTarget);
      __syncwarp((1u << kPipeProdLanes) - 1u);
      asm volatile("fence.proxy.async;" ::: "memory");
    }
    const int64_t sz = (dbg & kDbgWrapDz) ? (s & 127) : s, ss = (dbg & kDbgWrapSaved) ? (s & 127) : s;
#pragma unroll
    for (int it = kChain ? 0 : 2; it < 6; ++it, ++k) {
      if ((k & (uint32_t)(kPipeProdLanes - 1)) == (uint32_t)lane) {
        const uint8_t* src;
        uint32_t bytes;
        if (it <= IT_C23) {
          src = dz_ws + (size_t)(2 * sz + rank) * kDzTileBytes + (size_t)G.b_off + (size_t)it * 32768;
          bytes = 32768;
        } else {
          const int t = (it - 2) >> 1;
          if ((it & 1) == 0) {           // B: dZ_{l+1}, this CTA's feature half
            src = dz_ws + (size_t)(2 * sz + t) * kDzTileBytes + (size_t)G.b_off + (size_t)rank * b_bytes;
            bytes = b_bytes;
          } else {                       // A: saved activations, this CTA's feature half
            src = saved + (size_t)(2 * ss + t) * kSavedTileBytes + (size_t)G.a_off + (size_t)rank * 32768;
            bytes = 32768;
          }
        }
        const uint32_t fb = smem_u32(&bars->full[p.s]);
        mbar_wait_spin(smem_u32(&bars->empty[p.s]), p.ph);
        if (dbg & kDbgNoWeightCopy) {
          mbar_arrive(fb);
        } else {
          mbar_arrive_expect_tx(fb, bytes);
          bulk_g2s(sbase + kPRing + p.s * kPSlotBytes, src, bytes, fb);
        }
      }
      p.next();
    }
  }
}

// ---- relay (warp 17 lane 0 of the PEER CTA): "my half of this item has landed" -> the leader's full barrier --------------
template <bool kChain>
__device__ __forceinline__ void pipe_relay(PipeBars* bars, int64_t s_first, int64_t s_step, int64_t n_super) {
  if (kChain) {
    mbar_wait_spin(smem_u32(&bars->wfull), 0);
    mbar_arrive_cluster(mapa_shared(smem_u32(&bars->wfull), 0));
  }
  const uint32_t f0 = smem_u32(&bars->full[0]), f0_l = mapa_shared(f0, 0);
  RingPos p{0, 0};
  for (int64_t s = s_first; s < n_super; s += s_step) {
#pragma unroll
    for (int it = kChain ? 0 : 2; it < 6; ++it) {
      mbar_wait_spin(f0 + 8u * p.s, p.ph);
      mbar_arrive_cluster(f0_l + 8u * p.s);
      p.next();
    }
  }
}

// ---- MMA issuer (warp 17 lane 0 of the LEADER CTA) ---------------------------------------------------------------------------
template <bool kChain>
__device__ __forceinline__ void pipe_mma(const PipeGroup& G, PipeBars* bars, uint32_t sbase, uint32_t tmem_base,
                                         int64_t s_first, int64_t s_step, int64_t n_super, uint32_t dbg) {
  const uint32_t idesc_c = make_idesc(G.n_chain, 0, 0, 1, 256);
  const uint32_t idesc_d = make_idesc(G.n_dw, 1, 1, 1, 256);
  // the whole warp runs this loop (see elect_one); the CTA pair owns all 512 TMEM columns, so its base is column 0 (checked
  // by the caller) and the accumulator addresses are literals
  (void)tmem_base;
  const uint32_t d1 = 0u, d2 = 256u;
  const uint64_t a_c0 = make_desc_k_nosw(sbase + kPRing, 2048, 128);    // chain A: [16 chunks][128 rows][16 B]
  const uint64_t b_c0 = make_desc_kmajor(sbase + kPW);                  // chain B: resident swizzled W chunks
  const uint64_t mn0 = make_desc_mn_nosw(sbase + kPRing, 128, 2048);    // dW A / B: [chunk][128 rows][16 B], K = rows
  const uint32_t f0 = smem_u32(&bars->full[0]), e0 = smem_u32(&bars->empty[0]);
  const bool no_mma = (dbg & kDbgNoMma) != 0;
  if (kChain) {
    mbar_wait_spin(smem_u32(&bars->wfull), 0);
    tc_fence_after();
  }
  RingPos p{0, 0};
  uint32_t it = 0, dacc = 0;
  const bool timing = (dbg & kDbgTiming) != 0 && s_first == 0;
  long long t_acc = 0, t_c = 0, t_d = 0, t0 = clock64();
  for (int64_t s = s_first; s < n_super; s += s_step, ++it) {
    if (kChain) {
      if (it > 0) {
        const long long ta = timing ? clock64() : 0;
        mbar_wait_spin(smem_u32(&bars->acc_empty), (it - 1) & 1u);   // both CTAs drained D1 of the previous super-tile
        tc_fence_after();
        if (timing) t_acc += clock64() - ta;
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {                                  // c01, c23: two K panels each
        const long long ta = timing ? clock64() : 0;
        mbar_wait_spin(f0 + 8u * p.s, p.ph);
        tc_fence_after();
        if (timing) t_c += clock64() - ta;
        if (elect_one()) {
          if (!no_mma) {
            const uint64_t a = a_c0 + (uint64_t)(p.s * (uint32_t)(kPSlotBytes >> 4));
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              umma_pair(d1, a + (uint64_t)(kk * (4096 >> 4)),
                        b_c0 + (uint64_t)((2 * h + (kk >> 2)) * (16384 >> 4) + 2 * (kk & 3)), idesc_c, (h > 0 || kk > 0) ? 1u : 0u);
          }
          umma_commit_pair(e0 + 8u * p.s);
          if (h == 1) umma_commit_pair(smem_u32(&bars->acc_full));
        }
        __syncwarp();
        p.next();
      }
    }
#pragma unroll
    for (int t = 0; t < 2; ++t) {                                    // (B0, A0), (B1, A1)
      const RingPos pb = p;
      p.next();
      const RingPos pa = p;
      p.next();
      const long long ta = timing ? clock64() : 0;
      mbar_wait_spin(f0 + 8u * pb.s, pb.ph);
      mbar_wait_spin(f0 + 8u * pa.s, pa.ph);
      tc_fence_after();
      if (timing) t_d += clock64() - ta;
      if (elect_one()) {
        if (!no_mma) {
          const uint64_t a = mn0 + (uint64_t)(pa.s * (uint32_t)(kPSlotBytes >> 4));
          const uint64_t b = mn0 + (uint64_t)(pb.s * (uint32_t)(kPSlotBytes >> 4));
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            umma_pair(d2, a + (uint64_t)(kk * (256 >> 4)), b + (uint64_t)(kk * (256 >> 4)), idesc_d, (dacc | (uint32_t)kk) ? 1u : 0u);
        }
        umma_commit_pair(e0 + 8u * pb.s);
        umma_commit_pair(e0 + 8u * pa.s);
      }
      __syncwarp();
      dacc = 1u;
    }
  }
  if (elect_one()) umma_commit_pair(smem_u32(&bars->fin));
  __syncwarp();
  if (timing && (threadIdx.x & 31) == 0)
    printf("pipe mma (layer %d): %u super-tiles, %lld cycles each; waits per super-tile: acc_empty %lld, chain items %lld, dW items %lld\n",
           (int)G.layer, it, (clock64() - t0) / it, t_acc / it, t_c / it, t_d / it);
}

// ---- bias-gradient warps (18, 19): column sums of the dZ pieces this CTA streams (its feature half) ---------------------------
// They follow EVERY item of the ring (one arrival per warp on every slot's empty barrier keeps its count uniform) and read
// the B items.
template <bool kChain>
__device__ __forceinline__ void pipe_bias(const PipeGroup& G, PipeBars* bars, uint32_t sbase, int tid64, int lane,
                                          int64_t s_first, int64_t s_step, int64_t n_super, float* s0_out, float* s1_out,
                                          uint32_t dbg) {
  // thread owns feature columns 2 tid64, 2 tid64 + 1 of this CTA's half = chunk tid64 / 4, 32-bit word tid64 % 4
  const bool col_ok = 2 * tid64 < (int)G.b_chunks * 8;
  const uint32_t rot = (uint32_t)(tid64 >> 2) & 7u;
  float s0 = 0.f, s1 = 0.f;
  RingPos p{0, 0};
  for (int64_t s = s_first; s < n_super; s += s_step) {
#pragma unroll 1
    for (int it = kChain ? 0 : 2; it < 6; ++it) {
      mbar_wait(smem_u32(&bars->full[p.s]), p.ph);
      if ((it == IT_B0 || it == IT_B1) && col_ok && !(dbg & kDbgNoBiasSum)) {
        const uint32_t pb = sbase + kPRing + p.s * kPSlotBytes + (uint32_t)(tid64 >> 2) * 2048u + (uint32_t)(tid64 & 3) * 4u;
        float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
#pragma unroll 8
        for (uint32_t i = 0; i < 128; i += 2) {
          const uint32_t w0 = lds32u(pb + (((i + rot) & 127u) << 4));
          const uint32_t w1 = lds32u(pb + (((i + 1 + rot) & 127u) << 4));
          a0 += __uint_as_float(w0 << 16);
          a1 += __uint_as_float(w0 & 0xffff0000u);
          b0 += __uint_as_float(w1 << 16);
          b1 += __uint_as_float(w1 & 0xffff0000u);
        }
        s0 += a0 + b0;
        s1 += a1 + b1;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bars->empty[p.s]));
      p.next();
    }
  }
  *s0_out = s0;
  *s1_out = s1;
}

// sign-mask bit of element e of a 32-column group (layout of neg_mask32, mlp_tc.cu): true on the positive side
__device__ __forceinline__ bool pipe_mask_bit(uint32_t mw, int e) { return !((mw >> (8 * (e & 3) + 7 - (e >> 2))) & 1u); }

// ---- the kernel -----------------------------------------------------------------------------------------------------------------
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreadsFwd, 1)
mlp_tc_bwd_pipe_kernel(const __grid_constant__ PipePlan plan, const uint8_t* __restrict__ packed,
                       const uint8_t* __restrict__ saved, uint8_t* __restrict__ dz_ws, int64_t M,
                       float* __restrict__ scratch, uint32_t* __restrict__ flags, float alpha, uint32_t dbg) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  if ((sbase & 1023u) != 0u) __trap();
  PipeBars* bars = reinterpret_cast<PipeBars*>(smem + kPBar);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = (int)cluster_id_x();

  int gi = 0;
  while (gi + 1 < plan.n_groups && pair >= plan.g[gi].first_pair + plan.g[gi].n_pairs) ++gi;
  const PipeGroup& G = plan.g[gi];
  const bool active = pair < G.first_pair + G.n_pairs;
  const bool chain = G.role == PIPE_ROLE_LAYER;
  const int64_t n_tiles = (M + kTileM - 1) / kTileM;
  const int64_t n_super = (n_tiles + 1) / 2;
  const int64_t s_first = active ? pair - G.first_pair : n_super, s_step = G.n_pairs;

  if (threadIdx.x == 0) {
    const uint32_t full_count = rank == 0 ? 2u : 1u;       // the leader's copy also counts the peer's relay
    mbar_init(smem_u32(&bars->wfull), full_count);
    for (int i = 0; i < kPSlots; ++i) { mbar_init(smem_u32(&bars->full[i]), full_count); mbar_init(smem_u32(&bars->empty[i]), 3); }
    mbar_init(smem_u32(&bars->acc_full), 1);
    mbar_init(smem_u32(&bars->acc_empty), 2 * 16);          // one arrive per epilogue warp of BOTH CTAs
    mbar_init(smem_u32(&bars->fin), 1);
    fence_barrier_init();
  }
  cluster_sync_all();
  if (warp == kWarpMma) tmem_alloc_pair(smem_u32(&bars->tmem_base), 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  float* part = scratch + (size_t)pair * kDwPartialFloats;   // this pair's split-K partial: [256 k][256 n] fp32 + 256 bias sums

  if (warp == kWarpProducer) {
    if (lane < kPipeProdLanes) {
      if (chain) pipe_producer<true>(G, bars, sbase, rank, lane, packed, saved, dz_ws, flags, s_first, s_step, n_super, dbg);
      else pipe_producer<false>(G, bars, sbase, rank, lane, packed, saved, dz_ws, flags, s_first, s_step, n_super, dbg);
    }
  } else if (warp == kWarpMma) {
    if (rank == 0) {
      if (tmem_base != 0u) __trap();                   // 512 of 512 columns: the allocation starts at column 0
      if (chain) pipe_mma<true>(G, bars, sbase, tmem_base, s_first, s_step, n_super, dbg);
      else pipe_mma<false>(G, bars, sbase, tmem_base, s_first, s_step, n_super, dbg);
    } else if (lane == 0) {
      if (chain) pipe_relay<true>(bars, s_first, s_step, n_super);
      else pipe_relay<false>(bars, s_first, s_step, n_super);
    }
  } else if (warp >= kWarpStore + 2) {
    // (only two bias warps; further store warps of the forward / chain launch shape have nothing to do here)
  } else if (warp >= kWarpStore) {
    const int tid64 = (warp - kWarpStore) * 32 + lane;
    float s0, s1;
    if (chain) pipe_bias<true>(G, bars, sbase, tid64, lane, s_first, s_step, n_super, &s0, &s1, dbg);
    else pipe_bias<false>(G, bars, sbase, tid64, lane, s_first, s_step, n_super, &s0, &s1, dbg);
    if (active && 2 * tid64 < (int)G.b_chunks * 8) {
      part[256 * 256 + rank * (int)G.b_chunks * 8 + 2 * tid64] = s0;
      part[256 * 256 + rank * (int)G.b_chunks * 8 + 2 * tid64 + 1] = s1;
    }
  } else {
    // ===== 16 epilogue warps: TMEM lane quarter q = warp % 4, column quarter cq = warp / 4 =====
    const int q = warp & 3, cq = warp >> 2;
    const int r = q * 32 + lane;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    if (chain) {
      const uint32_t acc_empty_leader = mapa_shared(smem_u32(&bars->acc_empty), 0);
      uint32_t it = 0;
      for (int64_t s = s_first; s < n_super; s += s_step, ++it) {
        const int64_t tile = 2 * s + rank;
        const uint32_t* saved_mask =
            reinterpret_cast<const uint32_t*>(saved + (size_t)tile * kSavedTileBytes + (size_t)kSavedPanels * kPanelBytes);
        uint32_t mw[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) mw[j] = __ldg(saved_mask + ((int)G.mask_row * 8 + cq * 2 + j) * 128 + r);
        const long long te0 = (dbg & kDbgTiming) ? clock64() : 0;
        mbar_wait(smem_u32(&bars->acc_full), it & 1u);
        tc_fence_after();
        const long long te1 = (dbg & kDbgTiming) ? clock64() : 0;
        uint32_t acc[2][32];
        tmem_ld32(taddr + cq * 64, acc[0]);
        tmem_ld32(taddr + cq * 64 + 32, acc[1]);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(acc_empty_leader);       // D1 may be overwritten by the next super-tile
        uint8_t* gout = dz_ws + (size_t)((dbg & kDbgWrapDz) ? 2 * (s & 127) + rank : tile) * kDzTileBytes + (size_t)G.out_off;
        if (!(dbg & kDbgNoStore)) {
          // experiments (bits 13..15): math but no store / raw accumulator words stored without math / half of the stores
          const bool x_math_only = (dbg & 8192u) != 0, x_raw = (dbg & 16384u) != 0, x_half = (dbg & 32768u) != 0;
#pragma unroll
          for (int j = 0; j < 2; ++j) {
#pragma unroll
            for (int g8 = 0; g8 < 4; ++g8) {
              if (x_raw) {
                stg128(gout + tcm_offset(r, cq * 8 + j * 4 + g8),
                       make_uint4(acc[j][8 * g8], acc[j][8 * g8 + 1], acc[j][8 * g8 + 2], acc[j][8 * g8 + 3]));
                continue;
              }
              float v[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float a = __uint_as_float(acc[j][8 * g8 + i]);
                v[i] = pipe_mask_bit(mw[j], 8 * g8 + i) ? a : alpha * a;
              }
              if (x_math_only ? (v[0] + v[3] + v[5] + v[7] == 1.2345e-33f) : (x_half ? (g8 & 1) == 0 : true))
              stg128(gout + tcm_offset(r, cq * 8 + j * 4 + g8),
                     make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7])));
              if (dbg >> 16) __nanosleep(dbg >> 16);       // experiment: pace the store burst (ns between a warp's stores)
            }
          }
        }
        if (G.signal_flag >= 0) {
          __syncwarp();
          if (lane == 0) flag_signal(flags + (size_t)s * kPipeFlagsPerSuper + G.signal_flag);
        }
        if ((dbg & kDbgTiming) && s_first == 0 && threadIdx.x == 0 && rank == 0 && it == 8)
          printf("pipe epilogue (layer %d, super-tile %u): waited %lld cycles for the accumulator, then %lld cycles of loads, math and stores\n",
                 (int)G.layer, it, te1 - te0, clock64() - te1);
      }
    }
    // ---- drain this CTA's half of the dW accumulator: rows k = 128 rank + lane, this warp's 64 columns ----
    if (active) {
      mbar_wait(smem_u32(&bars->fin), 0);
      tc_fence_after();
      if (s_first < n_super && !(dbg & kDbgNoDrain)) {
        const int k = (int)rank * 128 + r;
#pragma unroll 1
        for (int c0 = cq * 64; c0 < cq * 64 + 64 && c0 < (int)G.n_dw; c0 += 16) {
          uint32_t a[16];
          tmem_ld16(taddr + 256u + (uint32_t)c0, a);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; i += 4) stg128(part + (size_t)k * 256 + c0 + i, make_uint4(a[i], a[i + 1], a[i + 2], a[i + 3]));
        }
      }
      tc_fence_before();
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                  // the peer's shared memory / TMEM stay valid until both are done
  if (warp == kWarpMma) tmem_dealloc_pair(tmem_base, 512);
}

// Adds the per-pair split-K partials of every group into the flat gradient vector, in pair order (deterministic).
// grid = (n_groups, 64), 256 threads: thread = one float4 of the group's [256][256] block (+ the bias row, block y == 0).
__global__ void __launch_bounds__(256)
mlp_tc_bwd_pipe_reduce_kernel(const __grid_constant__ PipePlan plan, const __grid_constant__ NetGeom g,
                              const float* __restrict__ scratch, int64_t M, float* __restrict__ Gr) {
  const PipeGroup& G = plan.g[blockIdx.x];
  const int64_t n_tiles = (M + kTileM - 1) / kTileM;
  const int64_t n_super = (n_tiles + 1) / 2;
  const int live = (int)(n_super < G.n_pairs ? n_super : G.n_pairs);      // pairs that had a super-tile
  const float* base = scratch + (size_t)G.first_pair * kDwPartialFloats;
  const int e4 = blockIdx.y * 256 + threadIdx.x;                          // float4 index inside [256][64 float4]
  const int k = e4 >> 6, n0 = (e4 & 63) * 4;
  const bool head = G.role == PIPE_ROLE_DW_ONLY;                          // Dense 8 (+ sigma head in column 128), h8 rows
  if (n0 < (int)G.n_dw) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < live; ++s) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(base + (size_t)s * kDwPartialFloats + k * 256 + n0));
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    const float vals[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int n = n0 + i;
      float* tgt = nullptr;
      if (!head) {
        const LayerDesc& L = g.layers[G.dense];
        tgt = Gr + L.w_off + (int64_t)((G.dense == 4 ? g.dx : 0) + k) * L.out + n;
      } else if (n < 128) {
        tgt = Gr + g.layers[8].w_off + (int64_t)k * 128 + n;
      } else if (n == 128) {
        tgt = Gr + g.layers[10].w_off + k;
      }
      if (tgt) *tgt += vals[i];
    }
  }
  if (blockIdx.y == 0 && G.has_bias && threadIdx.x < G.n_dw) {
    float acc = 0.f;
    for (int s = 0; s < live; ++s) acc += __ldg(base + (size_t)s * kDwPartialFloats + 256 * 256 + threadIdx.x);
    const int n = threadIdx.x;
    float* tgt = nullptr;
    if (!head) tgt = Gr + g.layers[G.dense].b_off + n;
    else if (n < 128) tgt = Gr + g.layers[8].b_off + n;
    else if (n == 128) tgt = Gr + g.layers[10].b_off;
    if (tgt) *tgt += acc;
  }
}

// ---- host ---------------------------------------------------------------------------------------------------------------
uint32_t bwd_pack_layer_offset(int layer);      // mlp_tc_bwd.cu

static void fill_layer_group(PipeGroup* G, int l) {
  memset(G, 0, sizeof(*G));
  G->role = PIPE_ROLE_LAYER;
  G->layer = (int16_t)l;
  G->wait_flag = G->signal_flag = -1;
  G->n_chain = 256; G->n_dw = 256; G->b_chunks = 16; G->b_half_chunks = 32;
  G->mask_row = (int16_t)(l - 1);
  G->dense = (int16_t)l;
  G->has_bias = 1;
  G->w_off = (int32_t)bwd_pack_layer_offset(l);
  G->a_off = saved_panel_h(l) * kPanelBytes;
  G->b_off = dz_panel(l + 1) * kPanelBytes;
  G->out_off = dz_panel(l) * kPanelBytes;
}

// Dense 8 + sigma head, h8 rows: dW only (A = h8, B = dZ_L' incl. the sigma column: N = 144, 72 features = 9 chunks per CTA)
static void fill_head_group(PipeGroup* G) {
  memset(G, 0, sizeof(*G));
  G->role = PIPE_ROLE_DW_ONLY;
  G->layer = 8;
  G->wait_flag = G->signal_flag = -1;
  G->n_chain = 256; G->n_dw = 144; G->b_chunks = 9; G->b_half_chunks = kDzChunksL;
  G->dense = 8;
  G->has_bias = 1;
  G->a_off = saved_panel_h(8) * kPanelBytes;
  G->b_off = kDzPanelL * kPanelBytes;
}

static int pipe_launch(const PipePlan& plan, const NetGeom& g, const uint8_t* packed_bwd, const void* saved, uint8_t* dz_ws,
                       int64_t m, float* scratch, uint32_t* flags, float* grads, float alpha, cudaStream_t st) {
  if (device_first_use(2))
    NERF_CUDA(cudaFuncSetAttribute(mlp_tc_bwd_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPAlloc));
  int n_pairs = 0;
  for (int i = 0; i < plan.n_groups; ++i) n_pairs += plan.g[i].n_pairs;
  mlp_tc_bwd_pipe_kernel<<<2 * n_pairs, kThreadsFwd, kPAlloc, st>>>(plan, packed_bwd, (const uint8_t*)saved, dz_ws, m, scratch,
                                                                   flags, alpha, tc_debug_flags());
  NERF_CHECK_LAUNCH();
  mlp_tc_bwd_pipe_reduce_kernel<<<dim3(plan.n_groups, 64), 256, 0, st>>>(plan, g, scratch, m, grads);
  NERF_CHECK_LAUNCH();
  return NERF_OK;
}

// Diagnostic: ONE group over every CTA pair for Dense `layer` (1..7: chain + dW; 8: dW of the h8 rows only), reading
// dZ_{layer+1} from a workspace a previous nerf_mlp_bwd[_dx] filled.  Writes dZ_layer back into the workspace and ADDS the
// weight / bias gradient of that Dense layer to `grads`.
int mlp_tc_bwd_pipe_single(const nerf_net_cfg* cfg, const NetGeom& g, const uint8_t* packed_bwd, const void* saved,
                           int64_t m, void* workspace, int layer, float* grads, cudaStream_t st) {
  PipePlan plan;
  memset(&plan, 0, sizeof(plan));
  plan.n_groups = 1;
  if (layer == 8) fill_head_group(&plan.g[0]);
  else fill_layer_group(&plan.g[0], layer);
  plan.g[0].first_pair = 0;
  plan.g[0].n_pairs = (int16_t)(num_sms() / 2);
  uint8_t* dz_ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023);
  const int64_t tiles4 = ((m + kTileM - 1) / kTileM + 3) / 4 * 4;
  float* scratch = reinterpret_cast<float*>(dz_ws + tiles4 * (int64_t)kDzTileBytes);
  return pipe_launch(plan, g, packed_bwd, saved, dz_ws, m, scratch, nullptr, grads, cfg->leaky_alpha, st);
}

// Diagnostic: the layer groups hi, hi - 1, ..., lo (7 >= hi >= lo >= 1) in ONE launch, the CTA pairs split evenly over them;
// dZ_{hi+1} comes from the workspace, every other dZ is handed from group to group through the ready counters.
int mlp_tc_bwd_pipe_range(const nerf_net_cfg* cfg, const NetGeom& g, const uint8_t* packed_bwd, const void* saved, int64_t m,
                          void* workspace, int hi, int lo, float* grads, cudaStream_t st) {
  PipePlan plan;
  memset(&plan, 0, sizeof(plan));
  const int n_groups = hi - lo + 1, pairs = num_sms() / 2;
  plan.n_groups = n_groups;
  int first = 0;
  for (int i = 0; i < n_groups; ++i) {
    const int l = hi - i;
    PipeGroup& G = plan.g[i];
    fill_layer_group(&G, l);
    G.first_pair = (int16_t)first;
    G.n_pairs = (int16_t)(pairs / n_groups + (i < pairs % n_groups ? 1 : 0));
    first += G.n_pairs;
    G.wait_flag = (int16_t)(i == 0 ? -1 : l + 1);
    G.signal_flag = (int16_t)(i == n_groups - 1 ? -1 : l);
  }
  uint8_t* dz_ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023);
  const int64_t tiles4 = ((m + kTileM - 1) / kTileM + 3) / 4 * 4;
  float* scratch = reinterpret_cast<float*>(dz_ws + tiles4 * (int64_t)kDzTileBytes);
  uint32_t* flags = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(scratch) + kDwScratchBytes);
  NERF_CUDA(cudaMemsetAsync(flags, 0, (size_t)(tiles4 / 2) * kPipeFlagsPerSuper * sizeof(uint32_t), st));
  return pipe_launch(plan, g, packed_bwd, saved, dz_ws, m, scratch, flags, grads, cfg->leaky_alpha, st);
}

}  // namespace nerf

using namespace nerf;

extern "C" int nerf_debug_bwd_pipe_layer(const nerf_net_cfg* cfg, const void* packed, const void* saved, int64_t m,
                                         void* workspace, int32_t layer, float* grads, void* stream) {
  NetGeom g;
  TcPlan fplan;
  NERF_CHECK_ARG(make_geom(cfg, &g) && make_plan(g, &fplan) && g.view, "the pipe prototype handles the view-direction network");
  NERF_CHECK_ARG(packed && saved && workspace && grads, "null pointer");
  const int hi = layer / 100, lo = layer % 100;
  NERF_CHECK_ARG(m > 0 && (hi == 0 ? (layer >= 1 && layer <= 8) : (hi <= 7 && lo >= 1 && hi > lo)),
                 "layer must be 1..8, or 100 hi + lo for the groups hi..lo of one launch");
  const uint8_t* packed_bwd = (const uint8_t*)packed + ((fplan.total_bytes + 1023u) & ~1023u);
  if (hi) return mlp_tc_bwd_pipe_range(cfg, g, packed_bwd, saved, m, workspace, hi, lo, grads, (cudaStream_t)stream);
  return mlp_tc_bwd_pipe_single(cfg, g, packed_bwd, saved, m, workspace, layer, grads, (cudaStream_t)stream);
}
