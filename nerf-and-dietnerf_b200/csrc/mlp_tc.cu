// K2 in NERF_MODE_BF16: the coarse/fine NeRF MLP (src/NeRF.py:290-340, evaluated by model_predict,
// src/UtilsNeuralRadianceField.py:214-234) as ONE persistent tcgen05/TMEM kernel per call.
//
// Design (DESIGN.md "K2"):
//   * one CTA per SM, 320 threads: warps 0-3 = epilogue of row-tile A, warps 4-7 = epilogue of row-tile B,
//     warp 8 = weight producer (1-D bulk async copies, TMA engine), warp 9 = tcgen05.mma issuer (one elected lane);
//   * a CTA owns TWO 128-sample tiles at a time; both consume the same streamed weight chunk (32 KB, [N][64] bf16,
//     128-byte swizzle, pre-swizzled in HBM by nerf_pack_weights) so every weight byte fetched from L2 feeds 256 rows;
//   * activations never leave the SM: the fp32 accumulator (128 x 256) of each tile lives in TMEM (2 x 256 columns),
//     the epilogue warps read it with tcgen05.ld, add bias, apply LeakyReLU, round to bf16 and write the next layer's
//     A operand back into the SAME swizzled shared-memory panels (all MMAs of the layer have completed by then);
//   * the 33-d xyz encoding and the 24-d view encoding share one 64-column input panel (cols 0..32 xyz, 40..63 view)
//     that feeds layer 0, the skip connection of layer 4 and the view branch of layer 8 (zero weight rows elsewhere);
//   * layer 8 is N = 144: 128 last-hidden units + the sigma head (column 128); the 128->3 rgb head is evaluated on
//     CUDA cores from the fp32 registers of the last epilogue;
//   * training mode stores each layer's bf16 activation tile to HBM with bulk S2G copies in the swizzled tile-panel
//     format the backward kernels load back verbatim.
#include "mlp_tc.cuh"

namespace nerf {

// ---- weight packing -----------------------------------------------------------------------------------------------------
// One thread per bf16 element of every chunk, plus the fp32 tail (biases, rgb head).
__global__ void pack_weights_kernel(const __grid_constant__ TcPlan plan, NetGeom g, const float* __restrict__ P,
                                    uint8_t* __restrict__ packed) {
  const int chunk = blockIdx.y;
  if (chunk < plan.n_chunks) {
    // find the layer of this chunk
    int layer = 0;
    while (!(chunk >= plan.layer_first[layer] && chunk < plan.layer_first[layer] + plan.layer_nchunks[layer])) ++layer;
    const int n_rows = plan.layer_n[layer];
    const int e = blockIdx.x * blockDim.x + threadIdx.x;  // element index within [n_rows][64]
    if (e >= n_rows * 64) return;
    const int n = e >> 6, k = e & 63;
    const int src = plan.a_src[chunk];
    // source Dense layer and output column
    int dense = layer, col = n;
    bool valid = true;
    if (layer == 8) {
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
      else if (layer == 8) { if (k >= kInpViewCol && k < kInpViewCol + g.dv) row = g.hidden + (k - kInpViewCol); }
    }
    float v = 0.f;
    if (valid && row >= 0) {
      const LayerDesc& L = g.layers[dense];
      v = P[L.w_off + (int64_t)row * L.out + col];
    }
    *reinterpret_cast<__nv_bfloat16*>(packed + plan.chunk_off[chunk] + panel_offset(n, k)) = __float2bfloat16_rn(v);
  } else {
    // fp32 tail
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    float* bias = reinterpret_cast<float*>(packed + plan.bias_off);
    if (e < 9 * 256) {
      int l = e >> 8, n = e & 255;
      float v = 0.f;
      if (l < 8) v = P[g.layers[l].b_off + n];
      else if (n < 128) v = P[g.layers[8].b_off + n];
      else if (n == 128) v = P[g.layers[10].b_off];
      bias[e] = v;
    }
    if (e < 128) {
      const LayerDesc& L = g.layers[9];
      reinterpret_cast<float4*>(packed + plan.w_rgb_off)[e] =
          make_float4(P[L.w_off + e * 3 + 0], P[L.w_off + e * 3 + 1], P[L.w_off + e * 3 + 2], 0.f);
    }
    if (e < 4) reinterpret_cast<float*>(packed + plan.b_rgb_off)[e] = e < 3 ? P[g.layers[9].b_off + e] : 0.f;
  }
}

// ---- forward kernel --------------------------------------------------------------------------------------------------------
struct FwdBars {
  uint64_t full[kStages], empty[kStages], act_ready[2], acc_full[2];
  uint32_t tmem_base;
};

template <bool kSave>
__global__ void __launch_bounds__(kThreadsFwd, 1)
mlp_tc_fwd_kernel(const __grid_constant__ TcPlan plan, const uint8_t* __restrict__ packed,
                  const float* __restrict__ xyz_enc, const float* __restrict__ view_enc, int dx, int dv, int64_t M,
                  float* __restrict__ out4, uint8_t* __restrict__ saved, float alpha) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  FwdBars* bars = reinterpret_cast<FwdBars*>(smem + kSmemBar);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int64_t n_tiles = (M + kTileM - 1) / kTileM;
  const int64_t n_pairs = (n_tiles + 1) / 2;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(smem_u32(&bars->full[s]), 1); mbar_init(smem_u32(&bars->empty[s]), 1); }
    for (int t = 0; t < 2; ++t) { mbar_init(smem_u32(&bars->act_ready[t]), 128); mbar_init(smem_u32(&bars->acc_full[t]), 1); }
    fence_barrier_init();
  }
  if (warp == 9) tmem_alloc(smem_u32(&bars->tmem_base), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 8) {
    // ===== weight producer =====
    if (lane == 0) {
      uint32_t g = 0;
      for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
        for (int c = 0; c < plan.n_chunks; ++c, ++g) {
          const uint32_t s = g % kStages, ph = (g / kStages) & 1u;
          mbar_wait(smem_u32(&bars->empty[s]), ph ^ 1u);
          mbar_arrive_expect_tx(smem_u32(&bars->full[s]), plan.chunk_bytes[c]);
          bulk_g2s(sbase + kSmemStage + s * kStageBytes, packed + plan.chunk_off[c], plan.chunk_bytes[c],
                   smem_u32(&bars->full[s]));
        }
      }
    }
  } else if (warp == 9) {
    // ===== MMA issuer =====
    if (lane == 0) {
      uint32_t g = 0, act_cnt = 0;
      for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
        for (int l = 0; l < plan.n_layers; ++l) {
          const uint32_t idesc = make_idesc(plan.layer_n[l]);
          const int first = plan.layer_first[l], nch = plan.layer_nchunks[l];
          for (int ci = 0; ci < nch; ++ci, ++g) {
            const int c = first + ci;
            const uint32_t s = g % kStages, ph = (g / kStages) & 1u;
            mbar_wait(smem_u32(&bars->full[s]), ph);
            const uint32_t b_addr = sbase + kSmemStage + s * kStageBytes;
            const int src = plan.a_src[c];
#pragma unroll
            for (int t = 0; t < 2; ++t) {
              if (ci == 0) mbar_wait(smem_u32(&bars->act_ready[t]), act_cnt & 1u);
              tc_fence_after();
              const uint32_t a_addr = (src < 4) ? sbase + kSmemAct + (t * kActPanels + src) * kPanelBytes
                                                : sbase + kSmemInp + t * kPanelBytes;
              const uint32_t d_tmem = tmem_base + (uint32_t)t * 256u;
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(d_tmem, make_desc_kmajor(a_addr + k * 32), make_desc_kmajor(b_addr + k * 32), idesc,
                          (ci > 0 || k > 0) ? 1u : 0u);
              if (ci == nch - 1) umma_commit(smem_u32(&bars->acc_full[t]));
            }
            umma_commit(smem_u32(&bars->empty[s]));
          }
          ++act_cnt;
        }
      }
    }
  } else {
    // ===== epilogue warps: tile t = warp / 4, TMEM lane quarter q = warp % 4 =====
    const int t = warp >> 2, q = warp & 3;
    const int r = q * 32 + lane;                       // row inside the tile == TMEM lane
    const int gtid = threadIdx.x & 127;                // thread index inside this tile's group
    const int bar_id = 1 + t;
    uint8_t* act = smem + kSmemAct + t * kActPanels * kPanelBytes;
    uint8_t* inp = smem + kSmemInp + t * kPanelBytes;
    const uint32_t act_u32 = sbase + kSmemAct + t * kActPanels * kPanelBytes;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)t * 256u;
    const float* bias_all = reinterpret_cast<const float*>(packed + plan.bias_off);
    const float4* w_rgb = reinterpret_cast<const float4*>(packed + plan.w_rgb_off);
    const float* b_rgb = reinterpret_cast<const float*>(packed + plan.b_rgb_off);
    uint32_t acc_cnt = 0;
    for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
      const int64_t tile = pair * 2 + t;
      const int64_t row = tile * kTileM + r;
      const bool row_ok = row < M;
      // ---- prologue: build the input panel row (bf16): cols [0,dx) xyz, [40,40+dv) view, zeros elsewhere
      {
        float v[64];
#pragma unroll
        for (int i = 0; i < 64; ++i) v[i] = 0.f;
        if (row_ok) {
          const float* xr = xyz_enc + row * dx;
#pragma unroll
          for (int i = 0; i < kInpViewCol; ++i) if (i < dx) v[i] = __ldg(xr + i);
          const float* vr = view_enc + row * dv;
#pragma unroll
          for (int i = 0; i < 64 - kInpViewCol; ++i) if (i < dv) v[kInpViewCol + i] = __ldg(vr + i);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint4 pk = make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                                pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
          *reinterpret_cast<uint4*>(inp + r * 128 + ((j ^ (r & 7)) << 4)) = pk;
        }
      }
      fence_proxy_async();
      mbar_arrive(smem_u32(&bars->act_ready[t]));
      uint8_t* saved_tile = kSave ? saved + (size_t)tile * kSavedTileBytes : nullptr;
      uint32_t* saved_mask = reinterpret_cast<uint32_t*>(saved_tile + (size_t)kSavedPanels * kPanelBytes);
      if (kSave) {
        named_bar_sync(bar_id, 128);
        if (gtid == 0) {
          bulk_s2g(saved_tile, sbase + kSmemInp + t * kPanelBytes, kPanelBytes);
          bulk_commit();
        }
      }

      for (int l = 0; l < plan.n_layers; ++l) {
        mbar_wait(smem_u32(&bars->acc_full[t]), acc_cnt & 1u);
        ++acc_cnt;
        tc_fence_after();
        if (kSave) {
          // the previous layer's bulk stores still read the panels this epilogue overwrites
          if (gtid == 0) bulk_wait_read0();
          named_bar_sync(bar_id, 128);
        }
        const float* bias = bias_all + l * 256;
        if (l < 8) {
#pragma unroll 1
          for (int c0 = 0; c0 < 256; c0 += 32) {
            uint32_t acc[32];
            tmem_ld32(taddr + c0, acc);
            tmem_ld_wait();
            uint8_t* prow = act + (c0 >> 6) * kPanelBytes + r * 128;
            uint32_t mword = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + c0 + 8 * j));
              float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + c0 + 8 * j + 4));
              float f0 = leaky(__uint_as_float(acc[8 * j + 0]) + b0.x, alpha);
              float f1 = leaky(__uint_as_float(acc[8 * j + 1]) + b0.y, alpha);
              float f2 = leaky(__uint_as_float(acc[8 * j + 2]) + b0.z, alpha);
              float f3 = leaky(__uint_as_float(acc[8 * j + 3]) + b0.w, alpha);
              float f4 = leaky(__uint_as_float(acc[8 * j + 4]) + b1.x, alpha);
              float f5 = leaky(__uint_as_float(acc[8 * j + 5]) + b1.y, alpha);
              float f6 = leaky(__uint_as_float(acc[8 * j + 6]) + b1.z, alpha);
              float f7 = leaky(__uint_as_float(acc[8 * j + 7]) + b1.w, alpha);
              uint4 pk = make_uint4(pack_bf16x2(f0, f1), pack_bf16x2(f2, f3), pack_bf16x2(f4, f5), pack_bf16x2(f6, f7));
              const int chunk16 = ((c0 & 63) >> 3) + j;
              *reinterpret_cast<uint4*>(prow + ((chunk16 ^ (r & 7)) << 4)) = pk;
              if (kSave)
                mword |= ((f0 > 0.f) | ((f1 > 0.f) << 1) | ((f2 > 0.f) << 2) | ((f3 > 0.f) << 3) | ((f4 > 0.f) << 4) |
                        ((f5 > 0.f) << 5) | ((f6 > 0.f) << 6) | ((f7 > 0.f) << 7))
                       << (8 * j);
            }
            if (kSave) saved_mask[(l * 8 + (c0 >> 5)) * 128 + r] = mword;
          }
          tc_fence_before();
          fence_proxy_async();
          mbar_arrive(smem_u32(&bars->act_ready[t]));
          if (kSave) {
            named_bar_sync(bar_id, 128);
            if (gtid == 0) {
              bulk_s2g(saved_tile + (size_t)saved_panel_h(l + 1) * kPanelBytes, act_u32, kActPanels * kPanelBytes);
              bulk_commit();
            }
          }
        } else {
          // last layer: cols 0..127 = last hidden (LeakyReLU), col 128 = sigma (linear); rgb head on CUDA cores
          float rr = __ldg(b_rgb + 0), gg = __ldg(b_rgb + 1), bb = __ldg(b_rgb + 2);
#pragma unroll 1
          for (int c0 = 0; c0 < 128; c0 += 32) {
            uint32_t acc[32];
            tmem_ld32(taddr + c0, acc);
            tmem_ld_wait();
            uint8_t* prow = act + (c0 >> 6) * kPanelBytes + r * 128;
            float f[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              f[i] = leaky(__uint_as_float(acc[i]) + __ldg(bias + c0 + i), alpha);
              float4 w = __ldg(w_rgb + c0 + i);
              rr = fmaf(f[i], w.x, rr);
              gg = fmaf(f[i], w.y, gg);
              bb = fmaf(f[i], w.z, bb);
            }
            if (kSave) {
              uint32_t mword = 0;
#pragma unroll
              for (int i = 0; i < 32; ++i) mword |= (uint32_t)(f[i] > 0.f) << i;
              saved_mask[(8 * 8 + (c0 >> 5)) * 128 + r] = mword;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint4 pk = make_uint4(pack_bf16x2(f[8 * j], f[8 * j + 1]), pack_bf16x2(f[8 * j + 2], f[8 * j + 3]),
                                      pack_bf16x2(f[8 * j + 4], f[8 * j + 5]), pack_bf16x2(f[8 * j + 6], f[8 * j + 7]));
                const int chunk16 = ((c0 & 63) >> 3) + j;
                *reinterpret_cast<uint4*>(prow + ((chunk16 ^ (r & 7)) << 4)) = pk;
              }
            }
          }
          uint32_t sg[16];
          tmem_ld16(taddr + 128, sg);
          tmem_ld_wait();
          const float sigma = __uint_as_float(sg[0]) + __ldg(bias + 128);
          if (row_ok) reinterpret_cast<float4*>(out4)[row] = make_float4(rr, gg, bb, sigma);
          tc_fence_before();
          if (kSave) {
            fence_proxy_async();
            named_bar_sync(bar_id, 128);
            if (gtid == 0) {
              bulk_s2g(saved_tile + (size_t)kSavedPanelHL * kPanelBytes, act_u32, 2 * kPanelBytes);
              bulk_commit();
            }
          }
        }
      }
    }
    if (kSave && gtid == 0) bulk_wait0();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem_base, 512);
}

// ---- host side ----------------------------------------------------------------------------------------------------------------
int64_t mlp_tc_saved_bytes(const NetGeom& g, int64_t m) {
  int64_t tiles = (m + kTileM - 1) / kTileM;
  tiles = (tiles + 1) / 2 * 2;
  return tiles * (int64_t)kSavedTileBytes;
}

int64_t mlp_tc_workspace_bytes(const NetGeom& g, int64_t m, int backward) {
  (void)g;
  if (!backward) return 256;
  int64_t tiles = (m + kTileM - 1) / kTileM;
  tiles = (tiles + 1) / 2 * 2;
  return tiles * (int64_t)kDzTileBytes + 1024;
}

int mlp_tc_fwd(const nerf_net_cfg* cfg, const NetGeom& g, const float* params, const void* packed, const float* xyz_enc,
               const float* view_enc, int64_t m, float* out4, void* saved, void* workspace, cudaStream_t st) {
  (void)params; (void)workspace;
  TcPlan plan;
  if (!make_plan(g, &plan)) {
    set_error("NERF_MODE_BF16 supports hidden=256, last_hidden=128, n_angles in {1,2}, xyz width <= 40, view width <= 24");
    return NERF_E_UNSUPPORTED;
  }
  static bool attr_set = false;
  if (!attr_set) {
    NERF_CUDA(cudaFuncSetAttribute(mlp_tc_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemAlloc));
    NERF_CUDA(cudaFuncSetAttribute(mlp_tc_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemAlloc));
    attr_set = true;
  }
  int64_t n_pairs = ((m + kTileM - 1) / kTileM + 1) / 2;
  int grid = (int)(n_pairs < kNumSMs ? n_pairs : kNumSMs);
  if (saved)
    mlp_tc_fwd_kernel<true><<<grid, kThreadsFwd, kSmemAlloc, st>>>(plan, (const uint8_t*)packed, xyz_enc, view_enc, g.dx,
                                                                  g.dv, m, out4, (uint8_t*)saved, cfg->leaky_alpha);
  else
    mlp_tc_fwd_kernel<false><<<grid, kThreadsFwd, kSmemAlloc, st>>>(plan, (const uint8_t*)packed, xyz_enc, view_enc, g.dx,
                                                                   g.dv, m, out4, nullptr, cfg->leaky_alpha);
  NERF_CHECK_LAUNCH();
  return NERF_OK;
}

}  // namespace nerf

using namespace nerf;

extern "C" {

int64_t nerf_packed_bytes(const nerf_net_cfg* cfg) {
  NetGeom g;
  TcPlan plan;
  if (!make_geom(cfg, &g)) { set_error("nerf_packed_bytes: bad net config"); return NERF_E_ARG; }
  if (!make_plan(g, &plan)) { set_error("nerf_packed_bytes: config not supported by NERF_MODE_BF16"); return NERF_E_UNSUPPORTED; }
  return (int64_t)((plan.total_bytes + 1023u) & ~1023u) + bwd_pack_bytes();
}

int nerf_pack_weights(const nerf_net_cfg* cfg, const float* params, void* packed, void* stream) {
  NetGeom g;
  TcPlan plan;
  NERF_CHECK_ARG(make_geom(cfg, &g), "bad net config");
  NERF_CHECK_ARG(params && packed, "null pointer");
  if (!make_plan(g, &plan)) { set_error("nerf_pack_weights: config not supported by NERF_MODE_BF16"); return NERF_E_UNSUPPORTED; }
  dim3 grid((256 * 64 + 255) / 256, plan.n_chunks + 1);
  pack_weights_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(plan, g, params, (uint8_t*)packed);
  NERF_CHECK_LAUNCH();
  return bwd_pack_weights(g, params, (uint8_t*)packed + ((plan.total_bytes + 1023u) & ~1023u), (cudaStream_t)stream);
}

}  // extern "C"
