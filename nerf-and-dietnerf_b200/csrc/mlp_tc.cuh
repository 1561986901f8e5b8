// Shared definitions of the tensor-core MLP kernels (forward: mlp_tc.cu, backward: mlp_tc_bwd.cu).
#pragma once
#include <string.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace nerf {

using namespace tc;

constexpr int kTileM = 128;
constexpr int kPanelBytes = 128 * 128;            // [128 rows][64 bf16]
constexpr int kActPanels = 4;                     // 256 features
// The forward kernel runs on CTA PAIRS (cluster of 2, tcgen05 cta_group::2): one MMA covers 256 rows (128 per CTA) and each
// CTA streams only ITS half of every weight chunk ([<=128][64] bf16 = 16 KB): half the L2 traffic, half the B-operand
// shared-memory reads, and a 4-deep ring.  Measured motivation (NERF_TC_DEBUG=256 trace, tools/sm_probe.cu): with a
// 2 x 32 KB single-CTA ring every 512-cycle chunk of MMAs waited ~300 cycles for its weights (L2 latency ~700 cycles).
constexpr int kStageBytes = 16384;                // this CTA's half of one weight chunk
constexpr int kStages = 4;                        // forward ring
#ifndef NERF_PROD_LANES
#define NERF_PROD_LANES 1
#endif
// Producer lanes issuing the weight copies side by side (power of two, < kStages).  One thread gets ~3.5 M bulk copies per
// second out of the TMA path (tools/l2_bw_probe.cu); two lanes were tried here and change nothing (1393 TFLOP/s either
// way): at one 16 KB copy per 512 MMA cycles these kernels are not copy-issue-bound.  The layer-pipelined backward, which
// needs six 32 KB copies per 4096 MMA cycles from L2 and HBM, is where it matters.
constexpr int kProdLanes = NERF_PROD_LANES;
constexpr int kInpViewCol = 40;                   // first view-encoding column of the input panel
constexpr int kInpOneCol = 38;                    // input-panel columns 38, 39 hold the constant 1: the weight rows they
                                                  // meet carry bf16(b) and bf16(b - bf16(b)), so the bias rides in the MMA
constexpr int kBiasSlabBytes = 256 * 32;          // [256][16] bf16, un-swizzled K-major: the K = 16 slice (input-panel
                                                  // columns 32..47) that adds the bias of a layer without input-panel chunk
constexpr int kMaxChunks = 56;
// -DNERF_TC_TRACE=1 compiles the in-kernel handshake timing / event trace in (NERF_TC_DEBUG=256 then prints it); off by
// default because its 64-bit counters cost registers in kernels that run at the 96-register limit
#ifndef NERF_TC_TRACE
#define NERF_TC_TRACE 0
#endif
constexpr bool kTcTrace = NERF_TC_TRACE != 0;
#ifndef NERF_TMEM_BUFFERS
#define NERF_TMEM_BUFFERS 1
#endif
constexpr int kTmemBuffers = NERF_TMEM_BUFFERS;   // 2: tcgen05.ld of column group cc+1 in flight while cc is processed
constexpr int kEpiWarps = 16;                     // 8 per tile: 4 TMEM lane quarters x 2 column halves
constexpr int kEpiThreadsPerTile = 256;
constexpr int kWarpProducer = 16, kWarpMma = 17;
constexpr int kWarpStore = 18;                    // warps 18 ...: train-mode copy of tile 0 / 1's panels to HBM (RBCM)
#ifndef NERF_STORE_WARPS
#define NERF_STORE_WARPS 2
#endif
constexpr int kStoreWarps = NERF_STORE_WARPS;     // every store warp works on every copy (32 / kStoreWarps column chunks each)
constexpr int kThreadsFwd = (18 + kStoreWarps) * 32;   // 20 warps = 5 per SM sub-partition: the register cap stays at 96
// The forward that saves activations in the fp16 mode converts every value to bf16 on its way out (the backward is the bf16
// one): 12 more instructions per 16 bytes, and two store warps no longer keep up (771 vs 918 TFLOP/s at 524 k rows).  That
// variant runs four store warps (22 warps -> 80 registers, 24 bytes of spills): 871 TFLOP/s.  Every other variant keeps
// two: with bf16 operands four change nothing, and the chain kernel loses 5 % to the spills of the smaller register file
// (three store warps do not help: 21 warps put six on one SM sub-partition, the same 80-register cap as 22).
#ifndef NERF_FWD_HALF_STORE_WARPS
#define NERF_FWD_HALF_STORE_WARPS 4
#endif
__host__ __device__ constexpr int fwd_store_warps(bool save, bool half) { return save && half ? NERF_FWD_HALF_STORE_WARPS : kStoreWarps; }
__host__ __device__ constexpr int fwd_threads(bool save, bool half) { return (18 + fwd_store_warps(save, half)) * 32; }
// Saved activations of one 128-row tile (forward -> backward), bf16:
//   block 0            input panel (xyz | view encodings), [128 rows][64 cols] in the 128-byte-swizzled smem layout (16 KB)
//   blocks h_1 .. h_9  64 KB each, "row-block chunk-major" (RBCM): [row half 0/1][16-byte column chunk j][row 0..63][8 cols].
//                      (nine slots: the xyz-only network of src/NeRF.py:248-288 has nine 256-wide hidden layers; the
//                      view-direction network leaves slot 9 unwritten -- address space, not traffic)
//                      An epilogue warp (32 consecutive rows, one chunk) stores 512 contiguous bytes straight from its
//                      registers, and a 64-row K-slab of the block is ONE contiguous 32 KB bulk copy that lands in shared
//                      memory as the canonical un-swizzled MN-major UMMA operand of the dW kernel (core matrix = 8 rows x
//                      16 B, SBO = 1024 B between column chunks, LBO = 128 B between 8-row groups)
//   block h_L          last hidden (128 cols), RBCM with 16 chunks (32 KB)
//   then uint32 sign masks [10 rows][8 words][128 rows]: bit i of word w = (activation[32 w + i] > 0); row l - 1 = h_l,
//   row kMaskRowHL = the last hidden layer
constexpr int kHiddenSlots = 9;
constexpr int kSavedPanels = 1 + kHiddenSlots * kActPanels + 2;
constexpr int kMaskRowHL = kHiddenSlots;
constexpr int kSavedMaskBytes = (kHiddenSlots + 1) * 8 * 128 * 4;
constexpr int kSavedTileBytes = kSavedPanels * kPanelBytes + kSavedMaskBytes;
__host__ __device__ constexpr int saved_panel_h(int l) { return 1 + (l - 1) * kActPanels; }  // l = 1..9
constexpr int kSavedPanelHL = 1 + kHiddenSlots * kActPanels;
// byte offset of (row r, 16-byte chunk j) inside an RBCM block with n_chunks column chunks
__host__ __device__ constexpr uint32_t rbcm_offset(int r, int j, int n_chunks) {
  return (uint32_t)(r >> 6) * (uint32_t)(n_chunks * 1024) + (uint32_t)j * 1024u + (uint32_t)(r & 63) * 16u;
}
// "Tile chunk-major" (TCM) variant of the same blocks, used by the layer-pipelined backward: [16-byte column chunk j][row
// 0..127][8 cols].  A warp still stores 512 contiguous bytes, a 128-row tile of 64 features is one contiguous 16 KB piece
// with a uniform 128-byte stride between 8-row groups over all 128 rows (what a K-major UMMA A operand needs), and a
// feature half of the tile (16 chunks) is one contiguous 32 KB piece = the un-swizzled MN-major operand (SBO = 2048).
__host__ __device__ constexpr uint32_t tcm_offset(int r, int j) { return (uint32_t)j * 2048u + (uint32_t)r * 16u; }
// byte offset of (row r, chunk j) in a block with n_chunks chunks: RBCM (tcm == false) or TCM
__host__ __device__ constexpr uint32_t block_offset(bool tcm, int r, int j, int n_chunks) {
  return tcm ? tcm_offset(r, j) : rbcm_offset(r, j, n_chunks);
}

// Backward workspace of one tile (chain kernel -> dW kernel), RBCM blocks: dZ_1..dZ_9 (32 chunks each), dZ_L' (24
// chunks: 0..15 = dZ_L, chunk 16 = [d sigma, 0 ...], 17 zero, 18..23 unused), dOut (8 chunks: chunk 0 = d_out4, 1 zero)
constexpr int kDzPanels = kHiddenSlots * kActPanels + 3 + 1;
constexpr int kDzTileBytes = kDzPanels * kPanelBytes;
__host__ __device__ constexpr int dz_panel(int l) { return (l - 1) * kActPanels; }            // l = 1..9
constexpr int kDzPanelL = kHiddenSlots * kActPanels;
constexpr int kDzPanelOut = kHiddenSlots * kActPanels + 3;
constexpr int kDzChunksL = 24, kDzChunksOut = 8;

// shared memory map (offsets from a 1024-aligned base)
constexpr int kSmemAct = 0;                                          // [2 tiles][4 panels]
constexpr int kSmemInp = kSmemAct + 2 * kActPanels * kPanelBytes;    // [2 tiles]
constexpr int kSmemStage = kSmemInp + 2 * kPanelBytes;               // [kStages]
constexpr int kSmemBar = kSmemStage + kStages * kStageBytes;
constexpr int kSmemWrgb = kSmemBar + 192;                            // float4 [128] rgb-head kernel rows + float4 rgb bias
constexpr int kSmemWsig = kSmemWrgb + 129 * 16;                      // xyz-only network: 16-bit sigma-head kernel [256], fp32 bias
constexpr int kSmemTotal = kSmemWsig + 256 * 2 + 16;
// no alignment slack: the kernels check that the dynamic shared memory window is 1024-byte aligned and trap otherwise
constexpr int kSmemAlloc = kSmemTotal;
static_assert(kSmemAlloc <= 232448, "forward kernel exceeds the 227 KB shared-memory limit");

struct TcPlan {
  uint32_t chunk_off[kMaxChunks];    // byte offset of the chunk in the packed buffer
  uint32_t chunk_bytes[kMaxChunks];
  int8_t a_src[kMaxChunks];          // 0..3 = activation panel, 4 = input panel, 5 = bias slab (input-panel K slice 32..47)
  int8_t layer_first[12], layer_nchunks[12];
  int16_t layer_n[12];               // UMMA N of the layer
  int32_t n_layers, n_chunks;
  int32_t xyz_only;                  // 1: the xyz-only network (ten layers, sigma head off h8)
  uint32_t w_rgb_off;                // fp32 float4 [128] = (Wrgb[j][0], Wrgb[j][1], Wrgb[j][2], 0), then float4 (b_rgb, 0)
  uint32_t w_sig_off;                // xyz-only: 16-bit sigma-head kernel [256] (operand format of the pack), fp32 bias
  uint32_t total_bytes;
};

// Forward schedule of the two networks of the reference (MMA layers; every layer ends with the chunk that carries its
// bias: the input-panel chunk or the bias slab):
//   view network (src/NeRF.py:290-340), 9 layers: 0 inp | 1-3 | 4 (+inp: skip) | 5-7 | 8 = [h8 ; view] -> last hidden (128) +
//       sigma head in column 128 (N = 144, + inp chunk for the view columns); rgb head on CUDA cores in the last epilogue;
//   xyz-only network (:248-288), 10 layers: 0 inp | 1-3 | 4 (+inp) | 5-7 | 8 = h8 -> h9 (256) | 9 = h9 -> last hidden (128);
//       the sigma head reads h8: a dot product in the epilogue of layer 7; rgb head as above.
__host__ __device__ constexpr int tc_n_layers(bool xyz) { return xyz ? 10 : 9; }
__host__ __device__ constexpr uint32_t tc_layer_rows(bool xyz, int l) { return xyz ? (l == 9 ? 128u : 256u) : (l == 8 ? 144u : 256u); }
__host__ __device__ constexpr bool tc_layer_slab(bool xyz, int l) { return xyz ? !(l == 0 || l == 4) : !(l == 0 || l == 4 || l == 8); }

inline bool make_plan(const NetGeom& g, TcPlan* p) {
  if (g.hidden != 256 || g.last_hidden != 128 || g.dx > kInpOneCol || g.dv > 64 - kInpViewCol) return false;
  const bool xyz = !g.view;
  memset(p, 0, sizeof(*p));
  int c = 0;
  uint32_t off = 0;
  auto add = [&](int layer, int n, int src) {
    const uint32_t bytes = src == 5 ? (uint32_t)kBiasSlabBytes : (uint32_t)n * 128u;
    p->chunk_off[c] = off;
    p->chunk_bytes[c] = bytes;
    p->a_src[c] = (int8_t)src;
    off += bytes;
    if (p->layer_nchunks[layer] == 0) p->layer_first[layer] = (int8_t)c;
    p->layer_nchunks[layer]++;
    p->layer_n[layer] = (int16_t)n;
    ++c;
  };
  const int nl = tc_n_layers(xyz);
  for (int l = 0; l < nl; ++l) {
    const int n = (int)tc_layer_rows(xyz, l);
    if (l > 0) for (int k = 0; k < 4; ++k) add(l, n, k);
    add(l, n, tc_layer_slab(xyz, l) ? 5 : 4);
  }
  p->n_layers = nl;
  p->n_chunks = c;
  p->xyz_only = xyz ? 1 : 0;
  p->w_rgb_off = off;       off += 129 * 16;
  p->w_sig_off = off;       off += 256 * 2 + 16;     // xyz-only: 16-bit sigma-head kernel [256], then its fp32 bias
  p->total_bytes = off;
  return true;
}


__device__ __forceinline__ float leaky(float v, float alpha) { return fmaxf(v, 0.f) + alpha * fminf(v, 0.f); }

// Bottleneck-decomposition switches (env NERF_TC_DEBUG, profiling only; results are WRONG when any is set):
enum : uint32_t {
  kDbgNoStore = 1u,        // forward / chain: no activation, mask or dZ stores to HBM
  kDbgNoMma = 2u,          // forward / chain / dW: the MMA issuer only commits
  kDbgNoWeightCopy = 4u,   // forward / chain: the producer arrives without copying
  kDbgNoEpi = 8u,          // forward / chain: hidden-layer epilogues skip the TMEM loads, math and stores
  kDbgNoDrain = 16u,       // dW: no accumulator drain
  kDbgNoBiasSum = 32u,     // dW: no bias-gradient column sums
  kDbgNoChain = 64u,       // backward: skip the dX chain launch
  kDbgNoDw = 128u,         // backward: skip the dW launch
  kDbgNoRelay = 512u,      // forward (pair mode): the leader does not wait for the peer's weight halves (racy)
  kDbgStoreL2 = 1024u,     // forward: the activation copies land in a 64-tile window (L2-resident): isolates HBM from the SM store port
  kDbgTiming = 256u,       // dW: every CTA prints its cycle count
  kDbgWrapDz = 2048u,      // pipe: dZ loads / stores hit super-tile (s mod 128): the 38 MB window stays in L2 (what the
                           // pipeline's hand-over looks like to a stage that is timed alone)
  kDbgWrapSaved = 4096u,   // pipe: the saved activations wrap as well (everything from L2)
};
uint32_t tc_debug_flags();


// dW split-K partials (backward workspace, after the dZ tiles): one 256 x 256 fp32 block + 256 bias sums per dW CTA
constexpr int64_t kDwPartialFloats = 256 * 256 + 256;
constexpr int64_t kDwScratchBytes = (int64_t)kMaxSMs * kDwPartialFloats * 4;

// bytes of the per-(tile, block) dZ ready counters at the end of the backward workspace (overlapped mode)
int64_t mlp_tc_bwd_flag_bytes(int64_t m);

// backward half of the 16-bit weight pack (defined in mlp_tc_bwd.cu); half: fp16 instead of bf16
uint32_t bwd_pack_bytes();
int bwd_pack_weights(const NetGeom& g, const float* params, uint8_t* packed_bwd, cudaStream_t st, bool half = false);

// The packed weight buffer of one network: [bf16 forward | bf16 backward | fp16 forward], every region 1024-byte aligned.
// nerf_pack_weights fills the first two, nerf_pack_weights_fp16 the last two (the backward is bf16 in both 16-bit modes).
inline uint32_t pack_align(uint32_t b) { return (b + 1023u) & ~1023u; }
inline uint32_t pack_off_bwd(const TcPlan& plan) { return pack_align(plan.total_bytes); }
inline uint32_t pack_off_fwd_half(const TcPlan& plan) { return pack_off_bwd(plan) + pack_align(bwd_pack_bytes()); }
inline uint32_t pack_total_bytes(const TcPlan& plan) { return pack_off_fwd_half(plan) + pack_align(plan.total_bytes); }

}  // namespace nerf
