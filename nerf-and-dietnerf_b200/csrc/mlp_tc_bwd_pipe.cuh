// Plan and hand-over counters of the layer-pipelined backward kernel (mlp_tc_bwd_pipe.cu), shared with its host side in
// mlp_tc_bwd.cu.
#pragma once
#include <stdint.h>

namespace nerf {

// ---- ready counters in global memory ------------------------------------------------------------------------------------
__device__ __forceinline__ void flag_signal(uint32_t* p) {
  // release at gpu scope: this thread's (and, through the preceding __syncwarp, its warp's) global stores are visible
  // to whoever acquires the incremented counter
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}
__device__ __forceinline__ uint32_t flag_load(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void flag_wait(const uint32_t* p, uint32_t target) {
  if (flag_load(p) >= target) return;
  for (uint32_t it = 0;; ++it) {
    __nanosleep(100);
    if (flag_load(p) >= target) return;
    if (it > (1u << 25)) {     // seconds: the producer never started or died
      printf("nerf_b200: dZ hand-over flag timeout (block %d)\n", blockIdx.x);
      __trap();
    }
  }
}

// ---- pipe plan ----------------------------------------------------------------------------------------------------------
// flags[s * kPipeFlagsPerSuper + l]: ready counter of dZ_l of super-tile s (tiles 2s, 2s+1), l = 1..7; complete at
// kPipeFlagTarget = one signal per epilogue warp of both CTAs of the producing pair
constexpr int kPipeFlagsPerSuper = 8;
constexpr uint32_t kPipeFlagTarget = 32;

enum : int { PIPE_ROLE_LAYER = 0, PIPE_ROLE_DW_ONLY = 1 };

struct PipeGroup {
  int16_t role;            // PIPE_ROLE_*
  int16_t layer;           // l: consumes dZ_{l+1} (+ the saved input of Dense l), produces dZ_l and dW_l
  int16_t first_pair, n_pairs;
  int16_t wait_flag;       // flag index of the consumed block, -1: complete before the launch
  int16_t signal_flag;     // flag index of the produced block, -1: nobody inside the launch waits for it
  int16_t n_chain, n_dw;   // UMMA N of the chain step / of the dW step
  int16_t b_chunks;        // 16-byte feature chunks of the B slab this CTA streams (n_dw / 16)
  int16_t b_half_chunks;   // chunks per row half of the B block in the dZ workspace (its half stride in KB)
  int16_t mask_row;        // row of the saved sign masks for dZ_l (l - 1)
  int16_t dense;           // Keras Dense index the weight gradient belongs to (reduce kernel)
  int16_t has_bias;
  int32_t w_off;           // byte offset of W_l^T's first K chunk in the backward weight pack
  int32_t a_off;           // byte offset of the dW A block inside a saved tile
  int32_t b_off;           // byte offset of dZ_{l+1} inside a dZ-workspace tile
  int32_t out_off;         // byte offset of dZ_l inside a dZ-workspace tile
};
struct PipePlan {
  PipeGroup g[12];
  int32_t n_groups;
};

}  // namespace nerf
