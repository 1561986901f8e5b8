// placeholder, replaced below
#include "common.cuh"
namespace nerf {
int mlp_tc_fwd(const nerf_net_cfg*, const NetGeom&, const float*, const void*, const float*, const float*, int64_t,
               float*, void*, void*, cudaStream_t) { set_error("BF16 path not built"); return NERF_E_UNSUPPORTED; }
int mlp_tc_bwd(const nerf_net_cfg*, const NetGeom&, const float*, const void*, const float*, const float*, const void*,
               const float*, int64_t, float*, float*, void*, cudaStream_t) { set_error("BF16 path not built"); return NERF_E_UNSUPPORTED; }
int64_t mlp_tc_saved_bytes(const NetGeom&, int64_t) { return 0; }
int64_t mlp_tc_workspace_bytes(const NetGeom&, int64_t, int) { return 0; }
}
extern "C" {
int64_t nerf_packed_bytes(const nerf_net_cfg*) { return 0; }
int nerf_pack_weights(const nerf_net_cfg*, const float*, void*, void*) { nerf::set_error("BF16 path not built"); return NERF_E_UNSUPPORTED; }
}
