// Whole-path entry points of the C ABI: one call = NeRF.render (src/NeRF.py:109-134) or NeRF.train_step
// (src/NeRF.py:136-178; DietNeRF's ray loss src/DietNeRF.py:159-172) for one batch of rays.  Host code only: each call
// enqueues the kernels of the other entry points on the caller's stream, in the order the Python host package issues
// them, with every intermediate in a caller-provided workspace (nerf_*_workspace_bytes) -- what a TF custom-op shim or
// any non-Python caller binds instead of re-implementing that sequence.
#include <string.h>

#include "common.cuh"

using namespace nerf;

namespace {

inline int64_t align256(int64_t b) { return (b + 255) & ~(int64_t)255; }

// carves 256-byte aligned regions out of a workspace; with base == nullptr it only measures
struct Carver {
  uint8_t* base;
  int64_t off;
  template <class T>
  T* take(int64_t count) {
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += align256(count * (int64_t)sizeof(T));
    return p;
  }
};

bool render_cfg_ok(const nerf_render_cfg* rc) {
  return rc && rc->n_samples_coarse > 0 && rc->n_samples_coarse <= 1024 && rc->n_samples_fine >= 0 &&
         rc->n_samples_coarse + rc->n_samples_fine <= 1024 &&
         (rc->mode == NERF_MODE_FP32 || rc->mode == NERF_MODE_BF16 || rc->mode == NERF_MODE_FP16);
}

// the tensor-core modes exist for the networks nerf_packed_bytes accepts (view-direction variant, 256/128 wide)
bool mode_supported(const nerf_net_cfg* cfg, int32_t mode) { return mode == NERF_MODE_FP32 || nerf_packed_bytes(cfg) >= 0; }

// fp32 mode materialises the encodings and needs the GEMM scratch; the tensor-core modes encode inside the MLP kernel
struct EncodeWs {
  float *xyz, *view;
  void* mlp_ws;
};

void carve_encode(Carver& c, const NetGeom& g, const nerf_net_cfg* cfg, int32_t mode, int64_t m_max, bool backward,
                  EncodeWs* e) {
  e->xyz = e->view = nullptr;
  e->mlp_ws = nullptr;
  if (mode != NERF_MODE_FP32) return;
  e->xyz = c.take<float>(m_max * g.dx);
  if (g.dv) e->view = c.take<float>(m_max * g.dv);
  e->mlp_ws = c.take<uint8_t>(nerf_mlp_workspace_bytes(cfg, m_max, mode, backward ? 1 : 0));
}

// the coarse pass: stratified depths (src/NeRF.py:127,146) + model_predict on them -- one kernel in the tensor-core modes
int coarse_forward(const nerf_net_cfg* cfg, const nerf_render_cfg* rc, const nerf_rng_state* rng, const float* params,
                   const void* packed, const float* o, const float* d, float* z, int64_t n, float* raw, void* saved,
                   float* xyz, float* view, void* mlp_ws, void* stream) {
  const int32_t s = rc->n_samples_coarse;
  if (rc->mode != NERF_MODE_FP32)
    return nerf_mlp_fwd_rays_stratified(cfg, packed, o, d, rc->near_boundary, rc->far_boundary, rng->seed, rng->step,
                                        rng->ray_offset, n, s, z, raw, saved, rc->mode, stream);
  int r = nerf_stratified_z(rc->near_boundary, rc->far_boundary, n, s, nullptr, rng->seed, rng->step, rng->ray_offset, z, stream);
  if (r != NERF_OK) return r;
  r = nerf_encode_samples(cfg, o, d, z, n, s, xyz, view, stream);
  if (r != NERF_OK) return r;
  return nerf_mlp_fwd(cfg, params, packed, xyz, view, n * s, raw, saved, mlp_ws, rc->mode, stream);
}

// model_predict on the samples of n rays (src/UtilsNeuralRadianceField.py:204-207, :214-234)
int net_forward(const nerf_net_cfg* cfg, int32_t mode, const float* params, const void* packed, const float* o,
                const float* d, const float* z, int64_t n, int32_t s, float* raw, void* saved, float* xyz, float* view,
                void* mlp_ws, void* stream) {
  if (mode != NERF_MODE_FP32) return nerf_mlp_fwd_rays(cfg, packed, o, d, z, n, s, raw, saved, mode, stream);
  int r = nerf_encode_samples(cfg, o, d, z, n, s, xyz, view, stream);
  if (r != NERF_OK) return r;
  return nerf_mlp_fwd(cfg, params, packed, xyz, view, n * s, raw, saved, mlp_ws, mode, stream);
}

struct RenderWs {
  float *z_c, *raw, *w_c, *z_f, *z_all;
  EncodeWs enc;
};

int64_t carve_render(const NetGeom& g, const nerf_net_cfg* cfg, const nerf_render_cfg* rc, int64_t n, uint8_t* base,
                     RenderWs* w) {
  Carver c{base, 0};
  const int64_t sc = rc->n_samples_coarse, nf = rc->n_samples_fine;
  w->z_c = c.take<float>(n * sc);
  w->raw = c.take<float>(n * (sc + nf) * 4);
  w->w_c = w->z_f = w->z_all = nullptr;
  if (nf) {
    w->w_c = c.take<float>(n * sc);
    w->z_f = c.take<float>(n * nf);
    w->z_all = c.take<float>(n * (sc + nf));
  }
  carve_encode(c, g, cfg, rc->mode, n * (sc + nf), false, &w->enc);
  return c.off;
}

struct TrainWs {
  float *z_c, *raw_c, *rgb_c, *w_c, *d_rgb_c, *d_raw_c;
  float *z_f, *u, *raw_f, *d_raw_f, *d_z_f, *d_xyz_f, *d_w_c;
  int32_t* perm;
  void *saved_c, *saved_f, *ws_bwd, *ws_bwd_c;
  float *xyz_c, *view_c, *xyz_f, *view_f;
  void* ws_fwd;
};

int64_t carve_train(const NetGeom& g, const nerf_net_cfg* cfg, const nerf_render_cfg* rc, int64_t n, uint8_t* base,
                    TrainWs* w) {
  Carver c{base, 0};
  const int64_t sc = rc->n_samples_coarse, sf = rc->n_samples_fine;
  const int32_t mode = rc->mode;
  memset(w, 0, sizeof(*w));
  w->z_c = c.take<float>(n * sc);
  w->raw_c = c.take<float>(n * sc * 4);
  w->rgb_c = c.take<float>(n * 3);
  w->w_c = c.take<float>(n * sc);
  w->d_rgb_c = c.take<float>(n * 3);
  w->d_raw_c = c.take<float>(n * sc * 4);
  w->saved_c = c.take<uint8_t>(nerf_mlp_saved_bytes(cfg, n * sc, mode));
  if (sf) {
    w->z_f = c.take<float>(n * sf);
    w->u = c.take<float>(n * sf);
    w->perm = c.take<int32_t>(n * sf);
    w->raw_f = c.take<float>(n * sf * 4);
    w->d_raw_f = c.take<float>(n * sf * 4);
    w->d_z_f = c.take<float>(n * sf);
    w->d_xyz_f = c.take<float>(n * sf * g.dx);
    w->d_w_c = c.take<float>(n * sc);
    w->saved_f = c.take<uint8_t>(nerf_mlp_saved_bytes(cfg, n * sf, mode));
  }
  const int64_t m_max = n * (sc > sf ? sc : sf);
  w->ws_bwd = c.take<uint8_t>(nerf_mlp_workspace_bytes(cfg, m_max, mode, 1));
  // with a side stream the fine network's weight-gradient kernel runs under the coarse backward: the two must not share
  // the dZ workspace
  w->ws_bwd_c = (mode != NERF_MODE_FP32 && sf) ? c.take<uint8_t>(nerf_mlp_workspace_bytes(cfg, n * sc, mode, 1)) : w->ws_bwd;
  if (mode == NERF_MODE_FP32) {
    // the backward of the fp32 mode reads the encodings of BOTH networks: each keeps its own
    w->xyz_c = c.take<float>(n * sc * g.dx);
    if (g.dv) w->view_c = c.take<float>(n * sc * g.dv);
    if (sf) {
      w->xyz_f = c.take<float>(n * sf * g.dx);
      if (g.dv) w->view_f = c.take<float>(n * sf * g.dv);
    }
    w->ws_fwd = c.take<uint8_t>(nerf_mlp_workspace_bytes(cfg, m_max, mode, 0));
  }
  return c.off;
}

// the 16-bit weight pack the mode's kernels read (bf16 regions or fp16 regions of the packed buffer)
int pack_for_mode(const nerf_net_cfg* cfg, const float* params, void* packed, int32_t mode, void* stream) {
  return mode == NERF_MODE_FP16 ? nerf_pack_weights_fp16(cfg, params, packed, stream)
                                : nerf_pack_weights(cfg, params, packed, stream);
}

// `waiter` waits for everything enqueued on `signaller` so far (an event lives only for this hand-over)
int stream_wait(cudaStream_t waiter, cudaStream_t signaller) {
  cudaEvent_t ev;
  NERF_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  cudaError_t e = cudaEventRecord(ev, signaller);
  if (e == cudaSuccess) e = cudaStreamWaitEvent(waiter, ev, 0);
  cudaEventDestroy(ev);
  if (e != cudaSuccess) {
    set_error("stream_wait: %s", cudaGetErrorString(e));
    return NERF_E_CUDA;
  }
  return NERF_OK;
}

#define NERF_TRY(call)            \
  do {                            \
    int r__ = (call);             \
    if (r__ != NERF_OK) return r__; \
  } while (0)

}  // namespace

extern "C" {

int64_t nerf_render_workspace_bytes(const nerf_net_cfg* cfg, const nerf_render_cfg* rc, int64_t n_rays) {
  NetGeom g;
  if (!make_geom(cfg, &g) || !render_cfg_ok(rc) || n_rays < 0 || !mode_supported(cfg, rc->mode)) return -1;
  RenderWs w;
  return carve_render(g, cfg, rc, n_rays, nullptr, &w);
}

int nerf_render_fused_fwd(const nerf_net_cfg* cfg, const nerf_render_cfg* rc, const float* params_c, const void* packed_c,
                          const float* params_f, const void* packed_f, const float* origs4, const float* dirs4,
                          int64_t n_rays, const nerf_rng_state* rng, const nerf_render_outs* outs, void* workspace,
                          void* stream) {
  NetGeom g;
  NERF_CHECK_ARG(make_geom(cfg, &g), "bad network config");
  NERF_CHECK_ARG(render_cfg_ok(rc), "bad render config");
  if (!mode_supported(cfg, rc->mode)) {
    set_error("%s: this network has no tensor-core path (use NERF_MODE_FP32)", __func__);
    return NERF_E_UNSUPPORTED;
  }
  NERF_CHECK_ARG(origs4 && dirs4 && rng && outs && workspace, "null pointer");
  NERF_CHECK_ARG(n_rays >= 0, "bad ray count");
  const bool tc = rc->mode != NERF_MODE_FP32;
  const bool fine = rc->n_samples_fine > 0;
  NERF_CHECK_ARG(tc ? packed_c != nullptr : params_c != nullptr, "coarse network weights missing for this mode");
  NERF_CHECK_ARG(!fine || (tc ? packed_f != nullptr : params_f != nullptr), "fine network weights missing for this mode");
  NERF_CHECK_ARG(((uintptr_t)workspace & 255) == 0, "workspace must be 256-byte aligned");
  if (n_rays == 0) return NERF_OK;
  RenderWs w;
  carve_render(g, cfg, rc, n_rays, (uint8_t*)workspace, &w);
  const int32_t sc = rc->n_samples_coarse, nf = rc->n_samples_fine;
  const int64_t n = n_rays;
  // z = get_z_values(near, far, N, 1, S_c)[:, 0, :] and the coarse render_rays                         (:127-129)
  NERF_TRY(coarse_forward(cfg, rc, rng, params_c, packed_c, origs4, dirs4, w.z_c, n, w.raw, nullptr, w.enc.xyz, w.enc.view,
                          w.enc.mlp_ws, stream));
  if (!fine) {
    NERF_TRY(nerf_composite_fwd(w.raw, w.z_c, n, sc, outs->rgb, outs->weights, outs->cumprod, outs->alpha, outs->rgb_s,
                                outs->depth, outs->acc, stream));
    if (outs->z)
      NERF_CUDA(cudaMemcpyAsync(outs->z, w.z_c, (size_t)n * sc * sizeof(float), cudaMemcpyDeviceToDevice,
                                (cudaStream_t)stream));
    return NERF_OK;
  }
  // z_from_dist = get_z_vals_from_prob_dist_func(weights, z, N_f); z = sort(concat(z_from_dist, z))   (:131-132)
  float* z_all = outs->z ? outs->z : w.z_all;
  if (nf <= 256) {
    // coarse weights, draws, sort and merge in one launch (bit-identical to the three calls below)
    NERF_TRY(nerf_hierarchical_sample(w.raw, w.z_c, n, sc, nf, rng->seed, rng->step, rng->ray_offset, z_all, stream));
  } else {
    NERF_TRY(nerf_composite_fwd(w.raw, w.z_c, n, sc, nullptr, w.w_c, nullptr, nullptr, nullptr, nullptr, nullptr, stream));
    NERF_TRY(nerf_sample_pdf_fwd(w.w_c, w.z_c, n, sc, nf, nullptr, rng->seed, rng->step, rng->ray_offset, w.z_f, nullptr,
                                 nullptr, nullptr, stream));
    NERF_TRY(nerf_merge_sorted(w.z_f, nf, w.z_c, sc, n, z_all, stream));
  }
  // fine render_rays                                                                                   (:133)
  NERF_TRY(net_forward(cfg, rc->mode, params_f, packed_f, origs4, dirs4, z_all, n, sc + nf, w.raw, nullptr, w.enc.xyz,
                       w.enc.view, w.enc.mlp_ws, stream));
  return nerf_composite_fwd(w.raw, z_all, n, sc + nf, outs->rgb, outs->weights, outs->cumprod, outs->alpha, outs->rgb_s,
                            outs->depth, outs->acc, stream);
}

int64_t nerf_train_workspace_bytes(const nerf_net_cfg* cfg, const nerf_render_cfg* rc, int64_t n_rays) {
  NetGeom g;
  if (!make_geom(cfg, &g) || !render_cfg_ok(rc) || n_rays < 0 || !mode_supported(cfg, rc->mode))
    return -1;
  TrainWs w;
  return carve_train(g, cfg, rc, n_rays, nullptr, &w);
}

}  // extern "C"

static int train_step_impl(const nerf_net_cfg* cfg, const nerf_render_cfg* rc, const nerf_train_cfg* tc_cfg,
                           float* params_c, void* packed_c, float* params_f, void* packed_f, const float* origs4,
                           const float* dirs4, const float* target_rgb, int64_t n_rays, int64_t n_total_rays,
                           const nerf_rng_state* rng, float* grads, float* adam_m, float* adam_v, int64_t adam_t,
                           float* metrics4, void* workspace, void* side_stream, const nerf_peer_exchange* peer,
                           void* stream) {
  NetGeom g;
  NERF_CHECK_ARG(make_geom(cfg, &g), "bad network config");
  NERF_CHECK_ARG(render_cfg_ok(rc), "bad render config");
  if (!mode_supported(cfg, rc->mode)) {
    set_error("%s: this network has no tensor-core path (use NERF_MODE_FP32)", __func__);
    return NERF_E_UNSUPPORTED;
  }
  NERF_CHECK_ARG(tc_cfg && origs4 && dirs4 && target_rgb && rng && grads && workspace && params_c, "null pointer");
  NERF_CHECK_ARG(n_rays >= 0 && n_total_rays >= n_rays && n_total_rays > 0, "bad ray count");
  const bool tc = rc->mode != NERF_MODE_FP32;
  const bool fine = rc->n_samples_fine > 0;
  NERF_CHECK_ARG(!fine || params_f, "fine network parameters missing");
  NERF_CHECK_ARG(!tc || (packed_c && (!fine || packed_f)), "bf16 weight packs missing");
  NERF_CHECK_ARG((adam_m == nullptr) == (adam_v == nullptr), "adam_m and adam_v go together");
  NERF_CHECK_ARG(!adam_m || adam_t >= 1, "adam_t is the 1-based step");
  NERF_CHECK_ARG(((uintptr_t)workspace & 255) == 0, "workspace must be 256-byte aligned");
  NERF_CHECK_ARG(!peer || (adam_m && peer->pads_dev && peer->grads_dev && peer->reduced_sums && peer->world >= 1 &&
                           peer->rank >= 0 && peer->rank < peer->world),
                 "the sharded step needs the Adam state, the peer tables and a valid rank");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t np = g.n_params, n = n_rays;
  // optimizer step of one network (parameter offset `off` in the [coarse | fine] layout) on stream `s`: Adam on the local
  // gradients, or -- sharded -- the sum over the ranks straight from the peers' buffers fused with the same update
  auto step_net = [&](float* params, int64_t off, void* s) -> int {
    if (peer)
      return nerf_peer_reduce_adam(params, peer->grads_dev, peer->world, 4 + off, np, adam_m + off, adam_v + off,
                                   tc_cfg->learning_rate, tc_cfg->beta_1, tc_cfg->beta_2, tc_cfg->epsilon, adam_t, nullptr, s);
    return nerf_adam_step(params, grads + 4 + off, adam_m + off, adam_v + off, np, tc_cfg->learning_rate, tc_cfg->beta_1,
                          tc_cfg->beta_2, tc_cfg->epsilon, adam_t, s);
  };
  const int64_t n_all = np * (fine ? 2 : 1);
  float* sums = grads;                       // [sum sq err coarse, sum sq err fine, 0, 0]
  float* g_c = grads + 4;
  float* g_f = fine ? grads + 4 + np : nullptr;
  if (!tc_cfg->accumulate_grads) NERF_CUDA(cudaMemsetAsync(grads, 0, (size_t)(4 + n_all) * sizeof(float), st));
  const int32_t sc = rc->n_samples_coarse, sf = rc->n_samples_fine, mode = rc->mode;
  const bool use_side = tc && fine && side_stream != nullptr && side_stream != stream;
  bool forked = false, fine_stepped = false;
  const int32_t mode_all = mode;
  if (n > 0) {
    TrainWs w;
    carve_train(g, cfg, rc, n, (uint8_t*)workspace, &w);
    const bool through_z = fine && !tc_cfg->stop_grad_z;
    // coarse forward + loss                                                            (src/NeRF.py:146-151)
    NERF_TRY(coarse_forward(cfg, rc, rng, params_c, packed_c, origs4, dirs4, w.z_c, n, w.raw_c, w.saved_c, w.xyz_c, w.view_c,
                            w.ws_fwd, stream));
    NERF_TRY(nerf_composite_mse_fwd(w.raw_c, w.z_c, target_rgb, n, sc, n_total_rays, tc_cfg->coarse_loss_weight, w.rgb_c,
                                    w.w_c, sums, w.d_rgb_c, stream));
    const float* d_w_c = nullptr;
    if (fine) {
      // fine network on the importance samples only, loss, and the tape's way back      (:154-157, :160-163)
      NERF_TRY(nerf_sample_pdf_fwd(w.w_c, w.z_c, n, sc, sf, nullptr, rng->seed, rng->step, rng->ray_offset, w.z_f,
                                   nullptr, w.perm, w.u, stream));
      NERF_TRY(net_forward(cfg, mode, params_f, packed_f, origs4, dirs4, w.z_f, n, sf, w.raw_f, w.saved_f, w.xyz_f,
                           w.view_f, w.ws_fwd, stream));
      NERF_TRY(nerf_composite_mse_fwd_bwd(w.raw_f, w.z_f, target_rgb, n, sf, n_total_rays, 1.0f, nullptr, sums + 1,
                                          w.d_raw_f, through_z ? w.d_z_f : nullptr, stream));
      // tensor-core modes, the reference's encoding width: the chain kernel turns d(xyz encoding) into d z in its last
      // epilogue (nerf_mlp_bwd_rays) -- no d_xyz round trip, no nerf_encode_samples_bwd_z launch
      const bool dz_in_chain = tc && through_z && g.dx == 33 && g.view;
      if (use_side) {
        // input-gradient chain here, the weight-gradient kernel on the side stream; the fine network's Adam step and the
        // refresh of its bf16 pack follow it there and run under the sampler / compositing / coarse backward below
        if (dz_in_chain)
          NERF_TRY(nerf_mlp_bwd_rays(cfg, packed_f, w.saved_f, w.d_raw_f, origs4, dirs4, w.z_f, n, sf, g_f, w.d_z_f, 1, w.ws_bwd,
                                     mode, 3, side_stream, stream));
        else
        NERF_TRY(nerf_mlp_bwd_overlapped(cfg, params_f, packed_f, w.xyz_f, w.view_f, w.saved_f, w.d_raw_f, n * sf, g_f,
                                         through_z ? w.d_xyz_f : nullptr, w.ws_bwd, mode, side_stream, stream));
        if (adam_m) {
          if (peer) NERF_TRY(nerf_peer_barrier(peer->pads_dev, peer->rank, peer->world, peer->epoch, 0, side_stream));
          NERF_TRY(step_net(params_f, np, side_stream));
          NERF_TRY(pack_for_mode(cfg, params_f, packed_f, mode, side_stream));
          fine_stepped = true;
        }
        forked = true;
      } else if (dz_in_chain) {
        NERF_TRY(nerf_mlp_bwd_rays(cfg, packed_f, w.saved_f, w.d_raw_f, origs4, dirs4, w.z_f, n, sf, g_f, w.d_z_f, 1, w.ws_bwd,
                                   mode, 3, nullptr, stream));
      } else {
        NERF_TRY(nerf_mlp_bwd(cfg, params_f, packed_f, w.xyz_f, w.view_f, w.saved_f, w.d_raw_f, n * sf, g_f,
                              through_z ? w.d_xyz_f : nullptr, w.ws_bwd, mode, stream));
      }
      if (through_z) {
        if (!dz_in_chain)
          NERF_TRY(nerf_encode_samples_bwd_z(cfg, origs4, dirs4, w.z_f, w.d_xyz_f, n, sf, w.d_z_f, 1, stream));
        NERF_TRY(nerf_sample_pdf_bwd(w.w_c, w.z_c, w.u, w.perm, w.d_z_f, n, sc, sf, w.d_w_c, stream));
        d_w_c = w.d_w_c;
      }
    }
    // coarse backward
    NERF_TRY(nerf_composite_bwd(w.raw_c, w.z_c, w.d_rgb_c, d_w_c, n, sc, w.d_raw_c, nullptr, stream));
    const bool side_c = tc && side_stream != nullptr && side_stream != stream;
    NERF_TRY(nerf_mlp_bwd_overlapped(cfg, params_c, packed_c, w.xyz_c, w.view_c, w.saved_c, w.d_raw_c, n * sc, g_c, nullptr,
                                     forked ? w.ws_bwd_c : w.ws_bwd, mode, side_c ? side_stream : nullptr, stream));
    if (forked || side_c) NERF_TRY(stream_wait(st, (cudaStream_t)side_stream));   // the weight gradients join here
  }
  if (adam_m) {
    // optimizer.apply_gradients over the variables of both models (:164-167); moments laid out [coarse | fine]
    if (peer) {
      // every rank's [sums | coarse gradients] are final: barrier, then the loss sums and the coarse network straight
      // from the peers' buffers
      NERF_TRY(nerf_peer_barrier(peer->pads_dev, peer->rank, peer->world, peer->epoch, 1, stream));
      NERF_TRY(nerf_peer_reduce_adam(nullptr, peer->grads_dev, peer->world, 0, 4, nullptr, nullptr, 0.f, 0.f, 0.f, 0.f, 1,
                                     peer->reduced_sums, stream));
      sums = peer->reduced_sums;
    }
    NERF_TRY(step_net(params_c, 0, stream));
    if (fine && !fine_stepped) NERF_TRY(step_net(params_f, np, stream));
    if (tc) {
      NERF_TRY(pack_for_mode(cfg, params_c, packed_c, mode_all, stream));
      if (fine && !fine_stepped) NERF_TRY(pack_for_mode(cfg, params_f, packed_f, mode_all, stream));
    }
  }
  if (metrics4)
    NERF_TRY(nerf_train_metrics(sums, n_total_rays, tc_cfg->coarse_loss_weight, fine ? 1 : 0, metrics4, stream));
  return NERF_OK;
}

extern "C" {

int nerf_train_step_fused(const nerf_net_cfg* cfg, const nerf_render_cfg* rc, const nerf_train_cfg* tc_cfg,
                          float* params_c, void* packed_c, float* params_f, void* packed_f, const float* origs4,
                          const float* dirs4, const float* target_rgb, int64_t n_rays, int64_t n_total_rays,
                          const nerf_rng_state* rng, float* grads, float* adam_m, float* adam_v, int64_t adam_t,
                          float* metrics4, void* workspace, void* side_stream, void* stream) {
  return train_step_impl(cfg, rc, tc_cfg, params_c, packed_c, params_f, packed_f, origs4, dirs4, target_rgb, n_rays,
                         n_total_rays, rng, grads, adam_m, adam_v, adam_t, metrics4, workspace, side_stream, nullptr, stream);
}

int nerf_train_step_fused_sharded(const nerf_net_cfg* cfg, const nerf_render_cfg* rc, const nerf_train_cfg* tc_cfg,
                                  float* params_c, void* packed_c, float* params_f, void* packed_f, const float* origs4,
                                  const float* dirs4, const float* target_rgb, int64_t n_rays, int64_t n_total_rays,
                                  const nerf_rng_state* rng, float* grads, float* adam_m, float* adam_v, int64_t adam_t,
                                  float* metrics4, void* workspace, void* side_stream, const nerf_peer_exchange* peer,
                                  void* stream) {
  NERF_CHECK_ARG(peer != nullptr, "null peer exchange");
  return train_step_impl(cfg, rc, tc_cfg, params_c, packed_c, params_f, packed_f, origs4, dirs4, target_rgb, n_rays,
                         n_total_rays, rng, grads, adam_m, adam_v, adam_t, metrics4, workspace, side_stream, peer, stream);
}

}  // extern "C"
