/*
 * nerf_b200.h -- C ABI of the B200-native NeRF ray-render hot path (libnerf_b200.so).
 *
 * The reference (Sahar-E/NeRF-and-DietNeRF) has no FFI layer: its boundary for this path is the set of
 * Python functions in src/UtilsCV.py, src/UtilsNeuralRadianceField.py, src/NeRF.py and src/DietNeRF.py.
 * Each entry point below names the reference function (file:line, relative to the reference root) it
 * replaces.  The Python host package (nerf-and-dietnerf_b200/) keeps the reference's names and argument
 * order and calls these symbols through ctypes; INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless the name ends in _host;
 *     tensors are row-major, contiguous, fp32 unless stated;
 *   - `stream` is a cudaStream_t passed as void*; no call synchronises, allocates or frees device memory;
 *   - return value 0 = ok, negative = error (NERF_E_*); nerf_last_error() gives the thread-local message;
 *   - there is NO CPU fallback: without a CUDA device every compute call returns NERF_E_CUDA.
 *   - RNG: Philox4x32-10, key = seed, counter = (global ray index, draw/4, stream id, step); see
 *     oracle/philox.py for the exact stream (stream id 0 = stratified jitter, 1 = importance uniforms).
 */
#ifndef NERF_B200_H
#define NERF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NERF_OK 0
#define NERF_E_ARG (-1)     /* bad argument (shape, null pointer, unsupported config) */
#define NERF_E_CUDA (-2)    /* CUDA runtime error (message in nerf_last_error) */
#define NERF_E_UNSUPPORTED (-3)

/* MLP arithmetic modes.  FP32: SIMT fp32 GEMMs (1e-5 parity mode).  BF16: tcgen05/TMEM tensor-core
 * chain, bf16 operands, fp32 accumulate (1e-3 parity mode). */
#define NERF_MODE_FP32 0
#define NERF_MODE_BF16 1
/* FP16: the same tensor-core forward chain with fp16 operands (8x finer operand rounding than bf16, same throughput);
 * forward only -- the render mode that meets the 1e-3 bound on sharp networks (the reference itself renders in
 * float16).  Needs nerf_pack_weights_fp16. */
#define NERF_MODE_FP16 2

/* Network hyper-parameters: the `neural_net:` block of the reference YAML (src/NeRF.py:35-53,
 * src/ConfigurationKeys.py). */
typedef struct nerf_net_cfg {
  int32_t n_pos_enc_xyz;   /* n_pos_enc_dim_xyz  (L_xyz; 5 in every reference config) */
  int32_t n_pos_enc_view;  /* n_pos_enc_view_dir (L_view; 4 or 2) */
  int32_t n_angles;        /* n_angles_for_model: 0 (no view input), 1 ([x,z]) or 2 ([x,y,z]) */
  int32_t hidden;          /* hidden_layer_dim (256) */
  int32_t last_hidden;     /* last_hidden_layer_dim (128) */
  float leaky_alpha;       /* leaky_relu_alpha (0.05) */
} nerf_net_cfg;

const char* nerf_version(void);
const char* nerf_last_error(void);

/* Number of fp32 parameters of one net, laid out [W0 (in,out) row-major, b0, W1, b1, ...] in Keras layer
 * creation order (src/NeRF.py:263-286 / :312-337).  514332 for the standard config. */
int64_t nerf_param_count(const nerf_net_cfg* cfg);
/* Widths of the encoded inputs: 3+6*L_xyz and 2*L_view*(n_angles+1) (0 when n_angles==0). */
int32_t nerf_xyz_enc_dim(const nerf_net_cfg* cfg);
int32_t nerf_view_enc_dim(const nerf_net_cfg* cfg);

/* ---- rays ------------------------------------------------------------------------------------------ */
/* get_rays_directions (src/UtilsCV.py:467-499) + origin broadcast (src/NeRF.py:209,
 * src/UtilsNeuralRadianceField.py:177) for pixels [ray_begin, ray_begin+n_rays) of an h x w image
 * (ray index = y*w+x).  c2w_host: 16 floats, row-major 4x4, HOST memory.  dirs4/origs4: (n_rays,4). */
int nerf_ray_directions(const float* c2w_host, float fov, int32_t h, int32_t w, int64_t ray_begin,
                        int64_t n_rays, float* dirs4, float* origs4, void* stream);

/* get_z_values(z_start,z_end,N,1,S)[:,0,:] (src/UtilsCV.py:565-581; call sites src/NeRF.py:127,146).
 * jitter_or_null: (N,S) uniforms to use instead of the Philox stream (tests). z: (N,S). */
int nerf_stratified_z(float z_start, float z_end, int64_t n_rays, int32_t n_samples,
                      const float* jitter_or_null, uint64_t seed, uint32_t step, uint64_t ray_offset,
                      float* z, void* stream);

/* sample_along_rays (src/UtilsCV.py:584-599): out (N,S,4) = origin + dir * z. */
int nerf_sample_along_rays(const float* origs4, const float* dirs4, const float* z, int64_t n_rays,
                           int32_t n_samples, float* coords4, void* stream);

/* get_view_directions (src/UtilsCV.py:124-143): out (N*S, n_angles+1); n_angles must be 1 or 2. */
int nerf_view_directions(const float* dirs4, int64_t n_rays, int32_t n_samples, int32_t n_angles,
                         float* view_dirs, void* stream);

/* positional_encoding_for_xyz / _for_views (src/UtilsNeuralRadianceField.py:68-85 / :52-65).
 * x: (M,C) (C==3 for xyz); out: (M, C+2*L*C) resp. (M, 2*L*C). */
int nerf_posenc_xyz(const float* xyz, int64_t m, int32_t L, float* out, void* stream);
int nerf_posenc_views(const float* x, int64_t m, int32_t c, int32_t L, float* out, void* stream);
/* Gradient of positional_encoding_for_xyz w.r.t. xyz (what TF autodiff produces inside NeRF.train_step). */
int nerf_posenc_xyz_bwd(const float* xyz, const float* d_out, int64_t m, int32_t L, float* d_xyz, void* stream);

/* Fused K1 for the MLP input: sample_along_rays[..., :3] -> both encodings, never materialising the
 * (N,S,4) coordinates (src/UtilsNeuralRadianceField.py:204-205,229-231).  view_enc may be null when
 * cfg->n_angles == 0.  xyz_enc: (N*S, 3+6L); view_enc: (N*S, Dv). */
int nerf_encode_samples(const nerf_net_cfg* cfg, const float* origs4, const float* dirs4, const float* z,
                        int64_t n_rays, int32_t n_samples, float* xyz_enc, float* view_enc, void* stream);
/* Backward of the xyz half of nerf_encode_samples w.r.t. z: d_z (N,S) += sum_c d_xyz_c * dir_c. */
int nerf_encode_samples_bwd_z(const nerf_net_cfg* cfg, const float* origs4, const float* dirs4,
                              const float* z, const float* d_xyz_enc, int64_t n_rays, int32_t n_samples,
                              float* d_z, int32_t accumulate, void* stream);

/* ---- MLP (model_predict, src/UtilsNeuralRadianceField.py:214-234; nets src/NeRF.py:248-340) ------------ */
/* Bytes of activations nerf_mlp_fwd saves for nerf_mlp_bwd (0 rows -> 0). */
int64_t nerf_mlp_saved_bytes(const nerf_net_cfg* cfg, int64_t m, int32_t mode);
/* Scratch bytes nerf_mlp_fwd / nerf_mlp_bwd need. */
int64_t nerf_mlp_workspace_bytes(const nerf_net_cfg* cfg, int64_t m, int32_t mode, int32_t backward);
/* out4 (M,4) = [r,g,b,sigma] raw.  saved_or_null: activations for backward.  packed_or_null: bf16 weight
 * pack from nerf_pack_weights (required for NERF_MODE_BF16). */
int nerf_mlp_fwd(const nerf_net_cfg* cfg, const float* params, const void* packed_or_null,
                 const float* xyz_enc, const float* view_enc, int64_t m, float* out4, void* saved_or_null,
                 void* workspace, int32_t mode, void* stream);
/* Fused K1+K2 (NERF_MODE_BF16 only): render_rays' sample_along_rays -> get_view_directions -> both positional
 * encodings -> MLP (src/UtilsNeuralRadianceField.py:204-207) in ONE kernel: the encoded samples are written straight
 * into the shared-memory A operand of the first layer and never reach HBM.  Rows are ordered ray-major: M = N*S. */
int nerf_mlp_fwd_rays(const nerf_net_cfg* cfg, const void* packed, const float* origs4, const float* dirs4,
                      const float* z, int64_t n_rays, int32_t n_samples, float* out4, void* saved_or_null,
                      int32_t mode, void* stream);
/* The coarse pass of NeRF.render / train_step with the stratified sampling fused in as well: get_z_values
 * (src/UtilsCV.py:565-581; call sites src/NeRF.py:127,146) is evaluated in the MLP prologue from the Philox stream -- the
 * same draws and arithmetic as nerf_stratified_z, bit for bit -- so the coarse depths go straight from registers into the
 * positional encoding; they are also written to z_out (N,S) for the compositing and importance-sampling kernels. */
int nerf_mlp_fwd_rays_stratified(const nerf_net_cfg* cfg, const void* packed, const float* origs4, const float* dirs4,
                                 float z_start, float z_end, uint64_t seed, uint32_t step, uint64_t ray_offset,
                                 int64_t n_rays, int32_t n_samples, float* z_out, float* out4, void* saved_or_null,
                                 int32_t mode, void* stream);
/* grads (+= , same layout as params); d_xyz_enc_or_null (M, Dx) is written when non-null.  xyz_enc / view_enc may be
 * null in NERF_MODE_BF16 (the forward pass saved its bf16 input panel). */
int nerf_mlp_bwd(const nerf_net_cfg* cfg, const float* params, const void* packed_or_null,
                 const float* xyz_enc, const float* view_enc, const void* saved, const float* d_out4,
                 int64_t m, float* grads, float* d_xyz_enc_or_null, void* workspace, int32_t mode,
                 void* stream);
/* The two halves of nerf_mlp_bwd in NERF_MODE_BF16 (same arguments), for callers that time or overlap them:
 * _dx = the input-gradient chain (writes dZ of every layer into the workspace and d_xyz_enc), _dw = the weight
 * gradients from the saved activations and that workspace (deterministic split-K reduction).  _dw must follow _dx on
 * the same workspace.  TF autodiff of the Keras models in NeRF.train_step (src/NeRF.py:149-167) is what both replace. */
int nerf_mlp_bwd_dx(const nerf_net_cfg* cfg, const float* params, const void* packed_or_null,
                    const float* xyz_enc, const float* view_enc, const void* saved, const float* d_out4,
                    int64_t m, float* grads, float* d_xyz_enc_or_null, void* workspace, int32_t mode,
                    void* stream);
int nerf_mlp_bwd_dw(const nerf_net_cfg* cfg, const float* params, const void* packed_or_null,
                    const float* xyz_enc, const float* view_enc, const void* saved, const float* d_out4,
                    int64_t m, float* grads, float* d_xyz_enc_or_null, void* workspace, int32_t mode,
                    void* stream);
/* nerf_mlp_bwd over two streams of the caller's (NERF_MODE_BF16): the input-gradient chain on `stream`, the weight-gradient
 * kernel and its reduction on `side_stream` -- after the chain, so that it runs under whatever the caller enqueues next on
 * `stream` (the sampler / compositing backward and the head of the other network's chain in a train step).  d_xyz_enc is
 * complete at the end of `stream`, `grads` at the end of `side_stream`: the caller joins the two.  side_stream null or ==
 * stream, or NERF_MODE_FP32: identical to nerf_mlp_bwd.  With the environment variable NERF_BWD_OVERLAP=1 the two kernels
 * run AT THE SAME TIME on disjoint SMs and hand every dZ block over through a ready counter in the workspace (an
 * experiment kept for measurement: the weight-gradient kernel is bound by what one SM can pull in, so it loses more from
 * having fewer SMs than it gains from reading dZ out of L2 -- profiles/r02_a_overlap_probe.log); its split-K partition
 * follows the SMs it gets, so that variant agrees with nerf_mlp_bwd to fp32 rounding, not bit for bit.
 * Replaces TF autodiff of the Keras models in NeRF.train_step (src/NeRF.py:149-167). */
int nerf_mlp_bwd_overlapped(const nerf_net_cfg* cfg, const float* params, const void* packed_or_null,
                            const float* xyz_enc, const float* view_enc, const void* saved, const float* d_out4,
                            int64_t m, float* grads, float* d_xyz_enc_or_null, void* workspace, int32_t mode,
                            void* side_stream, void* stream);
/* nerf_mlp_fwd_rays for the pixels [ray_begin, ray_begin + n_rays) of an h x w image seen from c2w_host (16 floats,
 * row-major, HOST memory) -- the rays are generated in the kernel's prologue with nerf_ray_directions' arithmetic (bit for
 * bit), so get_rays_directions + get_z_values + sample_along_rays + both encodings + model_predict are ONE kernel and no ray
 * ever reaches HBM (north_star kernel (1) folded into kernel (2); src/NeRF.py:206-228 render_image's rays).  z_or_null:
 * (n_rays, n_samples) depths (the fine pass); null: stratified depths of [z_start, z_end] drawn in the prologue from the
 * Philox stream (seed, step, ray counter = pixel index) and written to z_out.  Inference only; tensor-core modes. */
int nerf_mlp_fwd_camera(const nerf_net_cfg* cfg, const void* packed, const float* c2w_host, float fov, int32_t h, int32_t w,
                        int64_t ray_begin, int64_t n_rays, int32_t n_samples, const float* z_or_null, float z_start,
                        float z_end, uint64_t seed, uint32_t step, float* z_out_or_null, float* out4, int32_t mode,
                        void* stream);
/* nerf_mlp_bwd of the rows of n_rays x n_samples ray samples (tensor-core modes, n_pos_enc_dim_xyz = 5), for callers
 * that need the gradient w.r.t. the DEPTHS instead of the gradient w.r.t. the xyz encoding (the reference does not
 * detach the importance samples, src/NeRF.py:155, so the fine loss reaches the coarse network through z): the chain
 * kernel's last epilogue contracts d(xyz encoding) with PE'(o + d z) and the ray direction itself and writes (or, with
 * accumulate_d_z, adds) 4 bytes per row -- nerf_encode_samples_bwd_z and its 132-byte-per-row round trip disappear.
 * parts: 1 = input-gradient chain only, 2 = weight gradients only, 3 = both (then side_stream_or_null works as in
 * nerf_mlp_bwd_overlapped).  Replaces TF autodiff of model_predict o positional_encoding_for_xyz o sample_along_rays
 * (src/UtilsNeuralRadianceField.py:52-69,204-234, src/UtilsCV.py:584-599). */
int nerf_mlp_bwd_rays(const nerf_net_cfg* cfg, const void* packed, const void* saved, const float* d_out4,
                      const float* origs4, const float* dirs4, const float* z, int64_t n_rays, int32_t n_samples,
                      float* grads, float* d_z, int32_t accumulate_d_z, void* workspace, int32_t mode, int32_t parts,
                      void* side_stream_or_null, void* stream);
/* Diagnostic (tests, tools/bwd_pipe_probe.py): ONE stage of the layer-pipelined backward kernel over every SM -- Dense
 * `layer` (1..7: dZ_layer = (dZ_{layer+1} W^T) * LeakyReLU' written back into the workspace, and dW/db of that layer ADDED
 * to `grads`; 8: only the weight gradient of Dense 8 / the sigma head w.r.t. the h8 rows), reading dZ_{layer+1} from a
 * workspace that a previous nerf_mlp_bwd[_dx] call on the same `saved` filled.  layer = 100 hi + lo (7 >= hi > lo >= 1): the
 * layer groups hi ... lo in ONE launch, the CTA pairs split evenly over them, dZ handed from group to group through ready
 * counters in L2 (blocks in the tile chunk-major layout, tools/tcm_layout.py). */
int nerf_debug_bwd_pipe_layer(const nerf_net_cfg* cfg, const void* packed, const void* saved, int64_t m,
                              void* workspace, int32_t layer, float* grads, void* stream);
/* bf16 weight pack for the tensor-core path (re-run after every optimizer step). */
int64_t nerf_packed_bytes(const nerf_net_cfg* cfg);
int nerf_pack_weights(const nerf_net_cfg* cfg, const float* params, void* packed, void* stream);
/* fills the fp16 region of the same packed buffer (only needed before NERF_MODE_FP16 calls). */
int nerf_pack_weights_fp16(const nerf_net_cfg* cfg, const float* params, void* packed, void* stream);

/* ---- alpha compositing (ray_marching, src/UtilsNeuralRadianceField.py:88-115) ------------------------- */
/* raw4 (N,S,4), z (N,S).  Outputs (any may be null): rgb (N,3), weights/cumprod/alpha (N,S),
 * rgb_s (N,S,3), depth (N) = sum w z (src/ExecutionRun.py:346), acc (N) = sum w. */
int nerf_composite_fwd(const float* raw4, const float* z, int64_t n_rays, int32_t n_samples, float* rgb,
                       float* weights, float* cumprod, float* alpha, float* rgb_s, float* depth, float* acc,
                       void* stream);
/* Gradients TF autodiff gives for ray_marching (cumprod gradient = div_no_nan form).  d_weights_or_null:
 * extra upstream gradient on `weights` (from the importance sampler).  d_z_or_null: written when non-null. */
int nerf_composite_bwd(const float* raw4, const float* z, const float* d_rgb, const float* d_weights_or_null,
                       int64_t n_rays, int32_t n_samples, float* d_raw4, float* d_z_or_null, void* stream);

/* ray_marching + keras MeanSquaredError in one launch (src/NeRF.py:150-151): rgb (N,3), weights (N,S) as above, plus
 * d_rgb = loss_weight * 2 (rgb-target) / (3*n_total_rays) and sq_err_sum += sum((rgb-target)^2) (caller zeroes). */
int nerf_composite_mse_fwd(const float* raw4, const float* z, const float* target, int64_t n_rays, int32_t n_samples,
                           int64_t n_total_rays, float loss_weight, float* rgb, float* weights, float* sq_err_sum,
                           float* d_rgb, void* stream);
/* ray_marching + MeanSquaredError + their gradient in one launch (the fine network's part of NeRF.train_step,
 * src/NeRF.py:156-157 and what the tape returns for it): d_raw4 (N,S,4), d_z (N,S) when non-null, rgb when non-null. */
int nerf_composite_mse_fwd_bwd(const float* raw4, const float* z, const float* target, int64_t n_rays, int32_t n_samples,
                               int64_t n_total_rays, float loss_weight, float* rgb_or_null, float* sq_err_sum,
                               float* d_raw4, float* d_z_or_null, void* stream);

/* ---- hierarchical sampling (get_z_vals_from_prob_dist_func, src/UtilsCV.py:502-539) ------------------- */
/* weights,z: (N,S); z_new: (N,Nf) sorted.  u_or_null: (N,Nf) uniforms replacing the Philox stream.
 * Optional outputs: idx (N,Nf) int32 searchsorted results in DRAW order, perm (N,Nf) int32 sort
 * permutation (z_new[k] = z_unsorted[perm[k]]), u_out (N,Nf) the uniforms used. */
int nerf_sample_pdf_fwd(const float* weights, const float* z, int64_t n_rays, int32_t n_samples,
                        int32_t n_new, const float* u_or_null, uint64_t seed, uint32_t step,
                        uint64_t ray_offset, float* z_new, int32_t* idx_or_null, int32_t* perm_or_null,
                        float* u_out_or_null, void* stream);
/* The hierarchical-sampling step of NeRF.render as ONE launch (src/NeRF.py:129-132 between the two networks): the coarse
 * weights of ray_marching (src/UtilsNeuralRadianceField.py:101-108: alpha, exclusive-cumprod transmittance) formed from
 * raw4_coarse (N,S,4), get_z_vals_from_prob_dist_func (src/UtilsCV.py:502-539, Philox stream of nerf_sample_pdf_fwd)
 * and z_all = sort(concat(z_from_dist, z_coarse)) (N,S+Nf).  Bit-identical to nerf_composite_fwd (weights) ->
 * nerf_sample_pdf_fwd -> nerf_merge_sorted; n_new <= 256. */
int nerf_hierarchical_sample(const float* raw4_coarse, const float* z_coarse, int64_t n_rays, int32_t n_samples,
                             int32_t n_new, uint64_t seed, uint32_t step, uint64_t ray_offset, float* z_all,
                             void* stream);
/* d_weights (N,S) written: gradient of z_new w.r.t. weights contracted with d_z_new (N,Nf, sorted order). */
int nerf_sample_pdf_bwd(const float* weights, const float* z, const float* u, const int32_t* perm,
                        const float* d_z_new, int64_t n_rays, int32_t n_samples, int32_t n_new,
                        float* d_weights, void* stream);
/* z = sort(concat(z_a, z_b)) per ray (src/NeRF.py:132); both inputs sorted ascending. out: (N,Sa+Sb). */
int nerf_merge_sorted(const float* z_a, int32_t sa, const float* z_b, int32_t sb, int64_t n_rays, float* out,
                      void* stream);
/* Same merge, also writing rank_a (N,Sa) int32: out[ray][rank_a[ray][j]] = z_a[ray][j].  Elements of z_a precede
 * equal elements of z_b (the order of a stable sort of concat(z_a, z_b)). */
int nerf_merge_sorted_rank(const float* z_a, int32_t sa, const float* z_b, int32_t sb, int64_t n_rays, float* out,
                           int32_t* rank_a, void* stream);
/* Gradient of that sort(concat) w.r.t. z_a when it sits inside the tape (DietNeRF's in-tape render_image,
 * src/DietNeRF.py:215-218 -> src/NeRF.py:132): d_a[ray][j] = d_out[ray][rank_a[ray][j]]. */
int nerf_merge_sorted_bwd(const float* d_out, const int32_t* rank_a, int32_t sa, int32_t sb, int64_t n_rays,
                          float* d_a, void* stream);

/* ---- loss / optimizer (src/NeRF.py:151,157,164-176) -------------------------------------------------- */
/* d_rgb = loss_weight * 2 (rgb-target) / (3*n_total); sums[0] += sum((rgb-target)^2) (caller zeroes). */
int nerf_mse_fwd_bwd(const float* rgb, const float* target, int64_t n_rays, int64_t n_total_rays,
                     float loss_weight, float* sq_err_sum, float* d_rgb, void* stream);
/* Metrics of the step from the two squared-error sums (src/NeRF.py:170-178; get_psnr src/UtilsNeuralRadianceField.py:
 * 123-132): out4 = [coarse_loss_weight*MSE_c + MSE_f, psnr_coarse, psnr_fine, MSE_c + MSE_f], MSE = sum / (3*n_total). */
int nerf_train_metrics(const float* sq_err_sums, int64_t n_total_rays, float coarse_loss_weight, int32_t has_fine,
                       float* out4, void* stream);
/* Keras-2.7 Adam update, t = 1-based step. */
int nerf_adam_step(float* params, const float* grads, float* m, float* v, int64_t n, float lr, float beta1,
                   float beta2, float eps, int64_t t, void* stream);

/* ---- gradient exchange of the ray-sharded train step over NVLink peer memory (replaces the all-reduce a
 * tf.distribute strategy would insert around optimizer.apply_gradients, src/NeRF.py:164-167) -------------------------- */
/* Barrier over the GPUs of one box.  pads_dev: DEVICE array of `world` pointers, entry r = rank r's signal pad (>= 1 KB of
 * zero-initialised uint32, peer-mapped symmetric memory).  `epoch` must grow from call to call on the same `slot` (0..3;
 * one slot per stream that issues barriers). */
int nerf_peer_barrier(void* const* pads_dev, int32_t rank, int32_t world, uint32_t epoch, int32_t slot, void* stream);
/* One-shot all-reduce (sum, rank order: bit-identical on every GPU) of floats [offset, offset + n) of the `world` flat
 * gradient buffers grads_dev[r] (DEVICE array of peer-mapped base pointers), fused with nerf_adam_step's update of
 * params / m / v (n floats each; same arithmetic).  params_or_null == NULL: reduction only.  reduced_or_null: local copy of
 * the summed slice.  Call after nerf_peer_barrier on the same stream; offset must be a multiple of 4. */
int nerf_peer_reduce_adam(float* params_or_null, void* const* grads_dev, int32_t world, int64_t offset, int64_t n, float* m,
                          float* v, float lr, float beta1, float beta2, float eps, int64_t t, float* reduced_or_null,
                          void* stream);

/* ---- the whole path in one call -------------------------------------------------------------------------- */
/* The `render:` block of the YAML plus the frustum and the MLP arithmetic (NeRF.__init__, src/NeRF.py:35-66). */
typedef struct nerf_render_cfg {
  float near_boundary, far_boundary;
  int32_t n_samples_coarse; /* n_render_samples_coarse */
  int32_t n_samples_fine;   /* n_render_samples_fine; 0 = coarse network only (no fine model) */
  int32_t mode;             /* NERF_MODE_* (FP16: render only) */
} nerf_render_cfg;
/* Position in the Philox stream: the draws of ray i use counter (ray_offset + i, ., ., step). */
typedef struct nerf_rng_state {
  uint64_t seed;
  uint64_t ray_offset;
  uint32_t step;
  uint32_t reserved;
} nerf_rng_state;
/* Outputs of NeRF.render (src/NeRF.py:134) for S = n_samples_coarse + n_samples_fine samples per ray, plus depth / acc;
 * any pointer may be null (a caller that only needs rgb, weights, depth, acc moves 24 instead of 44 bytes per sample). */
typedef struct nerf_render_outs {
  float* rgb;     /* (N,3) */
  float* weights; /* (N,S) */
  float* cumprod; /* (N,S) */
  float* alpha;   /* (N,S) */
  float* rgb_s;   /* (N,S,3) */
  float* z;       /* (N,S) */
  float* depth;   /* (N) = sum w z (src/ExecutionRun.py:346) */
  float* acc;     /* (N) = sum w */
} nerf_render_outs;
/* Loss and optimizer of the train step. */
typedef struct nerf_train_cfg {
  float coarse_loss_weight; /* loss = coarse_loss_weight * MSE_c + MSE_f: 1 for NeRF (src/NeRF.py:151-157), 2 for DietNeRF's
                               aliased sum (src/DietNeRF.py:164-171) */
  int32_t stop_grad_z;      /* 1: detach the importance samples (NOT the reference's behaviour) */
  int32_t accumulate_grads; /* 1: add to `grads` instead of zeroing it first (DietNeRF's consistency term came before) */
  float learning_rate, beta_1, beta_2, epsilon; /* Keras Adam (defaults 1e-3, 0.9, 0.999, 1e-7) */
} nerf_train_cfg;

/* NeRF.render (src/NeRF.py:109-134) for n_rays rays in one call: stratified z -> coarse render_rays -> inverse-CDF
 * samples -> sort(concat) -> fine render_rays.  params_* are needed in NERF_MODE_FP32, packed_* (nerf_pack_weights /
 * _fp16) in the tensor-core modes; the *_f pair is ignored when n_samples_fine == 0.  workspace: 256-byte aligned,
 * nerf_render_workspace_bytes (-1 for a bad config).  Enqueues the same kernels as the separate entry points. */
int64_t nerf_render_workspace_bytes(const nerf_net_cfg* cfg, const nerf_render_cfg* rc, int64_t n_rays);
int nerf_render_fused_fwd(const nerf_net_cfg* cfg, const nerf_render_cfg* rc, const float* params_c,
                          const void* packed_c, const float* params_f, const void* packed_f, const float* origs4,
                          const float* dirs4, int64_t n_rays, const nerf_rng_state* rng,
                          const nerf_render_outs* outs, void* workspace, void* stream);

/* NeRF.train_step (src/NeRF.py:136-178) / DietNeRF's ray loss (src/DietNeRF.py:159-172) on this GPU's n_rays rays of a
 * global batch of n_total_rays (gradients are normalised by the global count, so ranks only have to sum them).
 * grads: [sum sq err coarse, sum sq err fine, 0, 0 | d params_c | d params_f] (4 + 1 or 2 x nerf_param_count floats).
 * adam_m / adam_v (same length as the parameter part, [coarse | fine]) non-null: the Adam update of step adam_t (1-based)
 * is applied and the bf16 packs are refreshed; null: gradients only (the caller all-reduces, then nerf_adam_step).
 * metrics4_or_null: nerf_train_metrics of the sums in `grads` (per-rank sums unless n_total_rays == n_rays).
 * side_stream_or_null: a second stream of the caller's (tensor-core mode with a fine network): the fine network's
 * HBM-bound weight-gradient kernel, its Adam step and pack refresh are enqueued there, under the coarse backward, and
 * joined before the call's last kernels; everything the call enqueued is complete when `stream` reaches its end.
 * Same results with and without it.  After an error return the work already enqueued on the two streams is not
 * joined: synchronise both before reusing the buffers. */
int64_t nerf_train_workspace_bytes(const nerf_net_cfg* cfg, const nerf_render_cfg* rc, int64_t n_rays);
int nerf_train_step_fused(const nerf_net_cfg* cfg, const nerf_render_cfg* rc, const nerf_train_cfg* tc,
                          float* params_c, void* packed_c, float* params_f, void* packed_f, const float* origs4,
                          const float* dirs4, const float* target_rgb, int64_t n_rays, int64_t n_total_rays,
                          const nerf_rng_state* rng, float* grads, float* adam_m, float* adam_v, int64_t adam_t,
                          float* metrics4_or_null, void* workspace, void* side_stream_or_null, void* stream);

/* The ray-sharded train step of one rank as ONE call: nerf_train_step_fused with the optimizer step replaced by the
 * gradient exchange over NVLink peer memory (nerf_peer_barrier + nerf_peer_reduce_adam: sum over the ranks in rank order,
 * fused with Adam).  `grads` must be THIS rank's buffer of the symmetric set, i.e. grads_dev[rank]; the fine network's
 * exchange + Adam + pack refresh run on the side stream under the coarse backward, [loss sums | coarse gradients] at the
 * tail; metrics4 are formed from the summed loss sums (peer->reduced_sums).  epoch must grow from step to step (the
 * 1-based Adam step does); the caller alternates between two symmetric buffer sets from step to step (a buffer is rewritten
 * two barriers after its last remote read).  Same kernels and results as the host package's sharded call sequence. */
typedef struct nerf_peer_exchange {
  void* const* pads_dev;   /* nerf_peer_barrier: DEVICE array of `world` signal pads */
  void* const* grads_dev;  /* nerf_peer_reduce_adam: DEVICE array of the `world` peer-mapped gradient buffers */
  float* reduced_sums;     /* (4) device floats: the loss sums added over the ranks */
  int32_t rank, world;
  uint32_t epoch;
  uint32_t reserved;
} nerf_peer_exchange;
int nerf_train_step_fused_sharded(const nerf_net_cfg* cfg, const nerf_render_cfg* rc, const nerf_train_cfg* tc,
                                  float* params_c, void* packed_c, float* params_f, void* packed_f, const float* origs4,
                                  const float* dirs4, const float* target_rgb, int64_t n_rays, int64_t n_total_rays,
                                  const nerf_rng_state* rng, float* grads, float* adam_m, float* adam_v, int64_t adam_t,
                                  float* metrics4_or_null, void* workspace, void* side_stream_or_null,
                                  const nerf_peer_exchange* peer, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NERF_B200_H */
