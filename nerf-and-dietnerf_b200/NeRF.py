"""NeRF model: drop-in for the render/train step of the reference's src/NeRF.py, on hand-written sm_100a kernels.

Keeps the reference's surface -- ``NeRF(net_config, render_config, near_boundary, far_boundary)`` (:27),
``render`` (:109), ``call`` (:96), ``train_step`` (:136), ``render_rays`` (:180), ``render_image`` (:190),
``init_network`` (:59), ``get_nerf_model_path`` (:343) and the YAML keys of src/ConfigurationKeys.py -- but is not a
Keras model: parameters are two flat fp32 device vectors, the train step is an explicit forward + hand-written
backward (no autograd tape) over the C ABI, and ray batches can be sharded over the GPUs of one box with a single
NCCL all-reduce of the gradients per step.
"""
import contextlib
import ctypes
import math
from pathlib import Path
from typing import Dict

import numpy as np
import torch

from . import _lib
from ._lib import MODE_BF16, NetCfg, call, f32c, ptr
from .ConfigurationKeys import (HIDDEN_LAYER_DIM, LAST_HIDDEN_LAYER_DIM, LEAKY_RELU_ALPHA, N_ANGLES_FOR_MODEL,
                                N_POS_ENC_DIM_XYZ, N_POS_ENC_VIEW_DIR, N_RAYS_IN_BATCH_RENDER, N_RAYS_IN_BATCH_TRAIN,
                                N_RENDER_SAMPLES_COARSE, N_RENDER_SAMPLES_FINE)
from .network import DEFAULT_MODE, NerfMLP
from .optimizers import Adam
from .parallel import allreduce_sum_, shard_bounds
from .UtilsCV import get_rays_directions, get_z_vals_from_prob_dist_func, get_z_values, rng
from .UtilsNeuralRadianceField import get_psnr, split_to_batches
from . import UtilsNeuralRadianceField as _unrf

DIRNAME_TO_SAVE_WEIGHTS = 'saved_weights'
NAME_NERF_MODEL_FILE = 'NeRF_model_epoch_{:03}.h5'


def net_cfg_from_dict(net_config: Dict) -> NetCfg:
    """YAML ``neural_net:`` block -> struct nerf_net_cfg (strict key lookups like the reference, KeyError if missing)."""
    return NetCfg(int(net_config[N_POS_ENC_DIM_XYZ]), int(net_config[N_POS_ENC_VIEW_DIR]),
                  int(net_config[N_ANGLES_FOR_MODEL]), int(net_config[HIDDEN_LAYER_DIM]),
                  int(net_config[LAST_HIDDEN_LAYER_DIM]), float(net_config[LEAKY_RELU_ALPHA]))


class _StepWorkspace:
    """Device buffers of one train step for a fixed number of rays (allocated once, reused every step)."""

    def __init__(self, model, n, sc=None, sf=None, n_new=None):
        """``sc``/``sf``: samples per ray the coarse / fine network sees (default: the train step's, src/NeRF.py:146,155);
        ``n_new``: importance samples drawn when the fine network sees sort(concat(new, coarse)) (in-tape render)."""
        dev = model.device
        f = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
        sc = model.n_render_samples_coarse if sc is None else sc
        sf = model.n_render_samples_fine if sf is None else sf
        mc = model.model_coarse
        self.n = n
        self.z_c = f(n, sc)
        fused = mc.tensor_core                # tensor-core modes: encodings are computed inside the MLP kernel
        self.xyz_c = None if fused else f(n * sc, mc.dx)
        self.view_c = f(n * sc, mc.dv) if (mc.dv and not fused) else None
        self.raw_c = f(n, sc, 4)
        self.rgb_c = f(n, 3)
        self.w_c = f(n, sc)
        self.d_rgb_c = f(n, 3)
        self.d_raw_c = f(n, sc, 4)
        self.saved_c = torch.empty(max(mc.saved_bytes(n * sc), 16), dtype=torch.uint8, device=dev)
        self.sums = torch.zeros(2, dtype=torch.float32, device=dev)
        m_max = n * max(sc, sf)
        self.ws_fwd = torch.empty(max(mc.workspace_bytes(m_max, False), 16), dtype=torch.uint8, device=dev)
        self.ws_bwd = torch.empty(max(mc.workspace_bytes(m_max, True), 16), dtype=torch.uint8, device=dev)
        self.ws_bwd_c = self.ws_bwd           # coarse backward workspace (its own buffer when the fine dW overlaps it)
        if model.model_fine is not None:
            self.z_f = f(n, sf)
            self.u = f(n, sf)
            self.perm = torch.empty((n, sf), dtype=torch.int32, device=dev)
            self.xyz_f = None if fused else f(n * sf, mc.dx)
            self.view_f = f(n * sf, mc.dv) if (mc.dv and not fused) else None
            self.raw_f = f(n, sf, 4)
            self.rgb_f = f(n, 3)
            self.d_rgb_f = f(n, 3)
            self.d_raw_f = f(n, sf, 4)
            self.d_z_f = f(n, sf)
            self.d_xyz_f = f(n * sf, mc.dx)
            self.d_w_c = f(n, sc)
            self.saved_f = torch.empty(max(mc.saved_bytes(n * sf), 16), dtype=torch.uint8, device=dev)
            if n_new is not None:
                self.z_new = f(n, n_new)
                self.d_z_new = f(n, n_new)
                self.u = f(n, n_new)
                self.perm = torch.empty((n, n_new), dtype=torch.int32, device=dev)
                self.rank = torch.empty((n, n_new), dtype=torch.int32, device=dev)
            if mc.tensor_core:
                # the fine network's weight-gradient kernel runs on a side stream under the coarse backward: the two
                # must not share the dZ workspace
                self.ws_bwd_c = torch.empty(max(mc.workspace_bytes(n * sc, True), 16), dtype=torch.uint8, device=dev)


class NeRF:
    """The NeRF model (coarse + optional fine network) with its render and train step."""

    # loss = COARSE_LOSS_WEIGHT * MSE_coarse + MSE_fine  (1 for NeRF, src/NeRF.py:151-157; DietNeRF overrides)
    COARSE_LOSS_WEIGHT = 1.0
    # True: the one-stream backward issues its two halves as separate C-ABI calls (bench.py times them one by one)
    split_bwd_calls = False
    # True (default): on one GPU ``train_step_local`` IS one call of ``nerf_train_step_fused`` (the same kernels, enqueued by
    # the C side: half the host time per step); False: the host package's own call sequence (what a sharded run uses, whose
    # all-reduces sit between the calls)
    use_fused_step = True
    rays_in_kernel = True      # render_image_lean: rays generated inside the MLP kernel (tensor-core modes)

    def __init__(self, net_config: Dict, render_config: Dict, near_boundary: float, far_boundary: float, *,
                 mode: str = DEFAULT_MODE, device=None, seed=None, stop_grad_z: bool = False):
        """
        :param net_config:      ``neural_net`` block of the config file.
        :param render_config:   ``render`` block of the config file.
        :param mode:            (extension) "fp16" (default): tensor-core path with fp16 forward operands (train + render;
                                the reference's mixed_float16 arithmetic), bf16 gradients; "bf16": bf16 operands
                                everywhere; "fp32": SIMT parity path.
        :param stop_grad_z:     (extension, default False = reference behaviour) detach the importance samples.
        """
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.mode = mode
        self.net_cfg = net_cfg_from_dict(net_config)
        self.model_coarse = NeRF.init_network(net_config, mode=mode, device=self.device,
                                              seed=None if seed is None else 2 * seed)
        if render_config[N_RENDER_SAMPLES_FINE] > 0:
            self.model_fine = NeRF.init_network(net_config, mode=mode, device=self.device,
                                                seed=None if seed is None else 2 * seed + 1)
        else:
            self.model_fine = None

        self.batch_size_render = net_config[N_RAYS_IN_BATCH_RENDER]
        self.batch_size_train = net_config[N_RAYS_IN_BATCH_TRAIN]
        self.near_boundary = float(near_boundary)
        self.far_boundary = float(far_boundary)
        self.n_render_samples_coarse = int(render_config[N_RENDER_SAMPLES_COARSE])
        self.n_render_samples_fine = int(render_config[N_RENDER_SAMPLES_FINE])
        self.n_pos_enc_dim_xyz = self.net_cfg.n_pos_enc_xyz
        self.n_pos_enc_view_dir = self.net_cfg.n_pos_enc_view
        self.n_angles_for_model = self.net_cfg.n_angles
        self.stop_grad_z = stop_grad_z

        self.optimizer = None
        self.seed = 0 if seed is None else int(seed)
        self.step_counter = 0           # Philox `step` of the next train step
        self._ws = {}
        self._grads = None
        self._overlap_allreduce = False
        self._fine_allreduce = None
        self._fine_update = None        # set by train_step_local: the fine network's optimizer step, run early
        self._peer_step = None          # set by train_step_local while a step exchanges its gradients over peer memory
        self._fine_updated = False
        self._side = None
        self.overlap_dw = True          # False: every kernel of the step runs on one stream (per-kernel timing)
        # data-parallel state (set by distribute())
        self.world_size, self.rank, self._process_group = 1, 0, None
        self._peer = None               # parallel.PeerExchange once distribute() set one up

    # ---- construction helpers --------------------------------------------------------------------------------
    @staticmethod
    def init_network(net_config: Dict, mode: str = DEFAULT_MODE, device=None, seed=None) -> NerfMLP:
        """Initialise one network from the config (xyz+view when n_angles_for_model > 0, else xyz only)."""
        return NerfMLP(net_cfg_from_dict(net_config), mode=mode, device=device, seed=seed)

    @staticmethod
    def get_nerf_model_path(save_location: Path, epoch_number: int) -> Path:
        return Path(save_location) / DIRNAME_TO_SAVE_WEIGHTS / NAME_NERF_MODEL_FILE.format(epoch_number)

    def save_weights(self, filepath):
        """Keras ``Model.save_weights`` for this model: ``.h5`` -> the reference's Keras-2.7 HDF5 checkpoint layout
        (groups ``model`` / ``model_1``, layers ``dense`` ... ``dense_21``; see h5weights.py), anything else -> ``.npz``
        with the two flat vectors."""
        self.weights_snapshot().write(filepath)

    def weights_snapshot(self):
        """Host copy of the weights (one small D2H copy) whose ``write(path)`` can run on another thread while training
        continues."""
        pc = self.model_coarse.params.detach().cpu().numpy()
        pf = self.model_fine.params.detach().cpu().numpy() if self.model_fine is not None else None
        shapes = list(self.model_coarse.shapes)

        class _Snapshot:
            @staticmethod
            def write(filepath):
                from . import h5weights
                filepath = str(filepath)
                if filepath.endswith((".h5", ".hdf5")):
                    h5weights.save_flat_params(filepath, pc, pf, shapes)
                else:
                    import numpy as np
                    arrays = {"params_coarse": pc}
                    if pf is not None:
                        arrays["params_fine"] = pf
                    with open(filepath, "wb") as f:
                        np.savez(f, **arrays)
        return _Snapshot

    def load_weights(self, filepath):
        """Keras ``Model.load_weights``: reads a checkpoint written by ``save_weights`` OR by the reference's Keras
        (e.g. ``Results/.../saved_weights/NeRF_model_epoch_095.h5``)."""
        from . import h5weights
        filepath = str(filepath)
        if filepath.endswith((".h5", ".hdf5")):
            pc, pf = h5weights.load_flat_params(filepath)
        else:
            import numpy as np
            with np.load(filepath) as z:
                pc, pf = z["params_coarse"], (z["params_fine"] if "params_fine" in z.files else None)
        self.model_coarse.set_params(pc)
        if self.model_fine is not None:
            if pf is None:
                raise ValueError(f"{filepath} holds one network but the model has a fine network")
            self.model_fine.set_params(pf)
        return self

    def compile(self, optimizer=None, **kwargs):
        """Keras-style: attach the optimizer (``Adam(learning_rate)``)."""
        self.optimizer = optimizer if optimizer is not None else Adam(**kwargs)
        return self

    def distribute(self, process_group=None, peer_exchange=None):
        """Shard every train batch over the ranks of ``process_group`` (one process per GPU, NCCL).

        ``peer_exchange`` (default: on when the GPUs of the box can map each other's memory; env NERF_PEER_EXCHANGE=0/1
        overrides): the gradient exchange runs as our own one-shot reduce + Adam kernel over NVLink peer memory
        (parallel.PeerExchange) instead of NCCL all-reduces followed by the Adam kernels."""
        import os
        import torch.distributed as dist
        self._process_group = process_group
        self.world_size = dist.get_world_size(process_group)
        self.rank = dist.get_rank(process_group)
        self._peer = None
        env = os.environ.get("NERF_PEER_EXCHANGE")
        want = (env != "0") if env is not None else (peer_exchange is not False)
        if want and self.world_size > 1 and self.device.type == "cuda" and dist.get_backend(process_group) == "nccl":
            from .parallel import PeerExchange
            n = self.model_coarse.n_params + (self.model_fine.n_params if self.model_fine is not None else 0)
            try:
                self._peer = PeerExchange(4 + n, self.device, process_group)
            except Exception as e:                          # no peer access / no symmetric memory: NCCL it is
                if peer_exchange or env == "1":
                    raise
                if self.rank == 0:
                    print(f"NeRF.distribute: peer exchange unavailable ({type(e).__name__}: {e}); using NCCL all-reduce")
                self._peer = None
            # the ranks must agree (a rank that failed alone would wait in a barrier nobody else enters)
            ok = torch.tensor([1 if self._peer is not None else 0], device=self.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=process_group)
            if int(ok.item()) == 0:
                self._peer = None
        return self

    @property
    def trainable_variables(self):
        v = list(self.model_coarse.trainable_variables)
        if self.model_fine is not None:
            v += self.model_fine.trainable_variables
        return v

    def get_config(self):
        return {
            'near_boundary': self.near_boundary, 'far_boundary': self.far_boundary,
            'n_render_samples_coarse': self.n_render_samples_coarse,
            'n_render_samples_fine': self.n_render_samples_fine,
            'n_positional_encoding': self.n_pos_enc_dim_xyz, 'n_enc_phi_theta': self.n_pos_enc_view_dir,
            'n_angles_for_model': self.n_angles_for_model, 'mode': self.mode,
        }

    # ---- rendering ---------------------------------------------------------------------------------------------
    def render_rays(self, model, rays_orig, rays_dirs, z):
        """Render rays with ``model`` (5-tuple of ray_marching), autograd-aware.  Mirrors src/NeRF.py:180-188."""
        return _unrf.render_rays(model, rays_orig, rays_dirs, z, self.n_pos_enc_dim_xyz, self.n_pos_enc_view_dir,
                                 self.n_angles_for_model)

    def _coarse_pass(self, model, rays_orig, rays_dirs, n, n_c, seed, step, ray_offset, jitter, z, raw, saved=None, mode_id=None,
                     xyz=None, view=None, ws=None):
        """Stratified depths + coarse MLP on them, into the caller's ``z`` (n, n_c) and ``raw`` (n, n_c, 4) buffers.
        Tensor-core modes with the Philox stream: ONE kernel (``nerf_mlp_fwd_rays_stratified``: get_z_values, sample_along_rays,
        both encodings and the network; north_star's kernel (1) fused into kernel (2)).  Explicit ``jitter`` or the fp32
        mode: ``nerf_stratified_z`` first -- same depths, bit for bit."""
        mode_id = model.mode_id if mode_id is None else mode_id
        if model.tensor_core and jitter is None:
            call("nerf_mlp_fwd_rays_stratified", model.cfg_ref, ptr(model.packed_for(model.params, half=mode_id == _lib.MODE_FP16)),
                 ptr(rays_orig), ptr(rays_dirs), self.near_boundary, self.far_boundary, int(seed), int(step), int(ray_offset),
                 n, n_c, ptr(z), ptr(raw), ptr(saved), mode_id)
            return
        call("nerf_stratified_z", self.near_boundary, self.far_boundary, n, n_c, ptr(jitter), int(seed or 0), int(step or 0),
             int(ray_offset), ptr(z))
        if model.tensor_core:
            call("nerf_mlp_fwd_rays", model.cfg_ref, ptr(model.packed_for(model.params, half=mode_id == _lib.MODE_FP16)),
                 ptr(rays_orig), ptr(rays_dirs), ptr(z), n, n_c, ptr(raw), ptr(saved), mode_id)
        else:
            if xyz is None:
                xyz = torch.empty((n * n_c, model.dx), dtype=torch.float32, device=z.device)
                view = torch.empty((n * n_c, model.dv), dtype=torch.float32, device=z.device) if model.dv else None
            if ws is None:
                ws = model._buffer("ws_fwd", model.workspace_bytes(n * n_c, False))
            call("nerf_encode_samples", model.cfg_ref, ptr(rays_orig), ptr(rays_dirs), ptr(z), n, n_c, ptr(xyz), ptr(view))
            call("nerf_mlp_fwd", model.cfg_ref, ptr(model.params), ptr(model.packed_for(model.params)), ptr(xyz), ptr(view),
                 n * n_c, ptr(raw), ptr(saved), ptr(ws), mode_id)

    def _coarse_render(self, rays_orig, rays_dirs, n, n_c, seed, step, ray_offset, jitter, lean=False):
        """Coarse half of ``render``: (z, ray_marching outputs)."""
        z = torch.empty((n, n_c), dtype=torch.float32, device=self.device)
        raw = torch.empty((n, n_c, 4), dtype=torch.float32, device=self.device)
        mc = self.model_coarse
        self._coarse_pass(mc, rays_orig, rays_dirs, n, n_c, seed, step, ray_offset, jitter, z, raw, mode_id=mc.infer_mode_id)
        return z, (_unrf.ray_marching_lean(raw, z) if lean else _unrf.ray_marching(raw, z))

    def _render_rays_fused(self, model, rays_orig, rays_dirs, z, lean=False):
        """Inference-only fast path: fused encode -> MLP -> compositing, no autograd bookkeeping."""
        n, s = z.shape
        dev = z.device
        raw = torch.empty((n, s, 4), dtype=torch.float32, device=dev)
        if model.tensor_core:
            half = model.infer_mode_id == _lib.MODE_FP16
            call("nerf_mlp_fwd_rays", model.cfg_ref, ptr(model.packed_for(model.params, half=half)), ptr(rays_orig),
                 ptr(rays_dirs), ptr(z), n, s, ptr(raw), None, model.infer_mode_id)
        else:
            xyz = torch.empty((n * s, model.dx), dtype=torch.float32, device=dev)
            view = torch.empty((n * s, model.dv), dtype=torch.float32, device=dev) if model.dv else None
            call("nerf_encode_samples", model.cfg_ref, ptr(rays_orig), ptr(rays_dirs), ptr(z), n, s, ptr(xyz), ptr(view))
            ws = model._buffer("ws_fwd", model.workspace_bytes(n * s, False))
            call("nerf_mlp_fwd", model.cfg_ref, ptr(model.params), ptr(model.packed_for(model.params)), ptr(xyz),
                 ptr(view), n * s, ptr(raw), None, ptr(ws), model.mode_id)
        with torch.no_grad():
            if lean:
                return _unrf.ray_marching_lean(raw, z)
            return _unrf.ray_marching(raw, z)

    def render(self, rays_orig, rays_dirs, n_render_samples_c=None, n_render_samples_f=None, *, seed=None, step=0,
               ray_offset=0, jitter=None, u=None):
        """Render rays: returns (render_result, weights, cumprod, alpha, rgb, z) like src/NeRF.py:109-134.

        Keyword-only extensions pin the random stream (``seed/step/ray_offset``) or the draws themselves
        (``jitter`` (N,S_c), ``u`` (N,N_f)); the reference draws unseeded tf.random.uniform values.
        """
        rays_orig, rays_dirs = f32c(rays_orig, self.device), f32c(rays_dirs, self.device)
        n = rays_orig.shape[0]
        n_c = n_render_samples_c if n_render_samples_c else self.n_render_samples_coarse
        if seed is None and jitter is None:
            seed, step = rng.next_step()
        with torch.no_grad():
            z, out = self._coarse_render(rays_orig, rays_dirs, n, n_c, seed, step, ray_offset,
                                         None if jitter is None else f32c(jitter, self.device))
            if self.model_fine is not None:
                n_f = n_render_samples_f if n_render_samples_f else self.n_render_samples_fine
                z_from_dist = get_z_vals_from_prob_dist_func(out[1], z, n_f, u=u, seed=seed if u is None else None,
                                                             step=step, ray_offset=ray_offset)
                z_all = torch.empty((n, n_f + n_c), dtype=torch.float32, device=self.device)
                call("nerf_merge_sorted", ptr(z_from_dist), n_f, ptr(z), n_c, n, ptr(z_all))
                z = z_all
                out = self._render_rays_fused(self.model_fine, rays_orig, rays_dirs, z)
        return out + (z,)

    # ---- the whole path through ONE C-ABI call each (what a non-Python caller binds) -----------------------------------
    def _render_cfg(self, n_c=None, n_f=None, infer=True):
        mc = self.model_coarse
        mode = (mc.infer_mode_id if infer else mc.mode_id) if mc.tensor_core else _lib.MODE_FP32
        n_f = (n_f if n_f else self.n_render_samples_fine) if self.model_fine is not None else 0
        return _lib.RenderCfg(float(self.near_boundary), float(self.far_boundary),
                              int(n_c if n_c else self.n_render_samples_coarse), int(n_f), int(mode))

    def render_fused(self, rays_orig, rays_dirs, n_render_samples_c=None, n_render_samples_f=None, *, seed, step=0,
                     ray_offset=0, lean=False):
        """``render`` as one call of ``nerf_render_fused_fwd``: the same kernels in the same order, enqueued by the C
        side, so the results are bit-identical to ``render(..., seed=, step=, ray_offset=)``.  Returns the 6-tuple of
        ``render``, or with ``lean`` (rgb, weights, depth, acc, z)."""
        o, d = f32c(rays_orig, self.device), f32c(rays_dirs, self.device)
        n = o.shape[0]
        mc, mf = self.model_coarse, self.model_fine
        rc = self._render_cfg(n_render_samples_c, n_render_samples_f)
        s = rc.n_samples_coarse + rc.n_samples_fine
        f = lambda *shape: torch.empty(shape, dtype=torch.float32, device=self.device)
        rgb, weights, z = f(n, 3), f(n, s), f(n, s)
        extra = (f(n), f(n)) if lean else (f(n, s), f(n, s), f(n, s, 3))
        outs = _lib.RenderOuts(rgb=ptr(rgb), weights=ptr(weights), z=ptr(z))
        if lean:
            outs.depth, outs.acc = ptr(extra[0]), ptr(extra[1])
        else:
            outs.cumprod, outs.alpha, outs.rgb_s = ptr(extra[0]), ptr(extra[1]), ptr(extra[2])
        half = mc.tensor_core and rc.mode == _lib.MODE_FP16
        nbytes = int(_lib.load().nerf_render_workspace_bytes(mc.cfg_ref, ctypes.byref(rc), n))
        if nbytes < 0:
            raise _lib.NerfLibraryError("nerf_render_workspace_bytes: unsupported configuration")
        ws = mc._buffer("ws_render_fused", nbytes + 256)
        ws_ptr = (ws.data_ptr() + 255) & ~255
        rng_state = _lib.RngState(int(seed), int(ray_offset), int(step), 0)
        call("nerf_render_fused_fwd", mc.cfg_ref, ctypes.byref(rc), ptr(mc.params), ptr(mc.packed_for(mc.params, half=half)),
             ptr(mf.params) if mf is not None else None,
             ptr(mf.packed_for(mf.params, half=half)) if mf is not None else None, ptr(o), ptr(d), n,
             ctypes.byref(rng_state), ctypes.byref(outs), ws_ptr)
        if lean:
            return rgb, weights, extra[0], extra[1], z
        return rgb, weights, extra[0], extra[1], extra[2], z

    def train_step_fused(self, rays_orig, rays_dirs, real_rgb, *, n_total_rays=None, ray_offset=0, update=True, peer=None):
        """``train_step_local`` as one call of ``nerf_train_step_fused`` on the current stream, with the model's side
        stream handed to the C side for the fine network's weight gradients (``overlap_dw = False``: one stream).
        Single GPU with ``update``; with ``update=False`` only the gradients and sums are produced (flat buffer
        ``_grad_buffer()``) for a caller that all-reduces them itself.  ``peer`` (a ``parallel.PeerExchange``): the
        ray-sharded step of this rank as one call of ``nerf_train_step_fused_sharded`` -- the gradients go to the
        exchange's symmetric buffer and the optimizer step is the NVLink peer reduce fused with Adam."""
        if self.optimizer is None:
            raise RuntimeError("call compile(optimizer=Adam(lr)) before train_step")
        o, d, y = f32c(rays_orig, self.device), f32c(rays_dirs, self.device), f32c(real_rgb, self.device)
        n = o.shape[0]
        n_total = n if n_total_rays is None else int(n_total_rays)
        mc, mf = self.model_coarse, self.model_fine
        rc = self._render_cfg(infer=False)
        opt = self.optimizer
        tcfg = _lib.TrainCfg(float(self.COARSE_LOSS_WEIGHT), 1 if self.stop_grad_z else 0,
                             1 if getattr(self, "_keep_grads", False) else 0, opt.learning_rate, opt.beta_1, opt.beta_2,
                             opt.epsilon)
        n_all = mc.n_params + (mf.n_params if mf is not None else 0)
        g = self._grad_buffer() if peer is None else peer.grads
        opt._state(n_all, self.device)
        nbytes = int(_lib.load().nerf_train_workspace_bytes(mc.cfg_ref, ctypes.byref(rc), n))
        if nbytes < 0:
            raise _lib.NerfLibraryError("nerf_train_workspace_bytes: unsupported configuration")
        ws = mc._buffer("ws_train_fused", nbytes + 256)
        ws_ptr = (ws.data_ptr() + 255) & ~255
        rng_state = _lib.RngState(int(self.seed), int(ray_offset), int(self.step_counter), 0)
        out = torch.empty(4, dtype=torch.float32, device=self.device)
        side = self._side_stream() if mf is not None else None      # None with overlap_dw = False: one stream
        args = (mc.cfg_ref, ctypes.byref(rc), ctypes.byref(tcfg), ptr(mc.params),
                ptr(mc.packed_for(mc.params)), ptr(mf.params) if mf is not None else None,
                ptr(mf.packed_for(mf.params)) if mf is not None else None, ptr(o), ptr(d), ptr(y), n, n_total,
                ctypes.byref(rng_state), ptr(g), ptr(opt._m) if update else None, ptr(opt._v) if update else None,
                opt.iterations + 1, ptr(out), ws_ptr, side.cuda_stream if side is not None else None)
        if peer is None:
            call("nerf_train_step_fused", *args)
        else:
            if not update:
                raise ValueError("the sharded call applies the optimizer step: update=False has no meaning with peer")
            pcfg = _lib.PeerExchangeCfg(int(peer.pad_handle.buffer_ptrs_dev), int(peer.handles[peer.parity].buffer_ptrs_dev),
                                        ptr(peer.sums), peer.rank, peer.world, (opt.iterations + 1) & 0xFFFFFFFF, 0)
            call("nerf_train_step_fused_sharded", *args, ctypes.byref(pcfg))
            peer.flip()
        if update:
            opt.iterations += 1
            # the C side stepped the parameters in place AND refreshed the 16-bit packs this mode reads: only the other
            # precision's regions are stale now
            mc.mark_packed()
            if mf is not None:
                mf.mark_packed()
        self.step_counter += 1
        return self._metrics_dict(out)

    def call(self, inputs, training=None, mask=None):
        rays_orig, rays_dirs = inputs
        return self.render(rays_orig, rays_dirs)[0]

    __call__ = call

    def render_image(self, c2w, fov, h, w, batch_size_input=None, n_render_samples_c=None, n_render_samples_f=None, *,
                     seed=None, step=0):
        """Render an h x w image: returns (rgb (h,w,3), weights, cumprod, alpha (h,w,S), rgb_s (h,w,S,3), z (h,w,S))."""
        rays_dirs, rays_orig = get_rays_directions(h, w, fov, c2w, return_origins=True)
        rays_dirs = rays_dirs.reshape(h * w, 4)
        batch_size = batch_size_input if batch_size_input else self.batch_size_render
        if seed is None:
            seed, step = rng.next_step()
        parts = []
        offset = 0
        for o, d in zip(split_to_batches(rays_orig, batch_size), split_to_batches(rays_dirs, batch_size)):
            parts.append(self.render(o, d, n_render_samples_c, n_render_samples_f, seed=seed, step=step,
                                     ray_offset=offset))
            offset += o.shape[0]
        cat = [torch.cat([p[i] for p in parts], dim=0) for i in range(6)]
        return (cat[0].reshape(h, w, 3), cat[1].reshape(h, w, -1), cat[2].reshape(h, w, -1), cat[3].reshape(h, w, -1),
                cat[4].reshape(h, w, -1, 3), cat[5].reshape(h, w, -1))

    def render_image_lean(self, c2w, fov, h, w, batch_size_input=None, n_render_samples_c=None,
                          n_render_samples_f=None, *, seed=None, step=0, ray_begin=0, n_rays=None):
        """Video-path render (extension): only rgb (n,3), depth (n) = sum w z (src/ExecutionRun.py:346) and acc (n),
        for rays [ray_begin, ray_begin+n_rays) of the frame -- the unit a GPU takes when a frame is row-sharded."""
        n_total = h * w if n_rays is None else int(n_rays)
        batch_size = batch_size_input if batch_size_input else self.batch_size_render
        if seed is None:
            seed, step = rng.next_step()
        n_c = n_render_samples_c if n_render_samples_c else self.n_render_samples_coarse
        n_f = n_render_samples_f if n_render_samples_f else self.n_render_samples_fine
        rgbs, depths, accs = [], [], []
        mc, mf = self.model_coarse, self.model_fine
        if self.rays_in_kernel and mc.tensor_core and (mf is None or mf.tensor_core):
            # tensor-core modes: the MLP kernel generates its own rays from the camera (nerf_mlp_fwd_camera) -- ray generation,
            # stratified depths, positions, encodings and the network are one kernel, no ray buffer exists
            c2w_np = np.ascontiguousarray(c2w.detach().cpu().numpy() if isinstance(c2w, torch.Tensor) else np.asarray(c2w),
                                          dtype=np.float32)
            if c2w_np.shape != (4, 4):
                raise ValueError("c2w must be a 4x4 camera-to-world matrix")
            c2w_p = c2w_np.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
            f = lambda *shape: torch.empty(shape, dtype=torch.float32, device=self.device)

            def mlp(net, off, n, s, z, z_out, raw):
                half = net.infer_mode_id == _lib.MODE_FP16
                call("nerf_mlp_fwd_camera", net.cfg_ref, ptr(net.packed_for(net.params, half=half)), c2w_p, float(fov), int(h),
                     int(w), int(off), n, s, ptr(z), self.near_boundary, self.far_boundary, int(seed), int(step), ptr(z_out),
                     ptr(raw), net.infer_mode_id)
            with torch.no_grad():
                for s0 in range(0, n_total, batch_size):
                    n, off = min(batch_size, n_total - s0), ray_begin + s0
                    z, raw = f(n, n_c), f(n, n_c, 4)
                    mlp(mc, off, n, n_c, None, z, raw)
                    if mf is None:
                        rgb, wts, depth, acc = _unrf.ray_marching_lean(raw, z)
                    else:
                        z_all, raw_f = f(n, n_f + n_c), f(n, n_f + n_c, 4)
                        if n_f <= 256:
                            # coarse weights + importance draws + sort + merge with the coarse depths: one launch
                            call("nerf_hierarchical_sample", ptr(raw), ptr(z), n, n_c, n_f, int(seed), int(step), int(off),
                                 ptr(z_all))
                        else:
                            rgb, wts, depth, acc = _unrf.ray_marching_lean(raw, z)
                            z_f = get_z_vals_from_prob_dist_func(wts, z, n_f, seed=seed, step=step, ray_offset=off)
                            call("nerf_merge_sorted", ptr(z_f), n_f, ptr(z), n_c, n, ptr(z_all))
                        mlp(mf, off, n, n_f + n_c, z_all, None, raw_f)
                        rgb, wts, depth, acc = _unrf.ray_marching_lean(raw_f, z_all)
                    rgbs.append(rgb)
                    depths.append(depth)
                    accs.append(acc)
            if not rgbs:
                return f(0, 3), f(0), f(0)
            return torch.cat(rgbs), torch.cat(depths), torch.cat(accs)
        dirs, orig = get_rays_directions(h, w, fov, c2w, ray_begin=ray_begin, n_rays=n_total, return_origins=True)
        with torch.no_grad():
            for s0 in range(0, n_total, batch_size):
                o, d = orig[s0:s0 + batch_size], dirs[s0:s0 + batch_size]
                n = o.shape[0]
                off = ray_begin + s0
                z, (rgb, wts, depth, acc) = self._coarse_render(o, d, n, n_c, seed, step, off, None, lean=True)
                if self.model_fine is not None:
                    z_f = get_z_vals_from_prob_dist_func(wts, z, n_f, seed=seed, step=step, ray_offset=off)
                    z_all = torch.empty((n, n_f + n_c), dtype=torch.float32, device=self.device)
                    call("nerf_merge_sorted", ptr(z_f), n_f, ptr(z), n_c, n, ptr(z_all))
                    rgb, wts, depth, acc = self._render_rays_fused(self.model_fine, o, d, z_all, lean=True)
                rgbs.append(rgb)
                depths.append(depth)
                accs.append(acc)
        if not rgbs:        # an empty ray range (a rank whose shard of the frame is empty)
            f = lambda *shape: torch.empty(shape, dtype=torch.float32, device=self.device)
            return f(0, 3), f(0), f(0)
        return torch.cat(rgbs), torch.cat(depths), torch.cat(accs)

    # ---- training ----------------------------------------------------------------------------------------------------
    def _workspace(self, n, sc=None, sf=None, n_new=None):
        key = n if sc is None else (n, sc, sf, n_new)
        ws = self._ws.get(key)
        if ws is None:
            ws = _StepWorkspace(self, n, sc, sf, n_new)
            self._ws[key] = ws
        return ws

    def _grad_buffer(self):
        """Flat gradient buffer [sq_err_coarse, sq_err_fine, 0, 0 | grads_coarse | grads_fine].  The two squared-error sums
        ride in FRONT of the coarse gradients so that the buffer splits into the two all-reduces of a step: the fine
        gradients (final before the coarse backward starts, overlapped with it) and [sums | coarse gradients]."""
        if self._peer is not None:
            return self._peer.grads        # symmetric memory: the peers read it in place
        if self._grads is None:
            n = self.model_coarse.n_params + (self.model_fine.n_params if self.model_fine is not None else 0)
            self._grads = torch.zeros(4 + n, dtype=torch.float32, device=self.device)
        return self._grads

    def _grad_views(self):
        g = self._grad_buffer()
        nc = self.model_coarse.n_params
        g_f = g[4 + nc:] if self.model_fine is not None else None
        return g[0:2], g[4:4 + nc], g_f

    def forward_backward(self, rays_orig, rays_dirs, real_rgb, *, n_total_rays=None, ray_offset=0, seed=None, step=None,
                         jitter=None, u=None, keep_grads=False):
        """Loss and parameter gradients of one batch (src/NeRF.py:145-164 without the optimizer).

        ``keep_grads``: add to the gradients already in the flat buffer (DietNeRF's consistency term) instead of
        starting from zero.

        Returns (grads_coarse, grads_fine, sums) where sums = [sum sq err coarse, sum sq err fine] over THIS
        shard and the gradients are already divided by the GLOBAL element count 3*n_total_rays.
        """
        n = rays_orig.shape[0]
        n_total = n if n_total_rays is None else int(n_total_rays)
        w = self._workspace(n)
        mc, mf = self.model_coarse, self.model_fine
        sc, sf = self.n_render_samples_coarse, self.n_render_samples_fine
        seed = self.seed if seed is None else seed
        step = self.step_counter if step is None else step
        if not keep_grads:
            self._grad_buffer().zero_()
        sums, g_c, g_f = self._grad_views()
        self._fine_allreduce = None
        o, d, y = rays_orig, rays_dirs, real_rgb

        # coarse forward
        self._coarse_pass(mc, o, d, n, sc, seed, step, ray_offset, jitter, w.z_c, w.raw_c, saved=w.saved_c, xyz=w.xyz_c,
                          view=w.view_c, ws=w.ws_fwd)
        # ray_marching + MSE (+ its gradient) in one launch each: the loss is formed where the ray's colour is reduced
        call("nerf_composite_mse_fwd", ptr(w.raw_c), ptr(w.z_c), ptr(y), n, sc, n_total, self.COARSE_LOSS_WEIGHT,
             ptr(w.rgb_c), ptr(w.w_c), ptr(sums[0:1]), ptr(w.d_rgb_c))
        d_w_c = None
        if mf is not None:
            # fine forward on the n_f importance samples only (src/NeRF.py:155-156)
            call("nerf_sample_pdf_fwd", ptr(w.w_c), ptr(w.z_c), n, sc, sf, ptr(u), seed, step, ray_offset, ptr(w.z_f),
                 None, ptr(w.perm), ptr(w.u))
            self._mlp_fwd_train(mf, o, d, w.z_f, n, sf, w.xyz_f, w.view_f, w.raw_f, w.saved_f, w.ws_fwd)
            # fine compositing, loss and their backward
            through_z = not self.stop_grad_z
            call("nerf_composite_mse_fwd_bwd", ptr(w.raw_f), ptr(w.z_f), ptr(y), n, sf, n_total, 1.0, ptr(w.rgb_f),
                 ptr(sums[1:2]), ptr(w.d_raw_f), ptr(w.d_z_f) if through_z else None)
            dz_in_chain = through_z and self._dz_in_chain(mf)
            if dz_in_chain:
                side = self._mlp_bwd_rays(mf, w.saved_f, w.d_raw_f, o, d, w.z_f, n, sf, g_f, w.d_z_f, w.ws_bwd,
                                          side_stream=self._side_stream())
            else:
                side = self._mlp_bwd(mf, w.xyz_f, w.view_f, w.saved_f, w.d_raw_f, n * sf, g_f,
                                     w.d_xyz_f if through_z else None, w.ws_bwd, side_stream=self._side_stream())
            if self.world_size > 1 and self._overlap_allreduce and self._peer_step is None:
                # the fine network's gradients are final: their all-reduce runs under the coarse backward
                with torch.cuda.stream(side) if side is not None else contextlib.nullcontext():
                    self._fine_allreduce = allreduce_sum_(g_f, self._process_group, async_op=True)
            if self._fine_update is not None:
                # ... and so do its Adam update and the refresh of its bf16 weight pack
                with torch.cuda.stream(side) if side is not None else contextlib.nullcontext():
                    if self._fine_allreduce is not None:
                        self._fine_allreduce.wait()
                        self._fine_allreduce = None
                    self._fine_update(g_f)
                    self._fine_updated = True
            if through_z:
                # z_f -> xyz -> PE -> fine net, and z_f -> delta in the fine compositing, reach the coarse weights
                if not dz_in_chain:
                    call("nerf_encode_samples_bwd_z", mf.cfg_ref, ptr(o), ptr(d), ptr(w.z_f), ptr(w.d_xyz_f), n, sf,
                         ptr(w.d_z_f), 1)
                call("nerf_sample_pdf_bwd", ptr(w.w_c), ptr(w.z_c), ptr(w.u), ptr(w.perm), ptr(w.d_z_f), n, sc, sf,
                     ptr(w.d_w_c))
                d_w_c = w.d_w_c
        # coarse backward
        call("nerf_composite_bwd", ptr(w.raw_c), ptr(w.z_c), ptr(w.d_rgb_c), ptr(d_w_c), n, sc, ptr(w.d_raw_c), None)
        self._mlp_bwd(mc, w.xyz_c, w.view_c, w.saved_c, w.d_raw_c, n * sc, g_c, None, w.ws_bwd_c,
                      side_stream=self._side_stream())
        if self._side is not None:
            torch.cuda.current_stream().wait_stream(self._side)      # the weight gradients of both networks join here
        return g_c, g_f, sums

    def render_backward(self, rays_orig, rays_dirs, d_rgb, n_render_samples_c=None, n_render_samples_f=None, *,
                        seed, step=0, ray_offset=0, jitter=None, u=None):
        """Back-propagate ``d_rgb`` (N,3) = dL/d(rendered rgb) through ``render`` (src/NeRF.py:109-134) for these rays:
        what TF's tape does when render_image sits inside it (DietNeRF.calc_consistency_loss, src/DietNeRF.py:215-218).

        The render is recomputed in training mode from the same Philox stream (seed, step, ray_offset), so it is the
        render a previous ``render(..., seed=, step=, ray_offset=)`` call returned; the fine network sees
        sort(concat(z_from_dist, z_coarse)) (:132) and its loss reaches the coarse network through the importance
        sampler.  Parameter gradients are ADDED to the flat gradient buffer; returns the recomputed rgb (N,3).
        """
        o, d = f32c(rays_orig, self.device), f32c(rays_dirs, self.device)
        d_rgb = f32c(d_rgb, self.device)
        n = o.shape[0]
        mc, mf = self.model_coarse, self.model_fine
        sc = n_render_samples_c if n_render_samples_c else self.n_render_samples_coarse
        nf = (n_render_samples_f if n_render_samples_f else self.n_render_samples_fine) if mf is not None else 0
        sf = nf + sc
        w = self._workspace(n, sc, sf if mf is not None else 0, nf if mf is not None else None)
        _, g_c, g_f = self._grad_views()
        self._coarse_pass(mc, o, d, n, sc, seed, step, ray_offset, jitter, w.z_c, w.raw_c, saved=w.saved_c, xyz=w.xyz_c,
                          view=w.view_c, ws=w.ws_fwd)
        call("nerf_composite_fwd", ptr(w.raw_c), ptr(w.z_c), n, sc, ptr(w.rgb_c), ptr(w.w_c), None, None, None, None,
             None)
        if mf is None:
            call("nerf_composite_bwd", ptr(w.raw_c), ptr(w.z_c), ptr(d_rgb), None, n, sc, ptr(w.d_raw_c), None)
            self._mlp_bwd(mc, w.xyz_c, w.view_c, w.saved_c, w.d_raw_c, n * sc, g_c, None, w.ws_bwd_c,
                          side_stream=self._side_stream())
            if self._side is not None:
                torch.cuda.current_stream().wait_stream(self._side)
            return w.rgb_c
        call("nerf_sample_pdf_fwd", ptr(w.w_c), ptr(w.z_c), n, sc, nf, ptr(u), seed, step, ray_offset, ptr(w.z_new),
             None, ptr(w.perm), ptr(w.u))
        call("nerf_merge_sorted_rank", ptr(w.z_new), nf, ptr(w.z_c), sc, n, ptr(w.z_f), ptr(w.rank))
        self._mlp_fwd_train(mf, o, d, w.z_f, n, sf, w.xyz_f, w.view_f, w.raw_f, w.saved_f, w.ws_fwd)
        call("nerf_composite_fwd", ptr(w.raw_f), ptr(w.z_f), n, sf, ptr(w.rgb_f), None, None, None, None, None, None)
        # fine backward
        through_z = not self.stop_grad_z
        call("nerf_composite_bwd", ptr(w.raw_f), ptr(w.z_f), ptr(d_rgb), None, n, sf, ptr(w.d_raw_f),
             ptr(w.d_z_f) if through_z else None)
        dz_in_chain = through_z and self._dz_in_chain(mf)
        if dz_in_chain:
            self._mlp_bwd_rays(mf, w.saved_f, w.d_raw_f, o, d, w.z_f, n, sf, g_f, w.d_z_f, w.ws_bwd,
                               side_stream=self._side_stream())
        else:
            self._mlp_bwd(mf, w.xyz_f, w.view_f, w.saved_f, w.d_raw_f, n * sf, g_f, w.d_xyz_f if through_z else None,
                          w.ws_bwd, side_stream=self._side_stream())
        if through_z:
            if not dz_in_chain:
                call("nerf_encode_samples_bwd_z", mf.cfg_ref, ptr(o), ptr(d), ptr(w.z_f), ptr(w.d_xyz_f), n, sf,
                     ptr(w.d_z_f), 1)
            call("nerf_merge_sorted_bwd", ptr(w.d_z_f), ptr(w.rank), nf, sc, n, ptr(w.d_z_new))
            call("nerf_sample_pdf_bwd", ptr(w.w_c), ptr(w.z_c), ptr(w.u), ptr(w.perm), ptr(w.d_z_new), n, sc, nf,
                 ptr(w.d_w_c))
            # the rendered image is the fine network's: the coarse rgb carries no loss, only its weights do
            w.d_rgb_c.zero_()
            call("nerf_composite_bwd", ptr(w.raw_c), ptr(w.z_c), ptr(w.d_rgb_c), ptr(w.d_w_c), n, sc, ptr(w.d_raw_c),
                 None)
            self._mlp_bwd(mc, w.xyz_c, w.view_c, w.saved_c, w.d_raw_c, n * sc, g_c, None, w.ws_bwd_c,
                          side_stream=self._side_stream())
        if self._side is not None:
            torch.cuda.current_stream().wait_stream(self._side)
        return w.rgb_f

    def render_image_backward(self, c2w, fov, h, w, d_image, batch_size_input=None, n_render_samples_c=None,
                              n_render_samples_f=None, *, seed, step=0, ray_begin=0, n_rays=None):
        """``render_backward`` over the rays [ray_begin, ray_begin+n_rays) of an h x w image in batches, the way
        ``render_image`` (src/NeRF.py:206-228) walks them.  ``d_image``: (h,w,3) or (h*w,3) gradient of the loss w.r.t.
        the image ``render_image(..., seed=seed, step=step)`` returned.  Gradients are added to the flat buffer."""
        n_total = h * w - ray_begin if n_rays is None else int(n_rays)
        dirs, orig = get_rays_directions(h, w, fov, c2w, ray_begin=ray_begin, n_rays=n_total, return_origins=True)
        dirs = dirs.reshape(-1, 4)
        d_flat = f32c(d_image, self.device).reshape(h * w, 3)
        batch_size = batch_size_input if batch_size_input else self.batch_size_render
        for s0 in range(0, n_total, batch_size):
            e0 = min(n_total, s0 + batch_size)
            self.render_backward(orig[s0:e0], dirs[s0:e0], d_flat[ray_begin + s0:ray_begin + e0], n_render_samples_c,
                                 n_render_samples_f, seed=seed, step=step, ray_offset=ray_begin + s0)

    def _side_stream(self):
        """Stream of the fine network's weight-gradient kernel (HBM-bound): it runs under the small sampler / compositing
        backward kernels and the head of the coarse backward instead of in front of them."""
        if not self.overlap_dw:
            return None
        if self._side is None and self.model_coarse.tensor_core:
            self._side = torch.cuda.Stream(device=self.device)
        return self._side

    @staticmethod
    def _dz_in_chain(net):
        """Tensor-core view network with the reference's xyz encoding width: the chain kernel forms d z itself
        (``nerf_mlp_bwd_rays``), no ``nerf_encode_samples_bwd_z`` afterwards."""
        return net.tensor_core and net.dx == 33 and net.dv > 0

    @staticmethod
    def _mlp_bwd_rays(net, saved, d_raw, o, d, z, n, s, grads, d_z, ws, side_stream=None):
        """``_mlp_bwd`` for the rows of n x s ray samples with the gradient w.r.t. the depths ADDED to ``d_z`` by the chain
        kernel itself (see ``_dz_in_chain``)."""
        args = (net.cfg_ref, ptr(net.packed_for(net.params)), ptr(saved), ptr(d_raw), ptr(o), ptr(d), ptr(z), n, s, ptr(grads),
                ptr(d_z), 1, ptr(ws), net.mode_id)
        if side_stream is not None:
            call("nerf_mlp_bwd_rays", *args, 3, side_stream.cuda_stream)
            return side_stream
        if NeRF.split_bwd_calls:
            call("nerf_mlp_bwd_rays", *args, 1, None)
            call("nerf_mlp_bwd_rays", *args, 2, None)
            return None
        call("nerf_mlp_bwd_rays", *args, 3, None)
        return None

    @staticmethod
    def _mlp_bwd(net, xyz, view, saved, d_raw, m, grads, d_xyz, ws, side_stream=None):
        """TF autodiff of one Keras model (src/NeRF.py:149-167).  With ``side_stream`` the tensor-core path runs its two
        halves AT THE SAME TIME (``nerf_mlp_bwd_overlapped``): the input-gradient chain on the current stream, the
        weight-gradient kernel on ``side_stream`` on the SMs the chain leaves free, dZ handed over through L2.  The stream
        is returned for the caller to join before it reads ``grads``."""
        args = (net.cfg_ref, ptr(net.params), ptr(net.packed_for(net.params)), ptr(xyz), ptr(view), ptr(saved), ptr(d_raw), m,
                ptr(grads), ptr(d_xyz), ptr(ws), net.mode_id)
        if net.tensor_core and side_stream is not None:
            call("nerf_mlp_bwd_overlapped", *args, side_stream.cuda_stream)
            return side_stream
        if net.tensor_core and NeRF.split_bwd_calls:
            call("nerf_mlp_bwd_dx", *args)
            call("nerf_mlp_bwd_dw", *args)
            return None
        call("nerf_mlp_bwd", *args)
        return None

    def _mlp_fwd_train(self, net, o, d, z, n, s, xyz, view, raw, saved, ws):
        """Training-mode MLP forward (activations saved): fused encode+MLP kernel in bf16 mode, two kernels in fp32."""
        if net.tensor_core:
            call("nerf_mlp_fwd_rays", net.cfg_ref, ptr(net.packed_for(net.params)), ptr(o), ptr(d), ptr(z), n, s,
                 ptr(raw), ptr(saved), net.mode_id)
        else:
            call("nerf_encode_samples", net.cfg_ref, ptr(o), ptr(d), ptr(z), n, s, ptr(xyz), ptr(view))
            call("nerf_mlp_fwd", net.cfg_ref, ptr(net.params), ptr(net.packed_for(net.params)), ptr(xyz), ptr(view),
                 n * s, ptr(raw), ptr(saved), ptr(ws), net.mode_id)

    def _metrics(self, sums, n_total):
        """{"loss", "psnr_coarse", "psnr_fine"} as device scalars (src/NeRF.py:170-178), one tiny kernel."""
        out = torch.empty(4, dtype=torch.float32, device=self.device)
        has_fine = self.model_fine is not None
        call("nerf_train_metrics", ptr(sums), int(n_total), float(self.COARSE_LOSS_WEIGHT), 1 if has_fine else 0, ptr(out))
        return self._metrics_dict(out)

    def _metrics_dict(self, out):
        """The metrics dict from nerf_train_metrics' out4 = [loss, psnr_coarse, psnr_fine, MSE_c + MSE_f]."""
        metrics = {"loss": out[0], "psnr_coarse": out[1]}
        if self.model_fine is not None:
            metrics["psnr_fine"] = out[2]
        self._metrics_raw = out
        return metrics

    def train_step(self, data) -> Dict:
        """One optimisation step on a batch (rays_orig (B,4), rays_dirs (B,4), real_rgb (B,3)).

        Returns the metrics dict {"loss", "psnr_coarse", "psnr_fine"} as device scalars (src/NeRF.py:170-178).
        Host tensors are copied to the device here (pinned memory makes the copy asynchronous).  With
        ``distribute()`` every rank passes the SAME global batch and works on its contiguous shard.
        """
        rays_orig, rays_dirs, real_rgb = data
        n_total = rays_orig.shape[0]
        lo, hi = shard_bounds(n_total, self.world_size, self.rank)
        to_dev = lambda t: t[lo:hi].to(device=self.device, dtype=torch.float32, non_blocking=True).contiguous()
        return self.train_step_local(to_dev(rays_orig), to_dev(rays_dirs), to_dev(real_rgb), n_total, lo)

    def train_step_local(self, rays_orig, rays_dirs, real_rgb, n_total_rays, ray_offset=0) -> Dict:
        """Train step on THIS rank's shard of a global batch of ``n_total_rays`` rays (device tensors).

        Gradients are normalised by the global ray count, summed over ranks with ONE all-reduce (the two squared-error
        sums ride in the same buffer), and the identical Adam update is applied on every rank.
        """
        if self.optimizer is None:
            raise RuntimeError("call compile(optimizer=Adam(lr)) before train_step")
        mc, mf = self.model_coarse, self.model_fine
        if (self.use_fused_step and self.world_size == 1 and getattr(self, "_extra_grads", None) is None
                and hasattr(self.optimizer, "apply_one") and int(n_total_rays) == rays_orig.shape[0]):
            return self.train_step_fused(rays_orig, rays_dirs, real_rgb, n_total_rays=n_total_rays, ray_offset=ray_offset)
        if (self.use_fused_step and self.world_size > 1 and self._peer is not None and mc.tensor_core
                and getattr(self, "_extra_grads", None) is None and not getattr(self, "_keep_grads", False)
                and hasattr(self.optimizer, "apply_one")):
            # the sharded step as ONE C call: same kernels as the sequence below, half the host issue time
            return self.train_step_fused(rays_orig, rays_dirs, real_rgb, n_total_rays=n_total_rays, ray_offset=ray_offset,
                                         peer=self._peer)
        n_all = mc.n_params + (mf.n_params if mf is not None else 0)
        t_next = self.optimizer.iterations + 1
        early = mf is not None and hasattr(self.optimizer, "apply_one") and getattr(self, "_extra_grads", None) is None

        peer = self._peer if (self.world_size > 1 and hasattr(self.optimizer, "apply_one")
                              and getattr(self, "_extra_grads", None) is None) else None
        if peer is not None:
            self.optimizer._state(n_all, self.device)

        def fine_update(g_f):
            if peer is not None:
                # one-shot exchange over NVLink fused with Adam, on the side stream under the coarse backward
                peer.barrier(t_next, 0)
                peer.reduce_adam(mf.params, 4 + mc.n_params, mf.n_params, self.optimizer, mc.n_params, t_next)
            else:
                self.optimizer.apply_one(mf.params, g_f, mc.n_params, n_all, t_next)
            mf.mark_updated()
            mf.packed_for(mf.params)

        self._overlap_allreduce = True
        self._fine_update, self._fine_updated = (fine_update if early else None), False
        self._peer_step = peer
        try:
            self.forward_backward(rays_orig, rays_dirs, real_rgb, n_total_rays=n_total_rays, ray_offset=ray_offset,
                                  keep_grads=getattr(self, "_keep_grads", False))
        finally:
            self._overlap_allreduce = False
            self._fine_update = None
            self._peer_step = None
        g = self._grad_buffer()
        if peer is not None:
            # the tail of the step: barrier, then [loss sums] and [coarse gradients -> Adam] straight from the peers' buffers
            peer.barrier(t_next, 1)
            peer.reduce_adam(None, 0, 4, self.optimizer, 0, t_next, reduced=peer.sums)
            self.optimizer.iterations += 1
            peer.reduce_adam(mc.params, 4, mc.n_params, self.optimizer, 0, self.optimizer.iterations)
            if mf is not None and not self._fine_updated:
                peer.reduce_adam(mf.params, 4 + mc.n_params, mf.n_params, self.optimizer, mc.n_params, self.optimizer.iterations)
                mf.mark_updated()
            self._fine_updated = False
            mc.mark_updated()
            peer.flip()
            self.step_counter += 1
            return self._metrics(peer.sums[0:2], n_total_rays)
        if self.world_size > 1:
            if self._fine_allreduce is not None or self._fine_updated:
                allreduce_sum_(g[:4 + mc.n_params], self._process_group)   # [sums | coarse gradients]
                if self._fine_allreduce is not None:
                    self._fine_allreduce.wait()
                    self._fine_allreduce = None
            else:
                allreduce_sum_(g, self._process_group)
        self.apply_gradients(g)
        self.step_counter += 1
        return self._metrics(g[0:2], n_total_rays)

    def apply_gradients(self, g):
        mc, mf = self.model_coarse, self.model_fine
        n = mc.n_params + (mf.n_params if mf is not None else 0)
        if self._fine_updated:
            # the fine network was stepped on the side stream as soon as its gradients were final
            self.optimizer.iterations += 1
            self.optimizer.apply_one(mc.params, g[4:4 + mc.n_params], 0, n, self.optimizer.iterations)
            self._fine_updated = False
        else:
            self.optimizer.apply_flat([mc.params] + ([mf.params] if mf is not None else []), g[4:4 + n])
            if mf is not None:
                mf.mark_updated()
        mc.mark_updated()
