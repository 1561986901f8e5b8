"""The coarse/fine NeRF MLP as a device-resident flat parameter vector plus the C-ABI forward/backward.

Replaces the two Keras functional models of src/NeRF.py:248-288 (xyz only) and :290-340 (xyz + view direction).
Parameters live in ONE fp32 vector laid out [W0 (in,out) row-major, b0, W1, b1, ...] in Keras layer-creation
order (the order of ``dense … dense_10`` in the reference's saved .h5), initialised like Keras Dense
(glorot_uniform kernels, zero biases).  ``mode`` selects the arithmetic: "fp32" (SIMT fp32 GEMMs), "fp16" (default) or
"bf16" (tcgen05/TMEM tensor-core chain with 16-bit operands and fp32 accumulation).
"""
import math

import torch

from . import _lib
from ._lib import MODE_BF16, MODE_FP16, MODE_FP32, NetCfg, call, load, ptr

# mode -> (arithmetic of training / differentiable calls, arithmetic of inference-only calls).
#   "fp16" (default): forward MMAs with fp16 operands in training AND rendering (what the reference's mixed_float16 policy
#           computes in; 8x finer operand rounding than bf16: the 1e-3 render bound); the backward is the bf16 one (bf16
#           gradients keep fp32's range, so no loss scaling; tcgen05 wants one operand format per MMA, so the forward
#           saves its activations converted to bf16);
#   "bf16": bf16 operands everywhere;  "fp32": SIMT fp32 parity mode.
_MODES = {"fp32": (MODE_FP32, MODE_FP32), "bf16": (MODE_BF16, MODE_BF16), "fp16": (MODE_FP16, MODE_FP16)}
DEFAULT_MODE = "fp16"


def layer_shapes(cfg: NetCfg):
    dx = 3 + 6 * cfg.n_pos_enc_xyz
    h, hl = cfg.hidden, cfg.last_hidden
    trunk = [(dx, h)] + [(h, h)] * 3 + [(dx + h, h)] + [(h, h)] * 3
    if cfg.n_angles > 0:
        dv = 2 * cfg.n_pos_enc_view * (cfg.n_angles + 1)
        return trunk + [(h + dv, hl), (hl, 3), (h + dv, 1)]
    return trunk + [(h, h), (h, hl), (hl, 3), (h, 1)]


class _MlpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, params, xyz_enc, view_enc, model):
        m = xyz_enc.shape[0]
        out = torch.empty((m, 4), dtype=torch.float32, device=xyz_enc.device)
        need_grad = bool(ctx.needs_input_grad[0] or ctx.needs_input_grad[1])
        mode_id = model.mode_id if need_grad else model.infer_mode_id
        saved = (torch.empty(max(model.saved_bytes(m), 16), dtype=torch.uint8, device=xyz_enc.device)
                 if need_grad else None)
        ws = model._buffer("ws_fwd", model.workspace_bytes(m, False))
        call("nerf_mlp_fwd", model.cfg_ref, ptr(params), ptr(model.packed_for(params, half=mode_id == MODE_FP16)),
             ptr(xyz_enc), ptr(view_enc), m, ptr(out), ptr(saved), ptr(ws), mode_id)
        ctx.model = model
        ctx.save_for_backward(params, xyz_enc, view_enc if view_enc is not None else torch.empty(0), saved
                              if saved is not None else torch.empty(0))
        ctx.has_view = view_enc is not None
        return out

    @staticmethod
    def backward(ctx, d_out):
        params, xyz_enc, view_enc, saved = ctx.saved_tensors
        model = ctx.model
        m = xyz_enc.shape[0]
        grads = torch.zeros_like(params)
        d_xyz = torch.empty_like(xyz_enc) if ctx.needs_input_grad[1] else None
        ws = model._buffer("ws_bwd", model.workspace_bytes(m, True))
        call("nerf_mlp_bwd", model.cfg_ref, ptr(params), ptr(model.packed_for(params, half=model.mode_id == MODE_FP16)), ptr(xyz_enc),
             ptr(view_enc if ctx.has_view else None), ptr(saved), ptr(d_out.contiguous().float()), m, ptr(grads),
             ptr(d_xyz), ptr(ws), model.mode_id)
        return grads, d_xyz, None, None


class NerfMLP:
    """One NeRF network (coarse or fine)."""

    def __init__(self, cfg: NetCfg, mode: str = DEFAULT_MODE, device=None, seed=None):
        if mode not in _MODES:
            raise ValueError("mode must be 'fp32', 'bf16' or 'fp16'")
        load()
        self.cfg = cfg
        self.cfg_ref = _lib.ctypes.byref(cfg)
        self.mode = mode
        self.mode_id, self.infer_mode_id = _MODES[mode]
        self.tensor_core = self.mode_id != MODE_FP32
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.shapes = layer_shapes(cfg)
        self.n_params = int(load().nerf_param_count(self.cfg_ref))
        assert self.n_params == sum(i * o + o for i, o in self.shapes)
        self.dx = int(load().nerf_xyz_enc_dim(self.cfg_ref))
        self.dv = int(load().nerf_view_enc_dim(self.cfg_ref))
        if self.tensor_core and int(load().nerf_packed_bytes(self.cfg_ref)) < 0:
            raise _lib.NerfLibraryError("mode='bf16'/'fp16' does not support this network: "
                                        + load().nerf_last_error().decode() + " (use mode='fp32')")
        self.params = self._glorot_init(seed).to(self.device)
        self._buffers = {}
        self._packed = None
        self._packed_version = None
        self._packed_half_version = None

    # -- parameters -------------------------------------------------------------------------------------------
    def _glorot_init(self, seed):
        g = torch.Generator()
        if seed is not None:
            g.manual_seed(int(seed))
        parts = []
        for fan_in, fan_out in self.shapes:
            limit = math.sqrt(6.0 / (fan_in + fan_out))
            parts.append(((torch.rand(fan_in, fan_out, generator=g) * 2 - 1) * limit).reshape(-1))
            parts.append(torch.zeros(fan_out))
        return torch.cat(parts).float()

    @property
    def trainable_variables(self):
        """Views (kernel, bias, kernel, bias, ...) into the flat vector, Keras order."""
        out, off = [], 0
        for i, o in self.shapes:
            out.append(self.params[off:off + i * o].view(i, o))
            off += i * o
            out.append(self.params[off:off + o])
            off += o
        return out

    def set_params(self, flat):
        flat = torch.as_tensor(flat, dtype=torch.float32).reshape(-1)
        if flat.numel() != self.n_params:
            raise ValueError(f"expected {self.n_params} parameters, got {flat.numel()}")
        self.params = flat.to(self.device).contiguous().clone()
        self._packed_version = self._packed_half_version = None

    def mark_updated(self):
        """Call after an in-place parameter update so the bf16 weight pack is refreshed."""
        self._packed_version = self._packed_half_version = None

    def mark_packed(self):
        """After a call that updated the parameters in place AND refreshed the pack of this network's mode on the C side
        (nerf_train_step_fused): that region is current, the other precision's is stale."""
        key = (self.params.data_ptr(), self.params._version)
        if self.mode_id == MODE_FP16:
            self._packed_half_version, self._packed_version = key, None
        else:
            self._packed_version, self._packed_half_version = key, None

    # -- workspaces ------------------------------------------------------------------------------------------
    def saved_bytes(self, m):
        return int(load().nerf_mlp_saved_bytes(self.cfg_ref, int(m), self.mode_id))

    def workspace_bytes(self, m, backward):
        return int(load().nerf_mlp_workspace_bytes(self.cfg_ref, int(m), self.mode_id, 1 if backward else 0))

    def _buffer(self, name, nbytes):
        nbytes = max(int(nbytes), 16)
        buf = self._buffers.get(name)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self._buffers[name] = buf
        return buf

    def packed_for(self, params, half=None):
        """The packed 16-bit weight buffer, refreshed when the parameters changed (bf16 regions, or the fp16 regions with
        ``half``; default: what the network's mode reads)."""
        if not self.tensor_core:
            return None
        if half is None:
            half = self.mode_id == MODE_FP16
        key = (params.data_ptr(), params._version)
        if self._packed is None:
            self._packed = torch.empty(int(load().nerf_packed_bytes(self.cfg_ref)), dtype=torch.uint8,
                                       device=self.device)
        if half:
            if self._packed_half_version != key:
                call("nerf_pack_weights_fp16", self.cfg_ref, ptr(params), ptr(self._packed))
                self._packed_half_version = key
        elif self._packed_version != key:
            call("nerf_pack_weights", self.cfg_ref, ptr(params), ptr(self._packed))
            self._packed_version = key
        return self._packed

    # -- forward -----------------------------------------------------------------------------------------------
    def __call__(self, xyz_encoded, view_encoded=None, params=None):
        """(M, Dx)[, (M, Dv)] -> (M, 4) raw [r, g, b, sigma]; differentiable through torch.autograd."""
        if isinstance(xyz_encoded, dict):  # the reference calls model({XYZ_COORDS: ..., VIEW_DIRS: ...})
            view_encoded = xyz_encoded.get("view_dirs")
            xyz_encoded = xyz_encoded["xyz_coords"]
        if (view_encoded is None) != (self.cfg.n_angles == 0):
            raise ValueError("view directions must be given exactly when n_angles_for_model > 0")
        p = self.params if params is None else params
        x = xyz_encoded.contiguous().float()
        v = view_encoded.contiguous().float() if view_encoded is not None else None
        if x.shape[-1] != self.dx or (v is not None and v.shape[-1] != self.dv):
            raise ValueError("encoded input has the wrong width for this network")
        return _MlpFn.apply(p, x, v, self)
