"""ctypes binding of libnerf_b200.so (the C ABI declared in include/nerf_b200.h).

There is no CPU fallback: if the shared library is missing this module raises at import of the first op,
and every compute entry point returns an error without a CUDA device.  PyTorch is used only for device
memory and streams; tensors are handed over as raw device pointers.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int32, c_int64, c_uint32, c_uint64, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnerf_b200.so")

MODE_FP32 = 0
MODE_BF16 = 1
MODE_FP16 = 2


class NetCfg(Structure):
    """Mirror of struct nerf_net_cfg."""
    _fields_ = [("n_pos_enc_xyz", c_int32), ("n_pos_enc_view", c_int32), ("n_angles", c_int32),
                ("hidden", c_int32), ("last_hidden", c_int32), ("leaky_alpha", c_float)]


class RenderCfg(Structure):
    """Mirror of struct nerf_render_cfg."""
    _fields_ = [("near_boundary", c_float), ("far_boundary", c_float), ("n_samples_coarse", c_int32),
                ("n_samples_fine", c_int32), ("mode", c_int32)]


class RngState(Structure):
    """Mirror of struct nerf_rng_state."""
    _fields_ = [("seed", c_uint64), ("ray_offset", c_uint64), ("step", c_uint32), ("reserved", c_uint32)]


class RenderOuts(Structure):
    """Mirror of struct nerf_render_outs (device pointers, any may be null)."""
    _fields_ = [(n, c_void_p) for n in ("rgb", "weights", "cumprod", "alpha", "rgb_s", "z", "depth", "acc")]


class TrainCfg(Structure):
    """Mirror of struct nerf_train_cfg."""
    _fields_ = [("coarse_loss_weight", c_float), ("stop_grad_z", c_int32), ("accumulate_grads", c_int32),
                ("learning_rate", c_float), ("beta_1", c_float), ("beta_2", c_float), ("epsilon", c_float)]


class PeerExchangeCfg(Structure):
    """Mirror of struct nerf_peer_exchange."""
    _fields_ = [("pads_dev", c_void_p), ("grads_dev", c_void_p), ("reduced_sums", c_void_p), ("rank", c_int32),
                ("world", c_int32), ("epoch", c_uint32), ("reserved", c_uint32)]


class NerfLibraryError(RuntimeError):
    pass


_P = c_void_p
_CFG = POINTER(NetCfg)

# name -> (restype, argtypes); kept in the same order as include/nerf_b200.h
SIGNATURES = {
    "nerf_version": (c_char_p, []),
    "nerf_last_error": (c_char_p, []),
    "nerf_param_count": (c_int64, [_CFG]),
    "nerf_xyz_enc_dim": (c_int32, [_CFG]),
    "nerf_view_enc_dim": (c_int32, [_CFG]),
    "nerf_ray_directions": (c_int32, [POINTER(c_float), c_float, c_int32, c_int32, c_int64, c_int64, _P, _P, _P]),
    "nerf_stratified_z": (c_int32, [c_float, c_float, c_int64, c_int32, _P, c_uint64, c_uint32, c_uint64, _P, _P]),
    "nerf_sample_along_rays": (c_int32, [_P, _P, _P, c_int64, c_int32, _P, _P]),
    "nerf_view_directions": (c_int32, [_P, c_int64, c_int32, c_int32, _P, _P]),
    "nerf_posenc_xyz": (c_int32, [_P, c_int64, c_int32, _P, _P]),
    "nerf_posenc_views": (c_int32, [_P, c_int64, c_int32, c_int32, _P, _P]),
    "nerf_posenc_xyz_bwd": (c_int32, [_P, _P, c_int64, c_int32, _P, _P]),
    "nerf_encode_samples": (c_int32, [_CFG, _P, _P, _P, c_int64, c_int32, _P, _P, _P]),
    "nerf_encode_samples_bwd_z": (c_int32, [_CFG, _P, _P, _P, _P, c_int64, c_int32, _P, c_int32, _P]),
    "nerf_mlp_saved_bytes": (c_int64, [_CFG, c_int64, c_int32]),
    "nerf_mlp_workspace_bytes": (c_int64, [_CFG, c_int64, c_int32, c_int32]),
    "nerf_mlp_fwd": (c_int32, [_CFG, _P, _P, _P, _P, c_int64, _P, _P, _P, c_int32, _P]),
    "nerf_mlp_fwd_rays": (c_int32, [_CFG, _P, _P, _P, _P, c_int64, c_int32, _P, _P, c_int32, _P]),
    "nerf_mlp_fwd_rays_stratified": (c_int32, [_CFG, _P, _P, _P, c_float, c_float, c_uint64, c_uint32, c_uint64, c_int64, c_int32,
                                               _P, _P, _P, c_int32, _P]),
    "nerf_mlp_bwd": (c_int32, [_CFG, _P, _P, _P, _P, _P, _P, c_int64, _P, _P, _P, c_int32, _P]),
    "nerf_mlp_bwd_overlapped": (c_int32, [_CFG, _P, _P, _P, _P, _P, _P, c_int64, _P, _P, _P, c_int32, _P, _P]),
    "nerf_mlp_bwd_dx": (c_int32, [_CFG, _P, _P, _P, _P, _P, _P, c_int64, _P, _P, _P, c_int32, _P]),
    "nerf_mlp_bwd_dw": (c_int32, [_CFG, _P, _P, _P, _P, _P, _P, c_int64, _P, _P, _P, c_int32, _P]),
    "nerf_mlp_fwd_camera": (c_int32, [_CFG, _P, POINTER(c_float), c_float, c_int32, c_int32, c_int64, c_int64, c_int32, _P, c_float, c_float,
                                      c_uint64, c_uint32, _P, _P, c_int32, _P]),
    "nerf_mlp_bwd_rays": (c_int32, [_CFG, _P, _P, _P, _P, _P, _P, c_int64, c_int32, _P, _P, c_int32, _P, c_int32, c_int32, _P, _P]),
    "nerf_debug_bwd_pipe_layer": (c_int32, [_CFG, _P, _P, c_int64, _P, c_int32, _P, _P]),
    "nerf_packed_bytes": (c_int64, [_CFG]),
    "nerf_pack_weights": (c_int32, [_CFG, _P, _P, _P]),
    "nerf_pack_weights_fp16": (c_int32, [_CFG, _P, _P, _P]),
    "nerf_composite_fwd": (c_int32, [_P, _P, c_int64, c_int32, _P, _P, _P, _P, _P, _P, _P, _P]),
    "nerf_composite_bwd": (c_int32, [_P, _P, _P, _P, c_int64, c_int32, _P, _P, _P]),
    "nerf_composite_mse_fwd": (c_int32, [_P, _P, _P, c_int64, c_int32, c_int64, c_float, _P, _P, _P, _P, _P]),
    "nerf_composite_mse_fwd_bwd": (c_int32, [_P, _P, _P, c_int64, c_int32, c_int64, c_float, _P, _P, _P, _P, _P]),
    "nerf_sample_pdf_fwd": (c_int32, [_P, _P, c_int64, c_int32, c_int32, _P, c_uint64, c_uint32, c_uint64, _P, _P, _P,
                                      _P, _P]),
    "nerf_hierarchical_sample": (c_int32, [_P, _P, c_int64, c_int32, c_int32, c_uint64, c_uint32, c_uint64, _P, _P]),
    "nerf_sample_pdf_bwd": (c_int32, [_P, _P, _P, _P, _P, c_int64, c_int32, c_int32, _P, _P]),
    "nerf_merge_sorted": (c_int32, [_P, c_int32, _P, c_int32, c_int64, _P, _P]),
    "nerf_merge_sorted_rank": (c_int32, [_P, c_int32, _P, c_int32, c_int64, _P, _P, _P]),
    "nerf_merge_sorted_bwd": (c_int32, [_P, _P, c_int32, c_int32, c_int64, _P, _P]),
    "nerf_mse_fwd_bwd": (c_int32, [_P, _P, c_int64, c_int64, c_float, _P, _P, _P]),
    "nerf_train_metrics": (c_int32, [_P, c_int64, c_float, c_int32, _P, _P]),
    "nerf_adam_step": (c_int32, [_P, _P, _P, _P, c_int64, c_float, c_float, c_float, c_float, c_int64, _P]),
    "nerf_peer_barrier": (c_int32, [_P, c_int32, c_int32, c_uint32, c_int32, _P]),
    "nerf_peer_reduce_adam": (c_int32, [_P, _P, c_int32, c_int64, c_int64, _P, _P, c_float, c_float, c_float, c_float, c_int64,
                                        _P, _P]),
    "nerf_render_workspace_bytes": (c_int64, [_CFG, POINTER(RenderCfg), c_int64]),
    "nerf_render_fused_fwd": (c_int32, [_CFG, POINTER(RenderCfg), _P, _P, _P, _P, _P, _P, c_int64, POINTER(RngState),
                                        POINTER(RenderOuts), _P, _P]),
    "nerf_train_workspace_bytes": (c_int64, [_CFG, POINTER(RenderCfg), c_int64]),
    "nerf_train_step_fused": (c_int32, [_CFG, POINTER(RenderCfg), POINTER(TrainCfg), _P, _P, _P, _P, _P, _P, _P, c_int64,
                                        c_int64, POINTER(RngState), _P, _P, _P, c_int64, _P, _P, _P, _P]),
    "nerf_train_step_fused_sharded": (c_int32, [_CFG, POINTER(RenderCfg), POINTER(TrainCfg), _P, _P, _P, _P, _P, _P, _P, c_int64,
                                                c_int64, POINTER(RngState), _P, _P, _P, c_int64, _P, _P, _P,
                                                POINTER(PeerExchangeCfg), _P]),
}

_lib = None


def load():
    """Load libnerf_b200.so once; raise loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NerfLibraryError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


_empty_anchor = {}


def ptr(t):
    """Device pointer of a contiguous tensor (None -> NULL).  An EMPTY tensor has no storage (data_ptr() == 0), which the
    C side would take for a missing argument: it gets the address of a small per-device anchor buffer instead, and the
    entry points return early on a zero count."""
    if t is None:
        return None
    assert t.is_contiguous(), "tensor must be contiguous"
    if t.numel() == 0 and t.is_cuda:
        key = t.device.index
        if key not in _empty_anchor:
            _empty_anchor[key] = torch.zeros(64, dtype=torch.uint8, device=t.device)
        return _empty_anchor[key].data_ptr()
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def check(status, what):
    if status != 0:
        msg = load().nerf_last_error().decode()
        raise NerfLibraryError(f"{what} failed ({status}): {msg}")


# kernels launched per C-ABI call (bf16 mode; the fp32 MLP launches one GEMM per layer and is counted separately)
KERNELS_PER_CALL = {"nerf_render_fused_fwd": 5, "nerf_train_step_fused": 20, "nerf_train_step_fused_sharded": 23, "nerf_mlp_fwd": 1, "nerf_mlp_fwd_rays": 1, "nerf_mlp_fwd_rays_stratified": 1, "nerf_mlp_bwd": 3, "nerf_mlp_bwd_overlapped": 3, "nerf_mlp_bwd_dx": 1, "nerf_mlp_bwd_dw": 2, "nerf_mlp_bwd_rays": 3,
                    "nerf_pack_weights": 2}
launch_count = 0            # kernels launched through call() since import (bench.py reads the delta)
event_hook = None           # optional callable(name, args) -> context manager, used by bench.py to time single calls


def call(name, *args):
    """Call an int-returning entry point on the current torch CUDA stream and raise on error."""
    global launch_count
    lib = load()
    if event_hook is not None:
        with event_hook(name, args):
            status = getattr(lib, name)(*args, stream())
    else:
        status = getattr(lib, name)(*args, stream())
    launch_count += KERNELS_PER_CALL.get(name, 1)
    check(status, name)


def require_cuda(t, name="tensor"):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise NerfLibraryError(f"{name} must be a CUDA tensor: this framework has no CPU path")
    return t


def f32c(t, device=None):
    """float32 contiguous CUDA tensor from tensor / ndarray / list."""
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(t)
    if device is None:
        device = t.device if t.is_cuda else torch.device("cuda", torch.cuda.current_device())
    return t.to(device=device, dtype=torch.float32).contiguous()
