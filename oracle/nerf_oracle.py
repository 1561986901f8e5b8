"""CPU restatement (PyTorch fp32) of the reference's NeRF ray-render hot path.  TEST INFRASTRUCTURE.

Every function cites the reference lines it follows (paths relative to /root/reference).  The
arithmetic of the reference lives in TensorFlow/Keras 2.7 (pin: base image
``us-docker.pkg.dev/vertex-ai/training/tf-gpu.2-7`` in contain_dockerfile_for_base_environment/
Dockerfile:1, ``keras_version 2.7.0`` inside the saved .h5), which is neither vendored under
/root/reference nor installable here, so the op semantics (linspace, searchsorted, cumprod gradient,
Dense, LeakyReLU, Adam) are restated from TensorFlow's published behaviour.

PARITY PIN: the reference's own tests (tests/test_UtilsCV.py) hold no vector for this path.  The
oracle is instead pinned against an OUTPUT OF THE REFERENCE ITSELF: the trained weights
``Results/50px_alexander_71pics_sphere_nerf_save_dir_4/saved_weights/NeRF_model_epoch_095.h5`` and the
test-image PSNR the reference recorded for them (``saved_test_train_psnrs/psnrs_train_test_095.npy``,
27.83 dB @ epoch 95); tests/golden/make_golden.py renders that image with this oracle and
tests/test_oracle_pin.py checks the PSNR (see DESIGN.md "Oracle pin").  A second pin holds it against
PIXELS the reference rendered: frames decoded from the videos of that run (``video_save/*.avi``;
tests/golden/make_alexander50_videos.py -> alexander50_videos.npz) are reproduced by this oracle at
36-43 dB rgb / 36-41 dB equalised depth (tests/test_video.py; the rest is MJPG loss and unseeded
jitter).  Everything finer than that (bit-level behaviour of TF kernels) is unpinned and stated as such.

Random inputs are explicit arguments (``jitter``, ``u``) or come from oracle.philox so the CUDA
kernels can be driven with the identical stream.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product never does.
"""
import math
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import philox

EPS = 1e-7  # src/UtilsCV.py:30
F32 = torch.float32


# --------------------------------------------------------------------------------------------------
# Rays (src/UtilsCV.py:467-499, src/NeRF.py:207-209, src/UtilsNeuralRadianceField.py:165-178)
# --------------------------------------------------------------------------------------------------
def get_rays_directions(height: int, width: int, field_of_view: float, c2w) -> torch.Tensor:
    """src/UtilsCV.py:467-499.  (h, w, 4) un-normalised ray directions, w-component 0."""
    c2w = torch.as_tensor(np.asarray(c2w), dtype=F32)
    xs = torch.arange(width, dtype=F32)
    ys = torch.arange(height, dtype=F32)
    x_raster, y_raster = torch.meshgrid(xs, ys, indexing="xy")          # (h, w) each, :477
    x_raster = x_raster + 0.5                                           # :479
    y_raster = y_raster + 0.5
    x_ndc = x_raster / float(width)                                     # :482
    y_ndc = y_raster / float(height)
    x_screen = 2 * x_ndc - 1                                            # :485
    y_screen = 1 - 2 * y_ndc
    tan_half_fov = torch.tan(torch.tensor(field_of_view, dtype=F32) / 2)  # :488
    x_cam = x_screen * tan_half_fov
    y_cam = y_screen * tan_half_fov
    z = -torch.ones_like(x_raster)
    d_cam = torch.stack([x_cam, y_cam, z, torch.zeros_like(x_raster)], dim=-1)  # :492-495
    # einsum 'ij,...j' (:498): out[..., i] = sum_j c2w[i, j] * d[..., j], accumulated j = 0..3 in order
    out = torch.zeros(height, width, 4, dtype=F32)
    for i in range(4):
        acc = c2w[i, 0] * d_cam[..., 0]
        for j in range(1, 4):
            acc = acc + c2w[i, j] * d_cam[..., j]
        out[..., i] = acc
    return out


def rays_for_image(c2w, fov: float, h: int, w: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """src/NeRF.py:206-209 / src/UtilsNeuralRadianceField.py:174-177: (h*w,4) origins and directions."""
    c2w = torch.as_tensor(np.asarray(c2w), dtype=F32)
    dirs = get_rays_directions(h, w, fov, c2w).reshape(h * w, 4)
    orig = c2w[:, 3].expand(h * w, 4).contiguous()
    return orig, dirs


# --------------------------------------------------------------------------------------------------
# Stratified depths (src/UtilsCV.py:565-581)
# --------------------------------------------------------------------------------------------------
def linspace_tf(start: float, stop: float, num: int) -> torch.Tensor:
    """tf.linspace semantics: first = start, last = stop exactly, interior = start + delta * i (fp32)."""
    start32 = np.float32(start)
    stop32 = np.float32(stop)
    if num == 1:
        return torch.tensor([start32], dtype=F32)
    delta = np.float32(np.float32(stop32 - start32) / np.float32(num - 1))
    idx = np.arange(num, dtype=np.float32)
    vals = (start32 + (delta * idx).astype(np.float32)).astype(np.float32)
    vals[0] = start32
    vals[-1] = stop32
    return torch.from_numpy(vals)


def get_z_values(z_start: float, z_end: float, n_rays: int, n_samples: int, jitter: torch.Tensor) -> torch.Tensor:
    """src/UtilsCV.py:578-580 as called by src/NeRF.py:127,146 (height=n_rays, width=1, [:,0,:]).

    ``jitter`` (n_rays, n_samples) in [0,1) replaces tf.random.uniform.  Jitter is always applied.
    z = linspace + (U * (z_end - z_start)) / n_samples     (left-to-right, as the Python expression)
    """
    lin = linspace_tf(z_start, z_end, n_samples)
    span = torch.tensor(np.float32(float(z_end) - float(z_start)))
    return lin[None, :] + (jitter.to(F32) * span) / float(n_samples)


def stratified_jitter(seed: int, step: int, n_rays: int, n_samples: int, ray_offset: int = 0) -> torch.Tensor:
    return torch.from_numpy(philox.uniform(seed, philox.STREAM_JITTER, step, n_rays, n_samples, ray_offset))


def importance_uniforms(seed: int, step: int, n_rays: int, n_new: int, ray_offset: int = 0) -> torch.Tensor:
    return torch.from_numpy(philox.uniform(seed, philox.STREAM_IMPORTANCE, step, n_rays, n_new, ray_offset))


# --------------------------------------------------------------------------------------------------
# Sample positions / view inputs (src/UtilsCV.py:584-599, :124-143)
# --------------------------------------------------------------------------------------------------
def sample_along_rays(origin: torch.Tensor, direction_vectors: torch.Tensor, z_values: torch.Tensor) -> torch.Tensor:
    """src/UtilsCV.py:595-598.  (N,4),(N,4),(N,S) -> (N,S,4)."""
    return origin[..., None, :] + direction_vectors[..., None, :] * z_values[..., None]


def get_view_directions(coords_3d: torch.Tensor, rays_dirs: torch.Tensor, n_angles_for_model: int) -> torch.Tensor:
    """src/UtilsCV.py:133-142.  Raw (un-normalised) ray direction broadcast to every sample."""
    if n_angles_for_model == 1:
        indices = [0, 2]
    elif n_angles_for_model == 2:
        indices = [0, 1, 2]
    else:
        raise Exception("n_angles_for_model should be 1 or 2.")
    n, s = coords_3d.shape[0], coords_3d.shape[1]
    d = rays_dirs[:, None, :].expand(n, s, rays_dirs.shape[-1])
    return d[..., indices].reshape(-1, n_angles_for_model + 1)


# --------------------------------------------------------------------------------------------------
# Positional encodings (src/UtilsNeuralRadianceField.py:52-85)
# --------------------------------------------------------------------------------------------------
def _theta(x: torch.Tensor, n: int) -> torch.Tensor:
    pow2 = torch.pow(torch.tensor(2.0, dtype=F32), torch.arange(n, dtype=F32))
    # (pow_of_2 * pi) * x : left-to-right as in :60 / :79
    return (pow2 * torch.tensor(math.pi, dtype=F32)) * x[..., None]


def positional_encoding_for_views(x: torch.Tensor, n_positional_encoding: int) -> torch.Tensor:
    """src/UtilsNeuralRadianceField.py:60-65.  (M,C) -> (M, C*2L): component-major, freq, (sin,cos)."""
    theta = _theta(x, n_positional_encoding)
    sc = torch.stack((torch.sin(theta), torch.cos(theta)), dim=-1)
    return sc.reshape(x.shape[0], -1)


def positional_encoding_for_xyz(xyz: torch.Tensor, n_positional_encoding: int) -> torch.Tensor:
    """src/UtilsNeuralRadianceField.py:76-85.  (M,3) -> (M, 3+6L): per coordinate [c, s0,c0, ...]."""
    if n_positional_encoding == 0:
        return xyz.reshape(xyz.shape[0], -1)
    theta = _theta(xyz, n_positional_encoding)
    sc = torch.stack((torch.sin(theta), torch.cos(theta)), dim=-1).reshape(-1, 3, 2 * n_positional_encoding)
    cat = torch.cat([xyz[..., None], sc], dim=-1)
    return cat.reshape(xyz.shape[0], -1)


# --------------------------------------------------------------------------------------------------
# The two MLPs (src/NeRF.py:248-340)
# --------------------------------------------------------------------------------------------------
def layer_shapes(n_pos_enc_xyz: int = 5, n_pos_enc_view: int = 4, n_angles: int = 2,
                 hidden: int = 256, last_hidden: int = 128) -> List[Tuple[int, int]]:
    """Dense kernel shapes (in, out) in Keras creation order (src/NeRF.py:263-286 / :312-337)."""
    dx = 3 + 6 * n_pos_enc_xyz
    if n_angles > 0:
        dv = n_pos_enc_view * 2 * (n_angles + 1)
        return ([(dx, hidden)] + [(hidden, hidden)] * 3 + [(dx + hidden, hidden)] + [(hidden, hidden)] * 3 +
                [(hidden + dv, last_hidden), (last_hidden, 3), (hidden + dv, 1)])
    return ([(dx, hidden)] + [(hidden, hidden)] * 3 + [(dx + hidden, hidden)] + [(hidden, hidden)] * 3 +
            [(hidden, hidden), (hidden, last_hidden), (last_hidden, 3), (hidden, 1)])


def n_params(shapes: Sequence[Tuple[int, int]]) -> int:
    return sum(i * o + o for i, o in shapes)


def glorot_params(shapes: Sequence[Tuple[int, int]], seed: int = 0, bias_scale: float = 0.0) -> torch.Tensor:
    """Keras Dense defaults: glorot_uniform kernel, zeros bias.  Flat fp32 vector [W0,b0,W1,b1,...]."""
    g = torch.Generator().manual_seed(seed)
    parts = []
    for fan_in, fan_out in shapes:
        limit = math.sqrt(6.0 / (fan_in + fan_out))
        w = (torch.rand(fan_in, fan_out, generator=g, dtype=F32) * 2 - 1) * limit
        b = (torch.rand(fan_out, generator=g, dtype=F32) * 2 - 1) * bias_scale
        parts += [w.reshape(-1), b]
    return torch.cat(parts)


def unflatten(params: torch.Tensor, shapes: Sequence[Tuple[int, int]]):
    out, off = [], 0
    for i, o in shapes:
        w = params[off:off + i * o].reshape(i, o)
        off += i * o
        b = params[off:off + o]
        off += o
        out.append((w, b))
    return out


def _bf16(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).to(F32)


class _RoundBF16(torch.autograd.Function):
    """Straight-through bf16 rounding (used only to emulate the BF16 kernel path tightly in tests)."""

    @staticmethod
    def forward(ctx, x):
        return _bf16(x)

    @staticmethod
    def backward(ctx, g):
        return g


class _RoundFP16(torch.autograd.Function):
    """Straight-through fp16 rounding: the operand rounding of the default tensor-core mode (and of the reference's own
    mixed_float16 policy, contain_dockerfile.../ExecutionRun.py:221)."""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.float16).to(F32)

    @staticmethod
    def backward(ctx, g):
        return g


def mlp_forward(params: torch.Tensor, shapes, xyz_enc: torch.Tensor, view_enc: Optional[torch.Tensor],
                alpha: float = 0.05, emulate_bf16=False) -> torch.Tensor:
    """src/NeRF.py:312-339 (view variant) / :263-287 (xyz-only).  Returns (M,4) = [r,g,b,sigma] raw.

    emulate_bf16 = True rounds weights and every MMA operand to bf16 (fp32 accumulate, fp32 bias add and
    activation) the way the tcgen05 path does in mode "bf16"; emulate_bf16 = "fp16" rounds them to fp16 (mode "fp16");
    the default is plain fp32.
    """
    rnd = _RoundFP16.apply if emulate_bf16 == "fp16" else (_RoundBF16.apply if emulate_bf16 else (lambda t: t))
    layers = unflatten(params, shapes)

    def dense(x, idx, act=True):
        w, b = layers[idx]
        y = rnd(x) @ rnd(w) + b
        return torch.nn.functional.leaky_relu(y, alpha) if act else y

    h = dense(xyz_enc, 0)
    h = dense(h, 1)
    h = dense(h, 2)
    h = dense(h, 3)
    h = dense(torch.cat([xyz_enc, h], dim=-1), 4)          # Concatenate()([inputs_xyz, hidden]) :322
    h = dense(h, 5)
    h = dense(h, 6)
    h8 = dense(h, 7)
    if view_enc is not None:
        g = torch.cat([h8, view_enc], dim=-1)               # :329
        hl = dense(g, 8)                                    # :330
        rgb = dense(hl, 9, act=False)                       # :333
        sigma = dense(g, 10, act=False)                     # :336 (sigma sees the view encoding)
    else:
        h9 = dense(h8, 8)                                   # :278
        hl = dense(h9, 9)                                   # :279
        rgb = dense(hl, 10, act=False)                      # :281
        sigma = dense(h8, 11, act=False)                    # :284
    return torch.cat([rgb, sigma], dim=-1)                  # :339 / :287


def model_predict(params, shapes, n_enc_phi_theta: int, n_pos_enc_for_xyz: int, xyz: torch.Tensor,
                  view_dirs: Optional[torch.Tensor] = None, alpha: float = 0.05, emulate_bf16: bool = False):
    """src/UtilsNeuralRadianceField.py:229-234."""
    xyz_encoded = positional_encoding_for_xyz(xyz, n_pos_enc_for_xyz)
    dir_encoded = positional_encoding_for_views(view_dirs, n_enc_phi_theta) if view_dirs is not None else None
    return mlp_forward(params, shapes, xyz_encoded, dir_encoded, alpha, emulate_bf16)


# --------------------------------------------------------------------------------------------------
# Alpha compositing (src/UtilsNeuralRadianceField.py:88-115)
# --------------------------------------------------------------------------------------------------
# Throughput switch for bench.py's CPU baseline ONLY (never set by a parity test): True replaces the sequential
# left-to-right Python loops over the samples of a ray (the canonical, bit-defining order of this oracle) by
# torch.cumprod / cumsum / sum, i.e. what an optimised CPU implementation such as TF-CPU would run.  Same mathematics,
# different fp32 summation order.
FAST_REDUCTIONS = False


class _ExclusiveCumprodTF(torch.autograd.Function):
    """tf.math.cumprod(x, -1, exclusive=True) with TensorFlow's gradient
    (math_grad._CumprodGrad: div_no_nan(cumsum(out*grad, exclusive, reverse), x) -> 0 where x == 0).
    Forward is a sequential left-to-right fp32 product (the canonical order of this oracle)."""

    @staticmethod
    def forward(ctx, x):
        if FAST_REDUCTIONS:
            out = torch.cat([torch.ones_like(x[..., :1]), torch.cumprod(x[..., :-1], dim=-1)], dim=-1)
            ctx.save_for_backward(x, out)
            return out
        out = torch.empty_like(x)
        run = torch.ones_like(x[..., 0])
        for i in range(x.shape[-1]):
            out[..., i] = run
            run = run * x[..., i]
        ctx.save_for_backward(x, out)
        return out

    @staticmethod
    def backward(ctx, g):
        x, out = ctx.saved_tensors
        prod = out * g
        if FAST_REDUCTIONS:
            rev = torch.flip(torch.cumsum(torch.flip(prod, [-1]), dim=-1), [-1]) - prod
        else:
            rev = torch.zeros_like(prod)
            run = torch.zeros_like(prod[..., 0])
            for i in range(x.shape[-1] - 1, -1, -1):
                rev[..., i] = run
                run = run + prod[..., i]
        return torch.where(x == 0, torch.zeros_like(rev), rev / torch.where(x == 0, torch.ones_like(x), x))


def ray_marching(model_output: torch.Tensor, z_values: torch.Tensor):
    """src/UtilsNeuralRadianceField.py:99-115.  Returns rgb (N,3), weights, cumprod, alpha (N,S), rgb_s (N,S,3)."""
    model_output, z_values = model_output.to(F32), z_values.to(F32)
    sigma_a = torch.relu(model_output[..., 3])                                   # :100
    net_rgb_output = torch.sigmoid(model_output[..., :3])                        # :101
    delta = z_values[..., 1:] - z_values[..., :-1]                               # :104
    inf = torch.full(z_values.shape[:-1] + (1,), 1e9, dtype=F32)                # :105
    delta = torch.cat([delta, inf], dim=-1)                                      # :106
    alpha = 1.0 - torch.exp(-sigma_a * delta)                                    # :111
    cumprod = _ExclusiveCumprodTF.apply(1.0 - alpha)                             # :112
    weights = alpha * cumprod                                                    # :113
    # reduce_sum over samples, accumulated in sample order
    if FAST_REDUCTIONS:
        return (weights[..., None] * net_rgb_output).sum(-2), weights, cumprod, alpha, net_rgb_output
    rgb_image = torch.zeros(weights.shape[:-1] + (3,), dtype=F32)
    for i in range(weights.shape[-1]):
        rgb_image = rgb_image + weights[..., i, None] * net_rgb_output[..., i, :]
    return rgb_image, weights, cumprod, alpha, net_rgb_output


def depth_and_acc(weights: torch.Tensor, z_values: torch.Tensor):
    """depth = sum w z (src/ExecutionRun.py:346); acc = sum w (new output named by north_star)."""
    depth = torch.zeros_like(weights[..., 0])
    acc = torch.zeros_like(weights[..., 0])
    for i in range(weights.shape[-1]):
        depth = depth + weights[..., i] * z_values[..., i]
        acc = acc + weights[..., i]
    return depth, acc


# --------------------------------------------------------------------------------------------------
# Inverse-CDF importance sampling (src/UtilsCV.py:502-539)
# --------------------------------------------------------------------------------------------------
def _seq_sum(x: torch.Tensor) -> torch.Tensor:
    if FAST_REDUCTIONS:
        return x.sum(-1)
    run = torch.zeros_like(x[..., 0])
    for i in range(x.shape[-1]):
        run = run + x[..., i]
    return run


def _seq_cumsum(x: torch.Tensor) -> torch.Tensor:
    if FAST_REDUCTIONS:
        return torch.cumsum(x, dim=-1)
    outs = []
    run = torch.zeros_like(x[..., 0])
    for i in range(x.shape[-1]):
        run = run + x[..., i]
        outs.append(run)
    return torch.stack(outs, dim=-1)


def get_z_vals_from_prob_dist_func(weights: torch.Tensor, z_values: torch.Tensor, num_new_z_values: int,
                                   u: torch.Tensor, return_aux: bool = False):
    """src/UtilsCV.py:512-539.  ``u`` (N, num_new) in [0,1) replaces tf.random.uniform (:514).

    Canonical summation order (not defined by the reference, whose TF kernels are SIMD-width
    dependent): reduce_sum and cumsum are sequential left-to-right in fp32.  Differentiable w.r.t.
    ``weights`` exactly as in the reference (no stop_gradient; searchsorted / clamps / the
    ``where(den<1e-5, 1e-5, den)`` constant branch carry no gradient; sort back-propagates through
    its permutation).
    """
    weights, z_values = weights.to(F32), z_values.to(F32)
    s = weights.shape[-1]
    pdf = weights / (_seq_sum(weights)[..., None] + EPS)                          # :512
    cdf = _seq_cumsum(pdf)                                                       # :513
    u = u.to(F32)
    idx = torch.searchsorted(cdf.detach().contiguous(), u.contiguous(), right=False)   # :515 (side='left')
    bottom = torch.clamp(idx - 1, min=0)                                         # :517
    top = torch.clamp(idx, max=s - 1)                                            # :520
    cdf_lo = torch.gather(cdf, -1, bottom)                                       # :523
    cdf_hi = torch.gather(cdf, -1, top)
    avg_z = 0.5 * (z_values[..., 1:] + z_values[..., :-1])                        # :525
    z_lo = torch.gather(avg_z, -1, torch.clamp(bottom, 0, s - 2))                # :526-527
    z_hi = torch.gather(avg_z, -1, torch.clamp(top, 0, s - 2))
    den = cdf_hi - cdf_lo                                                        # :530
    den = torch.where(den < 1e-5, torch.full_like(den, 1e-5), den)               # :531
    t = (u - cdf_lo) / den                                                       # :533
    z_samples = z_lo + t * (z_hi - z_lo)                                         # :534
    z_sorted, perm = torch.sort(z_samples, dim=-1, stable=True)                  # :535
    if return_aux:
        return z_sorted, idx.to(torch.int32), perm.to(torch.int32), z_samples
    return z_sorted


# --------------------------------------------------------------------------------------------------
# render_rays / render / losses (src/UtilsNeuralRadianceField.py:181-211, src/NeRF.py:109-178)
# --------------------------------------------------------------------------------------------------
class NetCfg:
    def __init__(self, n_pos_enc_xyz=5, n_pos_enc_view=4, n_angles=2, hidden=256, last_hidden=128, alpha=0.05):
        self.n_pos_enc_xyz, self.n_pos_enc_view, self.n_angles = n_pos_enc_xyz, n_pos_enc_view, n_angles
        self.hidden, self.last_hidden, self.alpha = hidden, last_hidden, alpha
        self.shapes = layer_shapes(n_pos_enc_xyz, n_pos_enc_view, n_angles, hidden, last_hidden)
        self.n_params = n_params(self.shapes)


def render_rays(params, cfg: NetCfg, rays_orig, rays_dirs, z_values, emulate_bf16: bool = False):
    """src/UtilsNeuralRadianceField.py:204-211."""
    coords_3d = sample_along_rays(rays_orig, rays_dirs, z_values)[..., :3]
    view_dirs = None if cfg.n_angles == 0 else get_view_directions(coords_3d, rays_dirs, cfg.n_angles)
    xyz = coords_3d.reshape(-1, 3)
    pred = model_predict(params, cfg.shapes, cfg.n_pos_enc_view, cfg.n_pos_enc_xyz, xyz, view_dirs, cfg.alpha,
                         emulate_bf16)
    pred = pred.reshape(coords_3d.shape[:-1] + (4,))
    return ray_marching(pred, z_values)


def render(params_c, params_f, cfg: NetCfg, near: float, far: float, rays_orig, rays_dirs, n_c: int, n_f: int,
           jitter: torch.Tensor, u: Optional[torch.Tensor], emulate_bf16: bool = False):
    """src/NeRF.py:124-134.  Returns the 6-tuple (rgb, weights, cumprod, alpha, rgb_s, z)."""
    z = get_z_values(near, far, rays_orig.shape[0], n_c, jitter)
    out = render_rays(params_c, cfg, rays_orig, rays_dirs, z, emulate_bf16)
    if params_f is not None:
        z_from_dist = get_z_vals_from_prob_dist_func(out[1], z, n_f, u)
        z = torch.sort(torch.cat([z_from_dist, z], dim=-1), dim=-1).values       # :132
        out = render_rays(params_f, cfg, rays_orig, rays_dirs, z, emulate_bf16)
    return out + (z,)


def mse(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """keras.losses.MeanSquaredError: mean over every element of the batch."""
    return torch.mean((a - b) ** 2)


def get_psnr(mse_val: torch.Tensor) -> torch.Tensor:
    """src/UtilsNeuralRadianceField.py:131."""
    return -10.0 * torch.log(mse_val) / math.log(10.0)


def train_losses(params_c, params_f, cfg: NetCfg, near, far, rays_orig, rays_dirs, real_rgb, n_c, n_f,
                 jitter, u, dietnerf: bool = False, emulate_bf16: bool = False, stop_grad_z: bool = False):
    """Forward half of src/NeRF.py:145-157 (or src/DietNeRF.py:159-172 when dietnerf=True).

    NeRF:     loss = MSE_c + MSE_f.
    DietNeRF: the aliasing at DietNeRF.py:164-171 makes loss = 2*MSE_c + MSE_f  (tensors are
              immutable in TF so ``loss`` keeps MSE_c, then ``loss += loss_for_rays`` adds MSE_c+MSE_f).
    The fine net sees ONLY the n_f new samples (NeRF.py:155-156); z_from_dist is not detached.
    """
    z = get_z_values(near, far, rays_orig.shape[0], n_c, jitter)
    rgb_c, w_c = render_rays(params_c, cfg, rays_orig, rays_dirs, z, emulate_bf16)[:2]
    mse_c = mse(real_rgb, rgb_c)
    out = {"rgb_c": rgb_c, "w_c": w_c, "z_c": z, "mse_c": mse_c}
    loss = mse_c
    loss_for_rays = mse_c
    if params_f is not None:
        z_f = get_z_vals_from_prob_dist_func(w_c, z, n_f, u)
        if stop_grad_z:        # NOT the reference (it keeps the path, NeRF.py:155): a test knob that isolates the
            z_f = z_f.detach()  # well-conditioned part of the coarse gradient from the sampler path
        rgb_f, w_f = render_rays(params_f, cfg, rays_orig, rays_dirs, z_f, emulate_bf16)[:2]
        mse_f = mse(real_rgb, rgb_f)
        loss_for_rays = mse_c + mse_f
        loss = (mse_c + loss_for_rays) if dietnerf else loss_for_rays
        out.update({"rgb_f": rgb_f, "w_f": w_f, "z_f": z_f, "mse_f": mse_f})
    out["loss"] = loss
    out["loss_for_rays"] = loss_for_rays
    return out


def adam_step(p: torch.Tensor, g: torch.Tensor, m: torch.Tensor, v: torch.Tensor, t: int, lr: float,
              b1: float = 0.9, b2: float = 0.999, eps: float = 1e-7):
    """Keras 2.7 Adam (optimizer_v2/adam.py, non-amsgrad): lr_t = lr*sqrt(1-b2^t)/(1-b1^t);
    m += (g-m)(1-b1); v += (g^2-v)(1-b2); p -= lr_t*m/(sqrt(v)+eps).  t is 1-based."""
    lr_t = lr * math.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t)
    m = m + (g - m) * (1.0 - b1)
    v = v + (g * g - v) * (1.0 - b2)
    p = p - lr_t * m / (torch.sqrt(v) + eps)
    return p, m, v


def train_step(params_c, params_f, cfg, near, far, rays_orig, rays_dirs, real_rgb, n_c, n_f, jitter, u,
               dietnerf=False, emulate_bf16=False, stop_grad_z=False):
    """Loss + gradients of src/NeRF.py:145-167 via autograd.  Returns (metrics dict, grad_c, grad_f)."""
    pc = params_c.detach().clone().requires_grad_(True)
    pf = params_f.detach().clone().requires_grad_(True) if params_f is not None else None
    out = train_losses(pc, pf, cfg, near, far, rays_orig, rays_dirs, real_rgb, n_c, n_f, jitter, u, dietnerf,
                       emulate_bf16, stop_grad_z)
    out["loss"].backward()
    metrics = {"loss": out["loss"].detach(), "psnr_coarse": get_psnr(out["mse_c"].detach())}
    if pf is not None:
        metrics["psnr_fine"] = get_psnr(out["mse_f"].detach())
    return metrics, pc.grad, (pf.grad if pf is not None else None), out


def render_image(params_c, params_f, cfg, near, far, c2w, fov, h, w, batch_size, n_c, n_f, seed=0, step=0,
                 emulate_bf16=False):
    """src/NeRF.py:206-246 with the shared Philox stream (global ray index = y*w+x)."""
    orig, dirs = rays_for_image(c2w, fov, h, w)
    parts = []
    with torch.no_grad():
        for s in range(0, h * w, batch_size):
            e = min(h * w, s + batch_size)
            jit = stratified_jitter(seed, step, e - s, n_c, ray_offset=s)
            u = importance_uniforms(seed, step, e - s, n_f, ray_offset=s) if params_f is not None else None
            parts.append(render(params_c, params_f, cfg, near, far, orig[s:e], dirs[s:e], n_c, n_f, jit, u,
                                emulate_bf16))
    cat = [torch.cat([p[i] for p in parts], dim=0) for i in range(6)]
    rgb = cat[0].reshape(h, w, 3)
    weights, cumprod, alpha = (cat[i].reshape(h, w, -1) for i in (1, 2, 3))
    rgb_s = cat[4].reshape(h, w, -1, 3)
    z = cat[5].reshape(h, w, -1)
    return rgb, weights, cumprod, alpha, rgb_s, z


# --------------------------------------------------------------------------------------------------
# DietNeRF semantic-consistency term (src/DietNeRF.py:204-222, :261-279)
# --------------------------------------------------------------------------------------------------
def embedder_preprocess(images: torch.Tensor) -> torch.Tensor:
    """src/DietNeRF.py:279: tf.image.resize(images, (224,224)) * 2 - 1 (bilinear, half-pixel centres, antialias off).
    (B,H,W,3) -> (B,3,224,224)."""
    x = torch.nn.functional.interpolate(images.permute(0, 3, 1, 2), size=(224, 224), mode="bilinear",
                                        align_corners=False, antialias=False)
    return x * 2.0 - 1.0


def consistency_loss(embedding_source: torch.Tensor, embedding_target: torch.Tensor) -> torch.Tensor:
    """src/DietNeRF.py:270: (1 + keras.losses.cosine_similarity(s, t)) / 2, where Keras' cosine_similarity is
    -sum(l2_normalize(s) * l2_normalize(t)) (a loss: -1 when aligned)."""
    s = embedding_source / torch.sqrt(torch.clamp((embedding_source ** 2).sum(), min=1e-12))
    t = embedding_target / torch.sqrt(torch.clamp((embedding_target ** 2).sum(), min=1e-12))
    return (1.0 + (-(s * t).sum())) / 2.0


def consistency_loss_and_grads(params_c, params_f, cfg, near, far, c2w, fov, size, batch_size, n_samples, seed, step,
                               embed_fn, target_embedding, weight: float = 0.1, emulate_bf16: bool = False):
    """calc_consistency_loss (src/DietNeRF.py:204-222) with the tape open: render_image(pose, fov, size, size,
    batch_size, n_samples, n_samples)[0] -> preprocess -> embedder -> weight * consistency_loss, and its gradient
    w.r.t. both parameter vectors.  Jitter / importance draws come from the shared Philox stream keyed by the global
    ray index, as in render_image above.  Returns (loss, grad_c, grad_f, image)."""
    pc = params_c.detach().clone().requires_grad_(True)
    pf = params_f.detach().clone().requires_grad_(True) if params_f is not None else None
    orig, dirs = rays_for_image(c2w, fov, size, size)
    parts = []
    for s in range(0, size * size, batch_size):
        e = min(size * size, s + batch_size)
        jit = stratified_jitter(seed, step, e - s, n_samples, ray_offset=s)
        u = importance_uniforms(seed, step, e - s, n_samples, ray_offset=s) if pf is not None else None
        parts.append(render(pc, pf, cfg, near, far, orig[s:e], dirs[s:e], n_samples, n_samples, jit, u,
                            emulate_bf16)[0])
    image = torch.cat(parts, dim=0).reshape(size, size, 3)
    emb = embed_fn(embedder_preprocess(image[None]))[0]
    loss = weight * consistency_loss(emb, target_embedding)
    loss.backward()
    return loss.detach(), pc.grad, (pf.grad if pf is not None else None), image.detach()
