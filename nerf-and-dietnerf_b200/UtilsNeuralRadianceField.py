"""Host mirror of the reference's src/UtilsNeuralRadianceField.py (render-op library of the hot path).

Same names and argument order as the reference: split_to_batches :17-29, get_size_of_splits :32-49,
positional_encoding_for_views :52-65, positional_encoding_for_xyz :68-85, ray_marching :88-115,
get_psnr_for_image :118-120, get_psnr :123-132, prepare_ds :135-162, c2w_to_rays_prepare_ds :165-178,
render_rays :181-211, model_predict :214-234, get_num_of_batches :237-247.  Tensors are contiguous fp32
``torch.cuda`` tensors; compute is done by the sm_100a kernels behind include/nerf_b200.h.
"""
import math

import numpy as np
import torch

from . import _lib
from ._lib import call, f32c, ptr
from .UtilsCV import get_rays_directions, get_view_directions, sample_along_rays

N_COORDINATES = 3
N_COLOR_CHANNELS = 3
XYZ_COORDS = "xyz_coords"
VIEW_DIRS = "view_dirs"


def get_size_of_splits(batch_size, total_size):
    """Sizes of the batches when splitting ``total_size`` rows into batches of ``batch_size`` (last one ragged)."""
    n_full = total_size // batch_size
    if n_full == 0:
        return [total_size]
    rest = total_size - n_full * batch_size
    return [batch_size] * n_full + ([rest] if rest else [])


def split_to_batches(to_split, batch_size):
    """Split along dim 0 into batches of ``batch_size``; the last batch keeps the remainder."""
    assert batch_size > 0
    return list(torch.split(to_split, get_size_of_splits(batch_size, to_split.shape[0])))


class _PosEncXyz(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz, n):
        m = xyz.shape[0]
        out = torch.empty((m, 3 * (1 + 2 * n)), dtype=torch.float32, device=xyz.device)
        call("nerf_posenc_xyz", ptr(xyz), m, n, ptr(out))
        ctx.save_for_backward(xyz)
        ctx.n = n
        return out

    @staticmethod
    def backward(ctx, g):
        (xyz,) = ctx.saved_tensors
        d = torch.empty_like(xyz)
        call("nerf_posenc_xyz_bwd", ptr(xyz), ptr(g.contiguous().float()), xyz.shape[0], ctx.n, ptr(d))
        return d, None


def positional_encoding_for_xyz(xyz, n_positional_encoding):
    """(M,3) -> (M, 3+6L): per coordinate [c, sin(2^0 pi c), cos(2^0 pi c), ...]; L = 0 is the identity."""
    x = f32c(xyz).reshape(-1, 3)
    return _PosEncXyz.apply(x, int(n_positional_encoding))


def positional_encoding_for_views(x, n_positional_encoding):
    """(M,C) -> (M, 2LC): per component [sin(2^0 pi v), cos(2^0 pi v), ...] (no identity term)."""
    x = f32c(x)
    x = x.reshape(x.shape[0], -1)
    m, c = x.shape
    out = torch.empty((m, 2 * int(n_positional_encoding) * c), dtype=torch.float32, device=x.device)
    call("nerf_posenc_views", ptr(x), m, c, int(n_positional_encoding), ptr(out))
    return out


class _RayMarching(torch.autograd.Function):
    @staticmethod
    def forward(ctx, raw4, z, lean):
        n, s = z.shape
        dev = z.device
        rgb = torch.empty((n, 3), dtype=torch.float32, device=dev)
        weights = torch.empty((n, s), dtype=torch.float32, device=dev)
        if lean:
            cumprod = alpha = rgb_s = None
        else:
            cumprod = torch.empty((n, s), dtype=torch.float32, device=dev)
            alpha = torch.empty((n, s), dtype=torch.float32, device=dev)
            rgb_s = torch.empty((n, s, 3), dtype=torch.float32, device=dev)
        depth = torch.empty((n,), dtype=torch.float32, device=dev)
        acc = torch.empty((n,), dtype=torch.float32, device=dev)
        call("nerf_composite_fwd", ptr(raw4), ptr(z), n, s, ptr(rgb), ptr(weights), ptr(cumprod), ptr(alpha),
             ptr(rgb_s), ptr(depth), ptr(acc))
        ctx.save_for_backward(raw4, z)
        ctx.lean = lean
        if lean:
            ctx.mark_non_differentiable(depth, acc)
            return rgb, weights, depth, acc
        ctx.mark_non_differentiable(cumprod, alpha, rgb_s, depth, acc)
        return rgb, weights, cumprod, alpha, rgb_s, depth, acc

    @staticmethod
    def backward(ctx, d_rgb, d_weights, *unused):
        raw4, z = ctx.saved_tensors
        n, s = z.shape
        d_rgb = torch.zeros((n, 3), dtype=torch.float32, device=z.device) if d_rgb is None else d_rgb.contiguous().float()
        d_w = d_weights.contiguous().float() if d_weights is not None else None
        d_raw = torch.empty_like(raw4)
        d_z = torch.empty_like(z) if ctx.needs_input_grad[1] else None
        call("nerf_composite_bwd", ptr(raw4), ptr(z), ptr(d_rgb), ptr(d_w), n, s, ptr(d_raw), ptr(d_z))
        return d_raw, d_z, None


def _march(model_output, z_values, lean):
    z = f32c(z_values)
    lead = z.shape[:-1]
    s = z.shape[-1]
    raw = f32c(model_output).reshape(-1, s, 4)
    return _RayMarching.apply(raw, z.reshape(-1, s), lean), lead, s


def ray_marching(model_output, z_values):
    """Alpha compositing.  Returns (rgb_image, weights, cumprod, alpha, net_rgb_output) like the reference (:88-115).

    Gradients w.r.t. model_output and z_values flow through rgb_image and weights (the only outputs the reference's
    train step differentiates); cumprod/alpha/net_rgb_output are returned for inspection/plots.
    """
    (rgb, weights, cumprod, alpha, rgb_s, _, _), lead, s = _march(model_output, z_values, False)
    shp = tuple(lead)
    return (rgb.reshape(shp + (3,)), weights.reshape(shp + (s,)), cumprod.reshape(shp + (s,)),
            alpha.reshape(shp + (s,)), rgb_s.reshape(shp + (s, 3)))


def ray_marching_lean(model_output, z_values):
    """Extension: only (rgb, weights, depth = sum w z, acc = sum w); 24 B/sample instead of 44 B/sample of HBM."""
    (rgb, weights, depth, acc), lead, s = _march(model_output, z_values, True)
    shp = tuple(lead)
    return rgb.reshape(shp + (3,)), weights.reshape(shp + (s,)), depth.reshape(shp), acc.reshape(shp)


def get_psnr(mse):
    """Peak signal-to-noise ratio for signals with peak value 1:  -10 ln(mse) / ln 10."""
    if not isinstance(mse, torch.Tensor):
        mse = torch.tensor(float(mse))
    return -10.0 * torch.log(mse) / math.log(10.0)


def get_psnr_for_image(image_source, image_target):
    a = torch.as_tensor(image_source, dtype=torch.float32)
    b = torch.as_tensor(image_target, dtype=torch.float32, device=a.device)
    return get_psnr(torch.mean((a - b) ** 2))


def c2w_to_rays_prepare_ds(c2w, field_of_view, img):
    """One camera -> (rays_orig (h*w,4), rays_dirs (h*w,4), rgb (h*w,3)); ray index = y*w + x."""
    img = f32c(img)
    h, w = img.shape[0], img.shape[1]
    dirs, orig = get_rays_directions(h, w, float(field_of_view), c2w, return_origins=True)
    return orig, dirs.reshape(-1, 4), img.reshape(-1, 3)


class RayDataset:
    """GPU-resident table of every training ray with a fresh permutation per epoch.

    Stands in for the tf.data pipeline of prepare_ds (:135-162): all (origin, direction, rgb) triples live in HBM
    (44 B/ray; 202 MB for 70 images of 256x256) and an epoch is a random permutation cut into batches; the last
    batch is ragged, as with ``Dataset.batch`` (Keras' fit never reaches it because steps_per_epoch is floor).
    """

    def __init__(self, batch_size, origs, dirs, rgbs, seed=None):
        self.batch_size = int(batch_size)
        self.origs, self.dirs, self.rgbs = origs, dirs, rgbs
        self.n_rays = origs.shape[0]
        self.gen = torch.Generator(device=origs.device)
        if seed is not None:
            self.gen.manual_seed(int(seed))

    def __len__(self):
        return (self.n_rays + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        # one gather of the whole table per epoch; the batches are then contiguous views (no per-step kernels)
        perm = torch.randperm(self.n_rays, device=self.origs.device, generator=self.gen)
        origs, dirs, rgbs = self.origs[perm], self.dirs[perm], self.rgbs[perm]
        for s in range(0, self.n_rays, self.batch_size):
            e = s + self.batch_size
            yield origs[s:e], dirs[s:e], rgbs[s:e]


class DevicePrefetcher:
    """``Dataset.prefetch`` (src/UtilsNeuralRadianceField.py:159-160) for batches that live in (pinned) HOST memory:
    yields device tuples, copying batch i+1 on a copy stream while batch i trains, so the PCIe latency of the three
    small H2D copies of a step is off the compute stream."""

    _streams = {}          # one copy stream per device, shared: its allocator pool stays warm across epochs

    def __init__(self, batches, device=None):
        self.batches = batches
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        key = str(self.device)
        if key not in DevicePrefetcher._streams:
            DevicePrefetcher._streams[key] = torch.cuda.Stream(device=self.device)
        self.stream = DevicePrefetcher._streams[key]

    def _load(self, batch):
        if batch is None:
            return None
        self.stream.wait_stream(torch.cuda.current_stream())       # never run ahead of memory the consumer still uses
        with torch.cuda.stream(self.stream):
            dev = tuple(t.to(device=self.device, dtype=torch.float32, non_blocking=True) for t in batch)
            ready = torch.cuda.Event()
            ready.record(self.stream)
        return dev, ready

    def __iter__(self):
        it = iter(self.batches)
        nxt = self._load(next(it, None))
        while nxt is not None:
            dev, ready = nxt
            cur = torch.cuda.current_stream()
            cur.wait_event(ready)
            for t in dev:
                t.record_stream(cur)
            nxt = self._load(next(it, None))
            yield dev


def prepare_ds(batch_size, c2w_matrices, images, fov, seed=None):
    """Training dataset of shuffled ray batches (rays_orig, rays_dirs, rgb)."""
    o, d, c = [], [], []
    for c2w, img in zip(c2w_matrices, images):
        oo, dd, cc = c2w_to_rays_prepare_ds(c2w, fov, img)
        o.append(oo)
        d.append(dd)
        c.append(cc)
    return RayDataset(batch_size, torch.cat(o), torch.cat(d), torch.cat(c), seed)


def model_predict(model, n_enc_phi_theta, n_pos_enc_for_xyz, xyz, view_dirs=None):
    """Encode the inputs and apply the network: returns (M,4) raw (R, G, B, Sigma)."""
    xyz_encoded = positional_encoding_for_xyz(xyz, n_pos_enc_for_xyz)
    if view_dirs is not None:
        dir_encoded = positional_encoding_for_views(view_dirs, n_enc_phi_theta)
        return model({XYZ_COORDS: xyz_encoded, VIEW_DIRS: dir_encoded})
    return model(xyz_encoded)


def render_rays(model, rays_orig, rays_dirs, z_values, n_pos_enc_for_xyz, n_pos_enc_for_angles, n_angles_for_model):
    """Render rays with one network: returns (render_result, weights, cumprod, alpha, rgb) like the reference."""
    coords_3d = sample_along_rays(rays_orig, rays_dirs, z_values)[..., :3]
    view_dirs = None if n_angles_for_model == 0 else get_view_directions(coords_3d, rays_dirs, n_angles_for_model)
    xyz = coords_3d.reshape(-1, 3)
    predictions = model_predict(model, n_pos_enc_for_angles, n_pos_enc_for_xyz, xyz, view_dirs)
    predictions = predictions.reshape(tuple(coords_3d.shape[:-1]) + (N_COLOR_CHANNELS + 1,))
    return ray_marching(predictions, z_values)


def get_num_of_batches(n_rays_in_batch, n_c2w_mats, h, w):
    return (n_c2w_mats * h * w) // n_rays_in_batch
