"""Host mirror of the ray/sampling functions of the reference's src/UtilsCV.py for the render hot path.

Same function names and positional arguments as the reference (get_rays_directions :467-499,
get_z_vals_from_prob_dist_func :502-539, get_z_values :565-581, sample_along_rays :584-599,
get_view_directions :124-143); tensors are contiguous fp32 ``torch.cuda`` tensors and every function launches
hand-written sm_100a kernels through the C ABI (include/nerf_b200.h).  The reference's unseeded
``tf.random.uniform`` becomes an explicit Philox stream: keyword-only ``seed/step/ray_offset`` (defaults come
from ``rng.next_step()``) or explicit ``jitter`` / ``u`` tensors.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import call, f32c, ptr

EPS = 1e-7


class _Rng:
    """Process-wide default Philox stream for calls that do not pass seed/step explicitly."""

    def __init__(self):
        self.seed = 0
        self.step = 0

    def set_seed(self, seed: int):
        self.seed = int(seed)
        self.step = 0

    def next_step(self):
        s = self.step
        self.step += 1
        return self.seed, s


rng = _Rng()


def _seed_step(seed, step):
    if seed is None:
        return rng.next_step()
    return int(seed), int(step)


def get_rays_directions(height, width, field_of_view, c2w, ray_begin=0, n_rays=None, return_origins=False):
    """Direction vector (h, w, 4) of the ray through every pixel centre, in world coordinates.

    Replaces src/UtilsCV.py:467-499.  ``ray_begin/n_rays`` (extension) restrict the work to a contiguous range of
    ray indices y*w+x (used for multi-GPU row sharding); then the result is (n_rays, 4).
    """
    c2w_np = np.ascontiguousarray(c2w.detach().cpu().numpy() if isinstance(c2w, torch.Tensor) else np.asarray(c2w),
                                  dtype=np.float32)
    if c2w_np.shape != (4, 4):
        raise ValueError("c2w must be a 4x4 camera-to-world matrix")
    full = n_rays is None
    n = height * width if full else int(n_rays)
    dev = torch.device("cuda", torch.cuda.current_device())
    dirs = torch.empty((n, 4), dtype=torch.float32, device=dev)
    origs = torch.empty((n, 4), dtype=torch.float32, device=dev) if return_origins else None
    call("nerf_ray_directions", c2w_np.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), float(field_of_view),
         int(height), int(width), int(ray_begin), n, ptr(dirs), ptr(origs))
    if full:
        dirs = dirs.reshape(height, width, 4)
    return (dirs, origs) if return_origins else dirs


def get_z_values(z_start, z_end, height, width, n_samples, *, jitter=None, seed=None, step=0, ray_offset=0):
    """Stratified depths (height, width, n_samples): linspace + U[0,1)*(z_end-z_start)/n_samples.

    Replaces src/UtilsCV.py:565-581 (jitter is always applied, also at inference, like the reference).
    """
    n = int(height) * int(width)
    dev = torch.device("cuda", torch.cuda.current_device())
    z = torch.empty((n, int(n_samples)), dtype=torch.float32, device=dev)
    if jitter is not None:
        jitter = f32c(jitter).reshape(n, int(n_samples))
        seed, step = 0, 0
    else:
        seed, step = _seed_step(seed, step)
    call("nerf_stratified_z", float(z_start), float(z_end), n, int(n_samples), ptr(jitter), seed, step,
         int(ray_offset), ptr(z))
    return z.reshape(int(height), int(width), int(n_samples))


class _SamplePdf(torch.autograd.Function):
    @staticmethod
    def forward(ctx, weights, z_values, num_new, u, seed, step, ray_offset, want_aux):
        n, s = weights.shape
        z_new = torch.empty((n, num_new), dtype=torch.float32, device=weights.device)
        perm = torch.empty((n, num_new), dtype=torch.int32, device=weights.device)
        u_used = torch.empty((n, num_new), dtype=torch.float32, device=weights.device)
        idx = torch.empty((n, num_new), dtype=torch.int32, device=weights.device) if want_aux else None
        call("nerf_sample_pdf_fwd", ptr(weights), ptr(z_values), n, s, num_new, ptr(u), seed, step, ray_offset,
             ptr(z_new), ptr(idx), ptr(perm), ptr(u_used))
        ctx.save_for_backward(weights, z_values, u_used, perm)
        ctx.num_new = num_new
        if want_aux:
            ctx.mark_non_differentiable(idx, perm)
            return z_new, idx, perm
        return z_new

    @staticmethod
    def backward(ctx, d_z_new, *unused):
        weights, z_values, u_used, perm = ctx.saved_tensors
        n, s = weights.shape
        d_w = torch.empty_like(weights)
        call("nerf_sample_pdf_bwd", ptr(weights), ptr(z_values), ptr(u_used), ptr(perm),
             ptr(d_z_new.contiguous().float()), n, s, ctx.num_new, ptr(d_w))
        return d_w, None, None, None, None, None, None, None


def get_z_vals_from_prob_dist_func(weights, z_values, num_new_z_values, *, u=None, seed=None, step=0, ray_offset=0,
                                   return_aux=False):
    """Inverse-transform sampling of ``num_new_z_values`` depths per ray from the coarse weights, sorted.

    Replaces src/UtilsCV.py:502-539.  Differentiable w.r.t. ``weights`` exactly like the reference (which does not
    stop the gradient).  ``return_aux=True`` also returns the int32 searchsorted indices (draw order) and the
    sort permutation.
    """
    lead = weights.shape[:-1]
    s = weights.shape[-1]
    w2 = f32c(weights).reshape(-1, s)
    z2 = f32c(z_values).reshape(-1, s)
    if u is not None:
        u = f32c(u).reshape(-1, int(num_new_z_values))
        seed, step = 0, 0
    else:
        seed, step = _seed_step(seed, step)
    out = _SamplePdf.apply(w2, z2, int(num_new_z_values), u, seed, step, int(ray_offset), bool(return_aux))
    if return_aux:
        z_new, idx, perm = out
        shp = tuple(lead) + (int(num_new_z_values),)
        return z_new.reshape(shp), idx.reshape(shp), perm.reshape(shp)
    return out.reshape(tuple(lead) + (int(num_new_z_values),))


class _SampleAlongRays(torch.autograd.Function):
    @staticmethod
    def forward(ctx, origin, dirs, z):
        n, s = z.shape
        out = torch.empty((n, s, 4), dtype=torch.float32, device=z.device)
        call("nerf_sample_along_rays", ptr(origin), ptr(dirs), ptr(z), n, s, ptr(out))
        ctx.save_for_backward(dirs)
        return out

    @staticmethod
    def backward(ctx, g):
        (dirs,) = ctx.saved_tensors
        # d z = sum_c g_c * d_c ; origin/direction gradients are never needed on this path
        return None, None, (g * dirs[:, None, :]).sum(-1)


def sample_along_rays(origin, direction_vectors, z_values):
    """Sample coordinates origin + direction * z, shape (N, S, 4).  Replaces src/UtilsCV.py:584-599."""
    o, d, z = f32c(origin), f32c(direction_vectors), f32c(z_values)
    lead = z.shape[:-1]
    out = _SampleAlongRays.apply(o.reshape(-1, 4), d.reshape(-1, 4), z.reshape(-1, z.shape[-1]))
    return out.reshape(tuple(lead) + (z.shape[-1], 4))


def get_view_directions(coords_3d, rays_dirs, n_angles_for_model):
    """View-direction input of the network, (N*S, n_angles+1).  Replaces src/UtilsCV.py:124-143."""
    if n_angles_for_model not in (1, 2):
        raise Exception("n_angles_for_model should be 1 or 2.")
    d = f32c(rays_dirs).reshape(-1, 4)
    n = d.shape[0]
    s = int(np.prod(coords_3d.shape[:-1])) // max(n, 1)
    out = torch.empty((n * s, n_angles_for_model + 1), dtype=torch.float32, device=d.device)
    call("nerf_view_directions", ptr(d), n, s, int(n_angles_for_model), ptr(out))
    return out


# ---- depth-map visualisation (src/UtilsCV.py:700-760), used by ExecutionRun.render_video --------------------------------
RGB_TO_YIQ = np.array([[0.299, 0.587, 0.114], [0.596, -0.275, -0.321], [0.212, -0.523, 0.311]])


def rgb2yiq(im_rgb):
    return np.dot(im_rgb, RGB_TO_YIQ.T)


def yiq2rgb(im_yiq):
    return np.dot(im_yiq, np.linalg.inv(RGB_TO_YIQ).T)


def _equalized_levels(gray):
    """Grays of any range -> (levels in 0..255 as float64, histogram before, histogram after), or None for an image
    without contrast.  The algorithm of src/UtilsCV.py:724-743: stretch to [0, 255]; count the stretched values per
    unit-wide bin (bin = floor); running sum of the counts, re-based at its first non-zero entry and scaled to 255, is
    the look-up table; the table is read at the stretched value ROUNDED to nearest-even, not at its bin."""
    gray = np.asarray(gray)
    lo, hi = gray.min(), gray.max()
    if hi == lo:
        return None
    stretched = (gray - lo) / (hi - lo) * 255
    counts = np.bincount(np.floor(stretched).astype(np.int64).ravel(), minlength=256)
    running = np.cumsum(counts)
    base = running[running > 0][0]
    table = np.rint((running - base) / (running[-1] - base) * 255)
    levels = table[np.rint(stretched).astype(np.int64)]
    return levels, counts, np.bincount(levels.astype(np.int64).ravel(), minlength=256)


def histogram_equalize(im_orig):
    """-> [im_eq in [0,1], hist_orig, hist_eq] (256-bin histograms); RGB images are equalised on the Y channel of YIQ.
    Host NumPy, the contract of src/UtilsCV.py:700-721; the input is not modified.  An all-zero image comes back as it
    is with ``None`` histograms (the reference's early return); so does any other constant image, where the reference
    divides 0 by 0."""
    im = np.array(im_orig, copy=True)
    luma = rgb2yiq(im)[:, :, 0] if im.ndim == 3 else im
    res = _equalized_levels(luma)
    if res is None:
        return (im if im.ndim == 3 or not im.any() else np.zeros_like(im)), None, None
    levels, hist_orig, hist_eq = res
    if im.ndim == 3:
        yiq = rgb2yiq(im)
        yiq[:, :, 0] = levels / 255
        return yiq2rgb(yiq), hist_orig, hist_eq
    return levels / 255, hist_orig, hist_eq


def histogram_equalize_frames(depth):
    """The gray branch of ``histogram_equalize`` for a stack of frames (F,h,w) where they are rendered: a torch tensor
    on any device in, uint8 levels (F,h,w) out -- ``level/255`` is the reference's ``im_eq`` and ``level`` is what its
    video writer stores (``uint8(round(im_eq*255))``, src/UtilsVideo.py:34).  Same float32 stretch and float64 table as
    the NumPy version, so the levels are identical; a video frame leaves the GPU as h*w bytes instead of h*w floats."""
    f = depth.shape[0]
    if depth.numel() == 0:                                       # a rank that was dealt no frame
        return torch.empty(depth.shape, dtype=torch.uint8, device=depth.device)
    g = depth.reshape(f, -1).to(torch.float32)
    g = g - g.min(dim=1, keepdim=True).values
    top = g.max(dim=1, keepdim=True).values
    flat = top == 0                                              # all-zero / constant frames stay zero
    g = g / torch.where(flat, torch.ones_like(top), top) * 255
    idx = g.floor().clamp_(0, 255).long()
    hist = torch.zeros((f, 256), dtype=torch.int64, device=g.device).scatter_add_(1, idx, torch.ones_like(idx))
    cum = hist.cumsum(1)
    first = cum.gather(1, (cum > 0).to(torch.int64).argmax(dim=1, keepdim=True))
    denom = (cum[:, -1:] - first).clamp_(min=1).to(torch.float64)
    table = torch.round((cum - first).to(torch.float64) / denom * 255)
    levels = table.gather(1, torch.round(g).long().clamp_(0, 255)).to(torch.uint8)
    levels = torch.where(flat, torch.zeros_like(levels), levels)
    return levels.reshape(depth.shape)
