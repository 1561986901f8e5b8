"""Shared test inputs: seeded synthetic rays, networks with a sharpened sigma head, config dicts."""
import math

import numpy as np
import torch

from oracle import nerf_oracle as O

NEAR, FAR = 0.5576, 2.5635   # Alexander scene after the loader's rescale


def net_config(n_angles=2, l_view=4, batch_train=256, batch_render=256):
    return {"hidden_layer_dim": 256, "last_hidden_layer_dim": 128, "leaky_relu_alpha": 0.05, "n_pos_enc_dim_xyz": 5,
            "n_pos_enc_view_dir": l_view, "n_angles_for_model": n_angles, "n_rays_in_batch_train": batch_train,
            "n_rays_in_batch_render": batch_render}


def render_config(n_c=64, n_f=128):
    return {"n_render_samples_coarse": n_c, "n_render_samples_fine": n_f}


def oracle_cfg(n_angles=2, l_view=4):
    return O.NetCfg(5, l_view, n_angles, 256, 128, 0.05)


def make_params(cfg, seed, sigma_gain=30.0, sigma_bias=0.5):
    """Glorot weights, random biases, and a sharpened sigma head: rays see empty space AND opaque surfaces
    (alpha saturates), which exercises the div_no_nan branch of the cumprod gradient and empty-ray collapse."""
    p = O.glorot_params(cfg.shapes, seed, bias_scale=0.1)
    layers = O.unflatten(p, cfg.shapes)
    w, b = layers[-1]            # sigma head is the last Dense in both variants
    w *= sigma_gain
    b += sigma_bias
    return p


def sphere_pose(theta, phi, radius=1.0):
    """Camera on a sphere looking at the origin (OpenGL convention: camera looks down -z)."""
    cam = np.array([radius * math.cos(phi) * math.sin(theta), radius * math.sin(phi),
                    radius * math.cos(phi) * math.cos(theta)])
    fwd = -cam / np.linalg.norm(cam)
    up = np.array([0.0, 1.0, 0.0])
    right = np.cross(fwd, up)
    right /= np.linalg.norm(right)
    up2 = np.cross(right, fwd)
    c2w = np.eye(4, dtype=np.float32)
    c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = right, up2, -fwd, cam
    return c2w


def random_rays(n, seed=0, fov=0.69):
    """n rays from a few sphere cameras (what prepare_ds would feed), as CPU tensors (N,4),(N,4)."""
    rng = np.random.default_rng(seed)
    per = 16
    o_all, d_all = [], []
    while sum(x.shape[0] for x in o_all) < n:
        c2w = sphere_pose(rng.uniform(0, 2 * math.pi), rng.uniform(-0.5, 0.5), rng.uniform(0.8, 1.2))
        o, d = O.rays_for_image(c2w, fov, per, per)
        o_all.append(o)
        d_all.append(d)
    o, d = torch.cat(o_all)[:n], torch.cat(d_all)[:n]
    perm = torch.from_numpy(rng.permutation(n))
    return o[perm].contiguous(), d[perm].contiguous()
