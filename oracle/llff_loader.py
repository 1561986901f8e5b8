"""Restatement of the reference's COLMAP/LLFF dataset loader, used ONLY to build the oracle-pin fixture.
TEST INFRASTRUCTURE (follows src/UtilsFiles.py:73-130 and src/UtilsCV.py:274-330; images are read with OpenCV
because imageio is not installed here -- JPEG decoders may differ by a unit in the last place of a pixel)."""
import os

import numpy as np


def _normalize(x):
    return x / np.linalg.norm(x, axis=-1)[..., None]                                  # src/UtilsCV.py:256


def _orthonormal(z, y):
    v2 = _normalize(z)                                                                # src/UtilsCV.py:267-271
    v0 = _normalize(np.cross(y, v2))
    v1 = _normalize(np.cross(v2, v0))
    return np.stack([v0, v1, v2], 1)


def _homogeneous(mats):
    return np.concatenate([mats, np.tile(np.reshape(np.eye(4)[-1, :], [1, 1, 4]), [mats.shape[0], 1, 1])], 1)


def recenter_poses(poses_hwf):
    """src/UtilsCV.py:286-298."""
    p = poses_hwf[:, :3, :4]
    t, r3, r2 = p[:, :3, 3].mean(0), p[:, :3, 2].mean(0), p[:, :3, 1].mean(0)
    avg = np.concatenate([_orthonormal(r3, r2), t[:, None]], 1)                       # poses_avg, :274-283
    avg = _homogeneous(avg[None])[0]
    poses = np.linalg.inv(avg) @ _homogeneous(p)
    poses_hwf[:, :3, :4] = poses[:, :3, :]
    return poses_hwf, avg


def spherify_poses(poses_hwf, bounds):
    """src/UtilsCV.py:320-330."""
    radius = np.sqrt(np.max(np.sum(np.square(poses_hwf[:, :3, 3]), -1)))
    scale = 1.0 / radius
    poses_hwf[:, :3, 3] *= scale
    return poses_hwf, bounds * scale, scale


def load_colmap(dataset_location):
    """images (n,h,w,3) float32 in [0,1], c2w (n,4,4) float32, fov, near, far  (src/UtilsFiles.py:73-130)."""
    import cv2
    raw = np.load(os.path.join(dataset_location, "poses_bounds.npy"))
    poses_hwf = raw[:, :-2].reshape([-1, 3, 5])
    poses_hwf = poses_hwf[:, :, [1, 0, 2, 3, 4]]                                      # [-y, x, z] -> [x, y, z], :111-112
    poses_hwf[:, :, 1] = -poses_hwf[:, :, 1]
    bounds = np.moveaxis(raw[:, -2:].transpose([1, 0]), -1, 0)
    poses_hwf, _ = recenter_poses(poses_hwf)
    poses_hwf, bounds, scale = spherify_poses(poses_hwf, bounds)
    names = sorted(n for n in os.listdir(dataset_location) if n.lower().endswith(("jpg", "png")))
    images = np.asarray([cv2.cvtColor(cv2.imread(os.path.join(dataset_location, n)), cv2.COLOR_BGR2RGB)[..., :3] / 255.0
                         for n in names], dtype=np.float32)
    h, w, focal = poses_hwf[0, :3, -1]
    near, far = float(bounds.min() * 0.9), float(bounds.max() * 1.0)                  # :87-88
    fov = float(np.arctan2(w / 2, focal) * 2)                                         # :91
    c2w = _homogeneous(poses_hwf[:, :3, :4]).astype(np.float32)
    return images, c2w, fov, near, far, float(scale)
