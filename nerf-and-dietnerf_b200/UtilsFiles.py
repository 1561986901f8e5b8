"""Dataset loaders and checkpoint helpers: host-side mirror of the parts of src/UtilsFiles.py the train/render loop
needs.  NumPy only; runs once per experiment.

  get_data_from_colmap   src/UtilsFiles.py:73-98    LLFF ``poses_bounds.npy`` + images
  load_llff_data         src/UtilsFiles.py:101-130
  get_data_from_blender  src/UtilsFiles.py:35-70    ``cam_data.json`` + images
  recenter_poses / spherify_poses / poses_avg       src/UtilsCV.py:274-330
  load_config, save_weights, save_psnr_values, get_psnr_values, get_psnr_save_path   src/UtilsFiles.py:153-209
"""
import json
import os
from pathlib import Path

import numpy as np

POSES_BOUNDS_NPY = 'poses_bounds.npy'
CAM_DATA_JSON_FILE_NAME = 'cam_data.json'
DIRNAME_TO_SAVE_PSNRS = 'saved_test_train_psnrs'
NAME_PSNR_FILE = 'psnrs_train_test_{:03}.npy'


def imread(path):
    """uint8 (h, w, 3|4) RGB(A).  The reference uses imageio; Pillow / OpenCV decode the same files (JPEG decoders may
    differ by one code value in a pixel)."""
    try:
        from PIL import Image
        with Image.open(str(path)) as im:
            return np.asarray(im.convert("RGBA" if im.mode in ("RGBA", "LA", "P") else "RGB"))
    except ImportError:
        import cv2
        img = cv2.imread(str(path), cv2.IMREAD_UNCHANGED)
        if img is None:
            raise FileNotFoundError(path)
        return cv2.cvtColor(img, cv2.COLOR_BGRA2RGBA if img.shape[-1] == 4 else cv2.COLOR_BGR2RGB)


# ---- pose normalisation (src/UtilsCV.py:250-330) ----------------------------------------------------------------------
def normalize_vectors(x):
    return x / np.linalg.norm(x, axis=-1)[..., None]


def get_orthonormal_mat_from_2_vecs(z, y):
    vec2 = normalize_vectors(z)
    vec0 = normalize_vectors(np.cross(y, vec2))
    vec1 = normalize_vectors(np.cross(vec2, vec0))
    return np.stack([vec0, vec1, vec2], 1)


def change_mats_to_homogeneous(mats):
    bottom = np.tile(np.reshape(np.eye(4)[-1, :], [1, 1, 4]), [mats.shape[0], 1, 1])
    return np.concatenate([mats, bottom], 1)


def poses_avg(poses):
    t = poses[:, :3, 3].mean(0)
    r3 = poses[:, :3, 2].mean(0)
    r2 = poses[:, :3, 1].mean(0)
    return np.concatenate([get_orthonormal_mat_from_2_vecs(r3, r2), t[:, None]], 1)


def recenter_poses(poses_hwf):
    """Express every pose in the frame of the average pose (in place on [:, :3, :4]); returns (poses, average c2w)."""
    average_c2w = change_mats_to_homogeneous(poses_avg(poses_hwf[:, :3, :4])[..., :4][None])[0]
    poses = np.linalg.inv(average_c2w) @ change_mats_to_homogeneous(poses_hwf[:, :3, :4])
    poses_hwf[:, :3, :4] = poses[:, :3, :]
    return poses_hwf, average_c2w


def spherify_poses(poses_hwf, bounds):
    """Scale the scene so the farthest camera sits on the unit sphere; bounds scale with it."""
    radius = np.sqrt(np.max(np.sum(np.square(poses_hwf[:, :3, 3]), -1)))
    scale = 1.0 / radius
    poses_hwf[:, :3, 3] *= scale
    bounds = bounds * scale
    return poses_hwf, bounds, scale


# ---- loaders ------------------------------------------------------------------------------------------------------------
def load_llff_data(path_to_images):
    raw = np.load(os.path.join(str(path_to_images), POSES_BOUNDS_NPY))
    poses_hwf = raw[:, :-2].reshape([-1, 3, 5])
    poses_hwf = poses_hwf[:, :, [1, 0, 2, 3, 4]]            # stored as [-y, x, z]: reorder and flip to [x, y, z]
    poses_hwf[:, :, 1] = -poses_hwf[:, :, 1]
    bounds = np.moveaxis(raw[:, -2:].transpose([1, 0]), -1, 0)
    poses_hwf, average_c2w_before_recenter = recenter_poses(poses_hwf)
    poses_hwf, bounds, scale = spherify_poses(poses_hwf, bounds)
    names = sorted(n for n in os.listdir(str(path_to_images)) if n.endswith(('JPG', 'jpg', 'png')))
    images = np.asarray([imread(os.path.join(str(path_to_images), n))[..., :3] / 255.0 for n in names], dtype=np.float32)
    return images, poses_hwf, bounds, average_c2w_before_recenter, scale


def get_data_from_colmap(dataset_location):
    """-> images (n,h,w,3) in [0,1], c2w (n,4,4) float32, field_of_view [rad], near, far, average c2w, scale."""
    images, poses, bds, average_c2w_before_recenter, scale = load_llff_data(dataset_location)
    h, w, focal = poses[0, :3, -1]
    poses = poses[:, :3, :4]
    near, far = float(np.min(bds) * 0.9), float(np.max(bds) * 1.0)
    field_of_view = np.arctan2(w / 2, focal) * 2
    poses = np.concatenate([poses, np.tile(np.reshape([0, 0, 0, 1], [1, 1, 4]), [poses.shape[0], 1, 1])], -2)
    return images.astype(np.float32), poses.astype(np.float32), float(field_of_view), near, far, \
        average_c2w_before_recenter, scale


def get_data_from_blender(dataset_location, near_boundary, far_boundary):
    """-> same tuple as get_data_from_colmap, for a ``cam_data.json`` dataset rendered by the reference's Blender
    script; near/far come from the YAML and are rescaled with the poses."""
    dataset_location = Path(dataset_location)
    with open(dataset_location / CAM_DATA_JSON_FILE_NAME, 'r') as f:
        meta = json.load(f)
    mats, images = [], []
    for frame in meta['frames']:
        mats.append(frame['transformation_matrix'])
        images.append(imread(dataset_location / frame['filename']))
    images, cams = np.asarray(images, dtype=np.float32), np.asarray(mats, dtype=np.float64)
    cams, average_c2w_before_recenter = recenter_poses(cams)
    cams, bounds, scale = spherify_poses(cams, np.array([near_boundary, far_boundary], dtype=np.float64))
    return images / 255.0, cams.astype(np.float32), float(meta['field_of_view']), float(bounds[0]), float(bounds[1]), \
        average_c2w_before_recenter, scale


# ---- config / checkpoints ----------------------------------------------------------------------------------------------
def load_config(config_file_path):
    import yaml
    if not os.path.exists(config_file_path):
        raise Exception(f"Config file '{config_file_path}' not found.")
    with open(config_file_path, 'r') as f:
        return yaml.safe_load(f)


def save_weights(neural_net, filepath):
    os.makedirs(os.path.dirname(str(filepath)) or ".", exist_ok=True)
    neural_net.save_weights(filepath)


def get_psnr_save_path(save_location, epoch_number):
    return Path(save_location) / DIRNAME_TO_SAVE_PSNRS / NAME_PSNR_FILE.format(epoch_number)


def save_psnr_values(psnrs_test_values, psnrs_train_values, filepath):
    os.makedirs(os.path.dirname(str(filepath)) or ".", exist_ok=True)
    np.save(str(filepath), (np.asarray(psnrs_test_values), np.asarray(psnrs_train_values)))


def get_psnr_values(path_to_existing_psnr_values):
    if path_to_existing_psnr_values and os.path.exists(path_to_existing_psnr_values):
        psnrs_test_values, psnrs_train_values = np.load(str(path_to_existing_psnr_values))
        return list(psnrs_test_values), list(psnrs_train_values)
    return [], []
