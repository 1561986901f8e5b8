"""Dataset loaders and run-directory helpers of the host loop (src/UtilsFiles.py:35-130, :232-281,
src/UtilsCV.py:274-330) on small synthetic datasets written to a temp directory -- CPU only."""
import importlib
import json
import math
import os

import numpy as np
import pytest

from helpers import sphere_pose
from oracle.llff_loader import load_colmap

files = importlib.import_module("nerf-and-dietnerf_b200.UtilsFiles")
poses_mod = importlib.import_module("nerf-and-dietnerf_b200.poses")


def _write_images(folder, n, h, w, ext):
    from PIL import Image
    rng = np.random.default_rng(0)
    imgs = rng.integers(0, 256, size=(n, h, w, 3), dtype=np.uint8)
    names = []
    for i, im in enumerate(imgs):
        name = f"img_{i:03d}.{ext}"
        Image.fromarray(im).save(os.path.join(folder, name))
        names.append(name)
    return imgs, names


def test_blender_loader(tmp_path):
    n, h, w = 6, 8, 10
    imgs, names = _write_images(tmp_path, n, h, w, "png")
    cams = [sphere_pose(0.7 * i, 0.2 * (i % 3), 4.0).astype(np.float64) for i in range(n)]
    meta = {"field_of_view": 0.69111, "frames": [{"filename": nm, "transformation_matrix": c.tolist()}
                                                for nm, c in zip(names, cams)]}
    with open(tmp_path / "cam_data.json", "w") as f:
        json.dump(meta, f)
    images, c2w, fov, near, far, avg, scale = files.get_data_from_blender(tmp_path, 2.0, 6.0)
    assert images.shape == (n, h, w, 3) and images.dtype == np.float32
    assert np.array_equal(np.round(images * 255).astype(np.uint8), imgs)          # PNG is lossless
    assert c2w.shape == (n, 4, 4) and c2w.dtype == np.float32 and fov == pytest.approx(0.69111)
    radius = np.linalg.norm(c2w[:, :3, 3], axis=-1)
    assert radius.max() == pytest.approx(1.0, abs=1e-6)                           # spherify: farthest camera on the unit sphere
    assert near == pytest.approx(2.0 * scale) and far == pytest.approx(6.0 * scale)
    # recentre: poses are expressed in the frame of the average pose, rotations stay orthonormal
    for m in c2w:
        assert np.allclose(m[:3, :3] @ m[:3, :3].T, np.eye(3), atol=1e-5)
    assert avg.shape == (4, 4)


def test_colmap_loader_matches_the_oracle_restatement(tmp_path):
    n, h, w, focal = 7, 12, 12, 30.0
    _write_images(tmp_path, n, h, w, "jpg")
    rng = np.random.default_rng(1)
    rows = []
    for i in range(n):
        c = sphere_pose(0.5 * i, 0.1 * i, 3.0 + 0.1 * i)[:3, :4].astype(np.float64)
        llff = np.concatenate([c[:, 1:2], -c[:, 0:1], c[:, 2:3], c[:, 3:4]], 1)     # stored as [-y, x, z, t]: see the loader
        llff = np.concatenate([llff, np.array([[h], [w], [focal]])], 1)             # hwf column
        rows.append(np.concatenate([llff.reshape(-1), [1.5 + 0.1 * rng.random(), 7.0 + rng.random()]]))
    np.save(tmp_path / "poses_bounds.npy", np.asarray(rows))
    images, c2w, fov, near, far, avg, scale = files.get_data_from_colmap(tmp_path)
    o_images, o_c2w, o_fov, o_near, o_far, o_scale = load_colmap(str(tmp_path))
    assert images.shape == (n, h, w, 3)
    assert np.abs(images - o_images).max() <= 3.0 / 255           # Pillow vs OpenCV JPEG decoders
    assert np.abs(c2w - o_c2w).max() < 1e-6 and fov == pytest.approx(o_fov) and near == pytest.approx(o_near)
    assert far == pytest.approx(o_far) and scale == pytest.approx(o_scale)
    assert fov == pytest.approx(2 * math.atan2(w / 2, focal))
    assert np.linalg.norm(c2w[:, :3, 3], axis=-1).max() == pytest.approx(1.0, abs=1e-6)


def test_save_location_and_psnr_files(tmp_path):
    run = importlib.import_module("nerf-and-dietnerf_b200.ExecutionRun")
    cfg = {"general_save_location": str(tmp_path / "Results"), "existing_save_dir_name": None}
    first = run.get_save_location("config_files/50px_scene.yaml", cfg)
    second = run.get_save_location("config_files/50px_scene.yaml", cfg)
    assert first.name == "50px_scene_save_dir_0" and second.name == "50px_scene_save_dir_1"
    cfg["existing_save_dir_name"] = "50px_scene_save_dir_1"
    assert run.get_save_location("config_files/50px_scene.yaml", cfg) == second
    cfg["existing_save_dir_name"] = "missing"
    with pytest.raises(Exception, match="does not exists"):
        run.get_save_location("config_files/50px_scene.yaml", cfg)
    path = files.get_psnr_save_path(first, 7)
    assert str(path).endswith("saved_test_train_psnrs/psnrs_train_test_007.npy")
    files.save_psnr_values([20.0, 21.5], [22.0, 23.5], path)
    test, train = files.get_psnr_values(path)
    assert test == [20.0, 21.5] and train == [22.0, 23.5]
    assert files.get_psnr_values(files.get_psnr_save_path(first, 8)) == ([], [])
    with pytest.raises(Exception, match="not found"):
        files.load_config(tmp_path / "nope.yaml")


def test_point_of_interest_of_a_spherical_capture():
    cams = np.stack([sphere_pose(0.4 * i, 0.15 * (i % 4) - 0.2, 1.0) for i in range(16)])
    cams[:, :3, 3] += np.array([0.1, -0.05, 0.2])                  # every camera looks at this point
    point, spherical = poses_mod.estimate_point_of_interest_in_scene(cams, num_iter=200, rng=np.random.RandomState(0))
    assert spherical and np.allclose(point, [0.1, -0.05, 0.2], atol=1e-4)
    # a forward-facing capture (parallel optical axes) is not spherical
    ff = np.tile(np.eye(4), (8, 1, 1))
    ff[:, 0, 3] = np.linspace(-1, 1, 8)
    _, spherical = poses_mod.estimate_point_of_interest_in_scene(ff, num_iter=50, rng=np.random.RandomState(0))
    assert not spherical


def test_every_reference_config_maps_to_a_supported_mode():
    """ExecutionRun.effective_mode (CPU: only the plan query of the C ABI): every YAML the reference ships runs in the
    default tensor-core mode -- the xyz-only network (n_angles_for_model: 0, src/NeRF.py:248-288, 5 configs) has a plan too
    since round 2 -- and a network outside the plans falls back to the SIMT fp32 kernels instead of dying in the
    constructor."""
    import importlib
    import json
    from pathlib import Path
    pkg = importlib.import_module("nerf-and-dietnerf_b200")
    pkg.load()
    ER = importlib.import_module("nerf-and-dietnerf_b200.ExecutionRun")
    census = json.loads((Path(__file__).parent / "golden" / "config_census.json").read_text())
    census = {k: v for k, v in census.items() if "unparseable" not in v}
    n_xyz_only = 0
    for name, cfg in census.items():
        net = cfg["neural_net"]
        run = ER.ExecutionRun.__new__(ER.ExecutionRun)
        run.mode, run.net_config, run.is_main = "fp16", net, False
        assert run.effective_mode() == "fp16", name
        n_xyz_only += net["n_angles_for_model"] == 0
        run.mode = "fp32"
        assert run.effective_mode() == "fp32"
    assert len(census) >= 46 and n_xyz_only == 5
    odd = dict(next(iter(census.values()))["neural_net"], hidden_layer_dim=128)
    run.mode, run.net_config = "fp16", odd
    assert run.effective_mode() == "fp32"
