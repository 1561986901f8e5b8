"""Video row of the host loop (src/ExecutionRun.py:299-448, src/UtilsVideo.py:16-39, pose geometry of
src/UtilsCV.py:146-247, :407-437, :612-760) -- CPU tests.

The golden data are frames DECODED FROM THE VIDEOS THE REFERENCE RECORDED for its 50 px Alexander run
(tests/golden/alexander50_videos.npz, built by tests/golden/make_alexander50_videos.py): the trajectories of
ExecutionRun must produce the recorded frame counts, the dataset video must reproduce the recorded one, and the ORACLE
rendering the reference's epoch-95 weights at those trajectory poses must reproduce the recorded frames (rgb and
equalised depth) up to MJPG loss and the unseeded jitter -- a pin of the oracle, the pose generators and the depth
visualisation against outputs of the reference's TensorFlow path."""
import importlib
import json
import os

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import nerf_oracle as O

GOLD = os.path.join(ROOT, "tests", "golden")
PATH_INDICES = [4, 7, 15, 20, 28, 37, 48, 41, 54, 62, 70]      # img_indices_for_path_video of the run's YAML
RUN_CONFIG = {      # Results/50px_alexander_71pics_sphere_nerf_save_dir_4/50px_alexander_71pics_sphere_nerf.yaml
    "existing_save_dir_name": None, "starting_epoch_number": 95, "dataset_type": "colmap",
    "tasks_to_perform": {"start_training": False},
    "neural_net": {"type_of_model": "NeRF", "hidden_layer_dim": 256, "last_hidden_layer_dim": 128,
                   "leaky_relu_alpha": 0.05, "n_pos_enc_dim_xyz": 5, "n_pos_enc_view_dir": 4, "n_angles_for_model": 2,
                   "n_rays_in_batch_train": 4096, "n_rays_in_batch_render": 4096},
    "render": {"n_render_samples_coarse": 64, "n_render_samples_fine": 128},
    "training": {"n_epochs": 95, "optimizer_lr": 4.0e-4, "test_img_idx": 19, "idx_train_img_to_plot": 4},
    "video": {"fps_plot_video": 5, "fps_render_video": 60, "fps_train_set_video": 5,
              "img_indices_for_path_video": PATH_INDICES},
}


def psnr_u8(a, b):
    return float(-10 * np.log10(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2) / 255.0 ** 2 + 1e-12))


def alexander_run(save_location=None, tasks=None, seed=0):
    """ExecutionRun on the reference's scene with the recorded run's YAML values (no file system, no GPU needed until
    a model is asked for)."""
    pkg = importlib.import_module("nerf-and-dietnerf_b200")
    data = np.load(os.path.join(GOLD, "alexander50_dataset.npz"))
    config = json.loads(json.dumps(RUN_CONFIG))
    if tasks:
        config["tasks_to_perform"].update(tasks)
    images = data["images_u8"].astype(np.float32) / 255.0
    run = pkg.ExecutionRun.from_arrays(config, images, data["c2w"], float(data["fov"]), float(data["near"]),
                                       float(data["far"]), seed=seed, save_location=save_location)
    return pkg, run


@pytest.fixture(scope="module")
def videos():
    return np.load(os.path.join(GOLD, "alexander50_videos.npz"))


@pytest.fixture(scope="module")
def run_and_pkg():
    pkg, run = alexander_run()
    return run, pkg


# ---- pose geometry ------------------------------------------------------------------------------------------------------
def test_rotation_helpers():
    P = importlib.import_module("nerf-and-dietnerf_b200").poses
    a, b = P.get_sphere_matrix(1, 30, 40, 10)[:3, :3], P.get_sphere_matrix(1, -70, 200, 5)[:3, :3]
    r = P.get_rotation_matrix_from_source_to_dest_mats(a, b)
    assert r.shape == (4, 4) and np.abs(r[:3, :3] @ a - b).max() < 1e-12 and np.array_equal(r[3], [0, 0, 0, 1])
    assert np.abs(P.get_rotation_matrix_from_source_to_dest_mats(a, a) - np.eye(4)).max() < 1e-12
    v1, v2 = np.array([1.0, 2.0, 3.0]), np.array([-2.0, 0.5, 1.0])
    n1, n2 = v1 / np.linalg.norm(v1), v2 / np.linalg.norm(v2)
    m = P.get_rotation_matrix_from_v1_to_v2(v1, v2)
    assert np.abs(m @ n1 - n2).max() < 1e-12 and abs(np.linalg.det(m) - 1) < 1e-12
    q = P.get_rotation_quaternion_from_vec1_to_vec2(v1, v2)
    assert np.abs(P.rotate_vec_with_quaternion(n1, q) - n2).max() < 1e-12
    # the reference's special cases: opposite vectors turn by pi (also when they lie on the x axis), equal ones do not turn
    assert np.abs(P.rotate_vec_with_quaternion(v1, P.get_rotation_quaternion_from_vec1_to_vec2(v1, -v1)) + v1).max() < 1e-12
    qx = P.get_rotation_quaternion_from_vec1_to_vec2([1, 0, 0], [-1, 0, 0])
    assert np.abs(P.rotate_vec_with_quaternion([1, 0, 0], qx) - [-1, 0, 0]).max() < 1e-12
    assert np.array_equal(P.get_rotation_quaternion_from_vec1_to_vec2(v1, 2 * v1), [1, 0, 0, 0])


def test_interpolated_trajectories():
    P = importlib.import_module("nerf-and-dietnerf_b200").poses
    c1, c2 = P.get_sphere_matrix(1, 30, 40, 10), P.get_sphere_matrix(0.8, -20, 100, 5)
    plain = P.get_c2w_matrices_between_2_c2w(c1, c2, 9)
    slow = P.get_c2w_matrices_between_2_c2w_with_stretch(c1, c2, 9)
    assert len(plain) == len(slow) == 9
    for seq in (plain, slow):
        assert np.abs(seq[0] - c1).max() < 1e-6 and np.abs(seq[-1] - c2).max() < 1e-6
        for m in seq:
            assert np.abs(m[:3, :3] @ m[:3, :3].T - np.eye(3)).max() < 1e-5
    # translation is a lerp, so the interpolation weight can be read back: even steps / stretched steps a(9)/(a+2)
    weight = lambda m: float(np.dot(m[:3, 3] - c1[:3, 3], c2[:3, 3] - c1[:3, 3]) / np.sum((c2[:3, 3] - c1[:3, 3]) ** 2))
    alpha = np.linspace(0, 1, 9)
    assert np.allclose([weight(m) for m in plain], alpha, atol=1e-5)
    expected = alpha / (alpha + 2)
    expected = (expected - expected.min()) / (expected.max() - expected.min())
    assert np.allclose([weight(m) for m in slow], expected, atol=1e-5)
    assert np.all(np.diff(expected) > 0) and expected[4] > 0.5              # fast start, slow finish


# ---- depth visualisation ----------------------------------------------------------------------------------------------
def test_histogram_equalize_worked_example():
    U = importlib.import_module("nerf-and-dietnerf_b200").UtilsCV
    # four pixels 0, .25, .5, 1 -> stretched 0, 63.75, 127.5, 255 -> histogram bins 0, 63, 127, 255 one count each ->
    # cum 1,2,3,4 from the first non-zero entry: levels (cum-1)/3*255 = 0, 85, 170, 255.  The table is read at the
    # ROUNDED gray: 63.75 -> bin 64 (cum 2 -> 85), 127.5 -> bin 128 (half-to-even; cum 3 -> 170)
    im = np.array([[0.0, 0.25], [0.5, 1.0]], dtype=np.float32)
    eq, hist_orig, hist_eq = U.histogram_equalize(im)
    assert np.array_equal(np.round(eq * 255), [[0, 85], [170, 255]])
    assert hist_orig.sum() == 4 and hist_orig[[0, 63, 127, 255]].tolist() == [1, 1, 1, 1]
    assert hist_eq[[0, 85, 170, 255]].tolist() == [1, 1, 1, 1]
    assert np.array_equal(im, np.array([[0.0, 0.25], [0.5, 1.0]], dtype=np.float32))        # input untouched
    zero = np.zeros((3, 3), dtype=np.float32)
    out, h0, h1 = U.histogram_equalize(zero)
    assert np.array_equal(out, zero) and h0 is None and h1 is None                            # the reference's early return
    rgb = np.random.RandomState(0).rand(12, 12, 3)
    out = U.histogram_equalize(rgb)[0]
    assert out.shape == rgb.shape
    y = U.rgb2yiq(out)[..., 0]
    assert abs(y.min()) < 1e-9 and abs(y.max() - 1) < 1e-9                                     # Y is equalised over [0,1]
    assert np.allclose(U.rgb2yiq(out)[..., 1:], U.rgb2yiq(rgb)[..., 1:], atol=1e-9)           # chroma untouched
    assert np.allclose(U.yiq2rgb(U.rgb2yiq(rgb)), rgb, atol=1e-12)


def test_histogram_equalize_frames_equals_the_numpy_restatement():
    """The torch version (what runs on the GPU inside render_frames) gives the same uint8 levels as the reference's
    algorithm in NumPy, frame by frame; all-zero and constant frames give zeros."""
    U = importlib.import_module("nerf-and-dietnerf_b200").UtilsCV
    rng = np.random.RandomState(1)
    depth = (rng.rand(6, 50, 50) ** 3 * 2 + 0.5).astype(np.float32)
    depth[3] = np.round(depth[3] * 4) / 4            # few distinct values: sparse histogram
    depth[4] = 0
    depth[5] = 0.37
    levels = U.histogram_equalize_frames(torch.from_numpy(depth))
    assert levels.dtype == torch.uint8 and levels.shape == depth.shape
    for i in range(4):
        ref = np.uint8(np.round(U.histogram_equalize(depth[i])[0] * 255))
        assert np.array_equal(levels[i].numpy(), ref), i
        assert levels[i].min() == 0 and levels[i].max() == 255
    assert int(levels[4].max()) == 0 and int(levels[5].max()) == 0


# ---- video files ------------------------------------------------------------------------------------------------------
def test_save_frames_as_video_round_trip(tmp_path):
    V = importlib.import_module("nerf-and-dietnerf_b200").UtilsVideo
    yy, xx = np.mgrid[0:48, 0:64]
    frames = [np.stack([(xx + 4 * k) / 100.0 % 1.0 * 0 + xx / 63.0, yy / 47.0, np.full_like(xx, k / 6.0, dtype=float)], -1)
              for k in range(6)]
    path = tmp_path / "new_dir" / "clip.avi"                         # the directory is created
    V.save_frames_as_video(path, frames, 30)
    got, fps = V.read_video_frames(path)
    assert got.shape == (6, 48, 64, 3) and fps == 30
    for k in range(6):
        assert psnr_u8(got[k], V.frame_to_uint8(frames[k])) > 35      # MJPG is lossy; channel order survives
    # uint8 frames are written as they are; gray frames become three equal channels
    u8 = [V.frame_to_uint8(f) for f in frames]
    V.save_frames_as_video(tmp_path / "u8.avi", u8, 30)
    assert np.array_equal(V.read_video_frames(tmp_path / "u8.avi")[0], got)
    gray = [np.uint8(yy * 5) for _ in range(3)]
    V.save_frames_as_video(tmp_path / "gray.avi", gray, 5)
    g, fps = V.read_video_frames(tmp_path / "gray.avi")
    assert g.shape == (3, 48, 64, 3) and fps == 5 and psnr_u8(g[0][..., 0], gray[0]) > 35
    assert np.abs(g[..., 0].astype(int) - g[..., 2].astype(int)).max() <= 2
    with pytest.raises(AssertionError):
        V.save_frames_as_video(tmp_path / "empty.avi", [], 5)
    assert np.array_equal(V.frame_to_uint8(np.array([0.0, 0.498, 0.5, 1.0])), [0, 127, 128, 255])


def test_trajectories_have_the_recorded_frame_counts(run_and_pkg, videos):
    """300 / 720 / 1320 frames at 60 fps, as in the reference's own recorded videos of this run."""
    run, pkg = run_and_pkg
    point, spherical = run._point_of_interest()
    assert spherical and np.linalg.norm(point - np.array([-0.013, 0.038, -0.697])) < 0.05
    l_to_r = run.get_l_to_r_c2w_matrices_to_render()
    sphere = run.get_sphere_c2w_matrices_to_render()
    path = run.get_path_c2w_matrices_to_render()
    assert len(l_to_r) == int(videos["render_l_to_r_rgb_video__n_frames"]) == 300
    assert len(sphere) == int(videos["render_rgb_sphere_video__n_frames"]) == 720
    assert len(path) == int(videos["render_rgb_path_video__n_frames"]) == 1320
    assert float(videos["render_l_to_r_rgb_video__fps"]) == run.video_properties["fps_render_video"] == 60
    test_c2w = run.camera_poses[19]
    # left to right: the test pose's rotation, x sweeps from +1 to -1 about its position (t - x)
    assert np.abs(l_to_r[:, :3, :3] - test_c2w[:3, :3]).max() == 0
    assert np.allclose(l_to_r[:, 0, 3], test_c2w[0, 3] - np.linspace(-1, 1, 300), atol=1e-6)
    assert np.allclose(l_to_r[:, 1:3, 3], test_c2w[1:3, 3])
    # sphere: first pose carries the test pose's rotation, every camera is at distance 1 from the point of interest
    assert np.abs(sphere[0, :3, :3] - test_c2w[:3, :3]).max() < 1e-6
    assert np.allclose(np.linalg.norm(sphere[:, :3, 3] - point, axis=1), 1.0, atol=1e-6)
    # path: 11 segments of 120 frames, each from one listed view to the next, the last back to the first
    for k, idx in enumerate(PATH_INDICES):
        nxt = PATH_INDICES[(k + 1) % len(PATH_INDICES)]
        assert np.abs(path[120 * k] - run.camera_poses[idx]).max() < 1e-5
        assert np.abs(path[120 * k + 119] - run.camera_poses[nxt]).max() < 1e-5


def test_dataset_video_reproduces_the_recorded_one(tmp_path, videos):
    pkg, run = alexander_run(save_location=tmp_path, tasks={"save_dataset_video": True})
    run.start()                                                     # the switchboard: only the dataset video is on
    path = tmp_path / "video_save" / "train_set_video.avi"
    assert os.path.exists(path) and sorted(os.listdir(tmp_path / "video_save")) == ["train_set_video.avi"]
    frames, fps = pkg.UtilsVideo.read_video_frames(path)
    assert len(frames) == int(videos["train_set_video__n_frames"]) == 70 and fps == float(videos["train_set_video__fps"]) == 5
    for j, i in enumerate(videos["train_set_video__indices"]):
        assert psnr_u8(frames[i], videos["train_set_video__frames"][j]) > 40, i       # same images, same encoder settings
    # frame 19 is dataset image 20: the test image (idx 19) is held out
    data = np.load(os.path.join(GOLD, "alexander50_dataset.npz"))
    assert psnr_u8(frames[19], data["images_u8"][20]) > psnr_u8(frames[19], data["images_u8"][19]) + 5


def test_oracle_frames_match_the_recorded_videos(run_and_pkg, videos):
    """The reference's weights rendered by the oracle at this framework's trajectory poses against frames of the videos
    the reference rendered itself.  Measured: 36-43 dB rgb, 36-41 dB equalised depth on the left-to-right and path videos
    (MJPG loss + jitter); the sphere video depends on the reference's unseeded RANSAC point: 34-38 dB at the poses
    checked here."""
    run, pkg = run_and_pkg
    pin = np.load(os.path.join(GOLD, "alexander50_pin.npz"))
    pc, pf = torch.from_numpy(pin["params_coarse"]), torch.from_numpy(pin["params_fine"])
    trajectories = {"render_l_to_r_rgb_video": run.get_l_to_r_c2w_matrices_to_render(),
                    "render_rgb_path_video": run.get_path_c2w_matrices_to_render(),
                    "render_rgb_sphere_video": run.get_sphere_c2w_matrices_to_render()}
    depth_of = {"render_l_to_r_rgb_video": "render_depths_l_to_r_video", "render_rgb_path_video": "render_depths_path_video"}
    for name, frame, floor_rgb in (("render_l_to_r_rgb_video", 75, 36.0), ("render_rgb_path_video", 119, 33.0),
                                   ("render_rgb_sphere_video", 360, 32.0)):
        c2w = np.asarray(trajectories[name][frame], dtype=np.float32)
        rgb, weights, _, _, _, z = O.render_image(pc, pf, O.NetCfg(), float(pin["near"]), float(pin["far"]), c2w,
                                                  float(pin["fov"]), 50, 50, 4096, 64, 128, seed=3, step=0)
        u8 = pkg.UtilsVideo.frame_to_uint8(rgb.numpy().reshape(50, 50, 3).clip(0, 1))
        j = list(videos[name + "__indices"]).index(frame)
        got = psnr_u8(u8, videos[name + "__frames"][j])
        print(f"{name}[{frame}]: rgb {got:.2f} dB")
        assert got > floor_rgb, (name, frame, got)
        if name in depth_of:
            depth, _ = O.depth_and_acc(weights, z)
            levels = pkg.UtilsCV.histogram_equalize_frames(depth.reshape(1, 50, 50))[0].numpy()
            k = list(videos[depth_of[name] + "__indices"]).index(frame)
            got_d = psnr_u8(levels, videos[depth_of[name] + "__frames"][k])
            print(f"{depth_of[name]}[{frame}]: depth {got_d:.2f} dB")
            assert got_d > 30.0, (name, frame, got_d)


# ---- the CUDA path against the recorded videos ---------------------------------------------------------------------------
@pytest.mark.gpu
def test_rendered_videos_match_the_reference_recordings(pkg, tmp_path, videos):
    """ExecutionRun.start() with the reference's epoch-95 checkpoint in place and the left-to-right video task on: the
    two files it writes have the recorded length and their frames agree with the frames the reference's TensorFlow
    path rendered (MJPG loss + unseeded jitter: 36-43 dB measured with the fp32 oracle).  The sphere and path videos
    are checked on the golden frames through render_frames (the same loop without the encoder)."""
    pin = np.load(os.path.join(GOLD, "alexander50_pin.npz"))
    _, run = alexander_run(save_location=tmp_path, tasks={"render_and_save_test_left_to_right_video": True})
    seed_model = pkg.NeRFModel(run.net_config, run.render_config, run.near_boundary, run.far_boundary, mode="bf16")
    seed_model.model_coarse.set_params(pin["params_coarse"])
    seed_model.model_fine.set_params(pin["params_fine"])
    ckpt = pkg.NeRFModel.get_nerf_model_path(tmp_path, 95)
    os.makedirs(ckpt.parent, exist_ok=True)
    seed_model.save_weights(ckpt)                    # where get_nerf looks for starting_epoch_number = 95
    run.start()
    assert sorted(os.listdir(tmp_path / "video_save")) == ["render_depths_l_to_r_video.avi", "render_l_to_r_rgb_video.avi"]
    rgb, fps = pkg.UtilsVideo.read_video_frames(tmp_path / "video_save" / "render_l_to_r_rgb_video.avi")
    depth, fps_d = pkg.UtilsVideo.read_video_frames(tmp_path / "video_save" / "render_depths_l_to_r_video.avi")
    assert rgb.shape == (300, 50, 50, 3) and depth.shape == (300, 50, 50, 3) and fps == fps_d == 60
    for j, i in enumerate(videos["render_l_to_r_rgb_video__indices"]):
        p_rgb = psnr_u8(rgb[i], videos["render_l_to_r_rgb_video__frames"][j])
        p_depth = psnr_u8(depth[i][..., 1], videos["render_depths_l_to_r_video__frames"][j])
        print(f"l_to_r frame {i}: rgb {p_rgb:.2f} dB, depth {p_depth:.2f} dB")
        assert p_rgb > 33.0 and p_depth > 31.0, (i, p_rgb, p_depth)

    model = run.get_nerf()
    assert torch.equal(model.model_coarse.params.cpu(), torch.from_numpy(pin["params_coarse"]))
    # (rgb floor, depth floor) per golden frame; the quarter turns of the orbit look at the scene from outside the
    # training views and sit on the reference's unseeded RANSAC centre: 26-27 dB rgb, 21-22 dB depth with the oracle
    checks = {"sphere": ("render_rgb_sphere_video", "render_depths_sphere_video", run.get_sphere_c2w_matrices_to_render(),
                         lambda i: (22.0, 18.0) if i in (90, 270) else (30.0, 23.0)),
              "path": ("render_rgb_path_video", "render_depths_path_video", run.get_path_c2w_matrices_to_render(),
                       lambda i: (33.0, 26.0))}
    for tag, (name_rgb, name_depth, poses, floors) in checks.items():
        idx = [int(i) for i in videos[name_rgb + "__indices"]]
        idx_d = [int(i) for i in videos[name_depth + "__indices"]]
        rgbs, levels = run.render_frames(model, np.asarray(poses, dtype=np.float32)[idx], equalize_depth=True)
        assert rgbs.dtype == np.uint8 and levels.dtype == np.uint8 and levels.shape == (len(idx), 50, 50)
        for j, i in enumerate(idx):
            p_rgb = psnr_u8(rgbs[j], videos[name_rgb + "__frames"][j])
            line = f"{tag} frame {i}: rgb {p_rgb:.2f} dB"
            assert p_rgb > floors(i)[0], (tag, i, p_rgb)
            if i in idx_d:
                p_depth = psnr_u8(levels[j], videos[name_depth + "__frames"][idx_d.index(i)])
                line += f", depth {p_depth:.2f} dB"
                assert p_depth > floors(i)[1], (tag, i, p_depth)
            print(line)
    # float depth is still what render_frames returns by default, and its equalisation on the host gives the same levels
    _, depth_f = run.render_frames(model, np.asarray(checks["path"][2], dtype=np.float32)[:1])
    assert depth_f.dtype == np.float32 and depth_f.shape == (1, 50, 50)
    host = np.uint8(np.round(pkg.UtilsCV.histogram_equalize(depth_f[0])[0] * 255))
    dev = pkg.UtilsCV.histogram_equalize_frames(torch.from_numpy(depth_f).cuda())[0].cpu().numpy()
    assert np.array_equal(host, dev)


def test_trajectories_on_a_forward_facing_and_a_blender_scene():
    """The two other branches of the trajectory builders (src/ExecutionRun.py:372-376, :401-411): cameras that do not
    look at a common point (forward-facing capture) get the left-to-right dolly in the frame of their average pose and
    the bare unit-sphere orbit; a Blender scene gets the orbit scaled and pushed to the cameras' distance."""
    pkg = importlib.import_module("nerf-and-dietnerf_b200")
    rng = np.random.RandomState(0)
    poses = np.tile(np.eye(4, dtype=np.float32), (12, 1, 1))
    poses[:, :2, 3] = rng.uniform(-0.5, 0.5, (12, 2))              # a plane of parallel cameras: optical axes never meet
    config = json.loads(json.dumps(RUN_CONFIG))
    config["training"]["test_img_idx"] = 3
    config["video"]["img_indices_for_path_video"] = [0, 5]
    images = np.zeros((12, 8, 8, 3), dtype=np.float32)
    run = pkg.ExecutionRun.from_arrays(config, images, poses, 0.6, 0.5, 2.5)
    run.dataset_type = "colmap"
    point, spherical = run._point_of_interest()
    assert not spherical
    l_to_r = run.get_l_to_r_c2w_matrices_to_render()
    average = pkg.UtilsFiles.change_mats_to_homogeneous(pkg.UtilsFiles.poses_avg(poses)[..., :4][None])[0]
    assert l_to_r.shape == (300, 4, 4)
    assert np.allclose(l_to_r[:, :3, :3], average[:3, :3], atol=1e-6)
    assert np.allclose(l_to_r[:, :3, 3] - average[:3, 3], np.linspace(-1, 1, 300)[:, None] * average[:3, 0], atol=1e-5)
    sphere = run.get_sphere_c2w_matrices_to_render()
    assert np.array_equal(sphere, pkg.poses.get_sphere_matrices(360))                   # untouched for a COLMAP scene
    path = run.get_path_c2w_matrices_to_render()
    assert path.shape == (240, 4, 4) and np.abs(path[0] - poses[0]).max() < 1e-6 and np.abs(path[119] - poses[5]).max() < 1e-6
    assert np.abs(path[120] - poses[5]).max() < 1e-6 and np.abs(path[239] - poses[0]).max() < 1e-6
    # Blender: the loader hands over the average pose before recentring and the spherify scale
    run_b = pkg.ExecutionRun(config=config, data=(images, poses, 0.6, 0.5, 2.5, np.diag([1.0, 1.0, 1.0, 1.0]) +
                                                   np.array([[0, 0, 0, 0], [0, 0, 0, 0], [0, 0, 0, 4.0], [0, 0, 0, 0]]), 0.25))
    run_b.dataset_type = "blender"
    run_b._poi = (None, False)
    sphere_b = run_b.get_sphere_c2w_matrices_to_render()
    unit = pkg.poses.get_sphere_matrices(360)
    assert np.allclose(sphere_b[:, :3, :3], unit[:, :3, :3])
    assert np.allclose(sphere_b[:, :3, 3], unit[:, :3, 3] * 1.0 + np.array([0, 0, -1.0]), atol=1e-6)   # 0.25 * 4 = 1
    # a run without a `video:` block or a save directory says so instead of failing somewhere inside
    bare = pkg.ExecutionRun.from_arrays({k: v for k, v in config.items() if k != "video"}, images, poses, 0.6, 0.5, 2.5)
    with pytest.raises(KeyError):
        bare.get_l_to_r_c2w_matrices_to_render()
    with pytest.raises(Exception, match="save"):
        run.render_video(poses[:1], "x", "a.avi", "b.avi", model=object())


def test_pose_helper_properties():
    """Property checks over random rotations (hypothesis): source->dest rotation, vector-to-vector rotation, slerp
    end points / orthonormality, quaternion <-> matrix round trip (all four branches of the conversion)."""
    from hypothesis import given, settings, strategies as st
    P = importlib.import_module("nerf-and-dietnerf_b200").poses
    angle = st.floats(min_value=-360.0, max_value=360.0, allow_nan=False)

    @settings(max_examples=60, deadline=None)
    @given(angle, angle, angle, angle, angle, angle, st.floats(min_value=0.0, max_value=1.0))
    def check(ax, ay, az, bx, by, bz, t):
        a, b = P.get_sphere_matrix(1.0, ax, ay, az), P.get_sphere_matrix(0.7, bx, by, bz)
        ra, rb = a[:3, :3], b[:3, :3]
        assert np.abs(ra @ ra.T - np.eye(3)).max() < 1e-12 and abs(np.linalg.det(ra) - 1) < 1e-12
        q = P.quaternion_from_rotation_matrix(ra)
        assert np.abs(P.rotation_matrix_from_quaternion(q / np.linalg.norm(q)) - ra).max() < 1e-9
        r = P.get_rotation_matrix_from_source_to_dest_mats(ra, rb)
        assert np.abs(r[:3, :3] @ ra - rb).max() < 1e-9
        m = P.interpolation_type_slerp_for_c2w(a, b, t)
        assert np.abs(m[:3, :3] @ m[:3, :3].T - np.eye(3)).max() < 1e-5
        assert np.allclose(m[:3, 3], a[:3, 3] * (1 - t) + b[:3, 3] * t, atol=1e-6)
        v1, v2 = ra[:, 0] * 2.5, rb[:, 2] * 0.3
        rot = P.get_rotation_matrix_from_v1_to_v2(v1, v2)
        assert np.abs(rot @ (v1 / np.linalg.norm(v1)) - v2 / np.linalg.norm(v2)).max() < 1e-6
    check()
    for m0, m1 in ((P.get_sphere_matrix(1, 10, 20, 30), P.get_sphere_matrix(1, 10, 20, 30)),):
        assert np.abs(P.interpolation_type_slerp_for_c2w(m0, m1, 0.4) - m0).max() < 1e-6     # identical poses: no 0/0


def test_histogram_equalize_frames_property():
    """Random small frames, including few-valued ones where the rounded gray sits on .5 ties: the torch version and
    the NumPy restatement always give the same levels."""
    from hypothesis import given, settings, strategies as st
    from hypothesis.extra import numpy as hnp
    U = importlib.import_module("nerf-and-dietnerf_b200").UtilsCV

    @settings(max_examples=80, deadline=None)
    @given(hnp.arrays(np.float32, hnp.array_shapes(min_dims=2, max_dims=2, min_side=1, max_side=9),
                      elements=st.one_of(st.floats(0.25, 3.0, width=32), st.sampled_from([0.5, 1.0, 1.5, 2.0, 2.5]))))
    def check(frame):
        levels = U.histogram_equalize_frames(torch.from_numpy(frame)[None])[0].numpy()
        eq = U.histogram_equalize(frame)[0]
        assert np.array_equal(levels, np.uint8(np.round(np.asarray(eq, dtype=np.float64) * 255)))
    check()


def test_task_switchboard_and_trajectories_for_every_reference_config(monkeypatch):
    """Every parseable YAML of the reference (tests/golden/config_census.json): ``start()`` runs exactly the tasks its
    ``tasks_to_perform`` switches on (of the ones in scope: training, three rendered videos, dataset video), in the
    reference's order, and the trajectory builders give fps*5, 2*fps*6 and len(indices)*fps*2 poses for its ``video:``
    block."""
    pkg = importlib.import_module("nerf-and-dietnerf_b200")
    census = json.load(open(os.path.join(GOLD, "config_census.json")))
    poses = np.asarray([pkg.poses.get_sphere_matrix(1.0, x, y, 0) for x in (-30, 0, 30) for y in range(0, 360, 15)],
                       dtype=np.float32)                                         # 72 views looking at the origin
    images = np.zeros((len(poses), 4, 4, 3), dtype=np.float32)
    order = ["_training", "render_l_to_r_test_video", "render_sphere_test_video_with_net_weights",
             "render_path_test_video_with_net_weights", "save_dataset_video"]
    keys = ["start_training", "render_and_save_test_left_to_right_video", "render_and_save_test_sphere_video",
            "render_and_save_test_path_video", "save_dataset_video"]
    checked = 0
    for name, cfg in sorted(census.items()):
        if "unparseable" in cfg or not cfg.get("video"):
            continue
        run = pkg.ExecutionRun.from_arrays(dict(cfg), images, poses, 0.6, 0.5, 2.5)
        calls = []
        for method in order:
            monkeypatch.setattr(run, method, lambda *a, _m=method, **k: calls.append(_m))
        run.start()
        expected = [m for m, k in zip(order, keys) if (cfg.get("tasks_to_perform") or {}).get(k, False)]
        assert calls == expected, name
        fps = cfg["video"]["fps_render_video"]
        run._poi = (np.zeros(3), True)                 # the synthetic cameras look at the origin; skip the RANSAC
        assert len(run.get_l_to_r_c2w_matrices_to_render()) == fps * 5, name
        assert len(run.get_sphere_c2w_matrices_to_render()) == 2 * int(fps * 6), name
        indices = cfg["video"].get("img_indices_for_path_video")
        if indices and max(indices) < len(poses):
            assert len(run.get_path_c2w_matrices_to_render()) == len(indices) * int(fps * 2), name
        checked += 1
    assert checked >= 40


class _StubModel:
    """A 'model' whose render is a function of the pose only -- lets start() run its video tasks without a GPU."""
    device = torch.device("cpu")

    def render_image_lean(self, c2w, fov, h, w):
        k = float(np.abs(np.asarray(c2w)[:3, 3]).sum())
        yy, xx = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
        rgb = torch.stack([(xx / w + k) % 1.0, (yy / h + 0.5 * k) % 1.0, torch.full_like(xx, k % 1.0)], -1)
        return rgb.reshape(-1, 3), (1.0 + xx * 0.1 + yy * yy * 0.01 + k).reshape(-1), torch.ones(h * w)


def test_start_writes_the_reference_video_files(tmp_path, monkeypatch, videos):
    """All four video tasks through start(): the files under video_save/ carry the names of the reference's recorded
    run, the frame counts follow fps (2 here: 10 / 24 / 8 frames), rgb and depth videos have equal lengths."""
    pkg, run = alexander_run(save_location=tmp_path, tasks={
        "render_and_save_test_left_to_right_video": True, "render_and_save_test_sphere_video": True,
        "render_and_save_test_path_video": True, "save_dataset_video": True})
    run.video_properties = {"fps_render_video": 2, "fps_train_set_video": 5, "img_indices_for_path_video": [4, 7]}
    monkeypatch.setattr(run, "get_nerf", lambda: _StubModel())
    run.start()
    written = sorted(os.listdir(tmp_path / "video_save"))
    recorded = sorted({k.split("__")[0] + ".avi" for k in videos.files})
    assert written == recorded and len(written) == 7
    for stem, n in (("l_to_r", 10), ("sphere", 24), ("path", 8)):
        names = [w for w in written if stem in w]
        assert len(names) == 2
        for name in names:
            frames, fps = pkg.UtilsVideo.read_video_frames(tmp_path / "video_save" / name)
            assert frames.shape == (n, 50, 50, 3) and fps == 2, name
    depth, _ = pkg.UtilsVideo.read_video_frames(tmp_path / "video_save" / "render_depths_sphere_video.avi")
    assert depth.min() < 16 and depth.max() > 239                      # equalised: the levels span the range
