"""Host train / render loop around the hot path: the slice of src/ExecutionRun.py that feeds it.

Kept: config -> dataset -> model (``get_nerf`` :216-232, ``_init_dietnerf`` :234-264) -> epoch loop (``_training``
:169-201: ``fit`` over ``steps_per_epoch = (n_images*h*w) // batch`` shuffled ray batches, then PSNR of the held-out test
image and of one training image, then a weight + PSNR checkpoint per epoch), ``render_video`` (:315-356: rgb + equalised depth = sum w z per
pose, two MJPG files under ``video_save/``) with its three trajectories (left-to-right :358-377, sphere :390-413, path
between dataset views :426-440), ``save_dataset_video`` (:442-448) and the part of ``start``'s switchboard (:117-152)
that drives them.  Not rebuilt: matplotlib plots and the plot video, GCS sync.

Differences from the reference, all in how the loop is driven, none in what a step computes: the ray table lives in HBM
(``RayDataset``) instead of a tf.data pipeline; metrics stay on the device and are read once per epoch; with
``torch.distributed`` initialised every rank runs the same loop on its shard of each batch (one gradient all-reduce
per step) and rank 0 writes the checkpoints.
"""
import os
import re
import shutil
import threading
import time
from pathlib import Path, PureWindowsPath

import numpy as np
import torch

from . import UtilsFiles
from .ConfigurationKeys import (BLENDER, COLMAP, DATASET_LOCATION, DATASET_TYPE, DIETNERF_MODEL, EXISTING_SAVE_DIR_NAME,
                                FAR_DEPTH_RENDER, GENERAL_SAVE_LOCATION, IDX_TRAIN_IMG_TO_PLOT, N_EPOCHS,
                                N_ANGLES_FOR_MODEL, N_RAYS_IN_BATCH_TRAIN, NEAR_DEPTH_RENDER, NEURAL_NET, OPTIMIZER_LR,
                                PICS_INDICES_TO_USE_IN_DATASET, RENDER, STARTING_EPOCH_NUMBER, START_TRAINING,
                                TASKS_TO_PERFORM, TEST_IMG_IDX, TRAINING, TYPE_OF_MODEL, VIDEO, FPS_RENDER_VIDEO,
                                FPS_TRAIN_SET_VIDEO, IMG_INDICES_FOR_PATH_VIDEO, RENDER_AND_SAVE_TEST_L_TO_R_VIDEO,
                                RENDER_AND_SAVE_TEST_SPHERE_VIDEO, RENDER_AND_SAVE_TEST_PATH_VIDEO, SAVE_DATASET_VIDEO)
from .DietNeRF import DietNeRF
from .NeRF import NeRF
from .network import DEFAULT_MODE
from .optimizers import Adam
from .poses import (estimate_point_of_interest_in_scene, get_c2w_matrices_between_2_c2w_with_stretch,
                    get_l_to_r_c2w_matrices, get_rotation_matrix_from_source_to_dest_mats, get_sphere_matrices)
from .UtilsCV import histogram_equalize_frames
from .UtilsNeuralRadianceField import get_num_of_batches, get_psnr_for_image, prepare_ds
from .UtilsVideo import save_frames_as_video

SAVE_DIR_NAME_FORMAT = '{}_save_dir_{}'
DIR_SAVE_VIDEOS = 'video_save'
# file names of src/UtilsPlots.py:19-25
FILENAME_TRAIN_SET_VIDEO = 'train_set_video.avi'
FILENAME_RENDER_DEPTHS_L_TO_R_VIDEO = 'render_depths_l_to_r_video.avi'
FILENAME_RENDER_L_TO_R_RGB_VIDEO = 'render_l_to_r_rgb_video.avi'
FILENAME_RENDER_DEPTHS_SPHERE_VIDEO = 'render_depths_sphere_video.avi'
FILENAME_RENDER_DEPTHS_PATH_VIDEO = 'render_depths_path_video.avi'
FILENAME_RENDER_RGB_SPHERE_VIDEO = 'render_rgb_sphere_video.avi'
FILENAME_RENDER_RGB_PATH_VIDEO = 'render_rgb_path_video.avi'


def get_save_location(path_to_config_file, config) -> Path:
    """src/UtilsFiles.py:232-281: the named existing directory, or ``<general>/<config stem>_save_dir_<next index>``."""
    general = Path(config[GENERAL_SAVE_LOCATION])
    existing = config[EXISTING_SAVE_DIR_NAME]
    if existing:
        if not (general / existing).exists():
            raise Exception('Save location', general / existing, 'does not exists.')
        return general / existing
    stem = Path(path_to_config_file).stem
    os.makedirs(general, exist_ok=True)
    taken = [int(re.findall(r'\d+', n)[-1]) for n in os.listdir(general) if re.match(rf'^{re.escape(stem)}_save_dir_\d+$', n)]
    new_dir = general / SAVE_DIR_NAME_FORMAT.format(stem, max(taken) + 1 if taken else 0)
    os.makedirs(new_dir)
    return new_dir


class ExecutionRun:
    """``ExecutionRun(path_to_config_file).start()`` as in main.py:26-27; ``from_arrays`` skips the file system."""

    def __init__(self, path_to_config_file=None, *, config=None, data=None, save_location=None, mode=None, seed=0):
        if config is None:
            config = UtilsFiles.load_config(path_to_config_file)
        self.config = config
        self.mode, self.seed = (DEFAULT_MODE if mode is None else mode), seed
        self.tasks_to_perform = config.get(TASKS_TO_PERFORM, {START_TRAINING: True})
        self.dataset_type = config.get(DATASET_TYPE)
        self.pics_indices_to_use_in_dataset = config.get(PICS_INDICES_TO_USE_IN_DATASET)
        self.net_config, self.render_config, self.training_config = config[NEURAL_NET], config[RENDER], config[TRAINING]
        self.video_properties = config.get(VIDEO) or {}
        data = self.get_data(config) if data is None else data
        (self.images, self.camera_poses, self.field_of_view, self.near_boundary, self.far_boundary,
         self.average_c2w_before_recenter, self.c2w_scale_parameter) = data
        distributed = torch.distributed.is_available() and torch.distributed.is_initialized()
        self.is_main = not distributed or torch.distributed.get_rank() == 0
        if save_location is None and path_to_config_file is not None:
            # ONE save directory per run: rank 0 resolves / creates it (and keeps the config copy, src/ExecutionRun.py:
            # 83-86), every other rank receives the path -- the ranks would otherwise race on `<stem>_save_dir_<n>` and
            # ranks > 0 would look for checkpoints in directories of their own
            if self.is_main:
                save_location = get_save_location(path_to_config_file, config)
                shutil.copyfile(path_to_config_file, Path(save_location) / Path(path_to_config_file).name)
            if distributed and torch.distributed.get_world_size() > 1:
                box = [str(save_location) if self.is_main else None]
                torch.distributed.broadcast_object_list(box, src=0)      # also the barrier before any rank reads it
                save_location = box[0]
        self.save_location = Path(save_location) if save_location is not None else None
        start = config.get(STARTING_EPOCH_NUMBER, -1)
        self._epoch_number = start if start and start > 0 else 0
        self.model = None                    # set by _training: the videos after it render with THESE weights
        self.history = []                    # one dict per epoch: epoch, seconds, psnr_test, psnr_train, loss

    @classmethod
    def from_arrays(cls, config, images, camera_poses, field_of_view, near_boundary, far_boundary, **kw):
        return cls(config=config, data=(np.asarray(images, dtype=np.float32), np.asarray(camera_poses, dtype=np.float32),
                                        float(field_of_view), float(near_boundary), float(far_boundary), None, 1.0), **kw)

    def get_data(self, config):
        location = Path(PureWindowsPath(config[DATASET_LOCATION]))        # the YAMLs use Windows separators
        if self.dataset_type == BLENDER:
            return UtilsFiles.get_data_from_blender(location, config[RENDER][NEAR_DEPTH_RENDER],
                                                    config[RENDER][FAR_DEPTH_RENDER])
        if self.dataset_type == COLMAP:
            return UtilsFiles.get_data_from_colmap(location)
        raise Exception(f"unknown dataset_type {self.dataset_type!r}")

    # ---- model ---------------------------------------------------------------------------------------------------------------
    def get_train_images_indices(self, idx_test):
        keep = set(self.pics_indices_to_use_in_dataset) if self.pics_indices_to_use_in_dataset else None
        return [n for n in range(len(self.images)) if n != idx_test and (keep is None or n in keep)]

    def _get_train_data_from_loaded_dataset(self):
        idx_test = self.training_config[TEST_IMG_IDX]
        idx = self.get_train_images_indices(idx_test)
        return idx_test, self.images[idx], self.camera_poses[idx]

    def effective_mode(self) -> str:
        """The arithmetic mode this run's networks get: the requested one, or "fp32" when the network has no tensor-core
        plan (``n_angles_for_model: 0``, the xyz-only network of src/NeRF.py:248-288 -- 5 of the reference's 47 configs):
        those run on the SIMT fp32 kernels instead of failing."""
        if self.mode == "fp32":
            return "fp32"
        from . import _lib
        from .NeRF import net_cfg_from_dict
        cfg = net_cfg_from_dict(self.net_config)
        if int(_lib.load().nerf_packed_bytes(_lib.ctypes.byref(cfg))) < 0:
            if self.is_main and not getattr(self, "_warned_mode", False):
                print(f"ExecutionRun: mode={self.mode!r} has no tensor-core plan for this network "
                      f"(n_angles_for_model={self.net_config[N_ANGLES_FOR_MODEL]}); running it in mode='fp32'")
                self._warned_mode = True
            return "fp32"
        return self.mode

    def get_nerf(self) -> NeRF:
        """A new model + Adam(lr); loads ``saved_weights/NeRF_model_epoch_<starting epoch>.h5`` when it exists."""
        kw = dict(mode=self.effective_mode(), seed=self.seed, stop_grad_z=getattr(self, "stop_grad_z", False))
        if self.net_config[TYPE_OF_MODEL] == DIETNERF_MODEL:
            model = self._init_dietnerf(kw)
        else:
            model = NeRF(self.net_config, self.render_config, self.near_boundary, self.far_boundary, **kw)
        model.compile(optimizer=Adam(self.training_config[OPTIMIZER_LR]))
        if torch.distributed.is_available() and torch.distributed.is_initialized() and \
                torch.distributed.get_world_size() > 1:
            model.distribute()
        self._loaded_checkpoint = None
        if self.save_location is not None:
            path = NeRF.get_nerf_model_path(self.save_location, self._epoch_number)
            if os.path.exists(path):
                model.load_weights(path)
                self._loaded_checkpoint = path
        return model

    def _init_dietnerf(self, kw):
        h, w = self.images[0].shape[0], self.images[0].shape[1]
        _, train_images, train_poses = self._get_train_data_from_loaded_dataset()
        n_batches = get_num_of_batches(self.net_config[N_RAYS_IN_BATCH_TRAIN], len(train_images), h, w)
        n_steps = n_batches * (self.training_config[N_EPOCHS] - self._epoch_number)
        n_steps *= DietNeRF.PERCENTAGE_OF_TRAIN_STEPS_WITH_CONSISTENCY_LOSS
        point, spherical = self._point_of_interest()
        rot = None
        if spherical:
            rot = np.eye(4)
            rot[:3, :3] = self.camera_poses[self.training_config[TEST_IMG_IDX]][:3, :3]
        return DietNeRF(self.net_config, self.render_config, self.near_boundary, self.far_boundary, train_images,
                        train_poses, self.field_of_view, int(n_steps), point if spherical else None, rot, **kw)

    # ---- training ------------------------------------------------------------------------------------------------------------
    def start(self):
        """The tasks of ``tasks_to_perform`` in the reference's order (:117-152); a task missing from the YAML is off."""
        if self.tasks_to_perform.get(START_TRAINING, False):
            self._training()
            self._epoch_number = self.training_config[N_EPOCHS]
        if self.tasks_to_perform.get(RENDER_AND_SAVE_TEST_L_TO_R_VIDEO, False):
            self.render_l_to_r_test_video()
        if self.tasks_to_perform.get(RENDER_AND_SAVE_TEST_SPHERE_VIDEO, False):
            self.render_sphere_test_video_with_net_weights()
        if self.tasks_to_perform.get(RENDER_AND_SAVE_TEST_PATH_VIDEO, False):
            self.render_path_test_video_with_net_weights()
        if self.tasks_to_perform.get(SAVE_DATASET_VIDEO, False):
            self.save_dataset_video()

    def fit(self, model, ds, steps_per_epoch):
        """One Keras ``fit`` epoch: ``steps_per_epoch`` batches of a fresh shuffle; returns the mean metrics.  The
        per-step metric scalars stay on the device and are reduced once at the end of the epoch."""
        seen = []
        for batch in ds:
            if len(seen) >= steps_per_epoch:
                break
            seen.append(model.train_step(batch))
        if not seen:
            return {}
        out = {}
        for k in seen[0]:
            vals = [m[k] for m in seen]
            tensors = [v.detach().reshape(()).float() for v in vals if isinstance(v, torch.Tensor)]
            total = float(torch.stack(tensors).sum()) if tensors else 0.0
            total += float(sum(v for v in vals if not isinstance(v, torch.Tensor)))
            out[k] = total / len(vals)
        return out

    def _epoch_psnrs(self, model, train_image, train_c2w, test_image, test_c2w):
        out = []
        for img, c2w in ((test_image, test_c2w), (train_image, train_c2w)):
            h, w = img.shape[0], img.shape[1]
            render = model.render_image_lean(c2w, self.field_of_view, h, w)[0].reshape(h, w, 3)
            out.append(float(get_psnr_for_image(render, torch.as_tensor(img, device=render.device))))
        return out

    def _training(self, on_epoch_end=None):
        idx_test, train_images, train_poses = self._get_train_data_from_loaded_dataset()
        idx_plot = self.training_config[IDX_TRAIN_IMG_TO_PLOT]
        psnrs_test, psnrs_train = ([], []) if self.save_location is None else UtilsFiles.get_psnr_values(
            UtilsFiles.get_psnr_save_path(self.save_location, self._epoch_number))
        h, w = self.images[0].shape[0], self.images[0].shape[1]
        batch = self.net_config[N_RAYS_IN_BATCH_TRAIN]
        ds = prepare_ds(batch, train_poses, train_images, self.field_of_view, seed=self.seed)
        n_batches = get_num_of_batches(batch, len(train_poses), h, w)
        model = self.get_nerf()
        self.model = model
        for epoch_number in range(self._epoch_number + 1, self.training_config[N_EPOCHS] + 1):
            torch.cuda.synchronize()
            t0 = time.time()
            metrics = self.fit(model, ds, n_batches)
            torch.cuda.synchronize()
            seconds = time.time() - t0
            p_test, p_train = self._epoch_psnrs(model, self.images[idx_plot], self.camera_poses[idx_plot],
                                                self.images[idx_test], self.camera_poses[idx_test])
            psnrs_test.append(p_test)
            psnrs_train.append(p_train)
            self.history.append({"epoch": epoch_number, "seconds": seconds, "psnr_test": p_test, "psnr_train": p_train,
                                 **metrics})
            if self.save_location is not None and self.is_main:
                # the files are written by a helper thread while the next epoch trains (the GPU does not sit idle
                # behind a few tens of ms of file formatting); the previous write is joined first
                self._join_checkpoint()
                snapshot = model.weights_snapshot()
                w_path = NeRF.get_nerf_model_path(self.save_location, epoch_number)
                p_path = UtilsFiles.get_psnr_save_path(self.save_location, epoch_number)
                hist = (list(psnrs_test), list(psnrs_train))

                def write(snapshot=snapshot, w_path=w_path, p_path=p_path, hist=hist):
                    os.makedirs(os.path.dirname(str(w_path)), exist_ok=True)
                    snapshot.write(w_path)
                    UtilsFiles.save_psnr_values(hist[0], hist[1], p_path)
                self._checkpoint_thread = threading.Thread(target=write)
                self._checkpoint_thread.start()
            if self.is_main:
                print(f"Done epoch {epoch_number} in {seconds:.2f} sec. ({seconds / max(n_batches, 1) * 1e3:.2f} ms / step) "
                      f"Test PSNR: {p_test:.3f}")
            if on_epoch_end is not None:
                on_epoch_end(self, model, epoch_number)
        self._join_checkpoint()
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            torch.distributed.barrier()              # the last checkpoint is complete before any rank can resume from it
        return model

    def _join_checkpoint(self):
        t = getattr(self, "_checkpoint_thread", None)
        if t is not None:
            t.join()
            self._checkpoint_thread = None

    # ---- rendering ---------------------------------------------------------------------------------------------------------
    def render_frames(self, model, c2w_matrices, h=None, w=None, equalize_depth=False):
        """The frame loop of ``render_video`` (:315-356) without the encoder: uint8 rgb (F,h,w,3) = round(255 rgb) and
        depth (F,h,w) = sum w z, on the host -- float32, or with ``equalize_depth`` the uint8 levels of the reference's
        ``histogram_equalize`` computed on the GPU.  With ``torch.distributed`` initialised the frames are dealt
        round-robin to the ranks (rendering needs no collective) and gathered once at the end; every rank returns all
        frames."""
        h = self.images[0].shape[0] if h is None else h
        w = self.images[0].shape[1] if w is None else w
        dist = torch.distributed
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        rank = dist.get_rank() if world > 1 else 0
        n_frames = len(c2w_matrices)
        mine = list(range(rank, n_frames, world))
        dev = model.device
        rgb_dev = torch.empty((len(mine), h, w, 3), dtype=torch.uint8, device=dev)
        depth_dev = torch.empty((len(mine), h, w), dtype=torch.float32, device=dev)
        for k, i in enumerate(mine):
            rgb, depth, _ = model.render_image_lean(c2w_matrices[i], self.field_of_view, h, w)
            rgb_dev[k] = (rgb.reshape(h, w, 3).clamp(0, 1) * 255).round().to(torch.uint8)
            depth_dev[k] = depth.reshape(h, w)
        if equalize_depth:
            depth_dev = histogram_equalize_frames(depth_dev)
        if world == 1:
            return rgb_dev.cpu().numpy(), depth_dev.cpu().numpy()
        per = (n_frames + world - 1) // world                      # pad every rank's block to the same length
        pad = lambda t: torch.cat([t, t.new_zeros((per - t.shape[0],) + tuple(t.shape[1:]))]) if t.shape[0] < per else t
        rgb_all = [torch.empty((per, h, w, 3), dtype=torch.uint8, device=dev) for _ in range(world)]
        depth_all = [torch.empty((per, h, w), dtype=depth_dev.dtype, device=dev) for _ in range(world)]
        dist.all_gather(rgb_all, pad(rgb_dev))
        dist.all_gather(depth_all, pad(depth_dev))
        rgbs = np.empty((n_frames, h, w, 3), dtype=np.uint8)
        depths = np.empty((n_frames, h, w), dtype=np.uint8 if equalize_depth else np.float32)
        for r in range(world):
            idx = list(range(r, n_frames, world))
            rgbs[idx] = rgb_all[r][:len(idx)].cpu().numpy()
            depths[idx] = depth_all[r][:len(idx)].cpu().numpy()
        return rgbs, depths

    def render_video(self, c2w_matrices, process_description, filename_rgb, filename_depths, loops=1, model=None):
        """Render one frame per pose with this run's weights and save ``video_save/<filename_rgb>`` and
        ``<filename_depths>`` (depth = sum w z, histogram-equalised per frame) at ``fps_render_video`` (:315-356).
        Returns the two paths; rank 0 writes the files."""
        if self.save_location is None:
            raise Exception('render_video needs a save location (the run was built without one)')
        if model is None:
            # right after training: the model that was just trained; otherwise a model loaded from this run's checkpoint
            # -- which must exist: rendering a video with freshly initialised weights helps nobody
            model = self.model
        if model is None:
            self._loaded_checkpoint = "unknown"
            model = self.get_nerf()
            if self._loaded_checkpoint is None:
                path = NeRF.get_nerf_model_path(self.save_location, self._epoch_number)
                raise Exception(f'render_video: no checkpoint {path} (set starting_epoch_number to a saved epoch, or train '
                                f'first); refusing to render with randomly initialised weights')
        fps = self.video_properties[FPS_RENDER_VIDEO]
        if self.is_main:
            print(process_description, f"({len(c2w_matrices)} frames)")
        rgbs, depths = self.render_frames(model, np.asarray(c2w_matrices, dtype=np.float32), equalize_depth=True)
        where_rgb = self.save_location / DIR_SAVE_VIDEOS / filename_rgb
        where_depths = self.save_location / DIR_SAVE_VIDEOS / filename_depths
        if self.is_main:
            save_frames_as_video(where_rgb, list(rgbs) * loops, fps)
            save_frames_as_video(where_depths, list(depths) * loops, fps)
        return where_rgb, where_depths

    def _point_of_interest(self):
        if getattr(self, "_poi", None) is None:
            self._poi = estimate_point_of_interest_in_scene(self.camera_poses, rng=np.random.RandomState(self.seed))
        return self._poi

    def get_l_to_r_c2w_matrices_to_render(self):
        """5 s of a left-to-right dolly (:358-377): on a spherical dataset around the test pose with its rotation,
        otherwise in the frame of the average pose."""
        matrices = get_l_to_r_c2w_matrices(self.video_properties[FPS_RENDER_VIDEO] * 5)
        _, is_spherical_dataset = self._point_of_interest()
        if is_spherical_dataset:
            test_c2w = self.camera_poses[self.training_config[TEST_IMG_IDX]]
            matrices[:, :3, 3] = test_c2w[:3, 3] - matrices[:, :3, 3]
            matrices[:, :3, :3] = test_c2w[:3, :3]
            return matrices
        average_c2w = UtilsFiles.change_mats_to_homogeneous(UtilsFiles.poses_avg(self.camera_poses)[..., :4][None])
        return average_c2w @ matrices

    def get_sphere_c2w_matrices_to_render(self):
        """6 s per orbit, two orbits (:390-413): the unit-sphere poses turned so the first one has the test pose's
        rotation and centred on the scene's point of interest (spherical dataset), or pushed out to the Blender
        cameras' distance."""
        matrices = get_sphere_matrices(int(self.video_properties[FPS_RENDER_VIDEO] * 6))
        point, is_spherical_dataset = self._point_of_interest()
        if is_spherical_dataset:
            rotation = get_rotation_matrix_from_source_to_dest_mats(
                matrices[0, :3, :3], self.camera_poses[self.training_config[TEST_IMG_IDX]][:3, :3])
            matrices = rotation @ matrices
            matrices[:, :3, 3] += point
        elif self.dataset_type == BLENDER:
            distance = self.c2w_scale_parameter * self.average_c2w_before_recenter[2, 3]
            matrices[:, :3, 3] *= distance
            matrices[:, :3, 3] += np.asarray([0, 0, -distance])
        return matrices

    def get_path_c2w_matrices_to_render(self):
        """2 s (slowing down) from each view of ``img_indices_for_path_video`` to the next and back to the first
        (:426-440)."""
        total_frames = int(self.video_properties[FPS_RENDER_VIDEO] * 2)
        c2ws = self.camera_poses[self.video_properties[IMG_INDICES_FOR_PATH_VIDEO]]
        matrices = []
        for c2w1, c2w2 in zip(c2ws, np.roll(c2ws, -1, axis=0)):
            matrices.extend(get_c2w_matrices_between_2_c2w_with_stretch(c2w1, c2w2, total_frames))
        return np.asarray(matrices)

    def render_l_to_r_test_video(self, model=None):
        return self.render_video(self.get_l_to_r_c2w_matrices_to_render(), 'Rendering images for l_to_r video',
                                 FILENAME_RENDER_L_TO_R_RGB_VIDEO, FILENAME_RENDER_DEPTHS_L_TO_R_VIDEO, model=model)

    def render_sphere_test_video_with_net_weights(self, model=None):
        return self.render_video(self.get_sphere_c2w_matrices_to_render(), 'Rendering images for sphere video',
                                 FILENAME_RENDER_RGB_SPHERE_VIDEO, FILENAME_RENDER_DEPTHS_SPHERE_VIDEO, model=model)

    def render_path_test_video_with_net_weights(self, model=None):
        return self.render_video(self.get_path_c2w_matrices_to_render(), 'Rendering images for path video',
                                 FILENAME_RENDER_RGB_PATH_VIDEO, FILENAME_RENDER_DEPTHS_PATH_VIDEO, model=model)

    def save_dataset_video(self):
        """The training images as a video at ``fps_train_set_video`` (:442-448)."""
        filename = self.save_location / DIR_SAVE_VIDEOS / FILENAME_TRAIN_SET_VIDEO
        _, train_images, _ = self._get_train_data_from_loaded_dataset()
        if self.is_main:
            save_frames_as_video(filename, train_images, self.video_properties[FPS_TRAIN_SET_VIDEO])
        return filename
