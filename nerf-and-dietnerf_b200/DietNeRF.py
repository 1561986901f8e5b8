"""DietNeRF: drop-in for src/DietNeRF.py on the same sm_100a kernels as NeRF.

``_rgb_render_loss`` / ``train_step`` / ``_create_metrics`` (src/DietNeRF.py:120-202) and the
semantic-consistency term (:204-283): every 13th step a 150x150 image is rendered from a random pose with
55 coarse + 55 importance samples per ray INSIDE the tape, embedded by a frozen ViT-B/32 and pulled towards the
embedding of a random training image with ``0.1 * (1 - cos) / 2``.  Here the render is the hand-written forward, the
embedder is a PyTorch module (``vit.py``; random-initialised, the TF-Hub weights are unreachable offline) that hands
back dL/d(image), and ``NeRF.render_image_backward`` pushes that gradient through the compositing / MLP / sampler
backward kernels into the same flat gradient buffer as the ray loss (one all-reduce per step also when sharded).
Without an embedder (``consistency_loss_fn=None`` and no ``target_images``) the term is 0, which is what the reference
computes on 12 of every 13 steps.
"""
from typing import Dict

import numpy as np
import torch

from .NeRF import NeRF
from .parallel import all_gather_rows, shard_bounds
from .poses import get_sphere_matrix, interpolation_type_slerp_for_c2w
from .vit import ViTB32, consistency_loss, embedder_preprocess


class DietNeRF(NeRF):
    K_INTERVAL_SIZE_FOR_CONSISTENCY_LOSS = 13
    CONSISTENCY_LOSS_WEIGHT = 0.1
    PERCENTAGE_OF_TRAIN_STEPS_WITH_CONSISTENCY_LOSS = 0.95
    IMG_SIZE_FOR_CS_LOSS = 150
    N_RENDER_SAMPLES_CS_LOSS = 55

    # src/DietNeRF.py:164-171: ``loss = loss_for_rays`` (MSE_c) ; ``loss_for_rays += MSE_f`` (new tensor) ;
    # ``loss += loss_for_rays``  ->  loss = 2*MSE_c + MSE_f when a fine network exists.
    COARSE_LOSS_WEIGHT = 2.0

    def __init__(self, net_config: Dict, render_config: Dict, near_boundary: float, far_boundary: float,
                 target_images=None, target_camera_poses=None, field_of_view=None,
                 max_steps_of_consistency_loss: int = -1, estimated_intersection=None,
                 rot_mat_to_in_front_of_point_of_interest=None, *, consistency_loss_fn=None, embedder=None,
                 resample_every_call: bool = False, numpy_seed=None, **kwargs):
        """Positional arguments as src/DietNeRF.py:41-52.  Keyword-only extensions:

        :param embedder:             module mapping (B,3,224,224) in [-1,1] to (B,D); default a random-initialised
                                     ViT-B/32 (``vit.ViTB32``), built lazily and shared like the reference's class attribute.
        :param consistency_loss_fn:  replaces the whole consistency term: ``fn(model) -> (loss, flat_grads or None)``.
        :param resample_every_call:  False (default) reproduces the reference under Keras ``fit``: its ``np.random``
                                     draws run at tf.function TRACE time, so ONE source pose and ONE target image are
                                     used for the whole run.  True draws a new pose/target every time,
                                     as the DietNeRF paper intends.
        """
        super().__init__(net_config, render_config, near_boundary, far_boundary, **kwargs)
        if self.model_fine is None:
            self.COARSE_LOSS_WEIGHT = 1.0   # without a fine net the aliasing leaves loss = MSE_c
        self.net_config = net_config
        self.render_config = render_config
        self.camera_poses = target_camera_poses
        self.fov = field_of_view
        self.max_steps_of_consistency_loss = max_steps_of_consistency_loss
        self.point_of_interest_in_scene = estimated_intersection
        self.rot_mat_to_in_front_of_point_of_interest = rot_mat_to_in_front_of_point_of_interest
        self.is_spherical_dataset = self.point_of_interest_in_scene is not None
        self.counter = 0
        self._use_consistency_loss = True
        self.consistency_loss_fn = consistency_loss_fn
        self.resample_every_call = resample_every_call
        self._np_rng = np.random if numpy_seed is None else np.random.RandomState(numpy_seed)
        self._frozen_draw = None
        self.last_consistency_pose = None
        self.target_images_embedding = None
        self._embedder = embedder
        if target_images is not None:
            self.image_height, self.image_width = int(target_images[0].shape[0]), int(target_images[0].shape[1])
            if consistency_loss_fn is None:
                imgs = torch.as_tensor(np.asarray(target_images), dtype=torch.float32, device=self.device)
                with torch.no_grad():
                    self.target_images_embedding = torch.cat(
                        [self.embedder(embedder_preprocess(imgs[i:i + 16])) for i in range(0, imgs.shape[0], 16)])

    embedder_preprocess = staticmethod(embedder_preprocess)
    consistency_loss = staticmethod(consistency_loss)
    _shared_embedder = {}

    @property
    def embedder(self):
        """Lazy, shared per device (src/DietNeRF.py:38,71-79 keeps it as a class attribute)."""
        if self._embedder is None:
            key = str(self.device)
            if key not in DietNeRF._shared_embedder:
                DietNeRF._shared_embedder[key] = ViTB32().to(self.device).eval()
            self._embedder = DietNeRF._shared_embedder[key]
        return self._embedder

    def set_use_consistency_loss(self, should_use: bool):
        self._use_consistency_loss = bool(should_use)

    def is_use_consistency_loss(self) -> bool:
        return self._use_consistency_loss

    def should_use_consistency_loss(self) -> bool:
        """src/DietNeRF.py:224-236: every 13th step, while enabled and (max_steps <= 0 or counter < max_steps).  (The
        95 % factor is applied by the caller when it computes max_steps, src/ExecutionRun.py:246-247.)"""
        if self.consistency_loss_fn is None and self.target_images_embedding is None:
            return False
        in_range = self.max_steps_of_consistency_loss <= 0 or self.counter < self.max_steps_of_consistency_loss
        return in_range and self._use_consistency_loss and \
            self.counter % self.K_INTERVAL_SIZE_FOR_CONSISTENCY_LOSS == 0

    def sample_random_source_pose(self):
        """src/DietNeRF.py:238-259: a pose on a sphere around the scene's point of interest (spherical datasets), else
        a double slerp between three random training poses."""
        r = self._np_rng
        if self.is_spherical_dataset:
            radius = r.uniform(0.7, 1.1, 1)[0]
            x_rot = r.uniform(-90, 0, 1)[0]
            y_rot = r.uniform(-180, 180, 1)[0]
            c2w = np.asarray(self.rot_mat_to_in_front_of_point_of_interest) @ get_sphere_matrix(radius, x_rot, y_rot, 0)
            c2w[:3, 3] += np.asarray(self.point_of_interest_in_scene)
            return c2w.astype(np.float32)
        poses = np.asarray(self.camera_poses)
        choice = r.choice(len(poses), 3, replace=False)
        alphas = r.uniform(0, 1, 2)
        p1 = interpolation_type_slerp_for_c2w(poses[choice[0]], poses[choice[1]], alphas[0])
        return interpolation_type_slerp_for_c2w(p1, poses[choice[2]], alphas[1])

    def _draw_target_and_pose(self):
        if self._frozen_draw is not None and not self.resample_every_call:
            return self._frozen_draw
        rand_index = int(self._np_rng.randint(0, len(self.target_images_embedding), 1)[0])
        pose = np.asarray(self.sample_random_source_pose(), dtype=np.float32)
        if self.world_size > 1:
            # every rank must render the same view: rank 0's draw wins
            import torch.distributed as dist
            buf = torch.tensor([float(rand_index)] + pose.reshape(-1).tolist(), dtype=torch.float64, device=self.device)
            src = dist.get_global_rank(self._process_group, 0) if self._process_group is not None else 0
            dist.broadcast(buf, src=src, group=self._process_group)
            rand_index, pose = int(buf[0].item()), buf[1:].reshape(4, 4).to(torch.float32).cpu().numpy()
        draw = (rand_index, pose)
        self._frozen_draw = draw
        return draw

    def calc_consistency_loss(self, *, pose=None, target_index=None):
        """src/DietNeRF.py:204-222.  Renders the 150x150 source image (55 coarse + 55 importance samples, batches of
        ``n_rays_in_batch_train``), embeds it, and ADDS the gradient of ``0.1 * (1 - cos) / 2`` w.r.t. every network
        parameter to the flat gradient buffer.  Sharded runs render / back-propagate a contiguous block of the image's
        rays per rank and gather the image, so the sum of the ranks' gradient buffers is the full gradient.
        Returns the loss as a device scalar."""
        if pose is None or target_index is None:
            drawn_index, drawn_pose = self._draw_target_and_pose()
            target_index = drawn_index if target_index is None else target_index
            pose = drawn_pose if pose is None else pose
        self.last_consistency_pose = pose
        target = self.target_images_embedding[target_index]
        size, n_s = self.IMG_SIZE_FOR_CS_LOSS, self.N_RENDER_SAMPLES_CS_LOSS
        seed, step = self.seed ^ 0x5EED5EED, self.counter          # its own Philox stream, reproducible in backward
        lo, hi = shard_bounds(size * size, self.world_size, self.rank)
        rgb = self.render_image_lean(pose, self.fov, size, size, self.batch_size_train, n_s, n_s, seed=seed, step=step,
                                     ray_begin=lo, n_rays=hi - lo)[0]
        if self.world_size > 1:
            rgb = all_gather_rows(rgb, size * size, self._process_group)
        image = rgb.reshape(size, size, 3).detach().requires_grad_(True)
        with torch.enable_grad():
            emb = self.embedder(embedder_preprocess(image[None]))[0]
            loss = self.CONSISTENCY_LOSS_WEIGHT * consistency_loss(emb, target)
        d_image, = torch.autograd.grad(loss, image)
        self.render_image_backward(pose, self.fov, size, size, d_image, self.batch_size_train, n_s, n_s, seed=seed,
                                   step=step, ray_begin=lo, n_rays=hi - lo)
        return loss.detach()

    def train_step(self, data) -> Dict:
        return self._train_step_with_consistency(lambda: super(DietNeRF, self).train_step(data))

    def train_step_sharded(self, rays_orig, rays_dirs, real_rgb, n_total_rays, ray_offset=0) -> Dict:
        """``train_step`` for callers that already hold THIS rank's shard on the device (see NeRF.train_step_local)."""
        return self._train_step_with_consistency(
            lambda: self.train_step_local(rays_orig, rays_dirs, real_rgb, n_total_rays, ray_offset))

    def _train_step_with_consistency(self, ray_step) -> Dict:
        self.counter += 1
        cosine_similarity_loss = 0.0
        extra = None
        self._keep_grads = False
        if self.should_use_consistency_loss():
            if self.consistency_loss_fn is not None:
                cosine_similarity_loss, extra = self.consistency_loss_fn(self)
            else:
                # the consistency gradients go into the flat buffer first; the ray loss then adds to them
                self._grad_buffer().zero_()
                cosine_similarity_loss = self.calc_consistency_loss()
                self._keep_grads = True
        self._extra_grads = extra
        try:
            metrics = ray_step()
        finally:
            self._keep_grads = False
        return self._create_metrics(metrics, cosine_similarity_loss)

    def apply_gradients(self, g):
        """``g`` = [sq_err_c, sq_err_f, 0, 0 | d params_c | d params_f] (NeRF._grad_buffer).  The ``consistency_loss_fn``
        hook's gradient vector is laid out like the PARAMETERS ([coarse | fine]) and is added behind the four loss slots.
        It is applied after the all-reduce, so it must already be the global gradient (identical on every rank)."""
        extra = getattr(self, "_extra_grads", None)
        if extra is not None:
            n = g.numel() - 4
            if extra.numel() != n:
                raise ValueError(f"consistency_loss_fn returned {extra.numel()} gradient values, the networks have {n} parameters")
            g[4:] += extra.to(device=g.device, dtype=g.dtype).reshape(-1)
        super().apply_gradients(g)

    def train_step_fused(self, *args, **kwargs):
        if self.consistency_loss_fn is not None:
            raise RuntimeError("train_step_fused applies Adam inside the C call: the consistency_loss_fn hook's extra "
                               "gradients cannot be added; use train_step")
        return super().train_step_fused(*args, **kwargs)

    def _metrics_dict(self, out):
        m = super()._metrics_dict(out)
        m["loss_for_rays"] = out[3]                        # MSE_c + MSE_f (src/DietNeRF.py:163,168)
        return m

    @staticmethod
    def _create_metrics(metrics, cosine_similarity_loss):
        """src/DietNeRF.py:174-190: the cosine term is added to the loss once in train_step (:147) and once more
        in the reported metric (:188)."""
        metrics = dict(metrics)
        metrics["cosine_similarity_loss"] = cosine_similarity_loss
        metrics["loss"] = metrics["loss"] + 2.0 * cosine_similarity_loss
        return metrics
