"""DietNeRF train-step shell: drop-in for ``_rgb_render_loss`` / ``train_step`` / ``_create_metrics`` of the
reference's src/DietNeRF.py:120-202 on the same sm_100a kernels as NeRF.

In scope here: the ray loss and its metrics.  The semantic-consistency term (in-tape 150x150
render -> ViT-B/32 embedding -> cosine loss, src/DietNeRF.py:204-283) is a "next" row: the hook
``consistency_loss_fn`` lets a caller supply it (it receives the model and must return (loss_value, flat_grads or
None)); without it the term is 0, which is what the reference computes on 12 of every 13 steps and on the last 5 % of
training (src/DietNeRF.py:224-236).
"""
from typing import Dict

from .NeRF import NeRF


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
                 rot_mat_to_in_front_of_point_of_interest=None, *, consistency_loss_fn=None, **kwargs):
        super().__init__(net_config, render_config, near_boundary, far_boundary, **kwargs)
        if self.model_fine is None:
            self.COARSE_LOSS_WEIGHT = 1.0   # without a fine net the aliasing leaves loss = MSE_c
        self.target_images = target_images
        self.target_camera_poses = target_camera_poses
        self.field_of_view = field_of_view
        self.max_steps_of_consistency_loss = max_steps_of_consistency_loss
        self.point_of_interest_in_scene = estimated_intersection
        self.rot_mat_to_in_front_of_point_of_interest = rot_mat_to_in_front_of_point_of_interest
        self.counter = 0
        self.consistency_loss_fn = consistency_loss_fn

    def should_use_consistency_loss(self) -> bool:
        """src/DietNeRF.py:224-236: every 13th step while counter < 0.95 * total steps."""
        if self.consistency_loss_fn is None:
            return False
        in_range = self.max_steps_of_consistency_loss < 0 or \
            self.counter < self.PERCENTAGE_OF_TRAIN_STEPS_WITH_CONSISTENCY_LOSS * self.max_steps_of_consistency_loss
        return in_range and self.counter % self.K_INTERVAL_SIZE_FOR_CONSISTENCY_LOSS == 0

    def train_step(self, data) -> Dict:
        self.counter += 1
        cosine_similarity_loss = 0.0
        extra = None
        if self.should_use_consistency_loss():
            cosine_similarity_loss, extra = self.consistency_loss_fn(self)
        self._extra_grads = extra
        metrics = super().train_step(data)
        return self._create_metrics(metrics, cosine_similarity_loss)

    def apply_gradients(self, g):
        if getattr(self, "_extra_grads", None) is not None:
            g[:self._extra_grads.numel()] += self._extra_grads
        super().apply_gradients(g)

    def _metrics(self, sums, n_total):
        m = super()._metrics(sums, n_total)
        mse_c = sums[0] / (3.0 * n_total)
        m["loss_for_rays"] = mse_c + (sums[1] / (3.0 * n_total) if self.model_fine is not None else 0.0)
        return m

    @staticmethod
    def _create_metrics(metrics, cosine_similarity_loss):
        """src/DietNeRF.py:174-190: the cosine term is added to the loss once in train_step (:147) and once more
        in the reported metric (:188)."""
        metrics = dict(metrics)
        metrics["cosine_similarity_loss"] = cosine_similarity_loss
        metrics["loss"] = metrics["loss"] + 2.0 * cosine_similarity_loss
        return metrics
