"""The shuffled training-ray stream: prepare_ds / RayDataset against the reference's tf.data pipeline
(src/UtilsNeuralRadianceField.py:135-178): every ray of every image exactly once per epoch, ray index y*w + x inside an
image, origin = the camera's translation, (n_img*h*w)//B full batches per epoch (get_num_of_batches, :237-247) plus the
ragged tail Dataset.batch would give, a different order every epoch, the SAME order on every rank for the same seed."""
import numpy as np
import pytest
import torch

from helpers import sphere_pose
from oracle import nerf_oracle as O


def _table(n):
    ids = torch.arange(n, dtype=torch.float32)
    origs = torch.stack([ids, ids + 0.25, ids + 0.5, torch.ones(n)], 1)
    dirs = torch.stack([-ids, ids * 2, ids * 3, torch.zeros(n)], 1)
    rgbs = torch.stack([ids, ids, ids], 1)
    return origs, dirs, rgbs


@pytest.mark.parametrize("n,batch", [(1000, 128), (4096, 4096), (77, 100), (2 * 50 * 50, 4096)])
def test_ray_dataset_is_a_permutation_per_epoch(n, batch):
    import importlib
    unrf = importlib.import_module("nerf-and-dietnerf_b200.UtilsNeuralRadianceField")
    origs, dirs, rgbs = _table(n)
    ds = unrf.RayDataset(batch, origs, dirs, rgbs, seed=3)
    assert len(ds) == -(-n // batch)
    orders = []
    for epoch in range(2):
        seen, sizes = [], []
        for o, d, c in ds:
            assert o.shape[1] == 4 and d.shape[1] == 4 and c.shape[1] == 3 and o.shape[0] == d.shape[0] == c.shape[0]
            # the three columns of a batch row belong to the SAME ray
            ids = o[:, 0]
            assert torch.equal(d[:, 1], ids * 2) and torch.equal(c[:, 0], ids) and torch.equal(o[:, 2], ids + 0.5)
            seen.append(ids)
            sizes.append(o.shape[0])
        # (n_img*h*w)//B full batches -- what Keras' fit consumes (steps_per_epoch) -- and one ragged tail
        assert sizes[:n // batch] == [batch] * (n // batch) and sum(sizes) == n and len(sizes) == len(ds)
        ids = torch.cat(seen)
        assert torch.equal(torch.sort(ids).values, torch.arange(n, dtype=torch.float32)), "a ray is missing or repeated"
        orders.append(ids)
    if n > 100:
        assert not torch.equal(orders[0], orders[1]), "the second epoch repeats the first epoch's order"
    # same seed -> same stream (every rank builds the same dataset and slices its shard of each batch)
    again = unrf.RayDataset(batch, origs, dirs, rgbs, seed=3)
    assert torch.equal(torch.cat([o[:, 0] for o, _, _ in again]), orders[0])
    other = unrf.RayDataset(batch, origs, dirs, rgbs, seed=4)
    if n > 100:
        assert not torch.equal(torch.cat([o[:, 0] for o, _, _ in other]), orders[0])


def test_get_num_of_batches_is_the_floor():
    import importlib
    unrf = importlib.import_module("nerf-and-dietnerf_b200.UtilsNeuralRadianceField")
    assert unrf.get_num_of_batches(4096, 70, 50, 50) == 42          # the recorded 50 px run: 42 steps per epoch
    assert unrf.get_num_of_batches(4096, 71, 256, 256) == 1136
    assert unrf.get_num_of_batches(2048, 72, 100, 100) == 351
    assert unrf.get_num_of_batches(100, 1, 3, 3) == 0


@pytest.mark.gpu
def test_prepare_ds_rays_are_the_images_pixels_in_row_major_order(pkg):
    """prepare_ds end to end on the GPU: un-shuffling one epoch by the rgb key gives, image by image, ray j = pixel
    (j // w, j % w): direction = the oracle's pinhole ray of that pixel, origin = c2w[:, 3], colour = img[y, x]."""
    h, w, fov = 12, 17, 0.46134
    rng = np.random.default_rng(0)
    c2ws = [sphere_pose(rng.uniform(0, 6.28), rng.uniform(-0.4, 0.4), 1.0) for _ in range(3)]
    # a unique colour per (image, pixel) so that a shuffled ray can be traced back
    imgs = []
    for i in range(3):
        k = torch.arange(h * w, dtype=torch.float32).reshape(h, w)
        imgs.append(torch.stack([k / (h * w), torch.full((h, w), i / 4.0), (k % 7) / 7.0], -1))
    ds = pkg.UtilsNeuralRadianceField.prepare_ds(64, c2ws, imgs, fov, seed=11)
    assert len(ds) == -(-3 * h * w // 64)
    o, d, c = (torch.cat(x).cpu() for x in zip(*[(a, b, e) for a, b, e in ds]))
    assert o.shape == (3 * h * w, 4)
    img_idx = torch.round(c[:, 1] * 4).long()
    pix = torch.round(c[:, 0] * (h * w)).long()
    key = img_idx * (h * w) + pix
    assert torch.equal(torch.sort(key).values, torch.arange(3 * h * w)), "not every pixel of every image exactly once"
    inv = torch.argsort(key)
    o, d, c = o[inv], d[inv], c[inv]
    for i in range(3):
        sl = slice(i * h * w, (i + 1) * h * w)
        ref_d = O.get_rays_directions(h, w, fov, c2ws[i]).reshape(-1, 4)        # ray index = y*w + x
        assert (d[sl] - ref_d).abs().max().item() < 1e-6
        assert torch.equal(o[sl], torch.tensor(c2ws[i][:, 3]).expand(h * w, 4))
        assert torch.equal(c[sl], imgs[i].reshape(-1, 3))
