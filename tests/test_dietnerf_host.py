"""CPU checks of the host-side pieces of DietNeRF's consistency term: pose helpers
(src/UtilsCV.py:101-121, :175-247), the embedder's preprocessing and the cosine loss (src/DietNeRF.py:261-279)."""
import importlib
import math

import numpy as np
import torch

from oracle import nerf_oracle as O

poses = importlib.import_module("nerf-and-dietnerf_b200.poses")
vit = importlib.import_module("nerf-and-dietnerf_b200.vit")


def test_sphere_matrix_looks_at_origin():
    for radius, xr, yr in ((1.0, 0.0, 0.0), (0.8, -30.0, 45.0), (1.1, -90.0, -170.0)):
        m = poses.get_sphere_matrix(radius, xr, yr, 0.0)
        r, t = m[:3, :3], m[:3, 3]
        assert np.allclose(r @ r.T, np.eye(3), atol=1e-12) and abs(np.linalg.det(r) - 1) < 1e-12
        assert abs(np.linalg.norm(t) - radius) < 1e-12
        # the camera looks down its -z axis: that axis points from the camera to the origin
        assert np.allclose(-r[:, 2], -t / radius, atol=1e-12)
    # reference composition order: z_rot @ (y_rot @ (x_rot @ translate))
    m = poses.get_sphere_matrix(1.0, -90.0, 0.0, 0.0)
    assert np.allclose(m[:3, 3], [0.0, 1.0, 0.0], atol=1e-12)


def test_quaternion_round_trip_and_slerp():
    rng = np.random.default_rng(0)
    for _ in range(20):
        a, b, c = rng.uniform(-180, 180, 3)
        r = (poses.get_z_rot_mat(c) @ poses.get_y_rot_mat(b) @ poses.get_x_rot_mat(a))[:3, :3]
        q = poses.quaternion_from_rotation_matrix(r)
        assert abs(np.linalg.norm(q) - 1) < 1e-12
        assert np.allclose(poses.rotation_matrix_from_quaternion(q), r, atol=1e-12)
    c1 = poses.get_sphere_matrix(1.0, -20.0, 10.0, 0.0)
    c2 = poses.get_sphere_matrix(0.8, -60.0, 100.0, 0.0)
    assert np.allclose(poses.interpolation_type_slerp_for_c2w(c1, c2, 0.0), c1, atol=1e-6)
    assert np.allclose(poses.interpolation_type_slerp_for_c2w(c1, c2, 1.0), c2, atol=1e-6)
    mid = poses.interpolation_type_slerp_for_c2w(c1, c2, 0.5)
    assert np.allclose(mid[:3, 3], 0.5 * (c1[:3, 3] + c2[:3, 3]), atol=1e-6)          # translations are lerped
    # slerp: the halfway rotation is equally far (geodesic angle) from both ends
    ang = lambda ra, rb: math.acos(np.clip((np.trace(ra.T @ rb) - 1) / 2, -1, 1))
    assert abs(ang(mid[:3, :3], c1[:3, :3]) - ang(mid[:3, :3], c2[:3, :3])) < 1e-5
    assert abs(ang(mid[:3, :3], c1[:3, :3]) - 0.5 * ang(c1[:3, :3], c2[:3, :3])) < 1e-5
    lst = poses.interpolation_type_slerp_for_c2w(c1, c2, np.linspace(0, 1, 4))
    assert isinstance(lst, list) and len(lst) == 4


def test_preprocess_and_cosine_loss_match_the_oracle():
    img = torch.rand(2, 150, 150, 3, generator=torch.Generator().manual_seed(0))
    a, b = vit.embedder_preprocess(img), O.embedder_preprocess(img)
    assert a.shape == (2, 3, 224, 224) and torch.equal(a, b)
    assert a.min().item() >= -1.0 and a.max().item() <= 1.0
    s, t = torch.randn(768), torch.randn(768)
    assert abs(vit.consistency_loss(s, t).item() - O.consistency_loss(s, t).item()) < 1e-7
    assert vit.consistency_loss(s, s).item() < 1e-6                      # aligned embeddings: loss 0
    assert abs(vit.consistency_loss(s, -s).item() - 1.0) < 1e-6          # opposite: loss 1


def test_vit_b32_shape_frozen_and_differentiable_in_the_image():
    m = vit.ViTB32(layers=1, seed=1).eval()
    assert all(not p.requires_grad for p in m.parameters())
    x = torch.rand(1, 3, 224, 224, requires_grad=True)
    e = m(x)
    assert e.shape == (1, 768)
    e.square().sum().backward()
    assert x.grad is not None and torch.isfinite(x.grad).all() and x.grad.abs().sum().item() > 0
    full = vit.ViTB32(seed=0)
    n = sum(p.numel() for p in full.parameters())
    assert 87_000_000 < n < 89_000_000          # ViT-B/32 feature extractor: ~87.5 M parameters
    assert torch.equal(vit.ViTB32(layers=1, seed=1).pos, m.pos)           # seeded init is reproducible
