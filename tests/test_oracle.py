"""CPU tests of the oracle itself: known-answer vectors of the shared RNG, and the semantics SURVEY Appendix A lists
(each checked on a hand-computable case), so that a parity failure on the GPU can be attributed to the kernel."""
import math

import numpy as np
import pytest
import torch

from oracle import nerf_oracle as O
from oracle import philox


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = philox.philox4x32_10(*ctr, *key)
        assert tuple(int(x) for x in got) == want


def test_uniform_stream_layout():
    u = philox.uniform(seed=(5 << 32) | 7, stream_id=1, step=3, n_rays=4, n_draws=10, ray_offset=100)
    assert u.shape == (4, 10) and u.dtype == np.float32 and (u >= 0).all() and (u < 1).all()
    w = philox.philox4x32_10(np.uint32(102), np.uint32(1), np.uint32(1), np.uint32(3), 7, 5)   # ray 2, draws 4..7
    assert np.array_equal(u[2, 4:8], philox.bits_to_uniform(np.array([int(x) for x in w], dtype=np.uint32)))
    assert np.array_equal(philox.uniform(9, 0, 0, 8, 64)[5], philox.uniform(9, 0, 0, 1, 64, ray_offset=5)[0])


def test_ray_directions_semantics():
    """A.1: pixel centres, one tan(fov/2) for both axes, not normalised, w = 0."""
    d = O.get_rays_directions(2, 2, math.pi / 2, np.eye(4))
    assert torch.allclose(d[0, 0], torch.tensor([-0.5, 0.5, -1.0, 0.0]), atol=1e-6)
    assert torch.allclose(d[1, 1], torch.tensor([0.5, -0.5, -1.0, 0.0]), atol=1e-6)
    c2w = np.eye(4)
    c2w[:3, 3] = [1, 2, 3]
    o, dd = O.rays_for_image(c2w, 0.5, 3, 5)
    assert o.shape == (15, 4) and torch.equal(o[7], torch.tensor([1.0, 2.0, 3.0, 1.0]))


def test_stratified_z_semantics():
    """A.2: linspace endpoints inclusive, bins (far-near)/S wide, jitter always on, last sample overshoots far."""
    z = O.get_z_values(2.0, 6.0, 1, 5, torch.zeros(1, 5))
    assert torch.equal(z[0], torch.tensor([2.0, 3.0, 4.0, 5.0, 6.0]))
    z1 = O.get_z_values(2.0, 6.0, 1, 5, torch.full((1, 5), 0.5))
    assert torch.allclose(z1[0] - z[0], torch.full((5,), 0.4))
    assert z1[0, -1] > 6.0


def test_posenc_layout_and_pi():
    """A.4: coordinate-major [c, s0, c0, ...] with the factor pi; views have no identity term."""
    x = torch.tensor([[0.5, 0.25, -1.0]])
    e = O.positional_encoding_for_xyz(x, 2)
    assert e.shape == (1, 15)
    assert torch.allclose(e[0, :5], torch.tensor([0.5, 1.0, 0.0, 0.0, -1.0]), atol=1e-6)   # x, sin(pi/2), cos(pi/2), sin(pi), cos(pi)
    assert abs(e[0, 5].item() - 0.25) < 1e-7 and abs(e[0, 6].item() - math.sin(math.pi / 4)) < 1e-6
    v = O.positional_encoding_for_views(x, 4)
    assert v.shape == (1, 24) and abs(v[0, 0].item() - 1.0) < 1e-6 and abs(v[0, 8].item() - math.sin(math.pi / 4)) < 1e-6
    assert torch.equal(O.positional_encoding_for_xyz(x, 0), x)


def test_mlp_structure():
    """A.5: 514 332 parameters, skip concat puts xyz first, sigma head sees the view encoding."""
    cfg = O.NetCfg()
    assert cfg.n_params == 514332 and O.NetCfg(5, 4, 0).shapes[8] == (256, 256)
    p = O.glorot_params(cfg.shapes, 0)
    xyz, view = torch.randn(4, 33), torch.randn(4, 24)
    base = O.mlp_forward(p, cfg.shapes, xyz, view)
    assert base.shape == (4, 4)
    assert not torch.allclose(O.mlp_forward(p, cfg.shapes, xyz, view + 1)[:, 3], base[:, 3])   # sigma depends on view
    # zeroing rows 0..32 of the skip layer's kernel removes the xyz path of the concat
    layers = O.unflatten(p.clone(), cfg.shapes)


def test_ray_marching_closed_form():
    """A.7: sigma = relu, c = sigmoid, last delta 1e9, exclusive cumprod, no white background."""
    raw = torch.tensor([[[0.0, 0.0, 0.0, 1.0], [10.0, -10.0, 0.0, 2.0], [0.0, 0.0, 0.0, -3.0]]])
    z = torch.tensor([[1.0, 1.5, 2.5]])
    rgb, w, T, a, c = O.ray_marching(raw, z)
    a0, a1 = 1 - math.exp(-0.5), 1 - math.exp(-2.0)
    assert torch.allclose(a[0], torch.tensor([a0, a1, 0.0]), atol=1e-6)              # relu(-3) = 0 -> alpha 0 even with delta 1e9
    assert torch.allclose(T[0], torch.tensor([1.0, 1 - a0, (1 - a0) * (1 - a1)]), atol=1e-6)
    assert torch.allclose(w[0], a[0] * T[0])
    assert abs(rgb[0, 0].item() - (w[0, 0] * 0.5 + w[0, 1] * torch.sigmoid(torch.tensor(10.0))).item()) < 1e-6
    raw[0, 2, 3] = 0.5                                                               # any positive sigma at the last sample is opaque
    assert abs(O.ray_marching(raw, z)[3][0, 2].item() - 1.0) < 1e-7


def test_cumprod_gradient_is_div_no_nan():
    """A.8: TF's cumprod gradient returns 0 where an input factor is exactly 0 (saturated alpha)."""
    x = torch.tensor([[0.5, 0.0, 0.25]], requires_grad=True)
    out = O._ExclusiveCumprodTF.apply(x)
    assert torch.equal(out.detach(), torch.tensor([[1.0, 0.5, 0.0]]))
    out.backward(torch.ones_like(out))
    assert torch.equal(x.grad, torch.tensor([[1.0, 0.0, 0.0]]))                      # (0.5 + 0)/0.5, div_no_nan -> 0, nothing after


def test_importance_sampling_semantics():
    """A.6: searchsorted left on a cdf without leading 0, mid-point bins, 1e-5 denominator floor, sorted output."""
    w = torch.tensor([[0.0, 1.0, 1.0, 0.0]])
    z = torch.tensor([[0.0, 1.0, 2.0, 3.0]])
    u = torch.tensor([[0.75, 0.25, 0.5, 0.0, 0.999999]])
    zs, idx, perm, unsorted = O.get_z_vals_from_prob_dist_func(w, z, 5, u, return_aux=True)
    assert idx.tolist() == [[2, 1, 1, 0, 2]]          # cdf = [0, .5, 1, 1]; #(cdf < u); u = 0.5 is NOT > cdf[1]
    # idx 2: b=1,t=2, lo=.5, hi=1, mids (1.5, 2.5): z = 1.5 + (u-.5)/.5
    assert abs(unsorted[0, 0].item() - 2.0) < 1e-6
    # idx 1: b=0,t=1, lo=0, hi=.5, mids (0.5, 1.5)
    assert abs(unsorted[0, 1].item() - 1.0) < 1e-6 and abs(unsorted[0, 2].item() - 1.5) < 1e-6
    assert abs(unsorted[0, 3].item() - 0.5) < 1e-6   # idx 0 -> both ends clamp to bin 0: mid[0]
    assert torch.equal(zs, torch.sort(unsorted, -1).values)
    # empty ray: every sample collapses onto the last mid-point
    ze = O.get_z_vals_from_prob_dist_func(torch.zeros(1, 4), z, 3, torch.tensor([[0.1, 0.5, 0.9]]))
    assert torch.equal(ze, torch.full((1, 3), 2.5))


def test_sampler_gradient_reaches_weights():
    """The reference does not stop the gradient at z_from_dist (src/NeRF.py:155)."""
    w = torch.rand(3, 8).requires_grad_(True)
    z = torch.sort(torch.rand(3, 8), -1).values
    O.get_z_vals_from_prob_dist_func(w, z, 16, torch.rand(3, 16)).sum().backward()
    assert w.grad.abs().sum() > 0


def test_losses_and_adam():
    """A.9: NeRF loss = MSE_c + MSE_f; DietNeRF's aliasing gives 2*MSE_c + MSE_f; Keras Adam first step = -lr*sign(g)."""
    cfg = O.NetCfg()
    pc, pf = O.glorot_params(cfg.shapes, 1, 0.1), O.glorot_params(cfg.shapes, 2, 0.1)
    o, d = O.rays_for_image(np.eye(4), 0.6, 4, 4)
    y = torch.rand(16, 3)
    jit, u = torch.rand(16, 8), torch.rand(16, 8)
    a = O.train_losses(pc, pf, cfg, 0.5, 2.5, o, d, y, 8, 8, jit, u)
    b = O.train_losses(pc, pf, cfg, 0.5, 2.5, o, d, y, 8, 8, jit, u, dietnerf=True)
    assert torch.allclose(a["loss"], a["mse_c"] + a["mse_f"])
    assert torch.allclose(b["loss"], 2 * b["mse_c"] + b["mse_f"]) and torch.allclose(b["loss_for_rays"], a["loss"])
    assert a["z_f"].shape == (16, 8)                         # the fine net sees ONLY the new samples in training
    p, m, v = O.adam_step(torch.zeros(3), torch.tensor([1e-2, -3.0, 0.0]), torch.zeros(3), torch.zeros(3), 1, 1e-3)
    assert torch.allclose(p, torch.tensor([-1e-3, 1e-3, 0.0]), atol=1e-6)
    assert abs(float(O.get_psnr(torch.tensor(0.01))) - 20.0) < 1e-5


def test_render_uses_all_192_samples():
    cfg = O.NetCfg()
    pc, pf = O.glorot_params(cfg.shapes, 1), O.glorot_params(cfg.shapes, 2)
    o, d = O.rays_for_image(np.eye(4), 0.6, 2, 2)
    out = O.render(pc, pf, cfg, 0.5, 2.5, o, d, 64, 128, torch.rand(4, 64), torch.rand(4, 128))
    assert out[5].shape == (4, 192) and torch.all(out[5][:, 1:] >= out[5][:, :-1])
