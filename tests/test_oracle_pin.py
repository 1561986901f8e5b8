"""Pins the CPU oracle against outputs of the reference itself: the reference's trained weights
(NeRF_model_epoch_095.h5) rendered by the oracle must reproduce the PSNR the reference recorded for them
(psnrs_train_test_095.npy: 27.83 dB on the held-out test image).  Fixture: tests/golden/alexander50_pin.npz, built by
tests/golden/make_golden.py from /root/reference.  The jitter of the reference is unseeded (and it computes in
float16), so agreement is statistical: |delta PSNR| < 0.15 dB."""
import os

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import nerf_oracle as O

FIX = os.path.join(ROOT, "tests", "golden", "alexander50_pin.npz")


@pytest.fixture(scope="module")
def pin():
    return np.load(FIX)


def test_oracle_reproduces_reference_psnr(pin):
    cfg = O.NetCfg()
    pc, pf = torch.from_numpy(pin["params_coarse"]), torch.from_numpy(pin["params_fine"])
    h, w = pin["test_image"].shape[:2]
    rgb, weights, _, _, _, z = O.render_image(pc, pf, cfg, float(pin["near"]), float(pin["far"]), pin["test_c2w"],
                                              float(pin["fov"]), h, w, 4096, 64, 128, seed=int(pin["seed"]), step=0)
    psnr = float(O.get_psnr(O.mse(rgb, torch.from_numpy(pin["test_image"]))))
    ref = float(pin["psnr_reference_test"][-1])
    assert abs(ref - 27.834) < 1e-3                     # the value the reference saved at epoch 95
    assert abs(psnr - ref) < 0.15, (psnr, ref)
    # and the committed golden render is what this oracle produces (other BLAS builds may differ in the last bits)
    assert np.abs(rgb.numpy() - pin["test_rgb_oracle"]).max() < 5e-4
    assert abs(psnr - float(pin["test_psnr_oracle"])) < 0.01
    depth, acc = O.depth_and_acc(weights, z)
    assert np.abs(depth.numpy() - pin["test_depth_oracle"]).max() < 5e-3


def test_psnr_curve_fixture(pin):
    """BASELINE.md quality anchors: 16.70 / 24.61 / 26.85 / 27.83 dB at epochs 1 / 10 / 50 / 95."""
    p = pin["psnr_reference_test"]
    assert p.shape == (95,)
    for epoch, val in ((1, 16.70), (10, 24.61), (50, 26.85), (95, 27.83)):
        assert abs(p[epoch - 1] - val) < 0.01
    assert abs(pin["psnr_reference_train"][-1] - 32.46) < 0.01


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree only exists in the build container")
def test_fixture_matches_reference_files(pin):
    from oracle.h5lite import load_keras_nerf_weights
    from oracle.llff_loader import load_colmap
    run = "/root/reference/Results/50px_alexander_71pics_sphere_nerf_save_dir_4"
    pc, pf, shapes = load_keras_nerf_weights(os.path.join(run, "saved_weights/NeRF_model_epoch_095.h5"))
    assert np.array_equal(pc, pin["params_coarse"]) and np.array_equal(pf, pin["params_fine"])
    assert [shapes[i] for i in range(11)] == O.NetCfg().shapes           # Keras layer order == oracle layer order
    images, c2w, fov, near, far, scale = load_colmap("/root/reference/Assets/AlexanderColmap/50px_71pics")
    assert np.allclose(c2w[19], pin["test_c2w"]) and abs(fov - float(pin["fov"])) < 1e-12
    # fov 0.46134 rad, near 0.55759, far 2.56349, scale 0.18674 after the loader's recenter + spherify
    assert abs(fov - 0.46134) < 1e-5 and abs(near - 0.55759) < 1e-5 and abs(far - 2.56349) < 1e-5
    assert abs(scale - 0.18674) < 1e-5
