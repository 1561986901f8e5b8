"""Builds tests/golden/alexander50_pin.npz -- the fixture that pins the oracle against an OUTPUT OF THE REFERENCE.

Run in the build container, where /root/reference exists (the GPU box never needs it):
    python tests/golden/make_golden.py

Inputs (read-only, from the reference):
  * Results/50px_alexander_71pics_sphere_nerf_save_dir_4/saved_weights/NeRF_model_epoch_095.h5   (trained by the reference)
  * Results/.../saved_test_train_psnrs/psnrs_train_test_095.npy   (PSNR the reference measured for those weights)
  * Assets/AlexanderColmap/50px_71pics/                           (poses_bounds.npy + images)
What it stores: the two flat parameter vectors (fp32, exact), pose / intrinsics / frustum of the test image (idx 19,
`test_img_idx` of the run's YAML) and of the plotted train image (idx 4), both ground-truth images, the reference's
recorded PSNRs, and the ORACLE's render of the test image for the fixed Philox stream (seed 95, step 0) plus its PSNR.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import nerf_oracle as O          # noqa: E402
from oracle.h5lite import load_keras_nerf_weights  # noqa: E402
from oracle.llff_loader import load_colmap   # noqa: E402

REF = "/root/reference"
RUN = os.path.join(REF, "Results/50px_alexander_71pics_sphere_nerf_save_dir_4")
TEST_IDX, TRAIN_PLOT_IDX = 19, 4             # test_img_idx / idx_train_img_to_plot of the run's YAML
SEED = 95


def main():
    pc, pf, shapes = load_keras_nerf_weights(os.path.join(RUN, "saved_weights/NeRF_model_epoch_095.h5"))
    images, c2w, fov, near, far, scale = load_colmap(os.path.join(REF, "Assets/AlexanderColmap/50px_71pics"))
    psnrs = np.load(os.path.join(RUN, "saved_test_train_psnrs/psnrs_train_test_095.npy"))
    cfg = O.NetCfg()
    assert pc.size == cfg.n_params and pf.size == cfg.n_params
    train_ids = [i for i in range(len(images)) if i != TEST_IDX]          # the test image is held out
    train_idx = train_ids[TRAIN_PLOT_IDX]
    h, w = images.shape[1:3]
    out = {}
    for tag, idx in (("test", TEST_IDX), ("train", train_idx)):
        rgb, weights, _, _, _, z = O.render_image(torch.from_numpy(pc), torch.from_numpy(pf), cfg, near, far, c2w[idx],
                                                  fov, h, w, 4096, 64, 128, seed=SEED, step=0)
        depth, acc = O.depth_and_acc(weights, z)
        psnr = float(O.get_psnr(O.mse(rgb, torch.from_numpy(images[idx]))))
        print(f"{tag} image {idx}: oracle PSNR {psnr:.3f} dB (reference recorded "
              f"{psnrs[0 if tag == 'test' else 1, -1]:.3f} dB at epoch 95)")
        out[f"{tag}_rgb_oracle"] = rgb.numpy()
        out[f"{tag}_depth_oracle"] = depth.numpy()
        out[f"{tag}_psnr_oracle"] = np.float32(psnr)
        out[f"{tag}_image"] = images[idx]
        out[f"{tag}_c2w"] = c2w[idx]
    np.savez_compressed(os.path.join(ROOT, "tests/golden/alexander50_pin.npz"), params_coarse=pc, params_fine=pf,
                        fov=np.float64(fov), near=np.float64(near), far=np.float64(far), scale=np.float64(scale),
                        seed=np.int64(SEED), psnr_reference_test=psnrs[0], psnr_reference_train=psnrs[1], **out)
    print("wrote tests/golden/alexander50_pin.npz")


if __name__ == "__main__":
    main()
