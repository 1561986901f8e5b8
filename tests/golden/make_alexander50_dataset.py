"""Builds tests/golden/alexander50_dataset.npz: the reference's own 50 px Alexander scene (71 COLMAP views) after the
loader's recentre / spherify step, so that the GPU box (which has no /root/reference) can repeat the training run the
reference recorded under Results/50px_alexander_71pics_sphere_nerf_save_dir_4 (tools/train_alexander50.py,
tests/test_parity_gpu.py::test_training_run_tracks_the_reference_psnr_curve).

Run in the build container:  python tests/golden/make_alexander50_dataset.py
Images are stored as uint8 (the JPEG code values); poses / fov / near / far come from the product's loader
(nerf-and-dietnerf_b200/UtilsFiles.py) and are cross-checked here against the oracle's restatement of the same loader.
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.llff_loader import load_colmap      # noqa: E402

SRC = "/root/reference/Assets/AlexanderColmap/50px_71pics"


def main():
    files = importlib.import_module("nerf-and-dietnerf_b200.UtilsFiles")
    images, c2w, fov, near, far, _, scale = files.get_data_from_colmap(SRC)
    o_images, o_c2w, o_fov, o_near, o_far, o_scale = load_colmap(SRC)
    assert np.abs(c2w - o_c2w).max() < 1e-6 and abs(fov - o_fov) < 1e-9 and abs(near - o_near) < 1e-9
    assert abs(far - o_far) < 1e-9 and np.abs(images - o_images).max() <= 2.0 / 255      # decoders differ by <= 1-2 codes
    u8 = np.round(images * 255.0).astype(np.uint8)
    assert np.array_equal(u8.astype(np.float32) / 255.0, images)
    np.savez_compressed(os.path.join(ROOT, "tests/golden/alexander50_dataset.npz"), images_u8=u8, c2w=c2w,
                        fov=np.float64(fov), near=np.float64(near), far=np.float64(far), scale=np.float64(scale))
    print(f"{u8.shape[0]} images {u8.shape[1:]} fov {fov:.5f} near {near:.5f} far {far:.5f} scale {scale:.5f}")


if __name__ == "__main__":
    main()
