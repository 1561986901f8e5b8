"""Render throughput probe: one 256x256 frame (64 coarse + 192 fine samples per ray) through NeRFModel.render_image_lean
(the video path: rgb, depth, acc per ray), in bf16 and fp16 operand modes.  Usage: python tools/render_probe.py"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("nerf-and-dietnerf_b200")

MFLOP_PER_RAY = 256 * 2 * 512152 / 1e6      # 262.2 MFLOP/ray


def main():
    net = {"hidden_layer_dim": 256, "last_hidden_layer_dim": 128, "leaky_relu_alpha": 0.05, "n_pos_enc_dim_xyz": 5,
           "n_pos_enc_view_dir": 4, "n_angles_for_model": 2, "n_rays_in_batch_train": 2048,
           "n_rays_in_batch_render": 16384}
    rcfg = {"n_render_samples_coarse": 64, "n_render_samples_fine": 128}
    c2w = np.eye(4, dtype=np.float32)
    c2w[2, 3] = 1.0
    for mode in ("bf16", "fp16"):
        for batch in (16384, 65536):
            model = pkg.NeRFModel(net, rcfg, 0.3333, 2.0, mode=mode, seed=0)
            f = lambda: model.render_image_lean(c2w, 0.69111, 256, 256, batch_size_input=batch, seed=1, step=0)
            for _ in range(3):
                f()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10):
                f()
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 10
            rays = 256 * 256
            print(f"render 256x256 [{mode}] ray batch {batch:6d}: {ms:7.3f} ms/frame  {rays / ms / 1e3:7.3f} M rays/s  "
                  f"{rays * MFLOP_PER_RAY * 1e6 / (ms * 1e-3) / 1e12:7.1f} TFLOP/s (MLP algorithmic)")


if __name__ == "__main__":
    main()
