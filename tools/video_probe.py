#!/usr/bin/env python
"""Video render throughput (BASELINE configs[3], config_files/256px_ficus_72pics_sphere.yaml): the frame loop of
render_video (src/ExecutionRun.py:315-356) for the sphere trajectory (UtilsCV.get_sphere_matrices), 256x256 frames,
64 + 192 samples per ray, 16384-ray batches, random-init weights.  Frames are dealt round-robin to the GPUs.

    python tools/video_probe.py [--frames 48]                  (one GPU)
    torchrun --nproc-per-node N tools/video_probe.py           (N GPUs)
"""
import argparse
import importlib
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=48)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--encode", action="store_true", help="also time render_video: depth equalisation on the GPU + both "
                    "MJPG files written by rank 0 (src/ExecutionRun.py:315-356 end to end)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    pkg = importlib.import_module("nerf-and-dietnerf_b200")
    config = {"neural_net": {"type_of_model": "NeRF", "hidden_layer_dim": 256, "last_hidden_layer_dim": 128,
                             "leaky_relu_alpha": 0.05, "n_pos_enc_dim_xyz": 5, "n_pos_enc_view_dir": 4,
                             "n_angles_for_model": 2, "n_rays_in_batch_train": 2048, "n_rays_in_batch_render": 16384},
              "render": {"n_render_samples_coarse": 64, "n_render_samples_fine": 128},
              "training": {"n_epochs": 1, "optimizer_lr": 5e-4, "test_img_idx": 0, "idx_train_img_to_plot": 1}}
    images = np.zeros((2, args.size, args.size, 3), dtype=np.float32)
    run = pkg.ExecutionRun.from_arrays(config, images, np.stack([np.eye(4, dtype=np.float32)] * 2), 0.69111, 0.06285,
                                       2.19989, mode="bf16", seed=0)
    model = run.get_nerf()
    poses = pkg.poses.get_sphere_matrices(args.frames // 2)
    run.render_frames(model, poses[:2 * world])                      # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rgbs, depths = run.render_frames(model, poses)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    assert rgbs.shape == (len(poses), args.size, args.size, 3) and np.isfinite(depths).all()
    if world > 1:
        # every rank holds the same gathered video
        import torch.distributed as dist
        chk = torch.tensor([float(rgbs.astype(np.int64).sum())], device="cuda", dtype=torch.float64)
        both = [torch.empty_like(chk) for _ in range(world)]
        dist.all_gather(both, chk)
        assert all(torch.equal(both[0], b) for b in both)
    if rank == 0:
        rays = len(poses) * args.size * args.size
        print(f"{len(poses)} frames {args.size}x{args.size} on {world} GPU(s): {dt:.3f} s = {len(poses) / dt:.1f} frames/s = "
              f"{rays / dt / 1e6:.2f} M rays/s (incl. uint8 conversion, gather and the copy of the video to the host)")
    if args.encode:
        import tempfile
        with tempfile.TemporaryDirectory() as tmp:
            from pathlib import Path
            run.save_location = Path(tmp)
            run.video_properties = {"fps_render_video": 60}
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            where_rgb, where_depth = run.render_video(poses, "video probe", "rgb.avi", "depth.avi", model=model)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if rank == 0:
                t1 = time.perf_counter()
                rgbs, levels = run.render_frames(model, poses, equalize_depth=True) if world == 1 else (None, None)
                t_frames = time.perf_counter() - t1
                print(f"render_video ({len(poses)} frames, equalised depth, 2 MJPG files of {os.path.getsize(where_rgb)} + "
                      f"{os.path.getsize(where_depth)} bytes): {dt:.3f} s = {len(poses) / dt:.1f} frames/s"
                      + (f"; frames alone {t_frames:.3f} s, so the OpenCV encoder takes {dt - t_frames:.3f} s"
                         if world == 1 else ""))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
