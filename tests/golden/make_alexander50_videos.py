"""Builds tests/golden/alexander50_videos.npz: frames DECODED FROM THE VIDEOS THE REFERENCE ITSELF RECORDED for its
50 px Alexander run (Results/50px_alexander_71pics_sphere_nerf_save_dir_4/video_save/*.avi, rendered by the reference's
TensorFlow path with NeRF_model_epoch_095.h5), so that the GPU box and the CPU suite (neither sees /root/reference at
test time) can hold the oracle's and the CUDA path's rendered frames -- and the video trajectories of ExecutionRun --
against outputs of the reference.

Run in the build container:  python tests/golden/make_alexander50_videos.py
Stored per video: frame count, fps, and a handful of decoded RGB frames (uint8, MJPG-decoded by OpenCV) at fixed frame
indices; the depth videos are gray (three equal channels), one channel is kept.
"""
import os

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
RUN = "/root/reference/Results/50px_alexander_71pics_sphere_nerf_save_dir_4/video_save"

PICKS = {   # file stem -> frame indices kept
    "render_l_to_r_rgb_video": [0, 75, 150, 225, 299],
    "render_depths_l_to_r_video": [0, 75, 150, 225, 299],
    "render_rgb_sphere_video": [0, 90, 180, 270, 360, 450, 540, 630],
    "render_depths_sphere_video": [0, 90, 180, 270],
    "render_rgb_path_video": [0, 60, 119, 120, 600, 1319],
    "render_depths_path_video": [0, 119, 600],
    "train_set_video": [0, 1, 18, 19, 20, 69],
}


def decode(path):
    cap = cv2.VideoCapture(path)
    fps = cap.get(cv2.CAP_PROP_FPS)
    frames = []
    while True:
        ok, bgr = cap.read()
        if not ok:
            break
        frames.append(cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB))
    return np.asarray(frames, dtype=np.uint8), fps


def main():
    out = {}
    for stem, picks in PICKS.items():
        frames, fps = decode(os.path.join(RUN, stem + ".avi"))
        kept = frames[picks]
        if "depths" in stem:
            assert np.abs(kept.astype(int) - kept[..., :1].astype(int)).max() <= 2       # gray up to chroma rounding
            kept = kept[..., 1]
        out[stem + "__n_frames"] = np.int64(len(frames))
        out[stem + "__fps"] = np.float64(fps)
        out[stem + "__indices"] = np.asarray(picks, dtype=np.int64)
        out[stem + "__frames"] = kept
        print(f"{stem}: {len(frames)} frames at {fps} fps, kept {picks}")
    np.savez_compressed(os.path.join(ROOT, "tests/golden/alexander50_videos.npz"), **out)
    print("wrote tests/golden/alexander50_videos.npz",
          os.path.getsize(os.path.join(ROOT, "tests/golden/alexander50_videos.npz")), "bytes")


if __name__ == "__main__":
    main()
