"""Video files of the render loop: the MJPG ``.avi`` writer of the reference (src/UtilsVideo.py:16-39) and a reader
for tests.  Host side, OpenCV's encoder; the frames come off the GPU as uint8 (``ExecutionRun.render_frames``), so the
writer accepts both the reference's float frames in [0, 1] and ready uint8 frames.  The plot videos of the reference
(``save_plot_video``, matplotlib) are outside the path.
"""
import os
from pathlib import Path

import numpy as np


def frame_to_uint8(frame):
    """Float frames in [0,1] -> ``uint8(round(255 f))`` as the reference does (:34); uint8 frames pass through."""
    frame = np.asarray(frame)
    if frame.dtype == np.uint8:
        return frame
    return np.uint8(np.round(frame * 255))


def save_frames_as_video(filename, frames, fps):
    """Write ``frames`` (list or array of (h,w,3) RGB or (h,w) gray frames, float in [0,1] or uint8) as an MJPG video
    at ``fps``; the parent directory is created when missing.  Gray frames are stored as three equal channels, which
    is what ``cv2.cvtColor(gray, COLOR_RGB2BGR)`` makes of them in the reference."""
    import cv2
    assert len(frames) > 0
    filename = Path(filename)
    os.makedirs(filename.parent, exist_ok=True)
    first = np.asarray(frames[0])
    writer = cv2.VideoWriter(str(filename), cv2.VideoWriter_fourcc(*"MJPG"), fps, (first.shape[1], first.shape[0]))
    if not writer.isOpened():
        raise Exception(f"cannot open a MJPG writer for {filename}")
    try:
        for frame in frames:
            u8 = frame_to_uint8(frame)
            to_bgr = cv2.COLOR_GRAY2BGR if u8.ndim == 2 else cv2.COLOR_RGB2BGR
            writer.write(np.ascontiguousarray(cv2.cvtColor(u8, to_bgr)))
    finally:
        writer.release()
    print(f"video with {len(frames)} frames at {fps} fps: {filename}")


def read_video_frames(filename):
    """-> (frames uint8 (F,h,w,3) RGB, fps): decodes a file written by ``save_frames_as_video`` (or by the reference)."""
    import cv2
    cap = cv2.VideoCapture(str(filename))
    if not cap.isOpened():
        raise Exception(f"cannot open {filename}")
    fps = cap.get(cv2.CAP_PROP_FPS)
    frames = []
    while True:
        ok, bgr = cap.read()
        if not ok:
            break
        frames.append(cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB))
    cap.release()
    return np.asarray(frames, dtype=np.uint8), fps
