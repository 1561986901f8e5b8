#!/usr/bin/env python
"""Whole-path C-ABI entry points against the host package's call sequence (same kernels): wall time per call with the
GPU kept busy (host overhead shows when the batch is small) and device time per train step.

    python tools/fused_probe.py
"""
import importlib
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    t_issue = time.perf_counter() - t0
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, 1e3 * t_issue / reps


def main():
    pkg = importlib.import_module("nerf-and-dietnerf_b200")
    net = {"type_of_model": "NeRF", "hidden_layer_dim": 256, "last_hidden_layer_dim": 128, "leaky_relu_alpha": 0.05,
           "n_pos_enc_dim_xyz": 5, "n_pos_enc_view_dir": 4, "n_angles_for_model": 2, "n_rays_in_batch_train": 2048,
           "n_rays_in_batch_render": 16384}
    render = {"n_render_samples_coarse": 64, "n_render_samples_fine": 128}
    g = torch.Generator(device="cuda").manual_seed(0)
    for n in (2500, 16384):
        model = pkg.NeRFModel(net, render, 0.5576, 2.5635, mode="bf16", seed=0)
        o = torch.rand(n, 4, device="cuda", generator=g)
        d = torch.rand(n, 4, device="cuda", generator=g) - 0.5
        seq = timed(lambda: model.render(o, d, seed=1, step=0), 50)
        one = timed(lambda: model.render_fused(o, d, seed=1, step=0), 50)
        print(f"render {n} rays x (64 + 192): call sequence {seq[0]:.3f} ms/call (host issue {seq[1]:.3f} ms), "
              f"nerf_render_fused_fwd {one[0]:.3f} ms/call (host issue {one[1]:.3f} ms)")
    for n in (2048, 4096):
        y = torch.rand(n, 3, device="cuda", generator=g)
        o = torch.rand(n, 4, device="cuda", generator=g)
        d = torch.rand(n, 4, device="cuda", generator=g) - 0.5
        a = pkg.NeRFModel(net, render, 0.5576, 2.5635, mode="bf16", seed=0).compile(optimizer=pkg.Adam(5e-4))
        b = pkg.NeRFModel(net, render, 0.5576, 2.5635, mode="bf16", seed=0).compile(optimizer=pkg.Adam(5e-4))
        c = pkg.NeRFModel(net, render, 0.5576, 2.5635, mode="bf16", seed=0).compile(optimizer=pkg.Adam(5e-4))
        c.overlap_dw = False
        res = {}
        for rep in range(2):             # interleaved twice: the power-cap state drifts over a run
            res["seq"] = timed(lambda: a.train_step_local(o, d, y, n), 30)
            res["side"] = timed(lambda: b.train_step_fused(o, d, y), 30)
            res["one"] = timed(lambda: c.train_step_fused(o, d, y), 30)
            print(f"train step {n} rays x (64 + 128) [pass {rep}]: host package {res['seq'][0]:.3f} ms/step (host issue "
                  f"{res['seq'][1]:.3f} ms), nerf_train_step_fused with a side stream {res['side'][0]:.3f} ms/step (host issue "
                  f"{res['side'][1]:.3f} ms), on one stream {res['one'][0]:.3f} ms/step")


if __name__ == "__main__":
    main()
