#!/usr/bin/env python
"""Repeat the training run the reference recorded (Results/50px_alexander_71pics_sphere_nerf_save_dir_4: 70 training
views of the 50 px Alexander scene, 4096 rays/step, 42 steps/epoch, Adam 4e-4, 95 epochs) with this framework's train
loop, and print the held-out-image / training-image PSNR per epoch next to the reference's recorded curve.

    python tools/train_alexander50.py [--epochs 95] [--mode bf16|fp32] [--out gpurun_out/train_alexander50.json]

Needs tests/golden/alexander50_dataset.npz (the scene after the loader) and alexander50_pin.npz (the recorded PSNRs).
"""
import argparse
import importlib
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

RUN_CONFIG = {      # Results/50px_alexander_71pics_sphere_nerf_save_dir_4/50px_alexander_71pics_sphere_nerf.yaml
    "existing_save_dir_name": None, "starting_epoch_number": -1, "dataset_type": "colmap",
    "tasks_to_perform": {"start_training": True},
    "neural_net": {"type_of_model": "NeRF", "hidden_layer_dim": 256, "last_hidden_layer_dim": 128,
                   "leaky_relu_alpha": 0.05, "n_pos_enc_dim_xyz": 5, "n_pos_enc_view_dir": 4, "n_angles_for_model": 2,
                   "n_rays_in_batch_train": 4096, "n_rays_in_batch_render": 4096},
    "render": {"n_render_samples_coarse": 64, "n_render_samples_fine": 128},
    "training": {"n_epochs": 95, "optimizer_lr": 4.0e-4, "test_img_idx": 19, "idx_train_img_to_plot": 4},
}


def run(epochs=95, mode="bf16", seed=0, save_location=None, model_type="NeRF", stop_grad_z=False, lr=None,
        start_epoch=-1, history=None):
    pkg = importlib.import_module("nerf-and-dietnerf_b200")
    data = np.load(os.path.join(ROOT, "tests", "golden", "alexander50_dataset.npz"))
    pin = np.load(os.path.join(ROOT, "tests", "golden", "alexander50_pin.npz"))
    config = json.loads(json.dumps(RUN_CONFIG))
    config["training"]["n_epochs"] = epochs
    config["starting_epoch_number"] = start_epoch
    if lr is not None:
        config["training"]["optimizer_lr"] = lr
    config["neural_net"]["type_of_model"] = model_type
    images = data["images_u8"].astype(np.float32) / 255.0
    runner = pkg.ExecutionRun.from_arrays(config, images, data["c2w"], float(data["fov"]), float(data["near"]),
                                          float(data["far"]), mode=mode, seed=seed, save_location=save_location)
    runner.stop_grad_z = stop_grad_z
    if history:
        runner.history = list(history)
    t0 = time.time()
    runner._training()
    torch.cuda.synchronize()
    wall = time.time() - t0
    hist = runner.history
    ref_test, ref_train = pin["psnr_reference_test"], pin["psnr_reference_train"]
    for h in hist:
        e = h["epoch"]
        if e <= len(ref_test):
            h["psnr_test_reference"], h["psnr_train_reference"] = float(ref_test[e - 1]), float(ref_train[e - 1])
    hist.sort(key=lambda h: h["epoch"])
    train_s = sum(h["seconds"] for h in hist)
    steps = 42 * len(hist)
    return {"epochs": len(hist), "mode": mode, "model": model_type, "wall_s": wall, "train_s": train_s,
            "ms_per_step": 1e3 * train_s / steps, "rays_per_s": steps * 4096 / train_s, "history": hist}, runner


def main():
    if "RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1:      # torchrun: shard every batch over the ranks
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=95)
    ap.add_argument("--mode", default="bf16")
    ap.add_argument("--model", default="NeRF", choices=["NeRF", "DietNeRF"])
    ap.add_argument("--out", default=None)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--stages", default=None, help="e.g. 70:5e-4,95:4e-4 -- train to epoch 70 at 5e-4, then RESUME from the "
                    "checkpoint (fresh Adam state, as the reference's get_nerf does) to epoch 95 at 4e-4")
    ap.add_argument("--stop-grad-z", action="store_true", help="detach the importance samples (NOT the reference's behaviour)")
    args = ap.parse_args()
    rank0 = int(os.environ.get("RANK", "0")) == 0
    with tempfile.TemporaryDirectory() as tmp:
        if int(os.environ.get("WORLD_SIZE", "1")) > 1:
            import torch.distributed as dist
            box = [tmp]
            dist.broadcast_object_list(box, src=0)       # every rank resumes from rank 0's checkpoints
            tmp = box[0]
        if args.stages:
            start, hist = -1, None
            for stage in args.stages.split(","):
                end, lr = stage.split(":")
                res, runner = run(int(end), args.mode, seed=args.seed, save_location=tmp, model_type=args.model,
                                  stop_grad_z=args.stop_grad_z, lr=float(lr), start_epoch=start, history=hist)
                start, hist = int(end), res["history"]
            args.epochs = start
        else:
            res, runner = run(args.epochs, args.mode, seed=args.seed, save_location=tmp, model_type=args.model,
                              stop_grad_z=args.stop_grad_z)
        # checkpoint round trip through the Keras .h5 layout
        path = runner.model.get_nerf_model_path(tmp, args.epochs)
        before = runner.model.model_fine.params.clone()
        runner.model.load_weights(path)
        assert torch.equal(before, runner.model.model_fine.params), "checkpoint round trip changed the weights"
        res["checkpoint_bytes"] = os.path.getsize(path)
        if int(os.environ.get("WORLD_SIZE", "1")) > 1:
            import torch.distributed as dist
            p = runner.model.model_fine.params
            both = [torch.empty_like(p) for _ in range(dist.get_world_size())]
            dist.all_gather(both, p)
            assert all(torch.equal(both[0], t) for t in both), "replicas diverged"
            res["world_size"] = dist.get_world_size()
            dist.barrier()
    if not rank0:
        return
    print(f"{'epoch':>5} {'test':>7} {'ref':>7} {'train':>7} {'ref':>7}")
    for h in res["history"]:
        if h["epoch"] in (1, 2, 5) or h["epoch"] % 10 == 0 or h["epoch"] == args.epochs:
            print(f"{h['epoch']:5d} {h['psnr_test']:7.2f} {h.get('psnr_test_reference', float('nan')):7.2f} "
                  f"{h['psnr_train']:7.2f} {h.get('psnr_train_reference', float('nan')):7.2f}")
    print(f"{res['epochs']} epochs in {res['train_s']:.1f} s of training ({res['ms_per_step']:.2f} ms/step, "
          f"{res['rays_per_s'] / 1e6:.2f} M rays/s incl. host loop); wall {res['wall_s']:.1f} s")
    if args.out:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        with open(args.out, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
