#!/usr/bin/env python
"""Headline benchmark of the NeRF ray-render hot path: rays/s of the full train step
(render forward of the coarse+fine networks, loss, fused backward, Adam) on synthetic ray batches.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config NAME] [--mode fp16|bf16|fp32]

Workload (BASELINE.json configs[2], config_files/256px_alexander_71pics_sphere_nerf.yaml, the largest single-GPU training
config): 4096 rays per step PER GPU, 64 coarse + 128 fine samples per ray, 8x256 MLPs with view branch, Glorot-initialised
weights, rays from sphere cameras (near/far 0.5576/2.5635, fov 0.46134), targets U[0,1).  Weak scaling: every rank runs its
own 4096-ray shard of a global batch of N*4096 rays, with the gradient all-reduce(s) of the step over NCCL.

Prints ONE JSON line (rank 0).  `value` = rays/s with the ray batch resident in HBM, timed on the device with CUDA
events over exactly K steps (max over ranks) -- a BURST after an idle second; `sustained` = the same loop run for >= 2 s
of device time without the idle (power-capped clocks); `e2e` = the same metric through the public API `NeRF.train_step`
with PINNED HOST batches (H2D copy of the batch and D2H read of the loss inside the timed region, every step).
Sub-records measured in the same run: `render` (one 256x256 frame, 64 coarse + 128 new = 192 fine samples), `composite`
(alpha compositing at 65 536 rays x 192 samples against the HBM peak), `strong` (N > 1: the YAML's 4096-ray batch split
over the N ranks).  `--impl reference` times the CPU port of the reference's train step (TensorFlow is not installable
here) on all host cores, on a bounded sample of the same workload.
"""
import argparse
import contextlib
import importlib
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (train batch per GPU, near, far, fov)   -- SURVEY 8d per-config shapes
    "100px_robot_72pics_sphere": (2048, 0.3333, 2.0, 0.69111),
    "256px_alexander_71pics_sphere_nerf": (4096, 0.5576, 2.5635, 0.46134),
    "50px_alexander_71pics_sphere_nerf": (4096, 0.5576, 2.5635, 0.46134),
    # DietNeRF (BASELINE configs[4]): the same ray loss (2*MSE_c + MSE_f) plus, every 13th step, the semantic-consistency
    # term: a 150x150 in-tape render (55 + 55 samples), random-init ViT-B/32 embedding, cosine loss, full backward
    "256px_alexander_71pics_sphere_dietnerf": (2048, 0.5576, 2.5635, 0.46134),
}
N_C, N_F = 64, 128
MAC_FWD = 512152                    # MLP forward MAC per sample
MAC_DX_COARSE, MAC_DX_FINE = 492160, 509056
FLOP_PER_RAY_TRAIN = 2 * (N_C * (2 * MAC_FWD + MAC_DX_COARSE) + N_F * (2 * MAC_FWD + MAC_DX_FINE))   # 586.6 MFLOP


def net_config(batch):
    return {"hidden_layer_dim": 256, "last_hidden_layer_dim": 128, "leaky_relu_alpha": 0.05, "n_pos_enc_dim_xyz": 5,
            "n_pos_enc_view_dir": 4, "n_angles_for_model": 2, "n_rays_in_batch_train": batch,
            "n_rays_in_batch_render": 16384}


def synthetic_batch(n, fov, seed, rays_fn):
    """n rays from random sphere cameras + uniform targets (CPU tensors).  rays_fn(c2w, fov, h, w) -> (orig, dirs):
    the product's own ray generator on the GPU arm, the oracle's on the CPU reference arm."""
    import numpy as np
    import torch
    rng = np.random.default_rng(seed)
    o_all, d_all, got = [], [], 0
    while got < n:
        th, ph, rad = rng.uniform(0, 2 * math.pi), rng.uniform(-0.6, 0.6), rng.uniform(0.8, 1.0)
        cam = np.array([rad * math.cos(ph) * math.sin(th), rad * math.sin(ph), rad * math.cos(ph) * math.cos(th)])
        fwd = -cam / np.linalg.norm(cam)
        right = np.cross(fwd, [0.0, 1.0, 0.0])
        right /= np.linalg.norm(right)
        up = np.cross(right, fwd)
        c2w = np.eye(4, dtype=np.float32)
        c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = right, up, -fwd, cam
        o, d = rays_fn(c2w, fov, 32, 32)
        o_all.append(o)
        d_all.append(d)
        got += o.shape[0]
    o, d = torch.cat(o_all)[:n], torch.cat(d_all)[:n]
    perm = torch.from_numpy(rng.permutation(n))
    y = torch.rand(n, 3, generator=torch.Generator().manual_seed(seed))
    return o[perm].contiguous(), d[perm].contiguous(), y


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples taken DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.samples, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def window(self, t0, t1):
        rows = [s for t, s in self.samples if t0 - 0.05 <= t <= t1 + 0.15] or [s for _, s in self.samples[-3:]]
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
            except (ValueError, IndexError):
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()


CPU_SAMPLE_RAYS = 1024              # bounded sample of the batch the CPU arm runs per step (a few seconds on 16 cores)


def cpu_reference_rate(cfg_name, n_rays, reps, threads=None):
    """rays/s of the CPU port of NeRF.train_step (forward + autograd backward + Adam) on the host cores.  The port is the
    oracle with its per-sample Python loops (the canonical summation order of the parity tests) switched to
    torch.cumprod / cumsum / sum (oracle.FAST_REDUCTIONS): the throughput an optimised CPU implementation gets, on a
    sample large enough (>= 1024 rays) that every op is a full-width tensor op."""
    import torch
    from oracle import nerf_oracle as O
    O.FAST_REDUCTIONS = True
    batch, near, far, fov = CONFIGS[cfg_name]
    # all host threads, also under torchrun (which exports OMP_NUM_THREADS=1)
    torch.set_num_threads(threads or os.cpu_count() or 1)
    ocfg = O.NetCfg()
    pc, pf = O.glorot_params(ocfg.shapes, 0), O.glorot_params(ocfg.shapes, 1)
    o, d, y = synthetic_batch(n_rays, fov, 0, O.rays_for_image)
    mc, vc, mf, vf = (torch.zeros_like(pc) for _ in range(4))
    times = []
    for t in range(1, reps + 2):
        jit, u = O.stratified_jitter(0, t, n_rays, N_C), O.importance_uniforms(0, t, n_rays, N_F)
        t0 = time.perf_counter()
        _, gc, gf, _ = O.train_step(pc, pf, ocfg, near, far, o, d, y, N_C, N_F, jit, u)
        pc, mc, vc = O.adam_step(pc, gc, mc, vc, t, 5e-4)
        pf, mf, vf = O.adam_step(pf, gf, mf, vf, t, 5e-4)
        times.append(time.perf_counter() - t0)
    times = sorted(times[1:])          # first repetition is the warm-up
    return n_rays / times[len(times) // 2], torch.get_num_threads()


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  TensorFlow/Keras 2.7 cannot be installed
    in this image (no wheel, no network), so the line-by-line oracle port is what runs, on all host
    threads, each step = a bounded 1024-ray sample of the batch (vectorised reductions, see cpu_reference_rate)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_sample = CPU_SAMPLE_RAYS
    t0 = time.perf_counter()
    rate, cores = cpu_reference_rate(args.config, n_sample, max(1, args.steps), None)
    batch = CONFIGS[args.config][0]
    line = {
        "impl": "reference", "metric": "rays/sec render fwd+bwd (train step)", "value": rate, "unit": "rays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * n_sample / rate,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"train_step {args.config}: {batch} rays/step/GPU, {N_C} coarse + {N_F} fine samples",
                   "sample": f"{n_sample}-ray sample of the batch per step"},
        "cpu_baseline": {"value": rate, "unit": "rays/s", "cores": cores, "kind": "port",
                         "sample": f"{n_sample} rays x {N_C + N_F} samples per step, CPU port (PyTorch-CPU fp32, vectorised "
                                   "cumprod/cumsum) of NeRF.train_step incl. Adam; TensorFlow unavailable offline"},
        "e2e": {"value": rate, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="256px_alexander_71pics_sphere_nerf", choices=sorted(CONFIGS))
    ap.add_argument("--mode", default="fp16", choices=["fp16", "bf16", "fp32"],
                    help="fp16 (the product's default): fp16 forward operands, bf16 backward; bf16: bf16 everywhere")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--settle", type=float, default=1.0, help="idle seconds before each timed pass (power-cap state)")
    ap.add_argument("--sustained-s", type=float, default=2.0, help="device seconds of the sustained pass (0 = skip)")
    ap.add_argument("--no-extras", action="store_true", help="skip the render / composite / strong sub-records")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device: the product has no CPU path"
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # rank 0 must print ONE JSON line on stdout, and NCCL writes there too: at NCCL_DEBUG=VERSION / WARN its version
        # banner is a plain printf (drop those levels), at INFO and above the log goes where NCCL_DEBUG_FILE points
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            del os.environ["NCCL_DEBUG"]
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    pkg = importlib.import_module("nerf-and-dietnerf_b200")
    pkg.load()

    batch, near, far, fov = CONFIGS[args.config]
    n_total = batch * world
    diet = args.config.endswith("dietnerf")
    rcfg = {"n_render_samples_coarse": N_C, "n_render_samples_fine": N_F}
    if diet:
        import numpy as np
        g = torch.Generator().manual_seed(0)
        targets = torch.rand(8, 64, 64, 3, generator=g).numpy()        # stand-ins for the training images (embedded once)
        poses = np.stack([np.eye(4, dtype=np.float32) for _ in range(8)])
        model = pkg.DietNeRFModel(net_config(batch), rcfg, near, far, targets, poses, fov, -1, np.zeros(3), np.eye(4),
                                  mode=args.mode, seed=0, numpy_seed=0, resample_every_call=True)
    else:
        model = pkg.NeRFModel(net_config(batch), rcfg, near, far, mode=args.mode, seed=0)
    model.compile(optimizer=pkg.Adam(5e-4))
    if world > 1:
        model.distribute()
    # this rank's shard of the global batch; a few distinct batches so consecutive steps do not see the same rays
    n_batches = 4

    def gpu_rays(c2w, fov_, h, w):
        dirs, orig = pkg.UtilsCV.get_rays_directions(h, w, fov_, c2w, return_origins=True)
        return orig.cpu(), dirs.reshape(-1, 4).cpu()
    host = [synthetic_batch(batch, fov, 1000 * b + rank, gpu_rays) for b in range(n_batches)]
    pinned = [tuple(t.pin_memory() for t in hb) for hb in host]
    devb = [tuple(t.cuda() for t in hb) for hb in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def train_local(o, d, y):
        if diet:       # DietNeRF.train_step: counter, consistency term every 13th step, then the sharded ray step
            return model.train_step_sharded(o, d, y, n_total, rank * batch)
        return model.train_step_local(o, d, y, n_total, rank * batch)

    def step_device(i):
        o, d, y = devb[i % n_batches]
        return train_local(o, d, y)

    # e2e: every step copies ITS batch from pinned host memory and its loss is read back to the host, all inside the timed
    # region, through the public API: batches flow through DevicePrefetcher (the product's stand-in for tf.data's
    # prefetch: batch i+1 crosses PCIe on a copy stream while batch i trains) into NeRF.train_step; the loss goes to a
    # pinned 4-byte slot on a second copy stream and is consumed two steps later (the way a training loop logs; a
    # blocking .item() per step would only add host launch latency).
    LAG = 2          # the host reads step i's loss while issuing step i + LAG: the GPU always has a full step queued
    loss_slots = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(LAG + 1)]
    loss_events = [torch.cuda.Event() for _ in range(LAG + 1)]
    losses = []
    d2h = torch.cuda.Stream()

    def read_loss(i):
        loss_events[i % (LAG + 1)].synchronize()
        losses.append(float(loss_slots[i % (LAG + 1)][0]))

    def run_e2e(k):
        feeder = pkg.UtilsNeuralRadianceField.DevicePrefetcher(pinned[i % n_batches] for i in range(k))
        for i, (od, dd, yd) in enumerate(feeder):
            m = train_local(od, dd, yd)
            done = torch.cuda.Event()
            done.record()
            d2h.wait_event(done)
            with torch.cuda.stream(d2h):
                loss_slots[i % (LAG + 1)].copy_(m["loss"].reshape(1), non_blocking=True)   # D2H read of the step's result
                loss_events[i % (LAG + 1)].record(d2h)
            m["loss"].record_stream(d2h)
            if i >= LAG:
                read_loss(i - LAG)
        for i in range(max(k - LAG, 0), k):
            read_loss(i)
        torch.cuda.current_stream().wait_stream(d2h)

    def timed(fn, k, whole=False):
        barrier()
        # every timed pass starts from the same power state: the passes run back to back on a GPU that is power-capped
        # under sustained load (sw_power_cap), so without the pause the later pass (e2e) is measured at lower clocks
        # than the earlier one (device-resident) -- 1.68 vs 1.80 ms/step for the same work
        time.sleep(args.settle)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        if whole:
            fn(k)
        else:
            for i in range(k):
                fn(i)
        e1.record()
        barrier()
        t1 = time.time()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item(), t0, t1

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t_load0 = time.time()
    # DietNeRF: the warm-up must contain a consistency step (every 13th), or its one-time allocations and the lazily built
    # embedder land inside the timed region
    n_warm = max(args.warmup, 13 if diet else 3)
    for i in range(n_warm):
        step_device(i)
    barrier()
    launches0 = pkg._lib.launch_count
    ms_dev, t0, t1 = timed(step_device, args.steps)
    launches = pkg._lib.launch_count - launches0

    # ---- sustained pass: the same loop for >= --sustained-s seconds of device time, no idle before it ---------------------
    sustained = None
    if args.sustained_s > 0:
        chunk = max(20, args.steps)
        barrier()
        t_s0 = time.time()
        evs = [torch.cuda.Event(enable_timing=True)]
        evs[0].record()
        n_done, dev_ms = 0, 0.0
        while dev_ms < args.sustained_s * 1e3 and n_done < 200000:
            for i in range(chunk):
                step_device(n_done + i)
            n_done += chunk
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            e.synchronize()
            evs.append(e)
            done = torch.tensor([evs[0].elapsed_time(e)], device="cuda")
            if world > 1:                      # every rank runs the same number of chunks
                dist.all_reduce(done, op=dist.ReduceOp.MAX)
            dev_ms = done.item()
        barrier()
        t_s1 = time.time()
        # steady state = the second half of the window (the first chunks still run at burst clocks)
        half = len(evs) // 2
        tail_ms = torch.tensor([evs[half].elapsed_time(evs[-1])], device="cuda")
        if world > 1:
            dist.all_reduce(tail_ms, op=dist.ReduceOp.MAX)
        tail_steps = (len(evs) - 1 - half) * chunk
        sustained = {"window_s": dev_ms * 1e-3, "steps": n_done, "ms_per_step": dev_ms / n_done,
                     "value": n_total * n_done / (dev_ms * 1e-3), "unit": "rays/s",
                     "second_half_ms_per_step": tail_ms.item() / max(tail_steps, 1),
                     "second_half_value": n_total * tail_steps / (tail_ms.item() * 1e-3) if tail_steps else None,
                     "clocks": sampler.window(t_s0, t_s1) if rank == 0 else None}

    # per-call device times of the MLP kernels over another timed pass (events on the launching stream)
    per_call = {}

    @contextlib.contextmanager
    def hook(name, args=()):
        if name == "nerf_mlp_bwd_rays":        # the fine network's backward with d z formed in the chain: parts 1 / 2
            name = {1: "nerf_mlp_bwd_dx", 2: "nerf_mlp_bwd_dw"}.get(args[14], name)
        if name in ("nerf_mlp_fwd", "nerf_mlp_fwd_rays", "nerf_mlp_fwd_rays_stratified", "nerf_mlp_bwd", "nerf_mlp_bwd_dx",
                    "nerf_mlp_bwd_dw",
                    "nerf_composite_fwd", "nerf_composite_bwd", "nerf_composite_mse_fwd", "nerf_composite_mse_fwd_bwd",
                    "nerf_sample_pdf_fwd", "nerf_sample_pdf_bwd"):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            yield
            b.record()
            per_call.setdefault(name, []).append((a, b))
        else:
            yield
    pkg._lib.event_hook = hook
    model.overlap_dw = False          # one stream, the backward's two halves as separate calls: each kernel is timed alone
    model.use_fused_step = False      # ... through the host package's call sequence (the same kernels as the one C call)
    type(model).split_bwd_calls = True
    timed(step_device, args.steps)
    model.overlap_dw = True
    model.use_fused_step = True
    type(model).split_bwd_calls = False
    pkg._lib.event_hook = None
    torch.cuda.synchronize()
    # the coarse forward (stratified sampling fused in) and the fine forward are the same kernel: one average over both
    fwd_events = per_call.pop("nerf_mlp_fwd_rays_stratified", []) + per_call.pop("nerf_mlp_fwd_rays", [])
    if fwd_events:
        per_call["nerf_mlp_fwd_rays"] = fwd_events
    call_ms = {k: sum(a.elapsed_time(b) for a, b in v) / len(v) for k, v in per_call.items()}
    call_n = {k: len(v) // args.steps for k, v in per_call.items()}

    run_e2e(8)        # warm-up: also brings the copy streams' allocator pools to their steady-state size
    losses.clear()
    ms_e2e, _, t_load1 = timed(run_e2e, args.steps, whole=True)
    assert len(losses) == args.steps and all(math.isfinite(v) for v in losses), "every step's loss must reach the host"
    # the device-timed region alone lasts ~0.1 s (one nvidia-smi sample); report the median over every sample taken
    # while the GPU ran back-to-back steps (warm-up, device-timed, sustained, per-call-timed and e2e passes)
    clocks = sampler.window(t_load0, t_load1) if rank == 0 else None
    if clocks is not None:
        clocks["window"] = ("warm-up through e2e pass; the GPU idles %.1f s before each burst-timed pass so that every "
                            "pass starts from the same power-cap state; the sustained pass has its own clocks record"
                            % args.settle)

    def dev_time(fn, iters, warm=2):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters

    # ---- strong scaling: the YAML's batch split over the ranks (src/NeRF.py:136-147 fixes B from the config) ------------
    strong = None
    if world > 1 and not diet and not args.no_extras:
        per = batch // world
        sb = [tuple(t[rank * per:(rank + 1) * per].contiguous() for t in db) for db in devb]

        def step_strong(i):
            o, d, y = sb[i % n_batches]
            return model.train_step_local(o, d, y, per * world, rank * per)
        for i in range(5):
            step_strong(i)
        ms_strong, _, _ = timed(step_strong, args.steps)
        strong = {"global_batch_rays": per * world, "rays_per_gpu": per, "ms_per_step": ms_strong / args.steps,
                  "value": per * world * args.steps / (ms_strong * 1e-3), "unit": "rays/s",
                  "note": "same optimisation problem at every N (the weak-scaling `value` grows the global batch)"}

    # ---- render: one 256x256 frame at the reference's render setting (BASELINE configs[3], every YAML: 64 coarse samples +
    #      128 importance draws, the fine network sees the 192 merged depths = 256 network rows = 262.2 MFLOP per ray,
    #      BASELINE.md), this rank's row block ---
    render = composite = None
    if not args.no_extras and not diet:
        h = w = 256
        import numpy as np
        c2w = np.eye(4, dtype=np.float32)
        c2w[:3, 3] = [0.0, 0.0, 1.0]
        lo, hi = (h * w * rank) // world, (h * w * (rank + 1)) // world
        with torch.no_grad():
            frame = lambda: model.render_image_lean(c2w, fov, h, w, 16384, N_C, N_F, seed=1, step=0, ray_begin=lo,
                                                    n_rays=hi - lo)
            barrier()
            ms_frame = torch.tensor([dev_time(frame, 3)], device="cuda")
        if world > 1:
            dist.all_reduce(ms_frame, op=dist.ReduceOp.MAX)
        ms_frame = ms_frame.item()
        flop_ray_render = 2 * MAC_FWD * (N_C + N_C + N_F)
        render = {"what": "render_image_lean, one 256x256 frame, n_render_samples_coarse 64 + n_render_samples_fine 128 as in "
                          "every reference YAML (the fine network sees the 192 merged depths: 256 network rows = 262.2 "
                          "MFLOP per ray), rows sharded over the ranks, no collective; records before r02_bk drew 192 new "
                          "samples (320 rows per ray)",
                  "ms_per_frame": ms_frame, "value": h * w / (ms_frame * 1e-3), "unit": "rays/s",
                  "mlp_tflops": flop_ray_render * h * w / (ms_frame * 1e-3) / 1e12}
        if rank == 0:
            # ---- compositing alone at the size of one frame x 192 samples: all outputs / lean / backward ------------------
            call, ptr = pkg._lib.call, pkg._lib.ptr
            n, sm = 65536, 192
            raw = torch.randn(n, sm, 4, device="cuda")
            z = torch.sort(torch.rand(n, sm, device="cuda"), -1).values.contiguous()
            rgb, wt, T = torch.empty(n, 3, device="cuda"), torch.empty(n, sm, device="cuda"), torch.empty(n, sm, device="cuda")
            al, rs = torch.empty(n, sm, device="cuda"), torch.empty(n, sm, 3, device="cuda")
            dep, acc = torch.empty(n, device="cuda"), torch.empty(n, device="cuda")
            d_rgb, d_raw, d_z = torch.randn(n, 3, device="cuda"), torch.empty_like(raw), torch.empty_like(z)
            full = dev_time(lambda: call("nerf_composite_fwd", ptr(raw), ptr(z), n, sm, ptr(rgb), ptr(wt), ptr(T), ptr(al),
                                         ptr(rs), None, None), 10)
            lean = dev_time(lambda: call("nerf_composite_fwd", ptr(raw), ptr(z), n, sm, ptr(rgb), ptr(wt), None, None, None,
                                         ptr(dep), ptr(acc)), 10)
            bwd = dev_time(lambda: call("nerf_composite_bwd", ptr(raw), ptr(z), ptr(d_rgb), ptr(wt), n, sm, ptr(d_raw),
                                        ptr(d_z)), 10)
            del raw, z, rgb, wt, T, al, rs, d_raw, d_z
            gbs = lambda byts, ms: byts / (ms * 1e-3) / 1e9
            composite = {"rays": n, "samples": sm, "unit": "GB/s (algorithmic bytes / device time)",
                         "fwd_all_outputs": gbs(n * (sm * 44 + 12), full), "fwd_lean": gbs(n * (sm * 24 + 20), lean),
                         "bwd": gbs(n * (sm * 44 + 12), bwd),
                         "bytes_per_sample": {"fwd_all_outputs": 44, "fwd_lean": 24, "bwd": 44},
                         "note": "inputs of a call (252-302 MB) exceed the 126 MB L2, no flush between the 10 timed calls; "
                                 "`peak` is the measured COPY bandwidth (reads = writes): a read-dominated call can pass it "
                                 "(ncu: 7.05 TB/s of DRAM reads on this part) and the lean forward's 50 MB of weights may "
                                 "still sit in the write-back L2 when the call ends (ncu: 26 MB of DRAM writes), so its "
                                 "fraction can exceed 1"}
    sampler.stop()

    if rank == 0:
        rays = n_total * args.steps
        value = rays / (ms_dev * 1e-3)
        e2e = rays / (ms_e2e * 1e-3)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            pass
        peak_burst = peaks.get("bf16_tflops", 1640.0)
        peak_sust = peaks.get("bf16_tflops_sustained", 1400.0)
        peak_src = "MEASURED_PEAKS.json bf16_tflops (burst) / bf16_tflops_sustained" if "bf16_tflops" in peaks else \
            "fallback (B200_PROFILING.md)"
        peak_hbm = peaks.get("hbm_gbs", 6500.0)
        # Dominant kernel by time: mlp_tc_bwd_dw_kernel, the weight gradients.  Its arithmetic intensity is fixed by the
        # 256x256 output it keeps in TMEM (128 FLOP/B), so its roofline is bandwidth: it streams every saved activation and
        # every dZ once.  Algorithmic bytes per sample: bf16 saved activations (64 + 8*256 + 128 columns) +
        # bf16 dZ (8*256 + 144 + 16 columns) = 8896 B.  The launch time is the nerf_mlp_bwd_dw call (dW kernel + its
        # fixed-order reduce), averaged over the coarse (64 samples/ray) and fine (128) calls of a step, timed with CUDA
        # events in THIS run (the per-call pass above).  `traffic` is NOT measured in this run: it is the dram read+write
        # of the same two launches in the committed ncu --set full capture (9431 B/sample), scaled to this batch.
        samples_per_launch = batch * (N_C + N_F) / 2
        dw_ms = call_ms.get("nerf_mlp_bwd_dw", float("nan"))
        dx_ms = call_ms.get("nerf_mlp_bwd_dx", float("nan"))
        fwd_ms = call_ms.get("nerf_mlp_fwd_rays", call_ms.get("nerf_mlp_fwd", float("nan")))
        fwd_flops = 2 * batch * (N_C + N_F) * MAC_FWD / 2
        dx_flops = 2 * batch * (N_C * MAC_DX_COARSE + N_F * MAC_DX_FINE) / 2
        dw_flops = fwd_flops
        ach = 8896 * samples_per_launch / (dw_ms * 1e-3) / 1e9
        step_tf = FLOP_PER_RAY_TRAIN * value / world / 1e12
        tensor = {
            "unit": "TFLOP/s", "peak_source": peak_src, "flop_per_ray": FLOP_PER_RAY_TRAIN,
            "burst": {"peak": peak_burst, "achieved": step_tf, "frac": step_tf / peak_burst,
                      "window": f"{args.steps} steps after {args.settle} s idle (`value`)"},
            "mlp_tc_fwd_kernel<save>": {"achieved": fwd_flops / (fwd_ms * 1e-3) / 1e12, "ms": fwd_ms},
            "mlp_tc_bwd_chain_kernel": {"achieved": dx_flops / (dx_ms * 1e-3) / 1e12, "ms": dx_ms},
            "mlp_tc_bwd_dw_kernel": {"achieved": dw_flops / (dw_ms * 1e-3) / 1e12, "ms": dw_ms},
            "step_tensor_frac": step_tf / peak_burst,
        }
        if sustained is not None:
            s_tf = FLOP_PER_RAY_TRAIN * sustained["value"] / world / 1e12
            s2 = sustained.get("second_half_value")
            tensor["sustained"] = {"peak": peak_sust, "achieved": s_tf, "frac": s_tf / peak_sust,
                                   "second_half_frac": (FLOP_PER_RAY_TRAIN * s2 / world / 1e12 / peak_sust) if s2 else None,
                                   "window": "%.2f s of back-to-back steps, no idle before it" % sustained["window_s"]}
        for k in ("mlp_tc_fwd_kernel<save>", "mlp_tc_bwd_chain_kernel", "mlp_tc_bwd_dw_kernel"):
            tensor[k]["frac_of_burst_peak"] = tensor[k]["achieved"] / peak_burst
        if render is not None:
            render["mlp_frac_of_burst_peak"] = render["mlp_tflops"] / world / peak_burst
        if composite is not None:
            composite["peak"] = peak_hbm
            composite["frac"] = {k: composite[k] / peak_hbm for k in ("fwd_all_outputs", "fwd_lean", "bwd")}
        roofline = {"bound": "hbm", "kernel": "mlp_tc_bwd_dw_kernel (nerf_mlp_bwd_dw call)",
                    "achieved": ach, "peak": peak_hbm, "unit": "GB/s", "frac": ach / peak_hbm,
                    "traffic": 9431 * samples_per_launch,
                    "traffic_source": "not measured in this run: dram__bytes_read+write of these two launches in "
                                      "profiles/r02_al_mlp_full_summary.txt (ncu --set full of the same command): "
                                      "2.432 + 4.985 GB for 262144 + 524288 samples = 9431 B/sample",
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6.5 TB/s",
                    "algorithmic_bytes_per_launch": 8896 * samples_per_launch,
                    "tensor_kernels": tensor,
                    "avg_call_ms": call_ms, "calls_per_step": call_n}
        if args.mode == "fp32":
            roofline["bound"] = "fp32 SIMT parity mode: the tensor-core rooflines above do not apply"
        cpu = None
        if not args.no_cpu_baseline:
            rate, cores = cpu_reference_rate(args.config, CPU_SAMPLE_RAYS, 2)
            cpu = {"value": rate, "unit": "rays/s", "cores": cores, "kind": "port",
                   "sample": f"{CPU_SAMPLE_RAYS}-ray sample of the batch, 2 timed steps after 1 warm-up, CPU port (PyTorch-CPU "
                             "fp32, vectorised cumprod/cumsum reductions) of NeRF.train_step incl. Adam; TensorFlow "
                             "unavailable offline"}
        line = {
            "metric": "rays/sec render fwd+bwd (train step)", "value": value, "unit": "rays/s", "n_gpus": world,
            "steps": args.steps, "warmup": n_warm, "ms_per_step": ms_dev / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp16": "fp16 forward operands / bf16 backward operands, fp32 accumulate", "bf16": "bf16",
                      "fp32": "f32"}[args.mode], "data": "synthetic",
            "config": {"workload": f"train_step {args.config}: {batch} rays/step/GPU, {N_C} coarse + {N_F} fine samples, "
                                   "8x256 MLPs + view branch, Adam"
                                   + (", + every 13th step the DietNeRF consistency term (150x150 in-tape render at "
                                      "55+55 samples, random-init ViT-B/32, cosine loss, backward)" if diet else ""),
                       "global_batch_rays": n_total, "parallelism": f"ray-sharded dp{world}",
                       "settle_s_before_each_timed_pass": args.settle,
                       "l2": "working set per step (saved activations + dZ, ~9 GB at 4096 rays) >> 126 MB L2; "
                             "4 distinct ray batches rotate"},
            "e2e": {"value": e2e, "unit": "rays/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": batch * (16 + 16 + 12), "d2h_bytes_per_step": 4,
                    "how": "DevicePrefetcher (pinned host batch -> device on a copy stream, one batch ahead) -> "
                           "train step -> loss to a pinned host slot (asynchronous read-back consumed two steps later); "
                           "every copy of every step inside the timed region"},
            "sustained": sustained,
            "strong": strong,
            "render": render,
            "composite": composite,
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roofline,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
