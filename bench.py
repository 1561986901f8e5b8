#!/usr/bin/env python
"""Headline benchmark of the NeRF ray-render hot path: rays/s of the full train step
(render forward of the coarse+fine networks, loss, fused backward, Adam) on synthetic ray batches.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config NAME] [--mode bf16|fp32]

Workload (BASELINE.json configs[1], config_files/100px_robot_72pics_sphere.yaml): 2048 rays per step PER GPU,
64 coarse + 128 fine samples per ray, 8x256 MLPs with view branch, Glorot-initialised weights, rays from sphere cameras
(near/far 0.3333/2.0, fov 0.69111), targets U[0,1).  Weak scaling: every rank runs its own 2048-ray shard of a
global batch of N*2048 rays, with ONE NCCL all-reduce of the 4.1 MB gradient vector per step.

Prints ONE JSON line (rank 0).  `value` = rays/s with the ray batch resident in HBM, timed on the device with CUDA
events over exactly K steps (max over ranks); `e2e` = the same metric through the public API `NeRF.train_step` with
PINNED HOST batches (H2D copy of the batch and D2H read of the loss inside the timed region, every step).
`--impl reference` times the CPU oracle port of the reference's train step (TensorFlow is not installable here) on
all host cores, on a bounded sample of the same workload.
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


def cpu_reference_rate(cfg_name, n_rays, reps, threads=None):
    """rays/s of the oracle port of NeRF.train_step (forward + autograd backward + Adam) on the host cores."""
    import torch
    from oracle import nerf_oracle as O
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
    threads, each step = a bounded 128-ray sample of the 2048-ray batch."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_sample = 128
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
                         "sample": f"{n_sample} rays x {N_C + N_F} samples per step, oracle port (PyTorch-CPU fp32) of "
                                   "NeRF.train_step; TensorFlow unavailable offline"},
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
    ap.add_argument("--config", default="100px_robot_72pics_sphere", choices=sorted(CONFIGS))
    ap.add_argument("--mode", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--settle", type=float, default=1.0, help="idle seconds before each timed pass (power-cap state)")
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

    # per-call device times of the MLP kernels over a second timed pass (events on the launching stream)
    per_call = {}

    @contextlib.contextmanager
    def hook(name):
        if name in ("nerf_mlp_fwd", "nerf_mlp_fwd_rays", "nerf_mlp_bwd", "nerf_mlp_bwd_dx", "nerf_mlp_bwd_dw",
                    "nerf_composite_fwd", "nerf_composite_bwd"):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            yield
            b.record()
            per_call.setdefault(name, []).append((a, b))
        else:
            yield
    pkg._lib.event_hook = hook
    model.overlap_dw = False          # one stream: each kernel is timed alone (the step itself overlaps the fine dW)
    timed(step_device, args.steps)
    model.overlap_dw = True
    pkg._lib.event_hook = None
    torch.cuda.synchronize()
    call_ms = {k: sum(a.elapsed_time(b) for a, b in v) / len(v) for k, v in per_call.items()}
    call_n = {k: len(v) // args.steps for k, v in per_call.items()}

    run_e2e(8)        # warm-up: also brings the copy streams' allocator pools to their steady-state size
    losses.clear()
    ms_e2e, _, t_load1 = timed(run_e2e, args.steps, whole=True)
    assert len(losses) == args.steps and all(math.isfinite(v) for v in losses), "every step's loss must reach the host"
    # the device-timed region alone lasts ~0.1 s (one nvidia-smi sample); report the median over every sample taken
    # while the GPU ran back-to-back steps (warm-up, device-timed, per-call-timed and e2e passes)
    clocks = sampler.window(t_load0, t_load1) if rank == 0 else None
    if clocks is not None:
        clocks["window"] = ("warm-up through e2e pass; the GPU idles %.1f s before each timed pass so that every pass "
                            "starts from the same power-cap state" % args.settle)
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
        peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
        peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained" if "bf16_tflops_sustained" in peaks else \
            "fallback (B200_PROFILING.md: ~1.4 PFLOP/s sustained)"
        # Dominant kernel by time (profiles/r01_i_launches_summary.txt: 35 % of the step): mlp_tc_bwd_dw_kernel, the weight
        # gradients.  Its arithmetic intensity is fixed by the 256x256 output it keeps in TMEM (128 FLOP/B), so its
        # roofline is HBM: it streams every saved activation and every dZ once.  Algorithmic bytes per sample (DESIGN.md
        # 4): bf16 saved activations (64 + 8*256 + 128 columns) + bf16 dZ (8*256 + 144 + 16 columns) = 8896 B.  The
        # launch time is the nerf_mlp_bwd_dw call (dW kernel + its 15 us fixed-order reduce), averaged over the coarse
        # (64 samples/ray) and fine (128) calls of a step; `traffic` is dram read+write of the same two launches from
        # the ncu --set full capture (profiles/r01_i_mlp_full_summary.txt: 1.285 + 2.586 GB for 131072 + 262144 rows
        # = 9845 B/sample).
        samples_per_launch = batch * (N_C + N_F) / 2
        dw_ms = call_ms.get("nerf_mlp_bwd_dw", float("nan"))
        dx_ms = call_ms.get("nerf_mlp_bwd_dx", float("nan"))
        fwd_ms = call_ms.get("nerf_mlp_fwd_rays", call_ms.get("nerf_mlp_fwd", float("nan")))
        peak_hbm = peaks.get("hbm_gbs", 6500.0)
        fwd_flops = 2 * batch * (N_C + N_F) * MAC_FWD / 2
        dx_flops = 2 * batch * (N_C * MAC_DX_COARSE + N_F * MAC_DX_FINE) / 2
        dw_flops = fwd_flops
        ach = 8896 * samples_per_launch / (dw_ms * 1e-3) / 1e9
        tensor = {
            "peak": peak_tf, "peak_source": peak_src, "unit": "TFLOP/s",
            "mlp_tc_fwd_kernel<save>": {"achieved": fwd_flops / (fwd_ms * 1e-3) / 1e12, "ms": fwd_ms},
            "mlp_tc_bwd_chain_kernel": {"achieved": dx_flops / (dx_ms * 1e-3) / 1e12, "ms": dx_ms},
            "mlp_tc_bwd_dw_kernel": {"achieved": dw_flops / (dw_ms * 1e-3) / 1e12, "ms": dw_ms},
            "step_tensor_frac": FLOP_PER_RAY_TRAIN * value / 1e12 / peak_tf,
        }
        for k in ("mlp_tc_fwd_kernel<save>", "mlp_tc_bwd_chain_kernel", "mlp_tc_bwd_dw_kernel"):
            tensor[k]["frac"] = tensor[k]["achieved"] / peak_tf
        roofline = {"bound": "hbm", "kernel": "mlp_tc_bwd_dw_kernel (nerf_mlp_bwd_dw call)",
                    "achieved": ach, "peak": peak_hbm, "unit": "GB/s", "frac": ach / peak_hbm,
                    "traffic": 9845 * samples_per_launch,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6.5 TB/s",
                    "algorithmic_bytes_per_launch": 8896 * samples_per_launch,
                    "tensor_kernels": tensor,
                    "avg_call_ms": call_ms, "calls_per_step": call_n}
        if args.mode != "bf16":
            roofline["bound"] = "fp32 SIMT parity mode: the bf16 rooflines above do not apply"
        cpu = None
        if not args.no_cpu_baseline:
            rate, cores = cpu_reference_rate(args.config, 128, 3)
            cpu = {"value": rate, "unit": "rays/s", "cores": cores, "kind": "port",
                   "sample": "128-ray sample of the batch, 3 timed steps after 1 warm-up, oracle port (PyTorch-CPU fp32) "
                             "of NeRF.train_step incl. Adam; TensorFlow unavailable offline"}
        line = {
            "metric": "rays/sec render fwd+bwd (train step)", "value": value, "unit": "rays/s", "n_gpus": world,
            "steps": args.steps, "warmup": n_warm, "ms_per_step": ms_dev / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.mode == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": f"train_step {args.config}: {batch} rays/step/GPU, {N_C} coarse + {N_F} fine samples, "
                                   "8x256 MLPs + view branch, Adam"
                                   + (", + every 13th step the DietNeRF consistency term (150x150 in-tape render at "
                                      "55+55 samples, random-init ViT-B/32, cosine loss, backward)" if diet else ""),
                       "global_batch_rays": n_total, "parallelism": f"ray-sharded dp{world}",
                       "settle_s_before_each_timed_pass": args.settle,
                       "l2": "working set per step (saved activations + dZ, ~4.5 GB at 2048 rays) >> 126 MB L2; "
                             "4 distinct ray batches rotate"},
            "e2e": {"value": e2e, "unit": "rays/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": batch * (16 + 16 + 12), "d2h_bytes_per_step": 4,
                    "how": "DevicePrefetcher (pinned host batch -> device on a copy stream, one batch ahead) -> "
                           "train step -> loss to a pinned host slot (asynchronous read-back consumed two steps later); "
                           "every copy of every step inside the timed region"},
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
