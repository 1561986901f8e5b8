"""Where a ray-sharded train step spends its time on every rank (torch.profiler / CUPTI on each rank of a torchrun launch):
step span, the wait inside the two peer barriers (= how long this rank was ahead of the slowest one), the one-shot reduce +
Adam kernels and the exposed tail after the last MLP kernel.  Rank 0 also prints its full kernel timeline of one step.
Run: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/timeline_sharded.py [--rays 4096]"""
import argparse
import importlib
import json
import os
import sys
import tempfile

import torch
import torch.distributed as dist
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=4096, help="rays per GPU")
    ap.add_argument("--steps", type=int, default=8)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = importlib.import_module("nerf-and-dietnerf_b200")
    pkg.load()
    per, n_total = args.rays, args.rays * world
    near, far, fov = 0.5576, 2.5635, 0.46134

    def rays(c2w, fov_, h, w):
        dirs, orig = pkg.UtilsCV.get_rays_directions(h, w, fov_, c2w, return_origins=True)
        return orig.cpu(), dirs.reshape(-1, 4).cpu()
    batches = [tuple(t.cuda() for t in B.synthetic_batch(per, fov, 1000 * b + rank, rays)) for b in range(3)]
    model = pkg.NeRFModel(B.net_config(per), {"n_render_samples_coarse": 64, "n_render_samples_fine": 128}, near, far, seed=0)
    model.compile(optimizer=pkg.Adam(5e-4))
    model.distribute()

    def run(k):
        for i in range(k):
            model.train_step_local(*batches[i % 3], n_total, rank * per)
    run(30)                                    # into the sustained power state
    dist.barrier()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        run(args.steps)
        torch.cuda.synchronize()
    path = os.path.join(tempfile.mkdtemp(), f"trace{rank}.json")
    prof.export_chrome_trace(path)
    ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
    ev.sort(key=lambda e: e["ts"])
    ends = [e["ts"] + e["dur"] for e in ev if "train_metrics" in e["name"]]
    spans = [b - a for a, b in zip(ends, ends[1:])]
    def total(pattern, stream_filter=None):
        return sum(e["dur"] for e in ev if pattern in e["name"]) / args.steps
    mlp = sum(e["dur"] for e in ev if "mlp_tc" in e["name"]) / args.steps
    summary = {"rank": rank, "step_us": sum(spans) / max(len(spans), 1), "mlp_kernels_us": mlp,
               "barrier_wait_us": total("peer_barrier"), "reduce_adam_us": total("peer_reduce_adam"),
               "other_kernels_us": sum(e["dur"] for e in ev if "mlp_tc" not in e["name"] and "peer_" not in e["name"]) / args.steps}
    out = [None] * world
    dist.all_gather_object(out, summary)
    if rank == 0:
        print(f"{world} GPUs x {per} rays, {args.steps} profiled steps per rank (sustained state); per step, microseconds:")
        print(f"{'rank':>4} {'step':>9} {'MLP kernels':>12} {'other':>8} {'barrier wait':>13} {'reduce+Adam':>12}")
        for s in out:
            print(f"{s['rank']:4d} {s['step_us']:9.1f} {s['mlp_kernels_us']:12.1f} {s['other_kernels_us']:8.1f} "
                  f"{s['barrier_wait_us']:13.1f} {s['reduce_adam_us']:12.1f}")
        # one steady-state step of rank 0, kernel by kernel
        if len(ends) > 3:
            lo, hi = ends[2], ends[3]
            print("\nrank 0, one step (start relative to the previous step's metrics kernel, gap to the previous kernel's end on any stream):")
            prev_end = lo
            for e in ev:
                if lo < e["ts"] + e["dur"] <= hi + 1:
                    print(f"{e['ts'] - lo:9.1f} {e['ts'] - prev_end:7.1f} {e['dur']:8.1f}  s{e['args'].get('stream', '?'):<3} {e['name'][:60]}")
                    prev_end = max(prev_end, e["ts"] + e["dur"])
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
