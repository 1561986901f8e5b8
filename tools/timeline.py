"""GPU timeline of a few train steps (torch.profiler / CUPTI): every kernel and memcpy with its start, duration and the
idle gap before it.  python tools/timeline.py [--feed device|pinned|prefetch] [--steps 3]"""
import argparse, importlib, json, os, sys, tempfile
import torch
from torch.profiler import ProfilerActivity, profile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B

ap = argparse.ArgumentParser()
ap.add_argument("--feed", default="device"); ap.add_argument("--steps", type=int, default=3)
args = ap.parse_args()
pkg = importlib.import_module("nerf-and-dietnerf_b200"); pkg.load()
batch, near, far, fov = B.CONFIGS["100px_robot_72pics_sphere"]
model = pkg.NeRFModel(B.net_config(batch), {"n_render_samples_coarse": 64, "n_render_samples_fine": 128}, near, far, mode="bf16", seed=0)
model.compile(optimizer=pkg.Adam(5e-4))
def gpu_rays(c2w, fov_, h, w):
    dirs, orig = pkg.UtilsCV.get_rays_directions(h, w, fov_, c2w, return_origins=True)
    return orig.cpu(), dirs.reshape(-1, 4).cpu()
o, d, y = (t.cuda() for t in B.synthetic_batch(batch, fov, 0, gpu_rays))
po, pd, py = (t.cpu().pin_memory() for t in (o, d, y))
def run(k):
    if args.feed == "device":
        for _ in range(k): model.train_step_local(o, d, y, batch, 0)
    elif args.feed == "pinned":
        for _ in range(k): model.train_step((po, pd, py))
    elif args.feed == "bench_e2e":          # bench.py's e2e loop: prefetcher + loss to a pinned slot, read one step later
        slots = [torch.zeros(1).pin_memory() for _ in range(2)]; evs = [torch.cuda.Event() for _ in range(2)]
        d2h = torch.cuda.Stream(); got = []
        def read(i):
            evs[i % 2].synchronize(); got.append(float(slots[i % 2][0]))
        for i, b in enumerate(pkg.UtilsNeuralRadianceField.DevicePrefetcher((po, pd, py) for _ in range(k))):
            m = model.train_step_local(*b, batch, 0)
            done = torch.cuda.Event(); done.record(); d2h.wait_event(done)
            with torch.cuda.stream(d2h):
                slots[i % 2].copy_(m["loss"].reshape(1), non_blocking=True); evs[i % 2].record(d2h)
            m["loss"].record_stream(d2h)
            if i > 0: read(i - 1)
        read(k - 1)
    elif args.feed == "bench_e2e_old":      # previous bench loop: copies and loss read on the compute stream
        slots = [torch.zeros(1).pin_memory() for _ in range(2)]; evs = [torch.cuda.Event() for _ in range(2)]; got = []
        def read(i):
            evs[i % 2].synchronize(); got.append(float(slots[i % 2][0]))
        for i in range(k):
            od, dd, yd = (t.cuda(non_blocking=True) for t in (po, pd, py))
            m = model.train_step_local(od, dd, yd, batch, 0)
            slots[i % 2].copy_(m["loss"].reshape(1), non_blocking=True); evs[i % 2].record()
            if i > 0: read(i - 1)
        read(k - 1)
    elif args.feed == "pinned_loss":
        slot = torch.zeros(1).pin_memory()
        for _ in range(k):
            m = model.train_step((po, pd, py))
            slot.copy_(m["loss"].reshape(1), non_blocking=True)
    else:
        for b in pkg.UtilsNeuralRadianceField.DevicePrefetcher((po, pd, py) for _ in range(k)):
            model.train_step_local(*b, batch, 0)
run(5); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    run(args.steps); torch.cuda.synchronize()
path = os.path.join(tempfile.mkdtemp(), "trace.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]; prev_end = t0; busy = 0.0
print(f"{'start us':>10} {'gap':>7} {'dur':>8}  stream  name")
for e in ev:
    gap = e["ts"] - prev_end
    print(f"{e['ts'] - t0:10.1f} {gap:7.1f} {e['dur']:8.1f}  {e['args'].get('stream', '?'):>6}  {e['name'][:70]}")
    prev_end = max(prev_end, e["ts"] + e["dur"]); busy += e["dur"]
print(f"span {prev_end - t0:.1f} us for {args.steps} steps = {(prev_end - t0) / args.steps:.1f} us/step")
ends = [e["ts"] + e["dur"] for e in ev if "train_metrics" in e["name"]]
if len(ends) > 2:
    per = [b - a for a, b in zip(ends, ends[1:])]
    print("steady-state step spans (metrics kernel to metrics kernel):", " ".join(f"{p:.0f}" for p in per))
