"""Where does a DietNeRF consistency step spend its time?  GPU timeline (torch.profiler) of calc_consistency_loss,
aggregated by kernel family."""
import collections, importlib, json, os, sys, tempfile, time
import numpy as np, torch
from torch.profiler import ProfilerActivity, profile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B
pkg = importlib.import_module("nerf-and-dietnerf_b200"); pkg.load()
batch, near, far, fov = B.CONFIGS["256px_alexander_71pics_sphere_dietnerf"]
targets = torch.rand(8, 64, 64, 3, generator=torch.Generator().manual_seed(0)).numpy()
poses = np.stack([np.eye(4, dtype=np.float32) for _ in range(8)])
m = pkg.DietNeRFModel(B.net_config(batch), {"n_render_samples_coarse": 64, "n_render_samples_fine": 128}, near, far, targets,
                      poses, fov, -1, np.zeros(3), np.eye(4), mode="bf16", seed=0, numpy_seed=0, resample_every_call=True)
m.compile(optimizer=pkg.Adam(5e-4))
for _ in range(2):
    m._grad_buffer().zero_(); m.calc_consistency_loss()
torch.cuda.synchronize()
t0 = time.perf_counter(); m._grad_buffer().zero_(); m.calc_consistency_loss(); torch.cuda.synchronize()
print(f"calc_consistency_loss wall: {1e3 * (time.perf_counter() - t0):.2f} ms")
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    m._grad_buffer().zero_(); m.calc_consistency_loss(); torch.cuda.synchronize()
path = os.path.join(tempfile.mkdtemp(), "trace.json"); prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
agg = collections.OrderedDict()
for e in ev:
    name = e["name"]
    fam = "nerf::" + name.split("nerf::")[1].split("(")[0].split("<")[0] if "nerf::" in name else "torch/library (ViT, resize, misc)"
    a = agg.setdefault(fam, [0, 0.0]); a[0] += 1; a[1] += e["dur"]
span = ev[-1]["ts"] + ev[-1]["dur"] - ev[0]["ts"]
busy = sum(a[1] for a in agg.values())
print(f"GPU span {span / 1e3:.2f} ms, kernel time {busy / 1e3:.2f} ms, {len(ev)} launches")
for fam, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {fam:45s} n={n:5d} {t / 1e3:8.3f} ms")
# idle gaps by phase: first nerf kernel .. first torch kernel etc.
