"""How much host time does one train step take (ctypes calls, torch bookkeeping) next to its GPU time?
python tools/host_overhead.py [--profile]"""
import argparse, cProfile, importlib, os, pstats, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B

ap = argparse.ArgumentParser(); ap.add_argument("--profile", action="store_true"); args = ap.parse_args()
pkg = importlib.import_module("nerf-and-dietnerf_b200"); pkg.load()
batch, near, far, fov = B.CONFIGS["100px_robot_72pics_sphere"]
model = pkg.NeRFModel(B.net_config(batch), {"n_render_samples_coarse": 64, "n_render_samples_fine": 128}, near, far, mode="bf16", seed=0)
model.compile(optimizer=pkg.Adam(5e-4))
def gpu_rays(c2w, fov_, h, w):
    dirs, orig = pkg.UtilsCV.get_rays_directions(h, w, fov_, c2w, return_origins=True)
    return orig.cpu(), dirs.reshape(-1, 4).cpu()
o, d, y = (t.cuda() for t in B.synthetic_batch(batch, fov, 0, gpu_rays))
po, pd, py = (t.cpu().pin_memory() for t in (o, d, y))
for _ in range(5): model.train_step_local(o, d, y, batch, 0)
torch.cuda.synchronize()
K = 50
t0 = time.perf_counter()
for _ in range(K): model.train_step_local(o, d, y, batch, 0)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"device-resident: host {1e3*(t1-t0)/K:.3f} ms/step issue time, total {1e3*(t2-t0)/K:.3f} ms/step")
t0 = time.perf_counter()
for _ in range(K):
    m = model.train_step((po, pd, py))
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"host batches   : host {1e3*(t1-t0)/K:.3f} ms/step issue time, total {1e3*(t2-t0)/K:.3f} ms/step")
if args.profile:
    pr = cProfile.Profile(); pr.enable()
    for _ in range(K): model.train_step((po, pd, py))
    pr.disable(); torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(25)

# ---- variants of the end-to-end feed ------------------------------------------------------------------------------------
U = pkg.UtilsNeuralRadianceField
def timed(label, fn):
    fn(5); torch.cuda.synchronize()
    t0 = time.perf_counter(); fn(K); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"{label:50s}: {1e3*(t2-t0)/K:.3f} ms/step")
timed("train_step(pinned host tuple)", lambda k: [model.train_step((po, pd, py)) for _ in range(k)])
def pf(k, wait=True):
    f = U.DevicePrefetcher((po, pd, py) for _ in range(k))
    for b in f:
        model.train_step_local(*b, batch, 0)
timed("DevicePrefetcher -> train_step_local", pf)
def pf_loss(k):
    slot = torch.zeros(1).pin_memory(); d2h = torch.cuda.Stream()
    f = U.DevicePrefetcher((po, pd, py) for _ in range(k))
    for b in f:
        m = model.train_step_local(*b, batch, 0)
        ev = torch.cuda.Event(); ev.record(); d2h.wait_event(ev)
        with torch.cuda.stream(d2h):
            slot.copy_(m["loss"].reshape(1), non_blocking=True)
        m["loss"].record_stream(d2h)
timed("DevicePrefetcher + loss D2H on a copy stream", pf_loss)
def same_stream_loss(k):
    slot = torch.zeros(1).pin_memory()
    for _ in range(k):
        m = model.train_step((po, pd, py))
        slot.copy_(m["loss"].reshape(1), non_blocking=True)
timed("train_step(pinned) + loss D2H same stream", same_stream_loss)
def one_buf(k):
    big = torch.cat([po, pd, py], 1).pin_memory()
    for _ in range(k):
        dv = big.cuda(non_blocking=True)
        model.train_step_local(dv[:, 0:4].contiguous(), dv[:, 4:8].contiguous(), dv[:, 8:11].contiguous(), batch, 0)
timed("one (N,11) pinned buffer, 1 H2D + 3 slices", one_buf)
