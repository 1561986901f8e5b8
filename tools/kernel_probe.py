"""Per-kernel timing probe (CUDA events, L2-exceeding inputs): MLP forward/backward TFLOP/s and compositing GB/s.
Usage (on the GPU box): python tools/kernel_probe.py [--rays 4096] [--mode bf16]"""
import argparse
import ctypes
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("nerf-and-dietnerf_b200")

MAC_FWD = 512152  # per sample, SURVEY 8d


def timeit(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=4096)
    ap.add_argument("--mode", default="bf16")
    ap.add_argument("--bwd", action="store_true")
    ap.add_argument("--only", default="all", choices=["all", "mlp", "composite", "sampler"])
    args = ap.parse_args()
    call, ptr = pkg._lib.call, pkg._lib.ptr
    cfg = pkg.NetCfg(5, 4, 2, 256, 128, 0.05)
    net = pkg.NerfMLP(cfg, mode=args.mode, seed=0)
    if args.only in ("all", "sampler"):
        for n in (2048, 65536):
            sc, nf = 64, 128
            w = torch.rand(n, sc, device="cuda") ** 4
            z = torch.sort(torch.rand(n, sc, device="cuda") * 2 + 0.5, -1).values.contiguous()
            z_new = torch.empty(n, nf, device="cuda"); u = torch.empty(n, nf, device="cuda")
            perm = torch.empty(n, nf, dtype=torch.int32, device="cuda")
            d_z = torch.randn(n, nf, device="cuda"); d_w = torch.empty(n, sc, device="cuda")
            f = lambda: call("nerf_sample_pdf_fwd", ptr(w), ptr(z), n, sc, nf, None, 1, 0, 0, ptr(z_new), None, ptr(perm), ptr(u))
            med, best = timeit(f)
            byts = n * (2 * 4 * sc + 3 * 4 * nf)          # w, z in; z_new, perm, u out
            print(f"sample_pdf_fwd      N={n:6d} S={sc} Nf={nf}: {med * 1e3:8.1f} us  {byts / med / 1e6:8.1f} GB/s")
            f = lambda: call("nerf_sample_pdf_bwd", ptr(w), ptr(z), ptr(u), ptr(perm), ptr(d_z), n, sc, nf, ptr(d_w))
            med, best = timeit(f)
            byts = n * (3 * 4 * sc + 3 * 4 * nf)
            print(f"sample_pdf_bwd      N={n:6d} S={sc} Nf={nf}: {med * 1e3:8.1f} us  {byts / med / 1e6:8.1f} GB/s")
    if args.only in ("all", "sampler"):
        # the step between the two networks of a render (64 coarse samples; 128 draws = the reference's render setting, 192):
        # three launches against one
        for nf in (128, 192):
            n, sc = 65536, 64
            raw = torch.randn(n, sc, 4, device="cuda"); raw[..., 3] *= 10
            z = torch.sort(torch.rand(n, sc, device="cuda") * 2 + 0.5, -1).values.contiguous()
            w = torch.empty(n, sc, device="cuda"); z_new = torch.empty(n, nf, device="cuda")
            z_all = torch.empty(n, sc + nf, device="cuda"); z_one = torch.empty(n, sc + nf, device="cuda")

            def three():
                call("nerf_composite_fwd", ptr(raw), ptr(z), n, sc, None, ptr(w), None, None, None, None, None)
                call("nerf_sample_pdf_fwd", ptr(w), ptr(z), n, sc, nf, None, 1, 0, 0, ptr(z_new), None, None, None)
                call("nerf_merge_sorted", ptr(z_new), nf, ptr(z), sc, n, ptr(z_all))
            one = lambda: call("nerf_hierarchical_sample", ptr(raw), ptr(z), n, sc, nf, 1, 0, 0, ptr(z_one))
            m3, _ = timeit(three)
            m1, _ = timeit(one)
            print(f"weights + sampler + merge, three launches N={n} S={sc} Nf={nf}: {m3 * 1e3:8.1f} us")
            print(f"nerf_hierarchical_sample,  one launch     N={n} S={sc} Nf={nf}: {m1 * 1e3:8.1f} us  equal={torch.equal(z_all, z_one)}")
    for s in ((64, 128, 192) if args.only in ("all", "mlp") else ()):
        m = args.rays * s
        xyz = torch.randn(m, 33, device="cuda")
        view = torch.randn(m, 24, device="cuda")
        out = torch.empty(m, 4, device="cuda")
        ws = torch.empty(max(net.workspace_bytes(m, False), 16), dtype=torch.uint8, device="cuda")
        saved = torch.empty(max(net.saved_bytes(m), 16), dtype=torch.uint8, device="cuda")
        packed = net.packed_for(net.params)
        for label, sv in (("infer", None), ("train(save acts)", saved)):
            f = lambda: call("nerf_mlp_fwd", net.cfg_ref, ptr(net.params), ptr(packed), ptr(xyz), ptr(view), m, ptr(out),
                             ptr(sv), ptr(ws), net.mode_id)
            med, best = timeit(f)
            print(f"mlp_fwd[{args.mode}] {label:18s} M={m:8d}: {med:8.3f} ms  {2 * MAC_FWD * m / med / 1e9:8.1f} TFLOP/s "
                  f"(best {2 * MAC_FWD * m / best / 1e9:.1f})")
        # the fused entry point NeRF.render / train_step use: rays + depths in, encodings computed in the prologue
        o4 = torch.randn(args.rays, 4, device="cuda"); d4 = torch.randn(args.rays, 4, device="cuda")
        zz = torch.sort(torch.rand(args.rays, s, device="cuda") * 2 + 0.5, -1).values.contiguous()
        for label, sv in (("rays infer", None), ("rays train", saved)):
            f = lambda: call("nerf_mlp_fwd_rays", net.cfg_ref, ptr(packed), ptr(o4), ptr(d4), ptr(zz), args.rays, s, ptr(out),
                             ptr(sv), net.mode_id)
            med, best = timeit(f)
            print(f"mlp_fwd[{args.mode}] {label:18s} M={m:8d}: {med:8.3f} ms  {2 * MAC_FWD * m / med / 1e9:8.1f} TFLOP/s "
                  f"(best {2 * MAC_FWD * m / best / 1e9:.1f})")
        if args.bwd:
            d_out = torch.randn(m, 4, device="cuda")
            grads = torch.zeros(net.n_params, device="cuda")
            d_xyz = torch.empty(m, 33, device="cuda")
            wsb = torch.empty(max(net.workspace_bytes(m, True), 16), dtype=torch.uint8, device="cuda")
            f = lambda: call("nerf_mlp_bwd", net.cfg_ref, ptr(net.params), ptr(packed), ptr(xyz), ptr(view), ptr(saved),
                             ptr(d_out), m, ptr(grads), ptr(d_xyz), ptr(wsb), net.mode_id)
            med, best = timeit(f)
            flops = 2 * (512152 + 509056) * m
            print(f"mlp_bwd[{args.mode}]                    M={m:8d}: {med:8.3f} ms  {flops / med / 1e9:8.1f} TFLOP/s")
    # compositing: one 256x256 frame
    for n, s in (((65536, 192), (65536, 64), (4096 * 16, 128)) if args.only in ("all", "composite") else ()):
        raw = torch.randn(n, s, 4, device="cuda")
        z = torch.sort(torch.rand(n, s, device="cuda"), -1).values
        rgb = torch.empty(n, 3, device="cuda"); w = torch.empty(n, s, device="cuda"); T = torch.empty(n, s, device="cuda")
        al = torch.empty(n, s, device="cuda"); rs = torch.empty(n, s, 3, device="cuda")
        dep = torch.empty(n, device="cuda"); acc = torch.empty(n, device="cuda")
        f = lambda: call("nerf_composite_fwd", ptr(raw), ptr(z), n, s, ptr(rgb), ptr(w), ptr(T), ptr(al), ptr(rs), None, None)
        med, best = timeit(f)
        byts = n * (s * 44 + 12)
        print(f"composite_fwd full  N={n} S={s}: {med * 1e3:8.1f} us  {byts / med / 1e6:8.1f} GB/s")
        f = lambda: call("nerf_composite_fwd", ptr(raw), ptr(z), n, s, ptr(rgb), ptr(w), None, None, None, ptr(dep), ptr(acc))
        med, best = timeit(f)
        byts = n * (s * 24 + 20)
        print(f"composite_fwd lean  N={n} S={s}: {med * 1e3:8.1f} us  {byts / med / 1e6:8.1f} GB/s")
        d_rgb = torch.randn(n, 3, device="cuda"); d_raw = torch.empty_like(raw); d_z = torch.empty_like(z)
        f = lambda: call("nerf_composite_bwd", ptr(raw), ptr(z), ptr(d_rgb), ptr(w), n, s, ptr(d_raw), ptr(d_z))
        med, best = timeit(f)
        byts = n * (s * 44 + 12)
        print(f"composite_bwd       N={n} S={s}: {med * 1e3:8.1f} us  {byts / med / 1e6:8.1f} GB/s")


if __name__ == "__main__":
    main()
