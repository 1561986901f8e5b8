"""Probe of the overlapped MLP backward (nerf_mlp_bwd_overlapped): chain + dW at the same time on disjoint SMs against the
sequential nerf_mlp_bwd, for a sweep of the SM split (env NERF_BWD_CHAIN_PAIRS) and stagger (NERF_BWD_STAGGER_NS).
Usage (GPU box): python tools/bwd_overlap_probe.py [--rays 4096] [--pairs 40,44,48,52,56] [--stagger 0,2500]"""
import argparse
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("nerf-and-dietnerf_b200")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=4096)
    ap.add_argument("--samples", default="64,128")
    ap.add_argument("--pairs", default="40,44,48,52,56")
    ap.add_argument("--stagger", default="0,2500")
    ap.add_argument("--iters", type=int, default=8)
    args = ap.parse_args()
    call, ptr = pkg._lib.call, pkg._lib.ptr
    cfg = pkg.NetCfg(5, 4, 2, 256, 128, 0.05)
    net = pkg.NerfMLP(cfg, mode="bf16", seed=0)
    os.environ["NERF_BWD_OVERLAP"] = "1"
    side = torch.cuda.Stream()
    main_s = torch.cuda.current_stream()
    for s in [int(x) for x in args.samples.split(",")]:
        n = args.rays
        m = n * s
        g = torch.Generator(device="cuda").manual_seed(5)
        o4 = torch.randn(n, 4, device="cuda", generator=g)
        d4 = torch.randn(n, 4, device="cuda", generator=g)
        z = torch.sort(torch.rand(n, s, device="cuda", generator=g) * 2 + 0.5, -1).values.contiguous()
        d_out = torch.randn(m, 4, device="cuda", generator=g)
        out = torch.empty(m, 4, device="cuda")
        packed = net.packed_for(net.params)
        saved = torch.empty(net.saved_bytes(m), dtype=torch.uint8, device="cuda")
        ws = torch.empty(net.workspace_bytes(m, True), dtype=torch.uint8, device="cuda")
        call("nerf_mlp_fwd_rays", net.cfg_ref, ptr(packed), ptr(o4), ptr(d4), ptr(z), n, s, ptr(out), ptr(saved), net.mode_id)
        grads = torch.zeros(net.n_params, device="cuda")
        d_xyz = torch.empty(m, 33, device="cuda")
        flops = 2 * (512152 + 509056) * m

        def run(overlap):
            grads.zero_()
            args_ = (net.cfg_ref, ptr(net.params), ptr(packed), None, None, ptr(saved), ptr(d_out), m, ptr(grads),
                     ptr(d_xyz), ptr(ws), net.mode_id)
            if overlap:
                call("nerf_mlp_bwd_overlapped", *args_, side.cuda_stream)
                main_s.wait_stream(side)
            else:
                call("nerf_mlp_bwd", *args_)

        def timed(overlap):
            for _ in range(2):
                run(overlap)
            torch.cuda.synchronize()
            evs = []
            for _ in range(args.iters):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                run(overlap)
                b.record()
                evs.append((a, b))
            torch.cuda.synchronize()
            ts = sorted(a.elapsed_time(b) for a, b in evs)
            return ts[len(ts) // 2], ts[0]

        med, best = timed(False)
        ref_g, ref_x = grads.clone(), d_xyz.clone()
        print(f"M={m:8d} sequential              : {med:7.3f} ms (best {best:.3f})  {flops / med / 1e9:7.1f} TFLOP/s", flush=True)
        for st in [int(x) for x in args.stagger.split(",")]:
            for pairs in [int(x) for x in args.pairs.split(",")]:
                os.environ["NERF_BWD_CHAIN_PAIRS"] = str(pairs)
                os.environ["NERF_BWD_STAGGER_NS"] = str(st)
                med, best = timed(True)
                rel = ((grads - ref_g).norm() / ref_g.norm()).item()
                same_x = torch.equal(d_xyz, ref_x)
                print(f"M={m:8d} overlapped pairs={pairs:2d} stagger={st:5d}: {med:7.3f} ms (best {best:.3f})  "
                      f"{flops / med / 1e9:7.1f} TFLOP/s  grads rel diff {rel:.2e}  d_xyz equal {same_x}", flush=True)


if __name__ == "__main__":
    main()
