"""Probe of the layer-pipelined backward kernel, one stage at a time (nerf_debug_bwd_pipe_layer): the stage's dZ_l against
the dZ_l the per-tile chain kernel wrote (bit for bit), its dW_l / db_l against the stand-alone dW kernel's, and its time.
Usage (GPU box): python tools/bwd_pipe_probe.py [--rays 4096] [--samples 64] [--layers 7,4,1,8]"""
import argparse
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("nerf-and-dietnerf_b200")

from tcm_layout import DZ_BLOCKS, DZ_TILE_BYTES, SAVED_BLOCKS, SAVED_TILE_BYTES, to_tcm  # noqa: E402

SHAPES = [(33, 256)] + [(256, 256)] * 3 + [(289, 256)] + [(256, 256)] * 3 + [(280, 128), (128, 3), (280, 1)]


def dense_slices():
    out, off = [], 0
    for i, o in SHAPES:
        out.append((off, off + i * o, off + i * o + o))
        off += i * o + o
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=4096)
    ap.add_argument("--samples", type=int, default=64)
    ap.add_argument("--layers", default="7,4,1,8")
    ap.add_argument("--iters", type=int, default=8)
    args = ap.parse_args()
    call, ptr = pkg._lib.call, pkg._lib.ptr
    cfg = pkg.NetCfg(5, 4, 2, 256, 128, 0.05)
    net = pkg.NerfMLP(cfg, mode="bf16", seed=0)
    n, s = args.rays, args.samples
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
    grads_ref = torch.zeros(net.n_params, device="cuda")
    d_xyz = torch.empty(m, 33, device="cuda")
    call("nerf_mlp_bwd", net.cfg_ref, ptr(net.params), ptr(packed), None, None, ptr(saved), ptr(d_out), m, ptr(grads_ref),
         ptr(d_xyz), ptr(ws), net.mode_id)
    torch.cuda.synchronize()
    tiles4 = ((m + 127) // 128 + 3) // 4 * 4
    base = (-ws.data_ptr()) % 1024
    region = slice(base, base + tiles4 * DZ_TILE_BYTES)
    # the stage reads and writes tile chunk-major (TCM) blocks: re-lay the chain kernel's RBCM output and the saved activations
    ws_tcm = ws.clone()
    ws_tcm[region] = to_tcm(ws[region], tiles4, DZ_TILE_BYTES, DZ_BLOCKS)
    saved_tcm = saved.clone()
    saved_tcm[:tiles4 * SAVED_TILE_BYTES] = to_tcm(saved[:tiles4 * SAVED_TILE_BYTES], tiles4, SAVED_TILE_BYTES, SAVED_BLOCKS)
    ws, saved = ws_tcm, saved_tcm
    sl = dense_slices()
    for layer in [int(x) for x in args.layers.split(",")]:
        ws2 = ws.clone()
        # same alignment offset as the original (clone keeps 256-byte alignment classes in practice; assert it)
        assert (-ws2.data_ptr()) % 1024 == base, "clone changed the 1024-byte phase of the workspace"
        hi, lo = (layer // 100, layer % 100) if layer >= 100 else (layer, layer)
        if 1 <= layer <= 7 or layer >= 100:  # wipe the dZ blocks the launch produces so that equality proves it wrote them
            v = ws2[region].view(tiles4, DZ_TILE_BYTES)
            v[:, (lo - 1) * 65536:hi * 65536] = 0x7f
        grads = torch.zeros(net.n_params, device="cuda")
        call("nerf_debug_bwd_pipe_layer", net.cfg_ref, ptr(packed), ptr(saved), m, ptr(ws2), layer, ptr(grads))
        torch.cuda.synchronize()
        same = torch.equal(ws2[region], ws[region])
        nbad = (ws2[region] != ws[region]).sum().item() if not same else 0
        if layer >= 100:
            parts = []
            for l in range(lo, hi + 1):
                w0, w1, b1 = sl[l]
                skip = 33 * 256 if l == 4 else 0
                parts.append((grads[w0 + skip:w1], grads[w1:b1], grads_ref[w0 + skip:w1], grads_ref[w1:b1]))
            gw, gb, rw, rb = (torch.cat([p[i] for p in parts]) for i in range(4))
        else:
            w0, w1, b1 = sl[layer]
            gw, gb = grads[w0:w1], grads[w1:b1]
            rw, rb = grads_ref[w0:w1], grads_ref[w1:b1]
        if layer == 4:      # the stage covers the h4 rows of Dense 4 (rows 33..288)
            gw, rw = gw[33 * 256:], rw[33 * 256:]
        if layer == 8:      # h8 rows of Dense 8; the sigma head's h8 rows are checked separately
            gw, rw = gw[:256 * 128], rw[:256 * 128]
            s0, s1, sb = sl[10]
            sig = ((grads[s0:s0 + 256] - grads_ref[s0:s0 + 256]).norm() / grads_ref[s0:s0 + 256].norm()).item()
            sigb = abs((grads[s1] - grads_ref[s1]).item()) / abs(grads_ref[s1].item())
        rel_w = ((gw - rw).norm() / rw.norm()).item()
        rel_b = ((gb - rb).norm() / rb.norm()).item()
        extra = f" sigma w {sig:.2e} b {sigb:.2e}" if layer == 8 else ""
        # timing
        for _ in range(2):
            call("nerf_debug_bwd_pipe_layer", net.cfg_ref, ptr(packed), ptr(saved), m, ptr(ws2), layer, ptr(grads))
        torch.cuda.synchronize()
        evs = []
        for _ in range(args.iters):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            call("nerf_debug_bwd_pipe_layer", net.cfg_ref, ptr(packed), ptr(saved), m, ptr(ws2), layer, ptr(grads))
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        ts = sorted(a.elapsed_time(b) for a, b in evs)
        med = ts[len(ts) // 2]
        flops = (4 * (hi - lo + 1) if layer != 8 else 2 * 144 / 256) * 256 * 256 * m
        print(f"layer {layer} M={m}: dZ region equal {same} (bad bytes {nbad})  dW rel {rel_w:.2e}  db rel {rel_b:.2e}{extra}  "
              f"{med * 1e3:7.1f} us  {flops / med / 1e9:7.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    main()
