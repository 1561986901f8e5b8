"""Per-CTA cycle counts of the weight-gradient kernel (NERF_TC_DEBUG=320: no chain launch, every dW CTA prints its cycles):
how evenly the (unit, row-range) split finishes.  Usage (GPU box): python tools/dw_balance_probe.py [--rows 524288]"""
import argparse
import importlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(rows):
    import torch
    sys.path.insert(0, ROOT)
    pkg = importlib.import_module("nerf-and-dietnerf_b200")
    call, ptr = pkg._lib.call, pkg._lib.ptr
    cfg = pkg.NetCfg(5, 4, 2, 256, 128, 0.05)
    net = pkg.NerfMLP(cfg, mode="fp16", seed=0)
    m = rows
    xyz = torch.randn(m, 33, device="cuda"); view = torch.randn(m, 24, device="cuda")
    saved = torch.zeros(max(net.saved_bytes(m), 16), dtype=torch.uint8, device="cuda")
    packed = net.packed_for(net.params)
    d_out = torch.randn(m, 4, device="cuda")
    grads = torch.zeros(net.n_params, device="cuda")
    d_xyz = torch.empty(m, 33, device="cuda")
    wsb = torch.zeros(max(net.workspace_bytes(m, True), 16), dtype=torch.uint8, device="cuda")
    for _ in range(3):          # the last of the three runs is the one that is read (warm, same clocks)
        print("---- run", flush=True)
        call("nerf_mlp_bwd", net.cfg_ref, ptr(net.params), ptr(packed), ptr(xyz), ptr(view), ptr(saved), ptr(d_out), m,
             ptr(grads), ptr(d_xyz), ptr(wsb), net.mode_id)
        torch.cuda.synchronize()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=524288)
    ap.add_argument("--child", action="store_true")
    args = ap.parse_args()
    if args.child:
        child(args.rows)
        return
    env = dict(os.environ, NERF_TC_DEBUG="320")
    out = subprocess.run([sys.executable, __file__, "--child", "--rows", str(args.rows)], env=env, capture_output=True, text=True)
    if out.returncode != 0:
        print(out.stdout[-2000:], out.stderr[-4000:])
        sys.exit(1)
    last = out.stdout.split("---- run")[-1]
    rows = [(int(a), int(b), int(c), int(d), int(e), int(f)) for a, b, c, d, e, f in
            re.findall(r"dw cta\s+(\d+) unit\s+(\d+) \(dense (\d+) kind (\d+)\) tiles\s+(\d+)\s+cycles\s+(\d+)", last)]
    if not rows:
        print(out.stdout[-2000:])
        sys.exit(1)
    mx = max(r[5] for r in rows)
    print(f"rows {args.rows}: {len(rows)} CTAs, slowest {mx} cycles, mean {sum(r[5] for r in rows) / len(rows):.0f} "
          f"({100.0 * sum(r[5] for r in rows) / len(rows) / mx:.1f} % of the slowest)")
    units = sorted(set(r[1] for r in rows))
    print("unit dense kind ctas tiles/cta   min cycles   max cycles   max/slowest")
    for u in units:
        rr = [r for r in rows if r[1] == u]
        cyc = [r[5] for r in rr]
        print(f"{u:4d} {rr[0][2]:5d} {rr[0][3]:4d} {len(rr):4d} {rr[0][4]:9d} {min(cyc):12d} {max(cyc):12d} {max(cyc) / mx:10.3f}")


if __name__ == "__main__":
    main()
