"""Diagnostic: how well-conditioned is the coarse network's gradient on the render path (it only flows through the
importance sampler) under bf16 operand rounding?  Compares oracle fp32 / oracle bf16-emulated / CUDA fp32 / CUDA bf16."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import FAR, NEAR, net_config, oracle_cfg, random_rays, render_config, make_params
from oracle import nerf_oracle as O
pkg = importlib.import_module("nerf-and-dietnerf_b200"); pkg.load()

def run(n, n_c, n_f, gain, seed_rays=11):
    ocfg = oracle_cfg()
    pc, pf = make_params(ocfg, 1, gain), make_params(ocfg, 2, gain)
    o, d = random_rays(n, seed_rays)
    d_rgb = torch.randn(n, 3, generator=torch.Generator().manual_seed(2)) / n
    jit = O.stratified_jitter(21, 3, n, n_c, ray_offset=64)
    u = O.importance_uniforms(21, 3, n, n_f, ray_offset=64)
    res = {}
    for emu in (False, True):
        pco, pfo = pc.clone().requires_grad_(True), pf.clone().requires_grad_(True)
        rgb = O.render(pco, pfo, ocfg, NEAR, FAR, o, d, n_c, n_f, jit, u, emulate_bf16=emu)[0]
        (rgb * d_rgb).sum().backward()
        res["oracle_bf16" if emu else "oracle_fp32"] = (pco.grad.clone(), pfo.grad.clone())
    for mode in ("fp32", "bf16"):
        m = pkg.NeRFModel(net_config(), render_config(n_c, n_f), NEAR, FAR, mode=mode, seed=7)
        m.model_coarse.set_params(pc); m.model_fine.set_params(pf)
        m._grad_buffer().zero_()
        m.render_backward(o.cuda(), d.cuda(), d_rgb.cuda(), n_c, n_f, seed=21, step=3, ray_offset=64)
        torch.cuda.synchronize()
        _, gc, gf = m._grad_views()
        res["cuda_" + mode] = (gc.cpu().clone(), gf.cpu().clone())
    names = list(res)
    print(f"n={n} n_c={n_c} n_f={n_f} gain={gain}")
    for i, a in enumerate(names):
        for b in names[i + 1:]:
            rc = ((res[a][0] - res[b][0]).norm() / res[b][0].norm()).item()
            rf = ((res[a][1] - res[b][1]).norm() / res[b][1].norm()).item()
            print(f"  {a:12s} vs {b:12s}: coarse {rc:.4f} fine {rf:.4f}")

for gain in (4.0, 1.0, 30.0):
    run(200, 55, 55, gain)
run(200, 64, 128, 4.0)
