"""Why does smoke()'s bf16 coarse-gradient check move?  Same inputs as __graft_entry__.smoke(): CUDA bf16 vs the
bf16-emulating oracle vs the fp32 oracle, with and without the sampler path (stop_grad_z)."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as E
from oracle import nerf_oracle as O
pkg = importlib.import_module("nerf-and-dietnerf_b200"); pkg.load()
net = {"hidden_layer_dim": 256, "last_hidden_layer_dim": 128, "leaky_relu_alpha": 0.05, "n_pos_enc_dim_xyz": 5,
       "n_pos_enc_view_dir": 4, "n_angles_for_model": 2, "n_rays_in_batch_train": 256, "n_rays_in_batch_render": 256}
rcfg = {"n_render_samples_coarse": 64, "n_render_samples_fine": 128}
near, far, n = 0.5576, 2.5635, 256
c2w = np.eye(4, dtype=np.float32); c2w[:3, 3] = [0.05, -0.1, 1.2]
ocfg = O.NetCfg()
orig, dirs = O.rays_for_image(c2w, 0.69, 16, 16)
target = torch.rand(n, 3, generator=torch.Generator().manual_seed(3))
jit = O.stratified_jitter(11, 0, n, 64); u = O.importance_uniforms(11, 0, n, 128)
for gain in (4.0, 30.0):
    pc, pf = E._smoke_params(O, ocfg, 1, gain), E._smoke_params(O, ocfg, 2, gain)
    _, g32c, g32f, _ = O.train_step(pc, pf, ocfg, near, far, orig, dirs, target, 64, 128, jit, u)
    _, gbc, gbf, _ = O.train_step(pc, pf, ocfg, near, far, orig, dirs, target, 64, 128, jit, u, emulate_bf16=True)
    res = {}
    for mode in ("fp32", "bf16"):
        for sg in (False, True):
            m = pkg.NeRFModel(net, rcfg, near, far, mode=mode, seed=11, stop_grad_z=sg)
            m.model_coarse.set_params(pc); m.model_fine.set_params(pf)
            gc, gf, _ = m.forward_backward(orig.cuda(), dirs.cuda(), target.cuda(), seed=11, step=0)
            res[(mode, sg)] = (gc.cpu().clone(), gf.cpu().clone())
    rel = lambda a, b: ((a - b).norm() / b.norm()).item()
    print(f"gain {gain}: |g_c| fp32 oracle {g32c.norm():.4e}")
    print(f"  oracle bf16 vs oracle fp32 : coarse {rel(gbc, g32c):.4f} fine {rel(gbf, g32f):.4f}")
    print(f"  cuda fp32   vs oracle fp32 : coarse {rel(res[('fp32', False)][0], g32c):.4f} fine {rel(res[('fp32', False)][1], g32f):.4f}")
    print(f"  cuda bf16   vs oracle bf16 : coarse {rel(res[('bf16', False)][0], gbc):.4f} fine {rel(res[('bf16', False)][1], gbf):.4f}")
    print(f"  cuda bf16   vs oracle fp32 : coarse {rel(res[('bf16', False)][0], g32c):.4f}")
    print(f"  stop_grad_z: cuda bf16 vs cuda fp32 : coarse {rel(res[('bf16', True)][0], res[('fp32', True)][0]):.4f} "
          f"fine {rel(res[('bf16', True)][1], res[('fp32', True)][1]):.4f}")

# ---- kernel-level check of the sampler backward on the SAME soft weights, and sensitivity of the draw to 1e-7 changes ----
print("\nsampler backward in isolation (gain 4 weights from the oracle's coarse pass):")
pc, pf = E._smoke_params(O, ocfg, 1, 4.0), E._smoke_params(O, ocfg, 2, 4.0)
out = O.train_losses(pc, pf, ocfg, near, far, orig, dirs, target, 64, 128, jit, u)
w_c, z_c = out["w_c"].detach(), out["z_c"].detach()
d_z = torch.randn(n, 128, generator=torch.Generator().manual_seed(9))
w_o = w_c.clone().requires_grad_(True)
zf_o = O.get_z_vals_from_prob_dist_func(w_o, z_c, 128, u)
(zf_o * d_z).sum().backward()
w_g = w_c.clone().cuda().requires_grad_(True)
zf_g = pkg.UtilsCV.get_z_vals_from_prob_dist_func(w_g, z_c.cuda(), 128, u=u.cuda())
(zf_g * d_z.cuda()).sum().backward()
print(f"  z_new equal: {torch.equal(zf_g.detach().cpu(), zf_o.detach())}; d_w rel err {((w_g.grad.cpu() - w_o.grad).norm() / w_o.grad.norm()).item():.3e}; "
      f"|d_w| max {w_o.grad.abs().max().item():.3e} median {w_o.grad.abs().median().item():.3e}")
per_ray = w_o.grad.norm(dim=1)
print(f"  per-ray |d_w|: max {per_ray.max().item():.3e}, median {per_ray.median().item():.3e}, top-5 share of the squared norm "
      f"{(per_ray.topk(5).values ** 2).sum().item() / (per_ray ** 2).sum().item():.3f}")
# perturb the weights by 1 ulp-level noise: how much does the gradient move?
w_p = (w_c * (1 + 1e-7 * torch.randn_like(w_c))).requires_grad_(True)
zf_p = O.get_z_vals_from_prob_dist_func(w_p, z_c, 128, u)
(zf_p * d_z).sum().backward()
print(f"  oracle, weights perturbed by 1e-7 relative: d_w rel change {((w_p.grad - w_o.grad).norm() / w_o.grad.norm()).item():.3e}")

# ---- where do CUDA fp32 and the fp32 oracle part ways at gain 4?  compare the intermediates of the train step ----
print("\nintermediates of the train step, CUDA fp32 vs oracle fp32 (gain 4):")
pco, pfo = pc.clone().requires_grad_(True), pf.clone().requires_grad_(True)
out = O.train_losses(pco, pfo, ocfg, near, far, orig, dirs, target, 64, 128, jit, u)
out["z_f"].retain_grad(); out["w_c"].retain_grad(); out["rgb_c"].retain_grad()
out["loss"].backward()
m = pkg.NeRFModel(net, rcfg, near, far, mode="fp32", seed=11)
m.model_coarse.set_params(pc); m.model_fine.set_params(pf)
gc, gf, _ = m.forward_backward(orig.cuda(), dirs.cuda(), target.cuda(), seed=11, step=0)
ws = m._workspace(n)
rel = lambda a, b: ((a - b).norm() / (b.norm() + 1e-30)).item()
print(f"  w_c   rel {rel(ws.w_c.cpu(), out['w_c'].detach()):.3e}   z_f equal {torch.equal(ws.z_f.cpu(), out['z_f'].detach())} max abs diff {(ws.z_f.cpu() - out['z_f'].detach()).abs().max().item():.3e}")
print(f"  d_z_f rel {rel(ws.d_z_f.cpu(), out['z_f'].grad):.3e}   |d_z_f| {out['z_f'].grad.norm().item():.3e}")
dwc_o = out["w_c"].grad           # total gradient on w_c: from rgb_c? no -- w_c feeds only the sampler here
print(f"  d_w_c (sampler path) rel {rel(ws.d_w_c.cpu(), dwc_o):.3e}   |d_w_c| {dwc_o.norm().item():.3e}")
per = (ws.d_w_c.cpu() - dwc_o).norm(dim=1); ref = dwc_o.norm(dim=1)
worst = per.argmax().item()
print(f"  worst ray {worst}: |diff| {per[worst].item():.3e} of |d_w_c| {ref[worst].item():.3e}; its d_z_f rel err "
      f"{rel(ws.d_z_f.cpu()[worst], out['z_f'].grad[worst]):.3e}, max|z_f diff| {(ws.z_f.cpu()[worst] - out['z_f'].detach()[worst]).abs().max().item():.3e}")
print(f"  grads: coarse {rel(gc.cpu(), pco.grad):.4f} fine {rel(gf.cpu(), pfo.grad):.4f}")
