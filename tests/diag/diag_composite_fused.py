"""Which output of the fused compositing kernels differs from the separate kernels (bit level)?  GPU box only."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
pkg = importlib.import_module("nerf-and-dietnerf_b200")
call = pkg._lib.call
for n, s in [(2048, 64), (1001, 128), (4096, 192), (333, 256), (77, 110), (64, 32)]:
    g = torch.Generator().manual_seed(n + s)
    raw = (torch.randn(n, s, 4, generator=g) * 2).cuda(); raw[..., 3] *= 3
    z = torch.sort(torch.rand(n, s, generator=g) * 2 + 0.5, -1).values.cuda()
    y = torch.rand(n, 3, generator=g).cuda()
    f = lambda *shape: torch.empty(shape, device="cuda")
    n_total, wgt = 3 * n, 2.0
    rgb0, w0, d_rgb0, sum0 = f(n, 3), f(n, s), f(n, 3), torch.zeros(1, device="cuda")
    call("nerf_composite_fwd", raw.data_ptr(), z.data_ptr(), n, s, rgb0.data_ptr(), w0.data_ptr(), None, None, None, None, None)
    call("nerf_mse_fwd_bwd", rgb0.data_ptr(), y.data_ptr(), n, n_total, wgt, sum0.data_ptr(), d_rgb0.data_ptr())
    d_raw0, d_z0 = f(n, s, 4), f(n, s)
    call("nerf_composite_bwd", raw.data_ptr(), z.data_ptr(), d_rgb0.data_ptr(), None, n, s, d_raw0.data_ptr(), d_z0.data_ptr())
    rgb1, w1, d_rgb1, sum1 = f(n, 3), f(n, s), f(n, 3), torch.zeros(1, device="cuda")
    call("nerf_composite_mse_fwd", raw.data_ptr(), z.data_ptr(), y.data_ptr(), n, s, n_total, wgt, rgb1.data_ptr(),
         w1.data_ptr(), sum1.data_ptr(), d_rgb1.data_ptr())
    rgb2, sum2, d_raw2, d_z2 = f(n, 3), torch.zeros(1, device="cuda"), f(n, s, 4), f(n, s)
    call("nerf_composite_mse_fwd_bwd", raw.data_ptr(), z.data_ptr(), y.data_ptr(), n, s, n_total, wgt, rgb2.data_ptr(),
         sum2.data_ptr(), d_raw2.data_ptr(), d_z2.data_ptr())
    rep = {k: (int((a != b).sum()), float((a - b).abs().max())) for k, a, b in
           [("rgb1", rgb1, rgb0), ("w1", w1, w0), ("d_rgb1", d_rgb1, d_rgb0), ("rgb2", rgb2, rgb0), ("d_raw2", d_raw2, d_raw0), ("d_z2", d_z2, d_z0)]}
    print(n, s, rep)
