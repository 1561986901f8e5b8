"""Multi-GPU check of the NVLink peer-memory gradient exchange (nerf_peer_barrier + nerf_peer_reduce_adam) against the NCCL
all-reduce path: same data, same steps -> parameters agree to fp32 summation order, replicas stay bit-identical, and the
step time of both.  Run: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/check_peer_exchange.py"""
import importlib
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = importlib.import_module("nerf-and-dietnerf_b200")
    pkg.load()
    per = int(os.environ.get("RAYS_PER_GPU", "4096"))
    n_total = per * world
    near, far, fov = 0.5576, 2.5635, 0.46134
    rcfg = {"n_render_samples_coarse": 64, "n_render_samples_fine": 128}

    def rays(c2w, fov_, h, w):
        dirs, orig = pkg.UtilsCV.get_rays_directions(h, w, fov_, c2w, return_origins=True)
        return orig.cpu(), dirs.reshape(-1, 4).cpu()
    batches = [tuple(t.cuda() for t in B.synthetic_batch(per, fov, 1000 * b + rank, rays)) for b in range(3)]
    models = {}
    # "peer": the sharded step as ONE C call (nerf_train_step_fused_sharded); "peer_seq": the host package's call sequence
    # with the same exchange -- must be bit-identical to it
    for name, peer in (("nccl", False), ("peer", True), ("peer_seq", True)):
        m = pkg.NeRFModel(B.net_config(per), rcfg, near, far, seed=0)
        m.compile(optimizer=pkg.Adam(5e-4))
        m.distribute(peer_exchange=peer)
        assert (m._peer is not None) == peer
        if name == "peer_seq":
            m.use_fused_step = False
        models[name] = m
    losses = {k: [] for k in models}
    for step in range(6):
        o, d, y = batches[step % 3]
        for name, m in models.items():
            losses[name].append(m.train_step_local(o, d, y, n_total, rank * per)["loss"].item())
    torch.cuda.synchronize()
    for net in ("model_coarse", "model_fine"):
        same = torch.equal(getattr(models["peer"], net).params, getattr(models["peer_seq"], net).params)
        assert same, f"{net}: the one-call sharded step and the call sequence differ"
    # (the reported loss is formed by nerf_train_metrics in the one-call step and by torch ops in the sequence: last-bit)
    assert all(abs(x - y) <= 1e-6 * abs(y) for x, y in zip(losses["peer"], losses["peer_seq"])), (losses["peer"], losses["peer_seq"])
    if rank == 0:
        print("one-call sharded step == call sequence: parameters bit-identical, reported losses equal to 1e-6")
    a, b = models["nccl"], models["peer"]
    for net in ("model_coarse", "model_fine"):
        pa, pb = getattr(a, net).params, getattr(b, net).params
        rel = ((pa - pb).norm() / pa.norm()).item()
        # replicas of the peer path: bit-identical across ranks
        ref = pb.clone()
        dist.broadcast(ref, src=0)
        same = torch.equal(ref, pb)
        flags = torch.tensor([1 if same else 0], device="cuda")
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(f"{net}: peer vs nccl parameters rel diff {rel:.2e} after 6 steps; replicas bit-identical: {bool(flags.item())}")
        # the two paths add the ranks in different orders (NCCL ring vs rank order); through the importance sampler and Adam's
        # g / sqrt(v) a last-bit difference of a near-zero gradient moves a weight by ~lr, so a few 1e-4 after 6 steps of
        # lr = 5e-4 on 8 GPUs is the expected size (0 on 2 GPUs, where a two-term sum has one order)
        assert rel < 5e-3 and flags.item() == 1
    if rank == 0:
        print("losses nccl", [f"{v:.6f}" for v in losses["nccl"]])
        print("losses peer", [f"{v:.6f}" for v in losses["peer"]])
    # timing
    for name, m in list(models.items()) * 3:          # interleaved repeats: the power cap drifts over a run
        for i in range(5):
            m.train_step_local(*batches[i % 3], n_total, rank * per)
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        k = 30
        for i in range(k):
            m.train_step_local(*batches[i % 3], n_total, rank * per)
        e1.record()
        dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / k], device="cuda")
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"{name}: {ms.item():.4f} ms/step at {per} rays/GPU x {world} GPUs = {n_total / ms.item() * 1e3:.0f} rays/s")
    # the exchange alone: barrier + one-shot reduce + Adam of one network's slice, against all-reduce + Adam kernel
    b = models["peer"]
    peer, opt, mc = b._peer, b.optimizer, b.model_coarse
    g_nccl = torch.zeros(4 + 2 * mc.n_params, device="cuda")

    def t_loop(fn, k=50):
        for _ in range(5):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / k * 1e3
    ep = [10 ** 6]

    def peer_tail():
        ep[0] += 1
        peer.barrier(ep[0], 1)
        peer.reduce_adam(mc.params, 4, mc.n_params, opt, 0, 7)

    def peer_barrier_only():
        ep[0] += 1
        peer.barrier(ep[0], 1)

    def nccl_tail():
        dist.all_reduce(g_nccl[:4 + mc.n_params])
        opt.apply_one(mc.params, g_nccl[4:4 + mc.n_params], 0, 2 * mc.n_params, 7)
    res = (t_loop(peer_barrier_only), t_loop(peer_tail), t_loop(nccl_tail))
    if rank == 0:
        print(f"exchange of one network's gradients alone: peer barrier {res[0]:.1f} us, barrier + one-shot reduce + Adam "
              f"{res[1]:.1f} us, NCCL all-reduce + Adam kernel {res[2]:.1f} us")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
