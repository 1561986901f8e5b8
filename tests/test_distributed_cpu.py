"""World-size-2 gloo test of the N>1 path's host logic (no GPU): contiguous ray shards, loss normalised by the GLOBAL
ray count, one all-reduce(sum) of the flat gradient buffer == the single-process gradient; and the Philox stream keyed
by the global ray index gives every rank the draws the single-process run uses."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, out_dir):
    import importlib
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from helpers import FAR, NEAR, make_params, oracle_cfg, random_rays
    from oracle import nerf_oracle as O
    parallel = importlib.import_module("nerf-and-dietnerf_b200.parallel")
    cfg = oracle_cfg()
    pc, pf = make_params(cfg, 1, 4.0), make_params(cfg, 2, 4.0)
    o, d = random_rays(n, 0)
    y = torch.rand(n, 3, generator=torch.Generator().manual_seed(0))
    lo, hi = parallel.shard_bounds(n, world, rank)
    jit = O.stratified_jitter(3, 0, hi - lo, 16, ray_offset=lo)
    u = O.importance_uniforms(3, 0, hi - lo, 16, ray_offset=lo)
    pcs, pfs = pc.clone().requires_grad_(True), pf.clone().requires_grad_(True)
    out = O.train_losses(pcs, pfs, cfg, NEAR, FAR, o[lo:hi], d[lo:hi], y[lo:hi], 16, 16, jit, u)
    # what NeRF.train_step_local does: squared errors summed over the shard, divided by the GLOBAL element count
    sq_c = ((out["rgb_c"] - y[lo:hi]) ** 2).sum()
    sq_f = ((out["rgb_f"] - y[lo:hi]) ** 2).sum()
    ((sq_c + sq_f) / (3.0 * n)).backward()
    # the step's flat buffer [sq_c, sq_f, 0, 0 | grads_coarse | grads_fine] and its two all-reduces: the fine slice
    # asynchronously (it overlaps the coarse backward on the GPU), then [sums | coarse gradients]
    flat = torch.cat([sq_c.detach().reshape(1), sq_f.detach().reshape(1), torch.zeros(2), pcs.grad, pfs.grad])
    n_c = pcs.numel()
    work = parallel.allreduce_sum_(flat[4 + n_c:], async_op=True)
    parallel.allreduce_sum_(flat[:4 + n_c])
    work.wait()
    if rank == 0:
        np.save(os.path.join(out_dir, "flat.npy"), flat.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gradient_equals_single_process(tmp_path):
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import FAR, NEAR, make_params, oracle_cfg, random_rays
    from oracle import nerf_oracle as O
    n = 24                                    # 2 ranks x 12 rays
    mp.spawn(_worker, args=(2, _free_port(), n, str(tmp_path)), nprocs=2, join=True)
    flat = torch.from_numpy(np.load(tmp_path / "flat.npy"))
    cfg = oracle_cfg()
    pc, pf = make_params(cfg, 1, 4.0), make_params(cfg, 2, 4.0)
    o, d = random_rays(n, 0)
    y = torch.rand(n, 3, generator=torch.Generator().manual_seed(0))
    jit, u = O.stratified_jitter(3, 0, n, 16), O.importance_uniforms(3, 0, n, 16)
    metrics, gc, gf, out = O.train_step(pc, pf, cfg, NEAR, FAR, o, d, y, 16, 16, jit, u)
    np_ = cfg.n_params
    assert ((flat[4:4 + np_] - gc).norm() / gc.norm()).item() < 1e-5
    assert ((flat[4 + np_:] - gf).norm() / gf.norm()).item() < 1e-5
    assert flat[2:4].abs().max().item() == 0.0
    loss = (flat[0] + flat[1]) / (3.0 * n)
    assert abs(loss.item() - metrics["loss"].item()) < 1e-6


def test_shard_bounds():
    import importlib
    parallel = importlib.import_module("nerf-and-dietnerf_b200.parallel")
    assert [parallel.shard_bounds(4096, 8, r) for r in (0, 7)] == [(0, 512), (3584, 4096)]
    spans = [parallel.shard_bounds(10, 4, r) for r in range(4)]
    assert spans == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert [parallel.shard_bounds(2, 4, r) for r in range(4)] == [(0, 1), (1, 2), (2, 2), (2, 2)]   # empty shards
    assert parallel.shard_bounds(7, 1, 0) == (0, 7)


def _consistency_worker(rank, world, port, size, n_s, out_dir):
    """The sharding DietNeRF.calc_consistency_loss uses, with the oracle standing in for the kernels: each rank renders
    its contiguous block of the image's rays, the blocks are all-gathered (ragged), every rank embeds the full image and
    back-propagates d(loss)/d(image) through ITS block only; the all-reduced gradient must equal the single-process one."""
    import importlib
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from helpers import FAR, NEAR, make_params, oracle_cfg, sphere_pose
    from oracle import nerf_oracle as O
    parallel = importlib.import_module("nerf-and-dietnerf_b200.parallel")
    vit = importlib.import_module("nerf-and-dietnerf_b200.vit")
    cfg = oracle_cfg()
    pc, pf = make_params(cfg, 1, 30.0), make_params(cfg, 2, 30.0)
    emb = vit.ViTB32(layers=1, seed=5).eval()
    target = torch.randn(768, generator=torch.Generator().manual_seed(1))
    orig, dirs = O.rays_for_image(sphere_pose(0.4, -0.2, 1.0), 0.6, size, size)
    n = size * size
    lo, hi = parallel.shard_bounds(n, world, rank)
    pcs, pfs = pc.clone().requires_grad_(True), pf.clone().requires_grad_(True)
    jit = O.stratified_jitter(9, 13, hi - lo, n_s, ray_offset=lo)
    u = O.importance_uniforms(9, 13, hi - lo, n_s, ray_offset=lo)
    rgb = O.render(pcs, pfs, cfg, NEAR, FAR, orig[lo:hi], dirs[lo:hi], n_s, n_s, jit, u)[0]
    image = parallel.all_gather_rows(rgb.detach(), n).reshape(size, size, 3).requires_grad_(True)
    loss = 0.1 * vit.consistency_loss(emb(vit.embedder_preprocess(image[None]))[0], target)
    d_image, = torch.autograd.grad(loss, image)
    (rgb * d_image.reshape(n, 3)[lo:hi]).sum().backward()
    flat = torch.cat([pcs.grad, pfs.grad])
    parallel.allreduce_sum_(flat)
    if rank == 0:
        np.save(os.path.join(out_dir, "cs.npy"), np.concatenate([[loss.item()], flat.numpy()]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_consistency_term_equals_single_process(tmp_path):
    import importlib
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import FAR, NEAR, make_params, oracle_cfg, sphere_pose
    from oracle import nerf_oracle as O
    vit = importlib.import_module("nerf-and-dietnerf_b200.vit")
    size, n_s = 5, 8                          # 25 rays: ragged shards of 13 and 12
    mp.spawn(_consistency_worker, args=(2, _free_port(), size, n_s, str(tmp_path)), nprocs=2, join=True)
    got = torch.from_numpy(np.load(tmp_path / "cs.npy")).float()
    cfg = oracle_cfg()
    pc, pf = make_params(cfg, 1, 30.0), make_params(cfg, 2, 30.0)
    emb = vit.ViTB32(layers=1, seed=5).eval()
    target = torch.randn(768, generator=torch.Generator().manual_seed(1))
    loss, gc, gf, _ = O.consistency_loss_and_grads(pc, pf, cfg, NEAR, FAR, sphere_pose(0.4, -0.2, 1.0), 0.6, size, size * size,
                                                   n_s, 9, 13, emb, target, 0.1)
    ref = torch.cat([gc, gf])
    assert abs(got[0].item() - loss.item()) < 1e-6
    assert ((got[1:] - ref).norm() / ref.norm()).item() < 1e-4


# ---- the video frame loop under torch.distributed (ExecutionRun.render_frames / render_video) ----------------------------
class _StubModel:
    """Stands in for NeRF on a CPU rank: the 'render' of a pose is a deterministic function of the pose, so that every
    frame can be told apart after the gather (the CUDA render itself is covered by the GPU tests)."""
    device = torch.device("cpu")

    def render_image_lean(self, c2w, fov, h, w):
        yy, xx = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
        k = float(np.asarray(c2w)[0, 3])
        rgb = torch.stack([(xx + k) / (w + 8.0), (yy + k) / (h + 8.0), torch.full_like(xx, (k % 7) / 7.0)], -1)
        depth = 0.5 + 0.1 * k + xx * 0.01 + (yy * 0.02) ** 2
        return rgb.reshape(-1, 3), depth.reshape(-1), torch.ones(h * w)


def _video_run(save_location=None):
    import importlib
    pkg = importlib.import_module("nerf-and-dietnerf_b200")
    config = {"neural_net": {"type_of_model": "NeRF"}, "render": {}, "training": {"test_img_idx": 0},
              "video": {"fps_render_video": 10}}
    images = np.zeros((2, 6, 8, 3), dtype=np.float32)
    return pkg, pkg.ExecutionRun.from_arrays(config, images, np.stack([np.eye(4, dtype=np.float32)] * 2), 0.6, 0.5, 2.5,
                                             save_location=save_location)


def _video_poses(n_frames):
    poses = np.tile(np.eye(4, dtype=np.float32), (n_frames, 1, 1))
    poses[:, 0, 3] = np.arange(n_frames)
    return poses


def _video_worker(rank, world, port, n_frames, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    pkg, run = _video_run(save_location=out_dir)
    assert run.is_main == (rank == 0)
    rgbs, depths = run.render_frames(_StubModel(), _video_poses(n_frames))
    _, levels = run.render_frames(_StubModel(), _video_poses(n_frames), equalize_depth=True)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), rgbs=rgbs, depths=depths, levels=levels)
    run.render_video(_video_poses(n_frames), "two-rank video", "rgb.avi", "depth.avi", model=_StubModel())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("n_frames", [5, 1])
def test_two_rank_video_frames_equal_single_process(tmp_path, n_frames):
    """Frames are dealt round-robin to the ranks and gathered once: every rank ends with ALL frames in pose order, equal
    to the single-process loop -- also when the count does not divide (5 over 2) or a rank gets nothing (1 over 2) -- and
    only rank 0 writes the two video files."""
    mp.spawn(_video_worker, args=(2, _free_port(), n_frames, str(tmp_path)), nprocs=2, join=True)
    pkg, run = _video_run()
    rgbs, depths = run.render_frames(_StubModel(), _video_poses(n_frames))
    _, levels = run.render_frames(_StubModel(), _video_poses(n_frames), equalize_depth=True)
    assert rgbs.shape == (n_frames, 6, 8, 3) and rgbs.dtype == np.uint8 and levels.dtype == np.uint8
    assert len({rgbs[i].tobytes() for i in range(n_frames)}) == n_frames            # frames are distinguishable
    for rank in range(2):
        got = np.load(tmp_path / f"rank{rank}.npz")
        assert np.array_equal(got["rgbs"], rgbs) and np.array_equal(got["depths"], depths)
        assert np.array_equal(got["levels"], levels)
    frames, fps = pkg.UtilsVideo.read_video_frames(tmp_path / "video_save" / "rgb.avi")
    assert frames.shape == (n_frames, 6, 8, 3) and fps == 10
    assert sorted(os.listdir(tmp_path / "video_save")) == ["depth.avi", "rgb.avi"]
