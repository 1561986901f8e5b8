#!/usr/bin/env python
"""2-GPU check of the sharded DietNeRF consistency term (run under torchrun --nproc-per-node 2):
every rank renders / back-propagates its block of the 150x150 image's rays, the image is all-gathered, and the SUM of
the ranks' gradient buffers must equal the single-GPU gradient; the drawn pose/target must agree across ranks.
Also runs a few sharded DietNeRF train steps and checks that the ranks' weights stay identical."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import FAR, NEAR, net_config, random_rays, render_config, sphere_pose  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    pkg = importlib.import_module("nerf-and-dietnerf_b200")
    targets = torch.rand(4, 32, 32, 3, generator=torch.Generator().manual_seed(4)).numpy()
    poses = np.stack([sphere_pose(0.3 * i, 0.1) for i in range(4)])

    def make(distributed):
        m = pkg.DietNeRFModel(net_config(batch_train=2048), render_config(), NEAR, FAR, targets, poses, 0.6, -1,
                              mode="bf16", seed=3, embedder=pkg.vit.ViTB32(layers=2, seed=5).cuda().eval())
        m.compile(optimizer=pkg.Adam(5e-4))
        m.IMG_SIZE_FOR_CS_LOSS = 149            # 22201 rays: the ranks' blocks are ragged
        if distributed:
            m.distribute()
        m.counter = 13
        return m

    single, sharded = make(False), make(True)
    pose = sphere_pose(0.4, -0.2, 1.0)
    single._grad_buffer().zero_()
    loss1 = single.calc_consistency_loss(pose=pose, target_index=1)
    sharded._grad_buffer().zero_()
    loss2 = sharded.calc_consistency_loss(pose=pose, target_index=1)
    g = sharded._grad_buffer().clone()
    dist.all_reduce(g)
    ref = single._grad_buffer()
    rel = ((g - ref).norm() / ref.norm()).item()
    print(f"rank {rank}: loss single {loss1.item():.6f} sharded {loss2.item():.6f}; grad rel diff (sum over ranks vs "
          f"single GPU) {rel:.2e}")
    assert abs(loss1.item() - loss2.item()) < 1e-6 and rel < 1e-3
    # the drawn pose / target agree across ranks (rank 0's draw is broadcast)
    sharded._np_rng = np.random.RandomState(100 + rank)
    idx, drawn = sharded._draw_target_and_pose()
    buf = torch.tensor([float(idx)] + drawn.reshape(-1).tolist(), device="cuda", dtype=torch.float64)
    lst = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(lst, buf)
    assert all(torch.equal(lst[0], t) for t in lst)
    # sharded train steps incl. a consistency step keep the replicas identical
    o, d = random_rays(4096, 2)
    y = torch.rand(4096, 3, generator=torch.Generator().manual_seed(1))
    sharded.counter = 11
    for _ in range(3):
        m = sharded.train_step((o, d, y))
    p = sharded.model_fine.params.clone()
    lst = [torch.empty_like(p) for _ in range(world)]
    dist.all_gather(lst, p)
    assert all(torch.equal(lst[0], t) for t in lst), "replicas diverged"
    print(f"rank {rank}: 3 sharded DietNeRF steps ok, loss {float(m['loss']):.5f}, cosine term "
          f"{float(m['cosine_similarity_loss']):.5f}; replicas identical")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
