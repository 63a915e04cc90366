"""Data-parallel host logic with world_size 2 on the gloo backend (CPU): the flat gradient bucket's
all-reduce-mean, and the "draw the global stream, keep your shard" RNG rule that makes N ranks consume the
random numbers one rank would."""
import os
import socket
import tempfile

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN_SCALE_LOSSES, GOLDEN_SIZES


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from sinddm_b200 import MultiScaleGaussianDiffusion, SinDDMNet
    from sinddm_b200 import dist as spdist
    r, lr, w = spdist.init_process_group(backend="gloo")
    assert (r, w) == (rank, world) and spdist.rank() == rank and spdist.world_size() == world

    # --- gradient bucket: every rank ends with the mean of the per-rank gradients
    torch.manual_seed(0)
    net = SinDDMNet(dim=16, multiscale=True)
    for i, p in enumerate(net.parameters()):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    bucket = spdist.GradientBucket(net.parameters())
    assert bucket.numel == sum(p.numel() for p in net.parameters())
    bucket.all_reduce_mean()
    for i, p in enumerate(net.parameters()):
        assert torch.allclose(p.grad, torch.full_like(p, 1.5 * (i + 1)))

    # --- RNG sharding: rank r's rows of the global draw
    dif = MultiScaleGaussianDiffusion(denoise_fn=net, n_scales=5, scale_factor=1.36, image_sizes=GOLDEN_SIZES,
                                      timesteps=100, train_full_t=True, scale_losses=GOLDEN_SCALE_LOSSES,
                                      results_folder=tempfile.mkdtemp())
    dif.set_data_parallel(rank, world)
    torch.manual_seed(123)
    local = dif._randn((3, 3, 5, 7), "cpu")
    torch.manual_seed(123)
    full = torch.randn((6, 3, 5, 7))
    assert torch.equal(local, full[rank * 3:(rank + 1) * 3])
    torch.save(local, os.path.join(out_dir, f"r{rank}.pt"))
    assert spdist.shard_batch(32, world) == 16
    with pytest.raises(ValueError):
        spdist.shard_batch(33, world)
    # --- trainer start-up: rank 0's parameters / EMA / RNG state win, image shards gather in rank order
    import numpy as np
    from PIL import Image
    from sinddm_b200 import MultiscaleTrainer
    torch.manual_seed(1000 + rank)                      # replicas deliberately different
    net2 = SinDDMNet(dim=16, multiscale=True)
    dif2 = MultiScaleGaussianDiffusion(denoise_fn=net2, n_scales=2, scale_factor=1.3, image_sizes=GOLDEN_SIZES[:2],
                                       timesteps=100, train_full_t=True, scale_losses=GOLDEN_SCALE_LOSSES[:1],
                                       results_folder=tempfile.mkdtemp())
    pyr = [(Image.fromarray(np.full((h, w, 3), 40 * (i + 1), np.uint8)),) * 2 for i, (w, h) in enumerate(GOLDEN_SIZES[:2])]
    tr = MultiscaleTrainer(dif2, None, n_scales=2, image_sizes=GOLDEN_SIZES[:2], train_batch_size=4,
                           results_folder=tempfile.mkdtemp(), device="cpu", pyramid=pyr)
    assert tr.local_batch == 2 and tr.data_list[1][0].shape[0] == 2
    tr._sync_replicas()
    flat = torch.cat([p.detach().reshape(-1) for p in tr.model.parameters()] +
                     [p.detach().reshape(-1) for p in tr.ema_model.parameters()] + [torch.rand(4)])
    both = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(both, flat)
    assert torch.equal(both[0], both[1]), "replicas / RNG differ after _sync_replicas"
    tr._check_replicas()                       # identical replicas: passes on both ranks
    first = next(tr.model.parameters())
    if rank == 1:
        with torch.no_grad():
            first.view(-1)[0] += 1e-7 * (1.0 + float(first.view(-1)[0].abs()))   # a last-bits drift on one rank
    try:
        tr._check_replicas()
        raised = False
    except RuntimeError as e:
        raised = "diverged" in str(e)
    assert raised, "replica drift was not detected"
    tr._sync_replicas()
    tr._check_replicas()
    imgs = tr._gather_images(torch.full((2, 3, 4, 5), float(rank)))
    assert imgs.shape[0] == 4 and torch.equal(imgs[:2], torch.zeros(2, 3, 4, 5)) and torch.equal(imgs[2:], torch.ones(2, 3, 4, 5))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo():
    port = _free_port()
    with tempfile.TemporaryDirectory() as td:
        mp.spawn(_worker, args=(2, port, td), nprocs=2, join=True)
        a, b = torch.load(os.path.join(td, "r0.pt")), torch.load(os.path.join(td, "r1.pt"))
        torch.manual_seed(123)
        assert torch.equal(torch.cat([a, b]), torch.randn((6, 3, 5, 7)))


def test_single_process_is_a_noop():
    from sinddm_b200 import dist as spdist
    assert spdist.world_size() == 1 and spdist.rank() == 0
    lin = torch.nn.Linear(3, 2)
    for p in lin.parameters():
        p.grad = torch.ones_like(p)
    spdist.GradientBucket(lin.parameters()).all_reduce_mean()
    assert all(torch.equal(p.grad, torch.ones_like(p)) for p in lin.parameters())
