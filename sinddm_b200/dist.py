"""Data-parallel plumbing: one process per GPU, torch.distributed (NCCL over NVLink on the B200 box, gloo in
the CPU tests).  The reference has no distributed code at all; this is the new part.

Training shards the batch: every rank runs the same scale `s`, draws the global (t, noise) stream and keeps
its rows (diffusion.set_data_parallel), then ONE all-reduce of a single flat fp32 gradient bucket
(1,106,772 values = 4.43 MB at dim=160) averages the gradients before the identical Adam / EMA / scheduler
steps on every rank.  Sampling shards independent images and needs no collective.
"""
from __future__ import annotations

import os
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


def env_world() -> "tuple[int, int, int]":
    """(rank, local_rank, world_size) from the torchrun environment; (0, 0, 1) outside torchrun."""
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)))


def init_process_group(backend: Optional[str] = None) -> "tuple[int, int, int]":
    """Initialise torch.distributed from the torchrun environment when WORLD_SIZE > 1."""
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def world_size() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank() -> int:
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def shard_batch(global_batch: int, world: int) -> int:
    """Rows per rank; the global batch must split evenly so the mean of shard means is the global mean."""
    if global_batch % world != 0:
        raise ValueError(f"global batch {global_batch} does not divide over {world} ranks")
    return global_batch // world


class GradientBucket:
    """One flat fp32 buffer holding every parameter gradient, all-reduced in a single collective."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.sizes = [p.numel() for p in self.params]
        self.numel = sum(self.sizes)
        self.flat: Optional[torch.Tensor] = None

    def _ensure(self, like: torch.Tensor) -> torch.Tensor:
        if self.flat is None or self.flat.device != like.device:
            self.flat = torch.empty(self.numel, dtype=torch.float32, device=like.device)
        return self.flat

    def all_reduce_mean(self, group=None) -> None:
        """grad <- mean over ranks of grad, for every parameter (no-op for world size 1)."""
        world = world_size()
        if world == 1:
            return
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.params]
        flat = self._ensure(grads[0])
        views = flat.split(self.sizes)
        torch._foreach_copy_(list(views), [g.reshape(-1) for g in grads])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(world)
        for p, v in zip(self.params, views):
            if p.grad is None:
                p.grad = v.view_as(p).clone()
            else:
                p.grad.copy_(v.view_as(p))
