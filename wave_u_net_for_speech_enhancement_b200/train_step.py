"""Data-parallel training step around the drop-in model.

The reference wraps the model in single-process ``nn.DataParallel`` (trainer/base_trainer.py:26-27): every step it scatters
the batch, replicates the parameters, runs per-replica forwards (BatchNorm statistics per replica) and sums the replica
gradients on GPU 0. Here it is one process per GPU: each rank steps its own shard of the batch and the gradients are
averaged with ONE all-reduce over a flat bucket (NCCL on GPUs, gloo in the CPU tests) before the caller's unchanged
optimizer step — the mean over ranks of per-shard mean losses is the reference's mean loss over the whole batch when the
shards are equal. BatchNorm running statistics stay per rank, as they stay per replica in the reference (only replica 0's
survive there; rank 0 writes the checkpoint here).
"""
from __future__ import annotations

from typing import Iterable

import torch
import torch.distributed as dist


def allreduce_gradients(params: Iterable[torch.nn.Parameter], world_size: int | None = None) -> None:
    """Average ``.grad`` of the given parameters over all ranks, in place, with a single collective."""
    if not dist.is_available() or not dist.is_initialized():
        return
    world = world_size or dist.get_world_size()
    if world == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat.div_(world)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


def train_step(model: torch.nn.Module, optimizer: torch.optim.Optimizer, loss_function, mixture: torch.Tensor,
               clean: torch.Tensor) -> float:
    """trainer/trainer.py:34-38 for one rank's shard, plus the gradient all-reduce; returns this rank's loss."""
    optimizer.zero_grad()
    enhanced = model(mixture)
    loss = loss_function(clean, enhanced)
    loss.backward()
    allreduce_gradients(model.parameters())
    optimizer.step()
    return float(loss.detach())
