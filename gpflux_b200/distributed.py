"""Data-parallel plumbing for the SVGP path: the minibatch N is sharded over ranks,
parameters and the M x M factorisations are replicated, and ONE all-reduce per step sums the flat
FP64 gradient buffer (Z, lengthscales, variance, q_mu, q_sqrt, likelihood variance).

The reference has no distributed code at all; this is a pure addition.  ``torch.distributed`` is
the transport (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Iterable, List, Tuple

import torch
import torch.distributed as dist


def shard_rows(n_rows: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous [start, stop) row range of ``rank``; remainders go to the first ranks."""
    base, rem = divmod(n_rows, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


class FlatGradients:
    """One collective per step for all parameter gradients.

    ``zero_()`` drops the old gradients (autograd then *assigns* the fresh ones instead of launching
    one accumulate kernel per parameter); ``allreduce_mean()`` packs them into one flat FP64 buffer
    (a single concatenation kernel), all-reduces it, and re-points every ``.grad`` at its slice of the
    reduced buffer (views, no copy back)."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        self._numel = sum(p.numel() for p in self.params)
        self.flat = None

    def zero_(self) -> None:
        for p in self.params:
            p.grad = None

    def numel(self) -> int:
        return self._numel

    def pack(self) -> torch.Tensor:
        grads = [(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in self.params]
        self.flat = torch.cat(grads)
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[off : off + n].view_as(p)
            off += n
        return self.flat

    def allreduce_mean(self, group=None) -> None:
        """Sum over ranks, then divide by the world size.  Each rank's loss is
        KL/num_data + mean over its own rows, so the mean over ranks is the gradient of the
        global-minibatch loss (the KL part is identical on every rank)."""
        if not dist.is_available() or not dist.is_initialized():
            return
        ws = dist.get_world_size(group)
        if ws == 1:
            return
        flat = self.pack()
        if flat.is_cuda and dist.get_backend(group) == "nccl":
            dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)  # the division rides inside NCCL
        else:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
            flat.div_(ws)


def allreduce_scalar_mean(x: torch.Tensor, group=None) -> torch.Tensor:
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return x
    y = x.detach().clone()
    dist.all_reduce(y, op=dist.ReduceOp.SUM, group=group)
    return y / dist.get_world_size(group)
