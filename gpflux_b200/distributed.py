"""Data-parallel plumbing for the SVGP path: the minibatch N is sharded over ranks,
parameters and the M x M factorisations are replicated, and the FP64 gradients (Z, lengthscales,
variance, q_mu, q_sqrt, likelihood variance) are summed over ranks - one all-reduce per layer
bucket, issued as soon as that layer's backward has produced its last gradient so that it overlaps
the backward of the layers below it.

The reference has no distributed code at all; this is a pure addition.  Transport: the library's
own NCCL communicator (``gpfx_allreduce_grads``, include/gpfx.h) on GPUs, ``torch.distributed``
(gloo) in the CPU tests.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_rows(n_rows: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous [start, stop) row range of ``rank``; remainders go to the first ranks."""
    base, rem = divmod(n_rows, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def local_loss_weight(n_local: int, n_global: int, world_size: int) -> float:
    """Factor for the DATA term of a rank's loss when shards are unequal: each rank's loss is
    KL/num_data + mean over its own rows, and ``allreduce_mean`` averages uniformly over ranks, so the
    data term must be weighted by n_local * world / n_global for the average to be the gradient of the
    global-minibatch loss (1.0 for equal shards; the KL part is identical on every rank)."""
    return float(n_local) * world_size / float(n_global)


# ------------------------------------------------------------------------------ library communicator
class GpfxComm:
    """The library's NCCL communicator (``gpfx_comm_*`` of include/gpfx.h).  The rendezvous token is
    created on rank 0 and broadcast with torch.distributed; every rank then calls ``gpfx_comm_init``
    with its CUDA device current."""

    def __init__(self, group=None):
        from . import _cabi

        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("GpfxComm needs an initialised torch.distributed process group for the rendezvous")
        self.lib = _cabi.load()
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        nbytes = self.lib.gpfx_comm_unique_id_bytes()
        token = (C.c_ubyte * nbytes)()
        if self.rank == 0:
            _cabi.check(self.lib.gpfx_comm_unique_id(token))
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else "cpu"
        t = torch.tensor(list(bytes(token)), dtype=torch.uint8, device=dev)
        dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        token = (C.c_ubyte * nbytes)(*t.cpu().tolist())
        self.handle = C.c_void_p()
        _cabi.check(self.lib.gpfx_comm_init(token, self.rank, self.world, C.byref(self.handle)))

    def allreduce_(self, flat: torch.Tensor, average: bool, stream: Optional[torch.cuda.Stream] = None) -> None:
        from . import _cabi

        if flat.dtype != torch.float64 or not flat.is_cuda or not flat.is_contiguous():
            raise ValueError("gpfx_allreduce_grads takes a contiguous CUDA float64 buffer")
        s = stream if stream is not None else torch.cuda.current_stream(flat.device)
        _cabi.check(self.lib.gpfx_allreduce_grads(self.handle, flat.data_ptr(), flat.numel(), int(average), s.cuda_stream))

    def destroy(self) -> None:
        if self.handle:
            self.lib.gpfx_comm_destroy(self.handle)
            self.handle = C.c_void_p()


# ------------------------------------------------------------------------------------ gradient buckets
class FlatGradients:
    """All parameter gradients live in ONE persistent flat FP64 buffer, cut into buckets.

    Every ``p.grad`` is a view of the buffer from construction on (it is never re-pointed, so a captured
    CUDA graph keeps writing where the all-reduce reads); ``zero_()`` clears the buffer in place and
    autograd accumulates into the views.  With ``overlap=True`` a bucket's all-reduce is issued on a
    communication stream the moment its last gradient has been accumulated (post-accumulate hooks), so
    the exchange of layer l overlaps the backward of layers < l; ``allreduce_mean()`` issues whatever
    is still pending and joins the communication stream back into the current one.

    ``buckets``: an iterable of parameters (one bucket) or a sequence of such iterables.
    ``comm``: a ``GpfxComm`` (library NCCL) or None for ``torch.distributed`` collectives."""

    def __init__(self, buckets, comm: Optional[GpfxComm] = None, overlap: bool = False, group=None):
        buckets = list(buckets)
        if buckets and isinstance(buckets[0], torch.Tensor):
            buckets = [buckets]
        self.buckets: List[List[torch.nn.Parameter]] = []
        seen = set()
        for b in buckets:
            ps = []
            for p in b:
                if p.requires_grad and id(p) not in seen:
                    seen.add(id(p))
                    ps.append(p)
            if ps:
                self.buckets.append(ps)
        self.params: List[torch.nn.Parameter] = [p for b in self.buckets for p in b]
        if not self.params:
            raise ValueError("no trainable parameters")
        p0 = self.params[0]
        if any(p.dtype != p0.dtype or p.device != p0.device for p in self.params):
            raise ValueError("FlatGradients needs all parameters on one device with one dtype")
        self._numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self._numel, dtype=p0.dtype, device=p0.device)
        self.comm = comm
        self.group = group
        self.overlap = bool(overlap) and p0.is_cuda
        self._views = {}
        self._bucket_slices = []
        self._bucket_of = {}
        off = 0
        for bi, b in enumerate(self.buckets):
            start = off
            for p in b:
                n = p.numel()
                self._views[id(p)] = self.flat[off : off + n].view_as(p)
                self._bucket_of[id(p)] = bi
                off += n
            self._bucket_slices.append((start, off))
        self._attach()
        self._pending = [len(b) for b in self.buckets]
        self._launched = [False] * len(self.buckets)
        self._events = [[] for _ in self.buckets]
        self._comm_stream = torch.cuda.Stream(device=p0.device) if self.overlap else None
        self._hooks = []
        self._paused = False
        if self.overlap:
            for p in self.params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))

    # -------------------------------------------------------------------------------- bookkeeping
    def _attach(self) -> None:
        for p in self.params:
            v = self._views[id(p)]
            if p.grad is None or p.grad.data_ptr() != v.data_ptr():
                p.grad = v

    def numel(self) -> int:
        return self._numel

    def bucket_numels(self) -> List[int]:
        return [b - a for a, b in self._bucket_slices]

    def zero_(self) -> None:
        """Clear all gradients in place (one memset) and re-arm the buckets."""
        self._attach()  # e.g. after optimizer.zero_grad(set_to_none=True)
        self.flat.zero_()
        self._pending = [len(b) for b in self.buckets]
        self._launched = [False] * len(self.buckets)
        self._events = [[] for _ in self.buckets]

    def pack(self) -> torch.Tensor:
        """The flat buffer (gradients are already in it; kept for callers of the old interface)."""
        self._attach()
        return self.flat

    # -------------------------------------------------------------------------------- collectives
    def _world(self) -> int:
        if self.comm is not None:
            return self.comm.world
        if not dist.is_available() or not dist.is_initialized():
            return 1
        return dist.get_world_size(self.group)

    def _reduce(self, buf: torch.Tensor, stream=None) -> None:
        ws = self._world()
        if self.comm is not None:
            self.comm.allreduce_(buf, average=True, stream=stream)
        elif buf.is_cuda and dist.get_backend(self.group) == "nccl":
            dist.all_reduce(buf, op=dist.ReduceOp.AVG, group=self.group)  # the division rides inside NCCL
        else:
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
            buf.div_(ws)

    def paused(self):
        """Context manager: backward passes inside it do not trigger this object's collectives (e.g. a
        rank-local reference evaluation of the same parameters)."""
        import contextlib

        @contextlib.contextmanager
        def _cm():
            self._paused = True
            try:
                yield self
            finally:
                self._paused = False

        return _cm()

    def _on_grad(self, p) -> None:
        bi = self._bucket_of[id(p)]
        if self._paused or self._launched[bi]:
            return
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.flat.device))
        self._events[bi].append(ev)
        self._pending[bi] -= 1
        if self._pending[bi] == 0:
            self._launch(bi)

    def _launch(self, bi: int) -> None:
        self._launched[bi] = True
        if self._world() == 1:
            return
        a, b = self._bucket_slices[bi]
        cs = self._comm_stream
        for ev in self._events[bi]:
            cs.wait_event(ev)
        if self.comm is not None:
            self._reduce(self.flat[a:b], stream=cs)
        else:
            with torch.cuda.stream(cs):
                self._reduce(self.flat[a:b])

    def allreduce_mean(self) -> None:
        """Mean over ranks of every gradient.  Each rank's loss is KL/num_data + mean over its own rows,
        so for equal shards the mean over ranks is the gradient of the global-minibatch loss (the KL
        part is identical on every rank; see ``local_loss_weight`` for unequal shards)."""
        ws = self._world()
        if ws == 1:
            return
        if not self.overlap:
            self._reduce(self.flat)
            return
        cur = torch.cuda.current_stream(self.flat.device)
        for bi in range(len(self.buckets)):
            if not self._launched[bi]:
                # gradients that never fired a hook (unused parameters) are zeros; order after everything
                # queued on the current stream
                ev = torch.cuda.Event()
                ev.record(cur)
                self._events[bi].append(ev)
                self._launch(bi)
        cur.wait_stream(self._comm_stream)

    @classmethod
    def for_model(cls, model, comm: Optional[GpfxComm] = None, overlap: bool = True, group=None) -> "FlatGradients":
        """One bucket per GPLayer of a DeepGP (in layer order) plus one for everything else (likelihood)."""
        buckets = [list(layer.parameters()) for layer in getattr(model, "f_layers", [])]
        rest = [p for p in model.parameters()]
        buckets.append(rest)  # duplicates are dropped by the constructor
        return cls(buckets, comm=comm, overlap=overlap, group=group)


def allreduce_scalar_mean(x: torch.Tensor, group=None) -> torch.Tensor:
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return x
    y = x.detach().clone()
    dist.all_reduce(y, op=dist.ReduceOp.SUM, group=group)
    return y / dist.get_world_size(group)
