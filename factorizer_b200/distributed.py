"""Gradient all-reduce for the batch-sharded training config (BASELINE config 5; the reference trains with Lightning's
DDP strategy: model_zoo/factorizer_brats23/configs/train_multigpu.yaml:3-6,25-38).  The hot path itself has no collective
(windows are independent, the volume batch is sharded); the glue parameters' gradients are the only exchange.

``BucketedGradAllReduce`` packs the gradients into a few flat buckets laid out in the order the backward pass produces
them (one multi-tensor copy per bucket) and launches a bucket's all-reduce from a post-accumulate hook as soon as its last
gradient has landed, so the collective runs on the process group's stream under the rest of the backward.  Nothing in it
touches the host between launches: a whole training step (forward, backward, the all-reduces, the optimizer) can be
captured in ONE CUDA graph, which torch's DistributedDataParallel (host-side reducer) does not allow.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist

__all__ = ["BucketedGradAllReduce"]


class BucketedGradAllReduce:
    """Average the gradients of ``params`` over the ranks of ``group`` in buckets of at most ``bucket_bytes``.

    Use per step: ``zero_grad()`` (instead of ``optimizer.zero_grad``), forward, backward, ``wait()``,
    ``optimizer.step()``.  During the backward autograd hands every parameter a fresh gradient tensor (no accumulation
    kernel); when the last gradient of a bucket has arrived they are packed into the flat bucket by ONE multi-tensor
    copy, ``.grad`` becomes the view into the bucket, and the bucket's all-reduce is launched.  Buckets are launched
    strictly in index order on every rank, whatever order the hooks fire in; ``wait()`` completes the buckets whose
    parameters received no gradient (zeros: every rank launches the same collectives), makes the current stream wait for
    all of them and, on backends without an averaging reduction (gloo), divides by the world size."""

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_bytes: int = 4 << 20,
                 group: Optional[dist.ProcessGroup] = None):
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("BucketedGradAllReduce needs an initialised torch.distributed process group")
        self.group = group
        self.world = dist.get_world_size(group)
        self._avg = dist.get_backend(group) == "nccl"
        plist = [p for p in params if p.requires_grad]
        if not plist:
            raise ValueError("no parameter requires a gradient")
        # backward reaches the last layers first: bucket 0 holds the last parameters
        plist = plist[::-1]
        groups: List[List[torch.nn.Parameter]] = []
        cur: List[torch.nn.Parameter] = []
        cur_bytes = 0
        for p in plist:
            nbytes = p.numel() * p.element_size()
            if cur and (cur_bytes + nbytes > bucket_bytes or p.dtype != cur[0].dtype or p.device != cur[0].device):
                groups.append(cur)
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_bytes += nbytes
        groups.append(cur)
        self.buckets: List[torch.Tensor] = []
        self._params = groups
        self._views: List[List[torch.Tensor]] = []
        self._bucket_of = {}
        for b, ps in enumerate(groups):
            flat = torch.zeros(sum(p.numel() for p in ps), dtype=ps[0].dtype, device=ps[0].device)
            views, o = [], 0
            for p in ps:
                views.append(flat[o:o + p.numel()].view_as(p))
                o += p.numel()
                self._bucket_of[p] = b
                p.grad = None
            self.buckets.append(flat)
            self._views.append(views)
        self._sizes = [len(ps) for ps in groups]
        self._missing = list(self._sizes)
        self._ready = [False] * len(groups)
        self._next = 0
        self._works = []
        self._handles = [p.register_post_accumulate_grad_hook(self._hook) for p in plist]

    # ---- per step ----
    def zero_grad(self) -> None:
        """Gradients to None (autograd then stores, not accumulates) and the bucket state reset; no device work."""
        for ps in self._params:
            for p in ps:
                p.grad = None
        self._missing = list(self._sizes)
        self._ready = [False] * len(self.buckets)
        self._next = 0
        self._works = []

    def _pack(self, b: int) -> None:
        views, ps = self._views[b], self._params[b]
        have = [(v, p.grad) for v, p in zip(views, ps) if p.grad is not None and p.grad.data_ptr() != v.data_ptr()]
        if have:
            torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
        for v, p in zip(views, ps):
            if p.grad is None:
                v.zero_()
            p.grad = v

    def _launch_ready(self) -> None:
        while self._next < len(self.buckets) and self._ready[self._next]:
            op = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
            self._works.append(dist.all_reduce(self.buckets[self._next], op=op, group=self.group, async_op=True))
            self._next += 1

    def _hook(self, p: torch.nn.Parameter) -> None:
        b = self._bucket_of[p]
        if self._ready[b] or self._missing[b] <= 0:
            raise RuntimeError("BucketedGradAllReduce: a gradient arrived for a bucket that was already reduced -- call zero_grad() "
                               "before every backward pass (gradient accumulation over several backward passes is not supported)")
        self._missing[b] -= 1
        if self._missing[b] == 0:
            self._pack(b)
            self._ready[b] = True
            self._launch_ready()

    def wait(self) -> None:
        for b in range(len(self.buckets)):
            if not self._ready[b]:
                self._pack(b)
                self._ready[b] = True
        self._launch_ready()
        for w in self._works:
            w.wait()
        self._works = []
        if not self._avg:
            for flat in self.buckets:
                flat.div_(self.world)

    # ---- teardown ----
    def remove(self) -> None:
        """Remove the hooks and give the parameters free-standing gradients again (``None``)."""
        for h in self._handles:
            h.remove()
        self._handles = []
        for p in self._bucket_of:
            p.grad = None

    @property
    def nbytes(self) -> int:
        return sum(f.numel() * f.element_size() for f in self.buckets)
