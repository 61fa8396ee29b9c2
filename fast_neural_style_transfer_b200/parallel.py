"""Multi-GPU sharding for the style-transfer path: one process per GPU (torch.distributed, NCCL over
NVLink 5 / NVSwitch).  Every operator on the path is per-image (InstanceNorm is per (n,c); Gram and the
squared errors are per-sample then summed), so:

  * inference shards the image batch across ranks with NO collective (`shard_batch`);
  * data-parallel training replicates StyleTransferNet + VGG-19 and has ONE exchange step per
    iteration: a SUM all-reduce of a flat fp32 bucket holding the 6 243 843 gradients (24.98 MB)
    (`GradientAllReduce`), followed by the caller's clip_grad_norm_ / Adam on every rank.

Exact equivalence with the single-process reference at the global batch needs the content and
style gradients SUMMED over ranks (batch sums, losses/losses.py:41,54) and the TV gradient AVERAGED (batch
mean, losses.py:71): scale the TV term by 1/world on every rank and all-reduce with SUM.  The NaN/Inf skip
of train.py:193 must be taken on a reduced flag (`all_finite`) or the ranks desynchronise.
"""
from __future__ import annotations

from typing import Iterable, List

import torch
import torch.distributed as dist


def shard_bounds(total: int, rank: int, world: int):
    """Contiguous, balanced [lo, hi) slice of `total` items for `rank`."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(x: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    lo, hi = shard_bounds(x.shape[0], rank, world)
    return x[lo:hi]


def tv_weight_scale(world: int) -> float:
    return 1.0 / world


def all_finite(value: torch.Tensor, world: int) -> bool:
    """True iff `value` is finite on every rank (one MIN all-reduce of a flag)."""
    flag = torch.isfinite(value.detach()).all().to(torch.float32).reshape(1)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return bool(flag.item() > 0)


class GradientAllReduce:
    """Flat-bucket SUM all-reduce of `module`'s gradients (parameters without a gradient contribute zeros)."""

    def __init__(self, module: torch.nn.Module, world: int, overlap: bool = False):
        """overlap: on the drop-in StyleTransferNet (CUDA-graph training path) the exchange is issued bucket by bucket from
        inside the backward -- the gradients of residual blocks 2-4 and the decoder (60 % of the bytes) as soon as the backward
        has passed block 2, those of blocks 0-1 after block 0, the encoder's last -- as asynchronous all-reduces on the
        communicator's stream, so they run under the rest of the backward; `all_reduce()` after `backward()` then only finds
        the buffer already reduced.  Collectives are issued in the same order on every rank (three per step).
        Default off: measured on 2 B200s the three stage graphs (weight-gradient branch joined at every cut) cost 0.09 ms per
        step, more than the ~0.05 ms all-reduce they hide there (profiles/r02_scale.md has the 8-GPU comparison)."""
        self.params: List[torch.nn.Parameter] = [p for p in module.parameters() if p.requires_grad]
        self.world = world
        self.numel = sum(p.numel() for p in self.params)
        self.flat = None
        self._pending: list = []
        self._reduced = None          # the flat buffer exchanged during the last backward (kept alive so its address stays unique)
        if hasattr(module, "_resolved_precision"):
            module.__dict__.pop("_fnst_bucket_hook", None)
            if overlap and world > 1:
                module.__dict__["_fnst_bucket_hook"] = self._bucket_ready

    def _bucket_ready(self, flat: torch.Tensor, lo: int, hi: int, last: bool) -> None:
        """Called by the staged backward when flat[lo:hi] holds final local gradients."""
        self._pending.append(dist.all_reduce(flat[lo:hi], op=dist.ReduceOp.SUM, async_op=True))
        if last:
            for work in self._pending:
                work.wait()                       # the current stream waits for the communicator's stream; the host does not block
            self._pending = []
            self._reduced = flat

    def _flat_view(self):
        """If every gradient is a view into one contiguous fp32 buffer in parameter order (the CUDA-graph backward
        returns them that way), return that buffer as a 1-D tensor -- the all-reduce then runs in place, no copies."""
        grads = [p.grad for p in self.params]
        if any(g is None or g.dtype != torch.float32 or not g.is_contiguous() for g in grads):
            return None
        storage = grads[0].untyped_storage()
        offset = grads[0].storage_offset()
        for g in grads:
            if g.untyped_storage().data_ptr() != storage.data_ptr() or g.storage_offset() != offset:
                return None
            offset += g.numel()
        return torch.empty(0, dtype=torch.float32, device=grads[0].device).set_(storage, grads[0].storage_offset(), (self.numel,))

    def all_reduce(self) -> torch.Tensor:
        flat = self._flat_view()
        if flat is not None:
            if self._reduced is not None and flat.data_ptr() == self._reduced.data_ptr():
                self._reduced = None               # exchanged bucket by bucket during the backward
                return flat
            self._reduced = None
            if self.world > 1:
                dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            return flat
        p0 = self.params[0]
        if self.flat is None or self.flat.device != p0.device:
            self.flat = torch.empty(self.numel, dtype=torch.float32, device=p0.device)
        off = 0
        for p in self.params:
            n = p.numel()
            if p.grad is None:
                self.flat[off:off + n].zero_()
            else:
                self.flat[off:off + n].copy_(p.grad.reshape(-1))
            off += n
        if self.world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
        off = 0
        for p in self.params:
            n = p.numel()
            g = self.flat[off:off + n].view_as(p)
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
            off += n
        return self.flat
