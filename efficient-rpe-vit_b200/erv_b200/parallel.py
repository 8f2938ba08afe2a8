"""Batch-sharded data parallelism: flat parameter / gradient buffers and the one collective
of the path, a sum all-reduce of the flat gradient.  Device agnostic on purpose: the same code runs over NCCL on
NVLink (one process per GPU) and over gloo on CPU in the tests."""
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


def world_size(group=None) -> int:
    return dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1


def shard_slice(global_batch: int, rank: int, world: int) -> slice:
    """Contiguous, equal shards; the global batch must divide evenly (the reference drops ragged batches,
    data/datasets.py DataLoader(drop_last=True))."""
    if global_batch % world != 0:
        raise ValueError(f"global batch {global_batch} is not divisible by world size {world}")
    per = global_batch // world
    return slice(rank * per, (rank + 1) * per)


class FlatParams:
    """Re-homes the trainable parameters of a module (and their .grad) as views of two flat fp32 buffers.  Every
    parameter starts on a 16-byte boundary (the kernels read parameter vectors with 128-bit loads; KERPLE's
    rel_pos_bias has 2(2N-1) elements and would misalign everything behind it); the padding stays zero."""

    ALIGN = 4  # floats

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        self.sizes = [p.numel() for p in self.params]
        self.offsets, total = [], 0
        for n in self.sizes:
            self.offsets.append(total)
            total += (n + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.flat = torch.zeros(total, device=dev, dtype=torch.float32)
        self.grad = torch.zeros(total, device=dev, dtype=torch.float32)
        for p, n, off in zip(self.params, self.sizes, self.offsets):
            self.flat[off:off + n].copy_(p.detach().reshape(-1))
            p.data = self.flat[off:off + n].view_as(p)
            p.grad = self.grad[off:off + n].view_as(p)

    def numel(self) -> int:
        return self.flat.numel()

    def zero_grad(self):
        self.grad.zero_()

    def broadcast(self, buffers: Optional[Iterable[torch.Tensor]] = None, src: int = 0, group=None):
        """Identical replicas: rank `src`'s parameters (and buffers such as `omega`) win."""
        if world_size(group) == 1:
            return
        dist.broadcast(self.flat, src=src, group=group)
        for b in buffers or ():
            dist.broadcast(b, src=src, group=group)

    def allreduce_grad(self, group=None):
        """Sum over ranks, in place; the 1/world factor is folded into the optimizer step."""
        if world_size(group) > 1:
            dist.all_reduce(self.grad, group=group)
