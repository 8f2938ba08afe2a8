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

    def __init__(self, params: Iterable[torch.nn.Parameter], grad_alloc=None):
        """grad_alloc(total) -> zero-filled fp32 tensor of >= total elements that will hold the flat gradient (used to place
        it in NVLink-mapped symmetric memory, PeerAllReduce); default: a plain device tensor."""
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
        self.grad = torch.zeros(total, device=dev, dtype=torch.float32) if grad_alloc is None else grad_alloc(total)[:total]
        self.reduce_hook = None  # set by PeerAllReduce: replaces the NCCL all-reduce of allreduce_grad()
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
            if self.reduce_hook is not None:
                self.reduce_hook()
            else:
                dist.all_reduce(self.grad, group=group)


class PeerAllReduce:
    """The flat gradient in NVLink-mapped symmetric memory + the one-shot sum kernel erv_allreduce_oneshot.

    At 28-57 k floats the NCCL all-reduce is pure latency (0.15 ms of a 1.75 ms step on 8 GPUs in round 1).  Here every rank
    allocates its gradient buffer with torch.distributed._symmetric_memory, every process maps every peer's buffer, and one
    8-CTA kernel per rank waits for the peers' gradients, adds the `world` buffers in rank order over NVLink (bit-identical
    sums on every rank) and leaves the result in the local gradient buffer.  CUDA-graph capturable (no host round trip; the
    launch counter lives on the device).  Use: `peer = PeerAllReduce.create(group)`, `fp = FlatParams(params, peer.alloc)`,
    `peer.attach(fp)`; create() returns None where the path does not apply (one rank, CPU tensors, more than 8 ranks,
    symmetric memory unavailable, ERV_NCCL_ALLREDUCE=1), and FlatParams then all-reduces through torch.distributed."""

    def __init__(self, group, device):
        self.group, self.device = group, device
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.buf = self.hdl = self.n = None

    @classmethod
    def create(cls, group=None, device=None):
        import os
        if world_size(group) < 2 or world_size(group) > 8 or os.environ.get("ERV_NCCL_ALLREDUCE"):
            return None
        if device is None or torch.device(device).type != "cuda" or dist.get_backend(group) != "nccl":
            return None
        try:
            import importlib
            importlib.import_module("torch.distributed._symmetric_memory")
        except Exception:
            return None
        return cls(group if group is not None else dist.group.WORLD, torch.device(device))

    def alloc(self, total: int) -> torch.Tensor:
        import importlib
        symm = importlib.import_module("torch.distributed._symmetric_memory")
        from . import _capi as C
        self.n = (total + 3) // 4 * 4
        nflag = C.load().erv_allreduce_flag_floats()
        self.buf = symm.empty(self.n + nflag, dtype=torch.float32, device=self.device)
        self.buf.zero_()
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)  # every rank's flags are zero before anyone can signal
        self.hdl = symm.rendezvous(self.buf, self.group)
        import ctypes
        self.ptrs = (ctypes.c_void_p * self.world)(*[int(q) for q in self.hdl.buffer_ptrs])
        self.scratch = torch.zeros(self.n, dtype=torch.float32, device=self.device)
        self.epoch = torch.zeros(1, dtype=torch.int32, device=self.device)
        return self.buf

    def attach(self, fp: "FlatParams"):
        if fp.grad.data_ptr() != self.buf.data_ptr():
            raise ValueError("the flat gradient was not allocated by this PeerAllReduce")
        fp.reduce_hook = self.allreduce

    def allreduce(self):
        from . import _capi as C
        C.check(C.load().erv_allreduce_oneshot(self.ptrs, self.n, self.n, C.ptr(self.scratch), self.rank, self.world,
                                               C.ptr(self.epoch), C.stream()), "allreduce_oneshot")


class BucketedReducer:
    """Backward + gradient all-reduce in two buckets, the first one under the rest of the backward.

    The flat gradient is split where `model.transformer_blocks[1]` starts.  Everything from there on (blocks 1.., final norm,
    head) has its gradient after the first part of the backward; its all-reduce is launched asynchronously (on the process
    group's own stream, forked from the step's stream) and runs under the backward of block 0 and the patch embedding; the
    second bucket (class token, positions, embedding, block 0) follows.  Both are called from the thread that runs the step,
    so they are recorded when the step is captured in a CUDA graph.  The backward is cut at the output of block 0, which the
    model records as `model._cut_tensor` when `model._cut_after == 0` (erv_b200.vit.BaseViT.features).

    Measured on 2 B200 (gpurun_out/r2_dp_*_n2.json): the split costs more than it hides at the reference dims (1.674 ms per
    step against 1.587 with one all-reduce: two latency-bound NCCL calls instead of one, and the cut adds an autograd pass), so
    it is opt-in (ERV_BUCKET_ALLREDUCE=1).  Otherwise, and when there is nothing to overlap (one rank, fewer than two blocks,
    a parameter order that does not follow the module order): `loss.backward()` + one all-reduce."""

    def __init__(self, model: torch.nn.Module, fp: FlatParams, group=None):
        import os
        self.model, self.fp, self.group = model, fp, group
        self.split = None
        blocks = getattr(model, "transformer_blocks", None)
        if world_size(group) == 1 or blocks is None or len(blocks) < 2 or not os.environ.get("ERV_BUCKET_ALLREDUCE"):
            return
        first = next(iter(blocks[1].parameters()), None)
        idx = next((i for i, q in enumerate(fp.params) if q is first), None) if first is not None else None
        if not idx:
            return
        late = {id(q) for b in list(blocks)[1:] for q in b.parameters()}
        head = getattr(model, "mlp_head", None)
        if head is not None:
            late |= {id(q) for q in head.parameters()}
        if any(id(q) not in late for q in fp.params[idx:]) or any(id(q) in late for q in fp.params[:idx]):
            return
        self.split = (idx, fp.offsets[idx])
        model._cut_after = 0

    def backward_and_reduce(self, loss: torch.Tensor):
        cut = getattr(self.model, "_cut_tensor", None)
        if self.split is None or cut is None:
            loss.backward()
            self.fp.allreduce_grad(self.group)
            return
        idx, off = self.split
        self.model._cut_tensor = None
        cut.retain_grad()
        torch.autograd.backward(loss, inputs=[*self.fp.params[idx:], cut], retain_graph=True)
        work = dist.all_reduce(self.fp.grad[off:], group=self.group, async_op=True)
        torch.autograd.backward(cut, grad_tensors=cut.grad, inputs=self.fp.params[:idx])
        cut.grad = None
        work.wait()
        dist.all_reduce(self.fp.grad[:off], group=self.group)
