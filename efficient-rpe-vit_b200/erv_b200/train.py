"""Data-parallel training step around the attention hot path.

One process per GPU, full model replica, batch sharded by rank.  Parameters and gradients live in two
flat fp32 buffers, so a step is: forward, backward, ONE all-reduce of the flat gradient over NCCL (NVLink 5 /
NVSwitch; skipped for world_size 1), ONE fused Adam kernel (erv_adam_step).  The whole step can be captured
in a CUDA graph, which removes the per-op launch latency that dominates at the reference's tiny dims.

The reference trains with torch.optim.Adam(lr=1e-3) on CrossEntropy (experiments/utils/training.py:53-69,
304-309) in a single process; its per-step .item() syncs are not reproduced -- step() returns a device tensor.
"""
from typing import Optional

import torch
import torch.nn.functional as F

from . import _capi as C
from . import ops
from .parallel import FlatParams, world_size


class Trainer:
    def __init__(self, model: torch.nn.Module, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, decoupled_weight_decay: bool = False, use_graph: bool = True,
                 autocast_dtype: Optional[torch.dtype] = None, process_group=None):
        self.model = model
        self.lr, self.betas, self.eps = lr, betas, eps
        self.weight_decay, self.decoupled = weight_decay, decoupled_weight_decay
        self.autocast_dtype = autocast_dtype
        self.group = process_group
        self.world = world_size(process_group)
        C.require_cuda(next(model.parameters()))
        self.fp = FlatParams(model.parameters())
        self.params, self.flat, self.gflat = self.fp.params, self.fp.flat, self.fp.grad
        dev = self.flat.device
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.step_count = torch.zeros((), device=dev, dtype=torch.int64)
        self.fp.broadcast(model.buffers(), src=0, group=self.group)
        self.use_graph = use_graph
        self._graph = None
        self._static = None

    # ---- one optimisation step on device-resident inputs -------------------------------------------------
    def _step_impl(self, images: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        self.gflat.zero_()
        prev, ops.GRAD_INPLACE = ops.GRAD_INPLACE, True  # fused kernels add straight into the flat gradient buffer
        try:
            if self.autocast_dtype is not None:
                with torch.autocast("cuda", dtype=self.autocast_dtype):
                    if hasattr(self.model, "loss"):
                        loss = self.model.loss(images, labels)
                    else:
                        loss = F.cross_entropy(self.model(images).float(), labels)
            elif hasattr(self.model, "loss"):
                loss = self.model.loss(images, labels)  # head + criterion fused when the model supports it
            else:
                loss = F.cross_entropy(self.model(images).float(), labels)
            loss.backward()
        finally:
            ops.GRAD_INPLACE = prev
        self.fp.allreduce_grad(self.group)
        self.step_count += 1
        C.check(C.load().erv_adam_step(C.ptr(self.flat), C.ptr(self.gflat), C.ptr(self.exp_avg), C.ptr(self.exp_avg_sq),
                                       self.flat.numel(), self.lr, self.betas[0], self.betas[1], self.eps,
                                       self.weight_decay, int(self.decoupled), 1.0 / self.world, 0,
                                       C.ptr(self.step_count), C.stream()), "adam_step")
        return loss.detach()

    def _capture(self, images: torch.Tensor, labels: torch.Tensor):
        self._static = (torch.empty_like(images), torch.empty_like(labels))
        self._static[0].copy_(images)
        self._static[1].copy_(labels)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # warm-up on a side stream: allocator, cuBLAS handles, lazy kernel loads
            keep = [t.clone() for t in (self.flat, self.exp_avg, self.exp_avg_sq, self.step_count)]
            for _ in range(3):
                self._step_impl(*self._static)
            for dst, src in zip((self.flat, self.exp_avg, self.exp_avg_sq, self.step_count), keep):
                dst.copy_(src)  # the warm-up steps must not train: the first replay is optimisation step 1
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._graph = torch.cuda.CUDAGraph()
        C.reset_launch_count()
        with torch.cuda.graph(self._graph):
            self._loss = self._step_impl(*self._static)
        self.graph_kernels = C.launch_count()  # erv kernels recorded in the graph = launched per replay

    def step(self, images: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        """images/labels may be CUDA tensors or pinned host tensors (copied asynchronously).  Returns the loss as
        a device tensor (no host sync)."""
        if not self.use_graph:
            return self._step_impl(images.to(self.flat.device, non_blocking=True),
                                   labels.to(self.flat.device, non_blocking=True))
        if self._graph is None:
            self._capture(images.to(self.flat.device), labels.to(self.flat.device))
        self._static[0].copy_(images, non_blocking=True)
        self._static[1].copy_(labels, non_blocking=True)
        self._graph.replay()
        return self._loss

    # ---- host-fed training: the next batch's host->device copy overlaps the current step ----------------------------
    def prefetch(self, images: torch.Tensor, labels: torch.Tensor):
        """Start the asynchronous copy of a (pinned) host batch into one of two device staging buffers on a side stream.
        The matching step_prefetched() call consumes it.  This is the data-loader pattern of the reference's training
        loop (pin_memory + non_blocking copies, experiments/utils/training.py:53-57) with the copy taken off the compute
        stream."""
        dev = self.flat.device
        if not hasattr(self, "_stage"):
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._stage = [(torch.empty(images.shape, dtype=images.dtype, device=dev),
                            torch.empty(labels.shape, dtype=labels.dtype, device=dev)) for _ in range(2)]
            self._stage_events = [torch.cuda.Event(), torch.cuda.Event()]
            self._stage_free = [torch.cuda.Event(), torch.cuda.Event()]
            self._stage_put, self._stage_get = 0, 0
        slot = self._stage_put % 2
        if self._stage_put >= 2:  # the step that read this slot two prefetches ago must have consumed it
            self._copy_stream.wait_event(self._stage_free[slot])
        with torch.cuda.stream(self._copy_stream):
            self._stage[slot][0].copy_(images, non_blocking=True)
            self._stage[slot][1].copy_(labels, non_blocking=True)
            self._stage_events[slot].record(self._copy_stream)
        self._stage_put += 1

    def step_prefetched(self) -> torch.Tensor:
        """One optimisation step on the oldest prefetched batch."""
        assert hasattr(self, "_stage") and self._stage_get < self._stage_put, "call prefetch() first"
        slot = self._stage_get % 2
        self._stage_get += 1
        cur = torch.cuda.current_stream()
        cur.wait_event(self._stage_events[slot])
        img, lab = self._stage[slot]
        if not self.use_graph:
            loss = self._step_impl(img, lab)
        else:
            if self._graph is None:
                self._capture(img, lab)
            self._static[0].copy_(img, non_blocking=True)  # device-to-device, a few microseconds
            self._static[1].copy_(lab, non_blocking=True)
            self._graph.replay()
            loss = self._loss
        self._stage_free[slot].record(cur)
        return loss

    # ---- optimizer state in torch.optim.Adam's layout (checkpoint interchange) -----------------
    def state_dict(self) -> dict:
        """The optimizer half of a reference checkpoint: what torch.optim.Adam(model.parameters()).state_dict() holds after
        the same steps (experiments/utils/training.py:393-398), so either side can resume the other's run."""
        steps = int(self.step_count)
        state = {}
        if steps > 0:
            for i, (p, n, off) in enumerate(zip(self.params, self.fp.sizes, self.fp.offsets)):
                state[i] = {"step": torch.tensor(float(steps)),
                            "exp_avg": self.exp_avg[off:off + n].view_as(p).clone(),
                            "exp_avg_sq": self.exp_avg_sq[off:off + n].view_as(p).clone()}
        group = {"lr": self.lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": self.weight_decay,
                 "amsgrad": False, "maximize": False, "foreach": None, "capturable": False, "differentiable": False,
                 "fused": None, "decoupled_weight_decay": self.decoupled, "params": list(range(len(self.params)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd: dict):
        """Accepts torch.optim.Adam / AdamW state (one parameter group, no amsgrad).  Copies in place, so a captured graph
        keeps working."""
        groups = sd["param_groups"]
        if len(groups) != 1 or len(groups[0]["params"]) != len(self.params):
            raise ValueError("expected one parameter group covering every trainable parameter, in model.parameters() order")
        g = groups[0]
        if g.get("amsgrad", False) or g.get("maximize", False):
            raise ValueError("amsgrad / maximize are not supported by the fused flat Adam step")
        self.lr, self.betas, self.eps = float(g["lr"]), tuple(g["betas"]), float(g["eps"])
        self.weight_decay = float(g.get("weight_decay", 0.0))
        self.decoupled = bool(g.get("decoupled_weight_decay", self.decoupled))
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        steps = set()
        for i, (p, n, off) in enumerate(zip(self.params, self.fp.sizes, self.fp.offsets)):
            st = sd["state"].get(g["params"][i])
            if st is None:
                steps.add(0)
                continue
            if tuple(st["exp_avg"].shape) != tuple(p.shape):
                raise ValueError(f"optimizer state {i} has shape {tuple(st['exp_avg'].shape)}, parameter {tuple(p.shape)}")
            self.exp_avg[off:off + n].copy_(st["exp_avg"].reshape(-1))
            self.exp_avg_sq[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
            steps.add(int(float(st["step"])))
        if len(steps) != 1:
            raise ValueError(f"parameters are at different step counts {sorted(steps)}; the flat step keeps one counter")
        self.step_count.fill_(steps.pop())
        if self._graph is not None:  # hyper-parameters are baked into a recorded step
            self._graph = None

    def kernels_per_step(self) -> Optional[int]:
        """erv kernels launched per step in graph mode (counted while the graph was recorded)."""
        return getattr(self, "graph_kernels", None)


def save_checkpoint(model: torch.nn.Module, optimizer, epoch: int, metrics: dict, filepath: str,
                    model_name: Optional[str] = None):
    """Same file layout as the reference's save_checkpoint (experiments/utils/training.py:373-412): `optimizer` is a
    Trainer or a torch optimizer; the file loads on either side."""
    ckpt = {"epoch": epoch, "model_state_dict": model.state_dict(), "optimizer_state_dict": optimizer.state_dict(),
            "metrics": metrics}
    name = getattr(model, "model_name", None) or model_name
    if name:
        ckpt["model_name"] = name
    for key in ("attention_type", "rpe_type"):
        if hasattr(model, key):
            ckpt[key] = getattr(model, key)
    torch.save(ckpt, filepath)


def load_checkpoint(model: torch.nn.Module, optimizer, filepath: str):
    """Reference load_checkpoint (experiments/utils/training.py:415-443): returns (epoch, metrics).  Parameters are copied
    in place, so a Trainer's flat buffers (and a captured graph) stay valid.  Pass the Trainer as `optimizer`."""
    ckpt = torch.load(filepath, map_location=next(model.parameters()).device)
    model.load_state_dict(ckpt["model_state_dict"])
    if optimizer is not None:
        optimizer.load_state_dict(ckpt["optimizer_state_dict"])
    return ckpt.get("epoch", 0), ckpt.get("metrics", {})
