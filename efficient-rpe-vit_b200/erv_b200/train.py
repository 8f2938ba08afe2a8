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
from .parallel import BucketedReducer, FlatParams, PeerAllReduce, world_size


class _ParamGroup(dict):
    """torch.optim-style parameter group: `for g in trainer.param_groups: g["lr"] = x` reaches the device-resident
    hyper-parameters (and therefore a step that has already been captured in a CUDA graph)."""

    def __init__(self, trainer, **kw):
        super().__init__(**kw)
        self._trainer = trainer

    def __setitem__(self, key, value):
        super().__setitem__(key, value)
        if key in ("lr", "betas", "eps", "weight_decay"):
            self._trainer._set_hyper(**{key: value})


class Trainer:
    """Hyper-parameters live in a 5-float device tensor `[lr, beta1, beta2, eps, weight_decay]` that the fused Adam kernel
    reads at run time (erv_adam_step_dev), so `trainer.lr = x`, `trainer.set_lr(x)` or `param_groups[0]["lr"] = x` take effect
    on the next step whether or not the step has been captured.  torch's LR schedulers insist on a torch.optim.Optimizer, so
    a schedule is attached as a plain callable instead: `trainer.lr_schedule = lambda step: ...` is evaluated on the host
    before every step with the number of steps taken so far (the reference's cosine / warm-up schedules,
    experiments/train.py:216-286, are such functions of the epoch)."""

    def __init__(self, model: torch.nn.Module, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, decoupled_weight_decay: bool = False, use_graph: bool = True,
                 autocast_dtype: Optional[torch.dtype] = None, process_group=None):
        self.model = model
        self._hp = {"lr": float(lr), "betas": (float(betas[0]), float(betas[1])), "eps": float(eps),
                    "weight_decay": float(weight_decay)}
        self.decoupled = decoupled_weight_decay
        self.lr_schedule = None
        self._host_steps = 0
        self.autocast_dtype = autocast_dtype
        self.group = process_group
        self.world = world_size(process_group)
        C.require_cuda(next(model.parameters()))
        # data-parallel runs on NVLink: the flat gradient lives in symmetric memory and is summed by one peer-memory kernel
        self.peer = PeerAllReduce.create(self.group, next(model.parameters()).device)
        self.fp = FlatParams(model.parameters(), grad_alloc=self.peer.alloc if self.peer else None)
        if self.peer:
            self.peer.attach(self.fp)
        self.params, self.flat, self.gflat = self.fp.params, self.fp.flat, self.fp.grad
        dev = self.flat.device
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.step_count = torch.zeros((), device=dev, dtype=torch.int64)
        self.hyper = torch.zeros(5, device=dev, dtype=torch.float32)
        self._push_hyper()
        self.param_groups = [_ParamGroup(self, lr=self.lr, betas=self.betas, eps=self.eps, weight_decay=self.weight_decay,
                                         params=self.params)]
        self.fp.broadcast(model.buffers(), src=0, group=self.group)
        self.use_graph = use_graph
        self._graph = None
        self._static = None
        # modules that redraw their random features every k-th training forward (favor_plus.py:168-171): the decision is a
        # host-side counter and the draw uses the host RNG + QR, neither of which can live inside a captured step
        self._redraw = [m for m in model.modules() if getattr(m, "feature_redraw_interval", None) is not None]
        self.reducer = BucketedReducer(model, self.fp, self.group)

    # ---- hyper-parameters -------------------------------------------------------------------------------
    lr = property(lambda self: self._hp["lr"], lambda self, v: self._set_hyper(lr=v))
    betas = property(lambda self: self._hp["betas"], lambda self, v: self._set_hyper(betas=v))
    eps = property(lambda self: self._hp["eps"], lambda self, v: self._set_hyper(eps=v))
    weight_decay = property(lambda self: self._hp["weight_decay"], lambda self, v: self._set_hyper(weight_decay=v))

    def set_lr(self, lr: float):
        self._set_hyper(lr=lr)

    def _set_hyper(self, **kw):
        for k, v in kw.items():
            self._hp[k] = (float(v[0]), float(v[1])) if k == "betas" else float(v)
            if hasattr(self, "param_groups"):
                dict.__setitem__(self.param_groups[0], k, self._hp[k])
        if hasattr(self, "hyper"):
            self._push_hyper()

    def _push_hyper(self):
        h = self._hp
        self.hyper.copy_(torch.tensor([h["lr"], h["betas"][0], h["betas"][1], h["eps"], h["weight_decay"]],
                                      dtype=torch.float32), non_blocking=False)

    def _before_step(self):
        """Host-side work that must not be recorded: the LR schedule and the feature redraw (done here with a host counter,
        rank 0's draw broadcast to every replica, the module's own counter-driven redraw switched off)."""
        if self.lr_schedule is not None:
            self._set_hyper(lr=self.lr_schedule(self._host_steps))
        for m in self._redraw:
            if self._host_steps % m.feature_redraw_interval == 0:
                m._create_random_features()
                if self.world > 1:
                    import torch.distributed as dist
                    dist.broadcast(m.omega, src=0, group=self.group)
            m.redraw_counter += 1
        self._host_steps += 1

    # ---- one optimisation step on device-resident inputs -------------------------------------------------
    def _step_impl(self, images: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        self.gflat.zero_()
        prev, ops.GRAD_INPLACE = ops.GRAD_INPLACE, True  # fused kernels add straight into the flat gradient buffer
        held = [(m, m.feature_redraw_interval) for m in self._redraw]
        for m, _ in held:
            m.feature_redraw_interval = None  # _before_step() owns the redraw; the module must not sync / redraw in here
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
            self.reducer.backward_and_reduce(loss)
        finally:
            ops.GRAD_INPLACE = prev
            for m, k in held:
                m.feature_redraw_interval = k
        self.step_count += 1
        C.check(C.load().erv_adam_step_dev(C.ptr(self.flat), C.ptr(self.gflat), C.ptr(self.exp_avg),
                                           C.ptr(self.exp_avg_sq), self.flat.numel(), C.ptr(self.hyper), int(self.decoupled),
                                           1.0 / self.world, C.ptr(self.step_count), C.stream()), "adam_step")
        return loss.detach()

    def _capture(self, images: torch.Tensor, labels: torch.Tensor):
        self._static = (torch.empty_like(images), torch.empty_like(labels))
        self._static[0].copy_(images)
        self._static[1].copy_(labels)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # warm-up on a side stream: allocator, cuBLAS handles, lazy kernel loads
            state = [self.flat, self.exp_avg, self.exp_avg_sq, self.step_count, *self.model.buffers()]
            keep = [t.clone() for t in state]
            for _ in range(3):
                self._step_impl(*self._static)
            for dst, src in zip(state, keep):
                dst.copy_(src)  # the warm-up steps must not train (nor move any buffer): the first replay is step 1
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
        self._before_step()
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
        self._before_step()
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
        self._set_hyper(lr=g["lr"], betas=tuple(g["betas"]), eps=g["eps"], weight_decay=g.get("weight_decay", 0.0))
        decoupled = bool(g.get("decoupled_weight_decay", self.decoupled))
        if decoupled != self.decoupled:  # the only hyper-parameter still recorded by value
            self._graph = None
        self.decoupled = decoupled
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
        self._host_steps = int(self.step_count)

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
