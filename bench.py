"""Headline benchmark: train images/sec (fwd+bwd+Adam) of the reference's model on the B200-native attention path.

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W   # reference arm: the CPU path, host cores

Workload = BASELINE.json configs[1]: performer_favor on CIFAR-10-shape synthetic 32x32x3, patch 4 (65 tokens incl.
CLS), M=256 random features, the reference's dims (dim 32, heads 2, depth 3, mlp 64, dropout 0.1), Adam lr 1e-3,
CrossEntropy.  Per-GPU batch is fixed (weak scaling).  Prints ONE JSON line on rank 0.

Timed region: K steps between barrier+synchronize, CUDA events on the launching stream, max over ranks.  Inputs
rotate through a device-resident pool larger than L2 (value) or through pinned host memory with the loss read back
every step (e2e; the K steps are repeated until the loop has run >= 0.5 s so the host-timed figure is stable).  The
roofline entry times the dominant attention kernel with CUDA events in a separate instrumented pass (same shapes,
eager launches) so the instrumentation does not perturb `value`.  `other_configs` times the other BASELINE configs
(1, 3, 4, 4b, 5) in the same run with fewer steps; the headline stays config 2.

The reference arm and the `cpu_baseline` leg run the UNMODIFIED reference (baseline/_ref, installed by
tools/install_reference.py: create_model + torch.optim.Adam + CrossEntropy, experiments/utils/training.py:53-69,304-309)
on the host cores at the GPU arm's batch; the oracle port is used only when baseline/_ref is absent (`kind` says which).
"""
import argparse
import gc
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "efficient-rpe-vit_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

os.environ.setdefault("NCCL_DEBUG", "WARN")  # keep NCCL's version banner off stdout (the result is ONE JSON line)

import torch  # noqa: E402

L2_BYTES = 126 * 1024 * 1024

WORKLOADS = {
    # name: (model, dataset, create_model kwargs, attention_config, autocast)
    "config2": dict(model="performer_favor", image=32, channels=3, patch=4, num_features=256, autocast=None,
                    desc="performer_favor, CIFAR-10-shape synthetic 32x32x3, patch 4 (N=65), M=256, dim 32 heads 2 depth 3"),
    "config1": dict(model="baseline", image=28, channels=1, patch=7, num_features=None, autocast=None,
                    desc="baseline softmax ViT, MNIST-shape synthetic 28x28x1, patch 7 (N=17), dim 32 heads 2 depth 3"),
    "config3": dict(model="performer_relu_most_general", image=32, channels=3, patch=4, num_features=None, autocast=None,
                    desc="performer_relu_most_general (KERPLE), CIFAR-10 shape, patch 4 (N=65), M=44"),
    "config3_p8": dict(model="performer_relu_most_general", image=32, channels=3, patch=8, num_features=None, autocast=None,
                       desc="performer_relu_most_general (KERPLE), CIFAR-10 shape, default patch 8 (N=17), M=44"),
    "config4": dict(model="performer_favor_circulant", image=224, channels=3, patch=16, num_features=None,
                    autocast="bf16", desc="performer_favor_circulant, 224x224x3 patch 16 (N=197), bf16 autocast"),
    "config4b": dict(model="baseline_rope", image=224, channels=3, patch=16, num_features=None, autocast="bf16",
                     desc="baseline_rope, 224x224x3 patch 16 (N=197), bf16 autocast"),
    "config5": dict(model="performer_favor_most_general", image=512, channels=3, patch=8, num_features=None,
                    autocast=None, desc="performer_favor_most_general (KERPLE), 512x512x3 patch 8 (N=4097)"),
}


def model_cfg(w):
    return dict(image_size=w["image"], in_channels=w["channels"], patch_size=w["patch"], num_classes=10, dim=32,
                depth=3, heads=2, mlp_dim=64, dropout=0.1, num_features=w["num_features"])


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons of this rank's GPU, polled through NVML every 5 ms while the timed region runs
    (nvidia-smi -lms as a fallback when the NVML binding is missing)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index, period=0.005):
        super().__init__(daemon=True)
        self.index, self.sm, self.reasons, self.sm_max, self.proc = index, [], set(), None, None
        self.period = period  # NVML calls serialise on a driver lock shared by all ranks: poll more slowly when there are several
        self._stop_evt = threading.Event()
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[index]) if visible and visible.split(",")[index].isdigit() else index
            self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self._nvml = pynvml
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nvml = None

    def _poll_nvml(self):
        n = self._nvml
        bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        while not self._stop_evt.is_set():
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM)))
                r = n.nvmlDeviceGetCurrentClocksEventReasons(self._h) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                self.reasons.update(k for k, b in bits.items() if r & b)
            except Exception:
                pass
            self._stop_evt.wait(self.period)

    def _poll_smi(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                r = [c.strip() for c in line.split(",")]
                if r and r[0].replace(".", "").isdigit():
                    self.sm.append(float(r[0]))
                if len(r) > 1 and r[1].replace(".", "").isdigit():
                    self.sm_max = max(self.sm_max or 0.0, float(r[1]))
                if len(r) >= 7:
                    self.reasons.update(n for n, v in zip(self.NAMES, r[3:7]) if v.lower().startswith("active"))
        except Exception:
            pass

    def run(self):
        if self._nvml is not None:
            self._poll_nvml()
        else:
            self._poll_smi()

    def mark(self):
        """Start of the timed region: samples and throttle reasons seen before this call are dropped."""
        self._mark = len(self.sm)
        self.reasons.clear()

    def stop(self):
        self._stop_evt.set()
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        sm = self.sm[getattr(self, "_mark", 0):] or self.sm
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(sm)}


REF_DIR = os.path.join(ROOT, "baseline", "_ref")


class _ReferenceTrainer:
    """The reference itself (baseline/_ref): create_model + Adam(lr=1e-3) + CrossEntropy, train mode, on CPU.  Runs in THIS
    process with baseline/_ref first on sys.path; only the reference arm / cpu_baseline leg construct it."""

    def __init__(self, w, seed=0):
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.") or k == "configs" or k.startswith("configs.")]:
            del sys.modules[k]
        sys.path.insert(0, REF_DIR)
        try:
            import importlib
            models = importlib.import_module("models")
            cfgs = importlib.import_module("configs")
            assert os.path.realpath(models.__file__).startswith(os.path.realpath(REF_DIR)), models.__file__
            import copy
            base = copy.deepcopy(cfgs.CIFAR10_CONFIG)
            base.update(image_size=w["image"], in_channels=w["channels"])
            torch.manual_seed(seed)
            acfg = {"num_features": w["num_features"]} if w["num_features"] else None
            self.model = models.create_model(w["model"], base, attention_config=acfg, patch_size=w["patch"]).train()
        finally:
            sys.path.remove(REF_DIR)
            for k in [k for k in sys.modules if k == "models" or k.startswith("models.") or k == "configs" or k.startswith("configs.")]:
                del sys.modules[k]  # the GPU arm must not see the reference's packages
        self.opt = torch.optim.Adam(self.model.parameters(), lr=1e-3)
        self.autocast = w["autocast"] == "bf16"

    def step(self, images, labels):
        if self.autocast:
            with torch.autocast("cpu", dtype=torch.bfloat16):
                loss = torch.nn.functional.cross_entropy(self.model(images).float(), labels)
        else:
            loss = torch.nn.functional.cross_entropy(self.model(images), labels)
        self.opt.zero_grad()
        loss.backward()
        self.opt.step()
        return float(loss.item())  # the reference's loop reads the loss every step (training.py:66-69)


def cpu_reference_steps(w, batch, steps, warmup, budget_s=None):
    """The reference's CPU path, all host threads: images/s over `steps` steps of `batch` images.
    Returns (images/s, ms/step, steps done, threads, kind)."""
    torch.set_num_threads(os.cpu_count() or 1)
    kind = "reference" if os.path.isdir(os.path.join(REF_DIR, "models")) else "port"
    if kind == "reference":
        tr = _ReferenceTrainer(w)
    else:
        from oracle import erv_oracle as O
        tr = O.CpuTrainer(w["model"], model_cfg(w), seed=0)
    g = torch.Generator().manual_seed(0)
    img = torch.randn(batch, w["channels"], w["image"], w["image"], generator=g)
    lab = torch.randint(0, 10, (batch,), generator=g)
    for _ in range(warmup):
        tr.step(img, lab)
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        tr.step(img, lab)
        done += 1
        if budget_s is not None and time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return done * batch / dt, dt / done * 1e3, done, torch.get_num_threads(), kind


def _kind_text(kind):
    return ("the UNMODIFIED reference (baseline/_ref: create_model + torch.optim.Adam + CrossEntropy, eager CPU)"
            if kind == "reference" else "oracle port of the reference's eager CPU path (baseline/_ref absent)")


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = args.cpu_batch if args.cpu_batch else args.batch  # same batch per step as the GPU arm unless overridden
    warm = min(args.warmup, 2)
    ips, ms, done, cores, kind = cpu_reference_steps(w, batch, args.steps, warm, budget_s=args.cpu_budget)
    small = None
    if batch != 128 and args.workload == "config2":  # second figure: the reference config's own small batch favours the CPU
        ips2, ms2, done2, _, _ = cpu_reference_steps(w, 128, 1000, 1, budget_s=8.0)
        small = {"batch_per_step": 128, "value": ips2, "ms_per_step": ms2, "steps": done2}
    line = {
        "impl": "reference", "metric": "train images/sec (fwd+bwd)", "value": ips, "unit": "images/s",
        "n_gpus": args.gpus, "steps": done, "warmup": warm, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if w["autocast"] == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": w["desc"], "batch_per_step": batch, "per_gpu_batch": batch},
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": kind,
                         "sample": f"{done} training steps (fwd+bwd+Adam) of {batch} images, {_kind_text(kind)}, all host "
                                   f"threads; stopped after {args.cpu_budget:.0f} s if not finished",
                         "batch_128": small},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def _peaks():
    hbm, tens, src = 6650.0, 1590.0, "fallback"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            d = json.load(f)
        hbm, tens, src = float(d["hbm_gbs"]), float(d["bf16_tflops"]), "measured"
    except Exception:
        pass
    return hbm, tens, src


def attention_flops(key, w, B, n_tok, M):
    """Algorithmic FLOPs of one attention call, H = 2, Dh = 16."""
    H, Dh = 2, 16
    bwd = key.endswith("_bwd")
    if key.startswith("linear"):
        return (16 if bwd else 8) * B * H * n_tok * Dh * M
    if key.startswith("softmax"):
        return (10 if bwd else 4) * B * H * n_tok * n_tok * Dh
    fwd = 2 * B * H * n_tok * n_tok * (M + Dh) + 4 * B * H * n_tok * Dh * M
    # backward of the tile route: P recomputed, dA = dnum V^T (+ dden), dV = A^T dnum, dphi_q, dphi_k, feature maps twice
    return 2 * B * H * n_tok * n_tok * (3 * M + 2 * Dh) + 8 * B * H * n_tok * Dh * M if bwd else fwd


def measure(wname, B, steps, warmup, env, do_e2e=True, use_graph=True):
    """Times `steps` training steps of one workload on this rank's GPU (all ranks call it together).  Returns a dict."""
    import torch.distributed as dist
    from erv_b200 import CIFAR10_CONFIG, _capi, create_model, ops
    from erv_b200.train import Trainer
    w = WORKLOADS[wname]
    world, rank, local, dev = env["world"], env["rank"], env["local"], env["dev"]

    torch.manual_seed(0)  # same construction seed on every rank
    base = dict(CIFAR10_CONFIG, image_size=w["image"], in_channels=w["channels"])
    acfg = {"num_features": w["num_features"]} if w["num_features"] else None
    model = create_model(w["model"], base, attention_config=acfg, patch_size=w["patch"]).to(dev).train()
    autocast = torch.bfloat16 if w["autocast"] == "bf16" else None
    trainer = Trainer(model, lr=1e-3, use_graph=use_graph, autocast_dtype=autocast)

    img_bytes = B * w["channels"] * w["image"] * w["image"] * 4
    pool_n = max(2, min(64, L2_BYTES // img_bytes + 2))  # device pool > L2 so inputs are never L2-resident
    g = torch.Generator(device=dev).manual_seed(1 + rank)
    pool = [(torch.randn(B, w["channels"], w["image"], w["image"], device=dev, generator=g),
             torch.randint(0, 10, (B,), device=dev, generator=g)) for _ in range(pool_n)]

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- value: inputs resident in HBM ------------------------------------------------------------------
    for i in range(warmup):
        trainer.step(*pool[i % pool_n])
    # the NVML sampler is created and started BEFORE the barrier: nvmlInit takes a different number of milliseconds on every
    # rank, and a rank that enters the timed loop late makes all the others wait for it inside their first all-reduce
    sampler = ClockSampler(local, 0.005 if world == 1 else 0.02)
    sampler.start()
    gc.collect()
    gc.disable()  # no collector pause on one rank while the others wait in a collective
    barrier()
    sampler.mark()
    _capi.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        loss = trainer.step(*pool[i % pool_n])
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop()
    gc.enable()
    eager_launches = _capi.launch_count()
    per_step = trainer.kernels_per_step()
    res = {"ms_per_step": ms_total / steps, "value": world * B * steps / (ms_total / 1e3), "clocks": clocks,
           "gpu_launches": per_step * steps if per_step is not None else eager_launches,
           "final_loss": float(loss.item()), "pool_n": pool_n, "img_bytes": img_bytes,
           "dtype": "bf16" if autocast is not None else "f32", "tokens": (w["image"] // w["patch"]) ** 2 + 1}

    # ---- e2e: host buffers in, loss out, every step ------------------------------------------------------
    # Every step's batch starts in pinned host memory and every step's loss is read back on the host.  The copy of batch
    # i+1 is issued on a side stream (double-buffered staging) as a data loader with pin_memory / non_blocking copies does;
    # the loss of step i is copied to pinned memory right behind the step and read by the host while step i+1 runs (one
    # step behind, like an asynchronous logger), so the GPU does not idle for a host round trip between steps.  The K-step
    # loop is repeated until it has run for >= 0.5 s.
    if do_e2e:
        host_pool = [(i.cpu().pin_memory(), l.cpu().pin_memory()) for i, l in pool[:min(pool_n, 8)]]
        loss_host = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]
        loss_ready = [torch.cuda.Event(), torch.cuda.Event()]
        for i in range(3):  # warm-up through the SAME path as the timed loop: staging buffers, copy stream, events, pinned slots
            trainer.prefetch(*host_pool[i % len(host_pool)])
            l = trainer.step_prefetched()
            loss_host[i % 2].copy_(l, non_blocking=True)
            loss_ready[i % 2].record()
        torch.cuda.synchronize()
        reps = max(1, int(0.5 / max(1e-6, ms_total / 1e3)) + 1)
        n_e2e = steps * reps
        gc.collect()
        gc.disable()
        barrier()
        t0 = time.perf_counter()
        trainer.prefetch(*host_pool[0])
        seen = 0.0
        sync_every_step = bool(os.environ.get("ERV_E2E_SYNC"))
        for i in range(n_e2e):
            l = trainer.step_prefetched()
            loss_host[i % 2].copy_(l, non_blocking=True)  # device -> pinned host, right behind the step
            loss_ready[i % 2].record()
            if i + 1 < n_e2e:
                trainer.prefetch(*host_pool[(i + 1) % len(host_pool)])
            if sync_every_step:  # ERV_E2E_SYNC=1: wait for this step's loss before launching the next step
                loss_ready[i % 2].synchronize()
            elif i > 0:  # the caller reads EVERY step's loss; the read of step i-1 happens while step i runs (an asynchronous logger)
                loss_ready[(i - 1) % 2].synchronize()
                seen += float(loss_host[(i - 1) % 2])
        loss_ready[(n_e2e - 1) % 2].synchronize()
        seen += float(loss_host[(n_e2e - 1) % 2])
        torch.cuda.current_stream().synchronize()
        gc.enable()
        assert seen == seen, "e2e loop read a NaN loss"
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        res["e2e"] = {"value": world * B * n_e2e / e2e_s, "unit": "images/s",
                      "h2d_bytes_per_step": world * (img_bytes + B * 8), "d2h_bytes_per_step": world * 4,
                      "steps_timed": n_e2e, "loss_read": "every step, by the host, one step behind (pinned copy + event)"}

    # ---- roofline: dominant attention kernel, CUDA events around its launches (separate eager pass) --------
    ops.PROFILE = {}
    for i in range(6):  # every rank runs the pass (the step contains the gradient all-reduce)
        # keep the GPU busy while the host enqueues the eager step, so that every timed kernel is already queued when its
        # predecessor ends and the events bracket its execution only (not the launch latency of an idle stream)
        torch.cuda._sleep(int(2.0e7))
        trainer._step_impl(*pool[i % pool_n])
    torch.cuda.synchronize()
    stats = {k: [a.elapsed_time(b) for a, b in v] for k, v in ops.PROFILE.items()}
    ops.PROFILE = None
    res["roofline"] = None
    if rank == 0:
        n_tok = res["tokens"]
        esize = 2 if autocast is not None else 4
        per_call = {k: statistics.mean(v[len(v) // 3:]) for k, v in stats.items() if v}  # drop the first third (warm-up)
        if per_call:
            dom = max(per_call, key=lambda k: per_call[k])
            factor = 8 if dom.endswith("_bwd") else 4  # fwd 4*B*N*C*s, bwd 8*B*N*C*s
            alg_bytes = factor * B * n_tok * 32 * esize
            M = w["num_features"] or 44
            flops = attention_flops(dom, w, B, n_tok, M)
            hbm, tens, src = _peaks()
            t = per_call[dom] * 1e-3
            gbs, tfs = alg_bytes / t / 1e9, flops / t / 1e12
            # tensor-pipe-bound only for softmax / KERPLE tiles at N >= 4097; everything else HBM-bound
            tensor_bound = n_tok >= 1024 and not dom.startswith("linear")
            traffic, traffic_source = None, None
            try:  # dram bytes per launch of this kernel from a COMMITTED ncu capture (not measured by this run)
                with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                    ent = json.load(f).get(f"{wname}:{B}", {})
                traffic, traffic_source = ent.get(dom), ent.get("source")
            except Exception:
                pass
            res["roofline"] = {
                "bound": "tensor" if tensor_bound else "hbm", "kernel": dom,
                "achieved": tfs if tensor_bound else gbs, "peak": tens if tensor_bound else hbm,
                "unit": "TFLOP/s" if tensor_bound else "GB/s",
                "frac": (tfs / tens) if tensor_bound else (gbs / hbm), "traffic": traffic,
                "traffic_source": traffic_source, "peak_source": src, "ms_per_launch": per_call[dom],
                "algorithmic_bytes": alg_bytes, "algorithmic_flops": flops, "hbm_gbs": gbs, "tflops": tfs,
                "attention_ms_per_step": sum(per_call.values()) * 3, "per_call_ms": per_call}
    if world > 1:
        barrier()
    trainer._graph = None
    del trainer, model, pool
    torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="erv", choices=["erv", "reference"])
    ap.add_argument("--workload", default="config2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=1024, help="per-GPU batch")
    ap.add_argument("--cpu-batch", type=int, default=0, help="images per CPU step (0 = the GPU arm's per-GPU batch)")
    ap.add_argument("--cpu-budget", type=float, default=150.0, help="reference arm: stop after this many seconds")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, w)
    args.warmup = max(args.warmup, 3)

    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    env = {"world": world, "rank": rank, "local": local, "dev": dev}

    B = args.batch
    head = measure(args.workload, B, args.steps, args.warmup, env, do_e2e=True, use_graph=not args.no_graph)

    # ---- the other BASELINE configs, same run, fewer steps (per-GPU batch in brackets) --------------------------
    other = {}
    if args.workload == "config2" and not args.no_other_configs:
        plan = [("config1", 128), ("config3", 1024), ("config3_p8", 1024), ("config4", 256), ("config4b", 256), ("config5", 2)]
        if world >= 8:
            plan.append(("config5", 8))
        k = max(3, min(args.steps, 20))
        for name, b in plan:
            try:
                r = measure(name, b, k, 3, env, do_e2e=False)
                rf = r["roofline"]
                other[f"{name}:b{b}"] = {
                    "workload": WORKLOADS[name]["desc"], "per_gpu_batch": b, "value": r["value"], "unit": "images/s",
                    "ms_per_step": r["ms_per_step"], "steps": k, "dtype": r["dtype"], "tokens": r["tokens"],
                    "roofline": None if rf is None else {kk: rf[kk] for kk in
                                                         ("bound", "kernel", "achieved", "peak", "unit", "frac",
                                                          "ms_per_launch", "hbm_gbs", "tflops", "per_call_ms")}}
            except Exception as exc:  # a config that cannot run must not take the headline down with it
                other[f"{name}:b{b}"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
                torch.cuda.empty_cache()

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb = args.cpu_batch if args.cpu_batch else B
        ips, ms, done, cores, kind = cpu_reference_steps(w, cb, 1000, 1, budget_s=20.0)
        cpu_baseline = {"value": ips, "unit": "images/s", "cores": cores, "kind": kind, "ms_per_step": ms,
                        "sample": f"{done} training steps (fwd+bwd+Adam) of {cb} images (~20 s), {_kind_text(kind)}, "
                                  "all host threads"}

    if rank == 0:
        line = {
            "metric": "train images/sec (fwd+bwd)", "value": head["value"], "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": head["dtype"], "data": "synthetic",
            "config": {"workload": w["desc"], "per_gpu_batch": B, "global_batch": B * world, "tokens": head["tokens"],
                       "parallelism": f"dp{world}", "optimizer": "Adam lr 1e-3 (fused, flat)", "cuda_graph": not args.no_graph,
                       "l2_policy": f"inputs rotate through a {head['pool_n']}-batch device pool "
                                    f"({head['pool_n'] * head['img_bytes'] >> 20} MiB > L2)"},
            "clocks": head["clocks"], "e2e": head.get("e2e"), "gpu_launches": head["gpu_launches"],
            "roofline": head["roofline"], "cpu_baseline": cpu_baseline, "final_loss": head["final_loss"],
            "other_configs": other,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # Captured graphs held NCCL work; they are dropped in measure().  Do not let a slow communicator abort keep the rank
        # alive after the result line has been printed.
        dist.barrier(device_ids=[local])
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
