"""Headline benchmark: train images/sec (fwd+bwd+Adam) of the reference's model on the B200-native attention path.

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W   # reference arm: the CPU path, host cores

Workload = BASELINE.json configs[1]: performer_favor on CIFAR-10-shape synthetic 32x32x3, patch 4 (65 tokens incl.
CLS), M=256 random features, the reference's dims (dim 32, heads 2, depth 3, mlp 64, dropout 0.1), Adam lr 1e-3,
CrossEntropy.  Per-GPU batch is fixed (weak scaling).  Prints ONE JSON line on rank 0.

Timed region: K steps between barrier+synchronize, CUDA events on the launching stream, max over ranks.  Inputs
rotate through a device-resident pool larger than L2 (value) or through pinned host memory with the loss read back
every step (e2e).  The roofline entry times the dominant attention kernel with CUDA events in a separate
instrumented pass (same shapes, eager launches) so the instrumentation does not perturb `value`.
"""
import argparse
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

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.sm, self.reasons, self.sm_max, self.proc = index, [], set(), None, None
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
            self._stop_evt.wait(0.005)

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

    def stop(self):
        self._stop_evt.set()
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


def cpu_reference_steps(w, batch, steps, warmup, budget_s=None):
    """The reference's CPU path (oracle port, all host threads): images/s over `steps` steps of `batch` images."""
    from oracle import erv_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = model_cfg(w)
    tr = O.CpuTrainer(w["model"], cfg, seed=0)
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
    return done * batch / dt, dt / done * 1e3, done, torch.get_num_threads()


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = args.cpu_batch
    ips, ms, done, cores = cpu_reference_steps(w, batch, args.steps, min(args.warmup, 2))
    line = {
        "impl": "reference", "metric": "train images/sec (fwd+bwd)", "value": ips, "unit": "images/s",
        "n_gpus": args.gpus, "steps": done, "warmup": min(args.warmup, 2), "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["desc"], "batch_per_step": batch},
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": f"{done} training steps (fwd+bwd+Adam) of {batch} images, oracle port of the "
                                   "reference's eager CPU path, all host threads"},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="erv", choices=["erv", "reference"])
    ap.add_argument("--workload", default="config2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=1024, help="per-GPU batch")
    ap.add_argument("--cpu-batch", type=int, default=128, help="images per CPU-baseline step")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, w)
    args.warmup = max(args.warmup, 3)

    import torch.distributed as dist
    from erv_b200 import CIFAR10_CONFIG, _capi, create_model, ops
    from erv_b200.train import Trainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    torch.manual_seed(0)  # same construction seed on every rank
    base = dict(CIFAR10_CONFIG, image_size=w["image"], in_channels=w["channels"])
    acfg = {"num_features": w["num_features"]} if w["num_features"] else None
    model = create_model(w["model"], base, attention_config=acfg, patch_size=w["patch"]).to(dev).train()
    autocast = torch.bfloat16 if w["autocast"] == "bf16" else None
    trainer = Trainer(model, lr=1e-3, use_graph=not args.no_graph, autocast_dtype=autocast)

    B = args.batch
    img_bytes = B * w["channels"] * w["image"] * w["image"] * 4
    pool_n = max(2, min(64, L2_BYTES // img_bytes + 2))  # device pool > L2 so inputs are never L2-resident
    g = torch.Generator(device=dev).manual_seed(1 + rank)
    pool = [(torch.randn(B, w["channels"], w["image"], w["image"], device=dev, generator=g),
             torch.randint(0, 10, (B,), device=dev, generator=g)) for _ in range(pool_n)]
    host_pool = [(i.cpu().pin_memory(), l.cpu().pin_memory()) for i, l in pool[:min(pool_n, 8)]]

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
    for i in range(args.warmup):
        trainer.step(*pool[i % pool_n])
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    _capi.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        loss = trainer.step(*pool[i % pool_n])
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop()
    eager_launches = _capi.launch_count()
    per_step = trainer.kernels_per_step()
    gpu_launches = per_step * args.steps if per_step is not None else eager_launches
    final_loss = float(loss.item())
    ms_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total / 1e3)

    # ---- e2e: host buffers in, loss out, every step ------------------------------------------------------
    # Every step's batch starts in pinned host memory and its loss is read back on the host; the copy of batch i+1 is
    # issued (side stream, double-buffered staging) before the host waits for the loss of batch i, as a data loader
    # with pin_memory / non_blocking copies does.
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()
    for i in range(2):
        trainer.step(*host_pool[i % len(host_pool)])
    barrier()
    t0 = time.perf_counter()
    trainer.prefetch(*host_pool[0])
    for i in range(args.steps):
        l = trainer.step_prefetched()
        if i + 1 < args.steps:
            trainer.prefetch(*host_pool[(i + 1) % len(host_pool)])
        loss_host.copy_(l, non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the caller reads the loss every step
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e = {"value": world * B * args.steps / e2e_s, "unit": "images/s",
           "h2d_bytes_per_step": world * (img_bytes + B * 8), "d2h_bytes_per_step": world * 4}

    # ---- roofline: dominant attention kernel, CUDA events around its launches (separate eager pass) --------
    roofline = None
    ops.PROFILE = {}
    for i in range(6):  # every rank runs the pass (the step contains the gradient all-reduce)
        trainer._step_impl(*pool[i % pool_n])
    torch.cuda.synchronize()
    stats = {k: [a.elapsed_time(b) for a, b in v] for k, v in ops.PROFILE.items()}
    ops.PROFILE = None
    if rank == 0:
        n_tok = (w["image"] // w["patch"]) ** 2 + 1
        esize = 2 if autocast is not None else 4
        per_call = {k: statistics.mean(v[len(v) // 3:]) for k, v in stats.items() if v}  # drop the first third (warm-up)
        if per_call:
            dom = max(per_call, key=lambda k: per_call[k])
            factor = 8 if dom.endswith("_bwd") else 4  # fwd 4*B*N*C*s, bwd 8*B*N*C*s
            alg_bytes = factor * B * n_tok * 32 * esize
            peak, src = 6650.0, "fallback"
            try:
                with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                    peak, src = float(json.load(f)["hbm_gbs"]), "measured"
            except Exception:
                pass
            achieved = alg_bytes / (per_call[dom] * 1e-3) / 1e9
            traffic = None  # dram bytes per launch of this kernel at this workload, from the committed ncu capture
            try:
                with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                    traffic = json.load(f).get(f"{args.workload}:{B}", {}).get(dom)
            except Exception:
                pass
            roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                        "frac": achieved / peak, "traffic": traffic, "peak_source": src, "ms_per_launch": per_call[dom],
                        "algorithmic_bytes": alg_bytes,
                        "attention_ms_per_step": sum(per_call.values()) * 3,
                        "per_call_ms": per_call}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ips, ms, done, cores = cpu_reference_steps(w, args.cpu_batch, 1000, 1, budget_s=15.0)
        cpu_baseline = {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                        "sample": f"{done} training steps (fwd+bwd+Adam) of {args.cpu_batch} images (~15 s), oracle "
                                  "port of the reference's eager CPU path, all host threads"}

    if rank == 0:
        line = {
            "metric": "train images/sec (fwd+bwd)", "value": value, "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if autocast is not None else "f32",
            "data": "synthetic",
            "config": {"workload": w["desc"], "per_gpu_batch": B, "global_batch": B * world, "tokens": None,
                       "parallelism": f"dp{world}", "optimizer": "Adam lr 1e-3 (fused, flat)", "cuda_graph": not args.no_graph,
                       "l2_policy": f"inputs rotate through a {pool_n}-batch device pool ({pool_n * img_bytes >> 20} MiB > L2)"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": gpu_launches, "roofline": roofline,
            "cpu_baseline": cpu_baseline, "final_loss": final_loss,
        }
        line["config"]["tokens"] = (w["image"] // w["patch"]) ** 2 + 1
        print(json.dumps(line), flush=True)
    if world > 1:
        # The captured graph holds NCCL work; drop it before tearing the communicator down, and do not let a slow
        # communicator abort keep the rank alive after the result line has been printed.
        dist.barrier(device_ids=[local])
        torch.cuda.synchronize()
        trainer._graph = None
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
