#!/usr/bin/env python3
"""Benchmark of the UNet watermark-mask inference hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A "step" = one forward of smp-Unet-resnet34 over one batch of 16 synthetic 512x512 RGB images per GPU
(BASELINE.json configs[1]) ending in the fused sigmoid+threshold uint8 mask.  Data-parallel, weak
scaling, no collective on the data path.  Prints ONE JSON line on rank 0.

  value     images/s, whole job, inputs already resident in HBM (uint8 NHWC), CUDA-event timed
  e2e       images/s through the public API (Unet.predict_mask) from pinned HOST uint8 batches:
            H2D copy + forward + D2H read of the uint8 masks inside the timed region
  roofline  the tcgen05 implicit-GEMM conv kernel (conv_halo_kernel, every conv launch of a step; conv_tc_kernel
            only serves shapes it does not cover): algorithmic conv FLOPs / (timed step x live conv share),
            vs the measured bf16 peak
  cpu_baseline / --impl reference
            the oracle restatement of the reference's CPU path (torch fp32, all host threads) on a
            bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "images/sec UNet-ResNet34 512x512 bf16 mask inference"
UNIT = "images/s"
# BASELINE.json configs (1-based, as VERDICT/SURVEY number them).  The default line is configs[1] = "config 2".
CONFIGS = {
    2: dict(encoder="resnet34", size=512, batch=16, scaling="weak", metric=METRIC,
            workload="configs[1]: smp Unet resnet34, 512x512, batch 16 per GPU, bf16 inference, sigmoid+threshold uint8 mask"),
    3: dict(encoder="resnet34", size=1024, batch=64, scaling="strong",
            metric="images/sec UNet-ResNet34 1024x1024 bf16 mask inference",
            workload="configs[2]: smp Unet resnet34, 1024x1024, batch 64 TOTAL split data-parallel over the GPUs "
                     "(64/N images per GPU), bf16 inference, sigmoid+threshold uint8 mask"),
    5: dict(encoder="resnet34", size=512, batch=16, scaling="weak", train=True,
            metric="images/sec UNet-ResNet34 512x512 training step (Dice+BCE, Adam, gradient all-reduce)",
            workload="configs[4]: smp Unet resnet34, 512x512, batch 16 per GPU, training step with Dice+BCE loss and NCCL "
                     "gradient allreduce"),
    4: dict(encoder="resnet50", size=768, batch=32, scaling="weak",
            metric="images/sec UNet-ResNet50 768x768 bf16 mask inference",
            workload="configs[3]: smp Unet resnet50, 768x768, batch 32 per GPU, bf16 inference, sigmoid+threshold uint8 mask"),
}
ENCODER, SIZE, BATCH = "resnet34", 512, 16
WORKLOAD = CONFIGS[2]["workload"]
SCALING = "weak"


def select_config(n: int, world: int):
    """Bind the module-level workload constants to BASELINE config n (per-GPU batch for the strong-scaling config 3)."""
    global ENCODER, SIZE, BATCH, WORKLOAD, METRIC, SCALING
    c = CONFIGS[n]
    ENCODER, SIZE, WORKLOAD, METRIC, SCALING = c["encoder"], c["size"], c["workload"], c["metric"], c["scaling"]
    BATCH = c["batch"] // world if c["scaling"] == "strong" else c["batch"]
    if BATCH < 1:
        raise SystemExit(f"config {n}: {c['batch']} images do not split over {world} GPUs")


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"bf16_tflops": float(d["bf16_tflops"]), "bf16_tflops_sustained": float(d.get("bf16_tflops_sustained", 0)),
                "hbm_gbs": float(d["hbm_gbs"]), "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0,
            "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU during the timed region (NVML, ~5 ms period)."""

    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz, self.ok = [], set(), None, False
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:  # noqa: BLE001
            self.ok = False
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    bits = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:  # noqa: BLE001
                    bits = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for b, name in self.REASONS.items():
                    if bits & b and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.005)

    def start(self):
        if self.ok:
            self.t.start()

    def stop(self):
        self._stop.set()
        if self.ok:
            self.t.join(timeout=1.0)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": int(statistics.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cpu_reference_run(steps: int, warmup: int, images_per_step: int):
    """The reference's CPU path (oracle restatement, torch fp32, all host threads): img/s."""
    import torch
    from oracle import unet_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = O.build(ENCODER, seed=0, random_bn=True)
    x = O.image_like_input(images_per_step, SIZE, seed=1)
    with torch.no_grad():
        for _ in range(warmup):
            O.binarize(model(x)[:, 0])
        t0 = time.perf_counter()
        for _ in range(steps):
            O.binarize(model(x)[:, 0])
        dt = time.perf_counter() - t0
    return images_per_step * steps / dt, dt / steps * 1e3, torch.get_num_threads()


def gpu_control_run(dev, u8_batch, steps: int):
    """Same-GPU control: the oracle module (torchvision ResNet + restated smp decoder) run by stock torch as bf16
    channels_last - i.e. cuDNN / cuBLAS kernels - on the same uint8 inputs, normalisation and threshold included.
    Not part of the product path; it is what `model.to('cuda').to(bfloat16)` of the reference would run."""
    import torch
    from oracle import unet_oracle as O
    try:
        torch.backends.cudnn.benchmark = True
        ref = O.build(ENCODER, seed=0, random_bn=True).to(dev).to(torch.bfloat16).to(memory_format=torch.channels_last)
        mean = torch.tensor(O.IMAGENET_MEAN, device=dev).view(1, 3, 1, 1) * 255.0
        inv = 1.0 / (torch.tensor(O.IMAGENET_STD, device=dev).view(1, 3, 1, 1) * 255.0)
        n = u8_batch.shape[0]
        chunk = n if n * SIZE * SIZE <= 16 * 1024 * 1024 else max(1, (16 * 1024 * 1024) // (SIZE * SIZE))

        def step():
            out = []
            for i in range(0, n, chunk):
                x = ((u8_batch[i:i + chunk].permute(0, 3, 1, 2).float() - mean) * inv).to(torch.bfloat16)
                x = x.contiguous(memory_format=torch.channels_last)
                out.append((ref(x)[:, 0] > 0).to(torch.uint8) * 255)
            return out
        with torch.no_grad():
            for _ in range(3):
                step()
            torch.cuda.synchronize(dev)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(steps):
                step()
            b.record()
            torch.cuda.synchronize(dev)
        ms = a.elapsed_time(b) / steps
        del ref
        torch.cuda.empty_cache()
        return {"value": n / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps, "images_per_step": n,
                "impl": "stock torch bf16 channels_last (cuDNN), oracle module, same GPU, same uint8 inputs, eager"}
    except Exception as e:  # noqa: BLE001 - the control must never take the bench line down
        return {"unavailable": f"{type(e).__name__}: {e}"[:200]}


def cpu_train_reference_run(steps: int, warmup: int, images_per_step: int):
    """Config 5 on the host cores: the oracle module trained the reference's way (train mode, Dice+BCE, Adam)."""
    import torch
    from oracle import unet_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = O.build(ENCODER, seed=0, random_bn=True).train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=1e-4)
    x = O.image_like_input(images_per_step, SIZE, seed=1)
    t = (torch.rand(images_per_step, 1, SIZE, SIZE, generator=torch.Generator().manual_seed(2)) > 0.85).float()

    def step():
        opt.zero_grad()
        loss = O.dice_bce_loss(model(x), t)
        loss.backward()
        opt.step()
        return loss.item()
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return images_per_step * steps / dt, dt / steps * 1e3, torch.get_num_threads()


def roofline_by_bound(launch_rows, peaks, step_ms):
    """Which roofline binds each launch of the step, layer by layer: a launch is HBM-bound when its algorithmic bytes
    (input + output + residual + weights, each once) / measured copy bandwidth exceed its algorithmic FLOPs / measured
    bf16 peak.  Per-launch times come from the eager CUDA-event pass and are scaled to the timed (graph-replayed) step.
    ``floor_ms`` = sum of the per-launch roofline floors = what an unfused layer-by-layer execution could reach."""
    pt, pb = peaks["bf16_tflops"] * 1e12, peaks["hbm_gbs"] * 1e9
    eager = sum(r[0] for r in launch_rows)
    scale = step_ms / eager if eager > 0 else 0.0
    out = {"hbm": [0.0, 0.0, 0], "tensor": [0.0, 0.0, 0]}          # [ms in the timed step, work, launches]
    floor = 0.0
    for ms, fl, by in launch_rows:
        t_t, t_b = fl / pt * 1e3, by / pb * 1e3
        floor += max(t_t, t_b)
        k = "hbm" if t_b > t_t else "tensor"
        out[k][0] += ms * scale
        out[k][1] += by if k == "hbm" else fl
        out[k][2] += 1
    h, t = out["hbm"], out["tensor"]
    gbs = h[1] / (h[0] * 1e-3) / 1e9 if h[0] > 0 else 0.0
    tfs = t[1] / (t[0] * 1e-3) / 1e12 if t[0] > 0 else 0.0
    return {"hbm_bound": {"launches": h[2], "time_share": h[0] / step_ms if step_ms > 0 else 0.0, "achieved_gbs": gbs,
                          "frac_of_hbm_peak": gbs / peaks["hbm_gbs"]},
            "tensor_bound": {"launches": t[2], "time_share": t[0] / step_ms if step_ms > 0 else 0.0,
                             "achieved_tflops": tfs, "frac_of_bf16_peak": tfs / peaks["bf16_tflops"]},
            "floor_ms": floor, "step_over_floor": step_ms / floor if floor > 0 else None,
            "how": "per launch: max(algorithmic FLOPs / bf16 peak, algorithmic bytes / HBM copy peak); eager CUDA-event "
                   "times scaled by (timed step / eager sum); floor_ms = sum of the per-launch floors (no cross-layer fusion)"}


def run_train(args):
    """BASELINE configs[4]: one optimisation step per 'step' (forward in train mode on the tcgen05 conv kernels,
    Dice+BCE, torch-autograd backward, bucketed NCCL gradient all-reduce overlapped with the backward, Adam)."""
    import torch
    import torch.distributed as dist
    from unet_watermark_b200 import _lib
    from unet_watermark_b200.training import TrainStep
    from unet_watermark_b200.unet_model import Unet
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    select_config(5, world)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()
    lib = _lib.load()
    torch.manual_seed(0)                                   # identical replicas on every rank
    model = Unet(ENCODER, encoder_weights=None).to(dev)
    ts = TrainStep(model)
    n_pool = 4
    gi = torch.Generator().manual_seed(100 + rank)
    host_x = [torch.randn(BATCH, 3, SIZE, SIZE, generator=gi).pin_memory() for _ in range(n_pool)]
    host_t = [(torch.rand(BATCH, SIZE, SIZE, generator=gi) > 0.85).long().pin_memory() for _ in range(n_pool)]   # long {0,1} masks (src/utils/dataset.py:119-122)
    dev_x = [h.to(dev) for h in host_x]
    dev_t = [h.to(dev) for h in host_t]

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier(device_ids=[local])
            torch.cuda.synchronize(dev)

    for i in range(max(args.warmup, 3)):
        ts.step(dev_x[i % n_pool], dev_t[i % n_pool])
    sync_all()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = lib.uwm_kernel_launch_count() + ts.replayed_native_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    exposed = []
    for i in range(args.steps):
        loss = ts.step(dev_x[i % n_pool], dev_t[i % n_pool], time_exchange=True)
        exposed.append(ts._pending)
    e1.record()
    sync_all()
    clocks = sampler.stop()
    launches = int(lib.uwm_kernel_launch_count() + ts.replayed_native_launches - l0)
    ms_total = e0.elapsed_time(e1)
    exposed_ms = sum(a.elapsed_time(b) for a, b in exposed) / max(len(exposed), 1)
    final_loss = float(loss)
    # end to end: every step uploads its images + masks from pinned host memory and reads the loss back (loss.item(), :107);
    # the upload of batch i+1 runs on the prefetcher's copy stream under step i (DataLoader(pin_memory) + .to(device))
    from unet_watermark_b200.training import DevicePrefetcher

    pf = DevicePrefetcher((), dev)

    def e2e_loop(nsteps):
        pf.batches = ((host_x[i % n_pool], host_t[i % n_pool]) for i in range(nsteps))
        for x, t in pf:
            _ = ts.step(x, t).item()

    e2e_loop(3)                                            # untimed: staging buffers, copy stream
    sync_all()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    e2e_loop(args.steps)                                   # the first upload is exposed and inside the timed region
    e3.record()
    sync_all()
    ms_e2e = e2.elapsed_time(e3)
    if world > 1:
        tt = torch.tensor([ms_total, ms_e2e, exposed_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total, ms_e2e, exposed_ms = (float(v) for v in tt)
    if rank == 0:
        eng_flops = 62.512e9 * (SIZE / 512.0) ** 2            # forward conv FLOPs per image
        value = world * BATCH * args.steps / (ms_total * 1e-3)
        tf = value * 3.0 * eng_flops / 1e12 / world            # forward + backward ~ 3x forward
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": SCALING, "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": WORKLOAD, "baseline_config": 5, "encoder": ENCODER, "image": [SIZE, SIZE],
                           "batch_per_gpu": BATCH, "global_batch": BATCH * world, "parallelism": f"dp{world}",
                           "loss": "0.5 * Dice(smooth 1e-5) + 0.5 * BCEWithLogits", "optimizer": "Adam lr 1e-4 wd 1e-4",
                           "forward": "tcgen05 conv kernels (all convs but the 3-channel stem and the 1-channel head) + train-mode BatchNorm (+ residual, ReLU) kernels",
                           "backward": "data gradients of the stride-1 convs on the tcgen05 conv kernel, BatchNorm / ReLU / residual / "
                                       "upsample backward on hand-written HBM-bound kernels, weight gradients cuDNN (aten.convolution_backward)",
                           "cuda_graph": bool(ts.use_graph),
                           "l2": f"inputs rotate through {n_pool} batches of {BATCH * 3 * SIZE * SIZE * 4 / 1e6:.0f} MB > 126 MB L2"},
                "tflops_per_gpu_3x_forward": tf, "frac_of_bf16_peak": tf / peaks["bf16_tflops"],
                "allreduce": {"bytes_per_step": ts.buckets.bytes, "buckets": len(ts.buckets.buckets),
                              "exposed_ms_per_step": exposed_ms,
                              "how": "CUDA events between the end of the backward and the last averaged bucket, mean per step, max over ranks"},
                "e2e": {"value": world * BATCH * args.steps / (ms_e2e * 1e-3), "unit": UNIT,
                        "h2d_bytes_per_step": world * BATCH * SIZE * SIZE * (3 * 4 + 8), "d2h_bytes_per_step": world * 4,
                        "ms_per_step": ms_e2e / args.steps,
                        "api": "pinned host fp32 images + int64 masks -> H2D on the prefetcher's copy stream (DevicePrefetcher, batch i+1 under step i) "
                               "-> TrainStep.step -> loss.item() every step"},
                "gpu_launches": launches, "clocks": clocks, "final_loss": final_loss,
                "roofline": {"bound": "tensor", "kernel": "conv_halo_kernel / conv_tc_kernel (forward convs of the training step)",
                             "achieved": tf, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": tf / peaks["bf16_tflops"],
                             "traffic": None, "peak_source": peaks["source"] + ", burst figure",
                             "how": "3 x forward conv FLOPs per image x images/s (whole step incl. torch backward, BatchNorm and Adam)"}}
        if world == 1 and not args.no_cpu_baseline:
            v, ms, cores = cpu_train_reference_run(steps=2, warmup=1, images_per_step=2)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "2 timed training steps x 2 images, fp32, oracle module, all host threads"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier(device_ids=[local])
        dist.destroy_process_group()
    return 0


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    select_config(args.config, int(os.environ.get("WORLD_SIZE", "1")))
    if args.config == 5:
        v, ms, cores = cpu_train_reference_run(args.steps, max(args.warmup, 1), 2)
        sample = f"2 of {BATCH} images per training step, fp32, oracle module trained the reference's way, {cores} threads"
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                          "scaling": SCALING, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": {"workload": WORKLOAD, "sample": sample},
                          "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
                          "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)
        return 0
    per_step = 2 if SIZE <= 512 else 1
    v, ms, cores = cpu_reference_run(args.steps, max(args.warmup, 1), per_step)
    sample = f"{per_step} of {BATCH} images per step, fp32, oracle port of the reference CPU path, {cores} threads"
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": SCALING,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": sample},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


def run_ours(args):
    import torch
    import torch.distributed as dist
    from unet_watermark_b200 import _lib
    from unet_watermark_b200.unet_model import Unet

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    select_config(args.config, world)
    affinity = None
    if world > 1 and hasattr(os, "sched_setaffinity"):
        # one disjoint CPU set per rank: the launch thread and the pinned-memory copies of eight ranks otherwise
        # migrate over the same cores (r01: e2e scaled 0.974 at 8 GPUs while the device-resident value scaled 0.992)
        try:
            cpus = sorted(os.sched_getaffinity(0))
            per = max(1, len(cpus) // world)
            mine = cpus[local * per:(local + 1) * per] or cpus
            os.sched_setaffinity(0, mine)
            affinity = [mine[0], mine[-1]]
        except OSError:
            affinity = None
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"          # keep stdout to the one JSON line (NCCL prints its version there)
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()
    lib = _lib.load()

    # random-init weights of the named architecture (no checkpoints offline), non-trivial BN statistics
    torch.manual_seed(0)
    model = Unet(ENCODER, encoder_weights=None)
    g = torch.Generator().manual_seed(1)
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.weight.data.uniform_(0.8, 1.2, generator=g)
            m.bias.data.normal_(0, 0.1, generator=g)
            m.running_mean.normal_(0, 0.1, generator=g)
            m.running_var.uniform_(0.7, 1.3, generator=g)
    model = model.to(dev).eval()

    # synthetic uint8 RGB batches; the rotating pool is larger than the 126 MB L2
    batch_bytes = BATCH * SIZE * SIZE * 3
    n_pool = max(2, min(12, -(-160_000_000 // batch_bytes)))
    gi = torch.Generator().manual_seed(100 + rank)
    host_pool = [torch.randint(0, 256, (BATCH, SIZE, SIZE, 3), dtype=torch.uint8, generator=gi).pin_memory()
                 for _ in range(n_pool)]
    dev_pool = [h.to(dev) for h in host_pool]
    eng = model.engine(BATCH, SIZE, SIZE, dev)
    flops_step = eng.flops_per_image * BATCH

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier(device_ids=[local])
            torch.cuda.synchronize(dev)

    # ---- device-resident throughput ---------------------------------------------------------
    # caller-owned mask buffers: with stable input/output pointers every step replays one cached CUDA graph
    mask_bufs = [torch.empty(BATCH, SIZE, SIZE, dtype=torch.uint8, device=dev) for _ in range(2)]
    for i in range(max(args.warmup, 3) + n_pool):      # n_pool extra steps: one graph capture per input buffer
        model.predict_mask(dev_pool[i % n_pool], 0.5, out=mask_bufs[i % 2])
    sync_all()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = lib.uwm_kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        mask = model.predict_mask(dev_pool[i % n_pool], 0.5, out=mask_bufs[i % 2])
    e1.record()
    sync_all()
    clocks = sampler.stop()
    launches = int(lib.uwm_kernel_launch_count() - l0)
    ms_total = e0.elapsed_time(e1)
    checksum = int(mask.sum().item())

    # ---- end to end through the public API with host buffers ---------------------------------
    # Every step copies its own input batch from pinned host memory and reads its masks back to pinned host
    # memory inside the timed region.  Double-buffered over three streams (H2D / compute / D2H) the way a
    # serving loop would run it: step i+1's upload and step i-1's download overlap step i's kernels.
    in_bufs = [torch.empty(BATCH, SIZE, SIZE, 3, dtype=torch.uint8, device=dev) for _ in range(2)]
    out_hosts = [torch.empty(BATCH, SIZE, SIZE, dtype=torch.uint8).pin_memory() for _ in range(2)]
    s_h2d, s_cmp, s_d2h = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    stage_ev = []              # per step: (h2d start, h2d end, compute start, compute end, d2h start, d2h end)

    def e2e_loop(nsteps, timed=False):
        up = [None, None]      # upload finished (per buffer)
        done = [None, None]    # compute finished: input buffer reusable, mask ready
        down = [None, None]    # download finished: mask buffer / host buffer reusable
        for i in range(nsteps):
            b = i % 2
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)] if timed else None
            with torch.cuda.stream(s_h2d):
                if done[b] is not None:
                    s_h2d.wait_event(done[b])
                if timed:
                    ev[0].record(s_h2d)
                in_bufs[b].copy_(host_pool[i % n_pool], non_blocking=True)          # H2D of this step's inputs
                up[b] = ev[1] if timed else torch.cuda.Event()
                up[b].record(s_h2d)
            with torch.cuda.stream(s_cmp):
                s_cmp.wait_event(up[b])
                if down[b] is not None:
                    s_cmp.wait_event(down[b])
                if timed:
                    ev[2].record(s_cmp)
                model.predict_mask(in_bufs[b], 0.5, out=mask_bufs[b])
                done[b] = ev[3] if timed else torch.cuda.Event()
                done[b].record(s_cmp)
            with torch.cuda.stream(s_d2h):
                s_d2h.wait_event(done[b])
                if timed:
                    ev[4].record(s_d2h)
                out_hosts[b].copy_(mask_bufs[b], non_blocking=True)                  # D2H of this step's masks
                down[b] = ev[5] if timed else torch.cuda.Event()
                down[b].record(s_d2h)
            if timed:
                stage_ev.append(ev)
        for s_ in (s_h2d, s_cmp, s_d2h):
            s_.synchronize()

    e2e_loop(4)
    sync_all()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for s_ in (s_h2d, s_cmp, s_d2h):
        s_.wait_stream(torch.cuda.current_stream())
    e2e_loop(args.steps)
    e3.record()
    sync_all()
    ms_e2e = e2.elapsed_time(e3)
    e2e_checksum = int(out_hosts[(args.steps - 1) % 2].sum().item())
    e2e_loop(min(args.steps, 10), timed=True)          # separate short pass: the per-stage events stay out of the timed loop
    sync_all()
    # per-stage device time of one step on its own stream (mean over the timed steps): names the exposed stage
    stage_ms = [sum(ev[2 * k].elapsed_time(ev[2 * k + 1]) for ev in stage_ev) / max(len(stage_ev), 1) for k in range(3)]
    # interval between consecutive completed downloads in the same pass: the pipeline's steady-state step time, i.e. the
    # timed K-step figure without its one-off fill (first upload) and drain (last download)
    steady_ms = (stage_ev[0][5].elapsed_time(stage_ev[-1][5]) / (len(stage_ev) - 1)) if len(stage_ev) > 1 else 0.0

    # ---- sustained leg: >= 3 s of back-to-back graph replays (device-resident), own clock record --------
    sustained = None
    if not args.no_sustained:
        est = max(ms_total / args.steps, 1e-3)
        n_sus = int(min(max(3000.0 / est, args.steps), 200000))
        sync_all()
        samp2 = ClockSampler(local)
        samp2.start()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for i in range(n_sus):
            model.predict_mask(dev_pool[i % n_pool], 0.5, out=mask_bufs[i % 2])
        s1.record()
        sync_all()
        ms_sus = s0.elapsed_time(s1)
        sustained = {"steps": n_sus, "ms_total": ms_sus, "ms_per_step": ms_sus / n_sus, "clocks": samp2.stop()}

    if world > 1:
        t = torch.tensor([ms_total, ms_e2e, sustained["ms_total"] if sustained else 0.0] + stage_ms + [steady_ms], device=dev,
                         dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, ms_e2e = float(t[0]), float(t[1])
        if sustained:
            sustained["ms_total"] = float(t[2]); sustained["ms_per_step"] = float(t[2]) / sustained["steps"]
        stage_ms = [float(v) for v in t[3:6]]
        steady_ms = float(t[6])

    # ---- GPU control: the oracle module as stock torch bf16 channels_last (cuDNN) on the same GPU ---------
    gpu_control = None
    if rank == 0 and not args.no_gpu_control:
        gpu_control = gpu_control_run(dev, dev_pool[0], steps=max(3, min(args.steps, 10)))

    # ---- roofline of the dominant kernels (the tcgen05 convs) -------------------------------
    conv_ms = conv_flops = other_ms = 0.0
    reps = 5
    per_kernel = {}
    launch_rows = []                                   # per launch of the plan: [mean ms, algorithmic FLOPs, algorithmic bytes]
    for r in range(reps + 1):
        prof = eng.profile(dev_pool[r % n_pool], 0.5)
        if r == 0:
            launch_rows = [[0.0, fl, by] for _, _, fl, by in prof]
            continue                                   # warm-up pass
        for i, (_, ms, _, _) in enumerate(prof):
            if i < len(launch_rows):
                launch_rows[i][0] += ms / reps
        for name, ms, fl, by in prof:
            if fl > 0:
                conv_ms += ms
                conv_flops += fl
            else:
                other_ms += ms
            a = per_kernel.setdefault(name, [0.0, fl, by])
            a[0] += ms / reps
    # The timed region replays one CUDA graph per step, so single launches cannot be bracketed by events there.
    # The conv kernels' share of a step is measured live with CUDA events around every launch of an eager pass
    # (conv_ms / (conv_ms + glue_ms); the ncu launch list in profiles/ gives the same share) and applied to the
    # timed step: conv time per step = ms_per_step * share.  The eager per-launch sum itself (which includes the
    # host launch gaps the graph removes) is reported next to it.
    conv_share = conv_ms / (conv_ms + other_ms) if conv_ms > 0 else 0.0
    step_ms = ms_total / args.steps
    conv_ms_step = step_ms * conv_share
    n_conv = sum(1 for v in per_kernel.values() if v[1] > 0)
    achieved = (conv_flops / reps) / (conv_ms_step * 1e-3) / 1e12 if conv_ms_step > 0 else 0.0
    achieved_eager = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if args.config == 2 and os.path.exists(tp):
        try:
            with open(tp) as f:
                traffic = json.load(f).get("conv_dram_bytes_per_step")
        except Exception:  # noqa: BLE001
            traffic = None
    roofline = {"bound": "tensor", "kernel": "conv_halo_kernel (tcgen05 implicit-GEMM convs, all conv launches of the step)",
                "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                "frac": achieved / peaks["bf16_tflops"], "traffic": traffic,
                "traffic_static": True if traffic is not None else None,   # read from profiles/roofline_traffic.json (ncu capture of this build), not measured in this run
                "peak_source": peaks["source"] + ", burst figure",
                "how": f"algorithmic conv FLOPs of one step ({flops_step / 1e12:.4f} TFLOP, {n_conv} conv launches) / "
                       f"(timed ms_per_step x conv share {conv_share:.4f}); share = CUDA-event time of the conv launches / "
                       f"all launches in an eager pass, mean of {reps}; traffic = DRAM bytes of the same launches (ncu)",
                "avg_launch_us": conv_ms_step * 1e3 / max(n_conv, 1),
                "conv_ms_per_step": conv_ms_step, "conv_share": conv_share,
                "eager_events": {"conv_ms_per_step": conv_ms / reps, "glue_ms_per_step": other_ms / reps,
                                 "achieved": achieved_eager},
                "hbm_peak_gbs": peaks["hbm_gbs"]}
    try:
        roofline["by_bound"] = roofline_by_bound(launch_rows, peaks, step_ms)
    except Exception as exc:  # noqa: BLE001 - an explanatory extra must never cost the bench line
        roofline["by_bound"] = {"error": str(exc)}

    if rank == 0:
        value = world * BATCH * args.steps / (ms_total * 1e-3)
        e2e_v = world * BATCH * args.steps / (ms_e2e * 1e-3)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": SCALING, "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": WORKLOAD, "baseline_config": args.config, "encoder": ENCODER, "image": [SIZE, SIZE],
                           "batch_per_gpu": BATCH, "cpu_affinity": affinity,
                           "global_batch": BATCH * world, "parallelism": f"dp{world}",
                           "weights": "random-init (seeded), BatchNorm folded",
                           "l2": f"inputs rotate through {n_pool} batches ({n_pool * BATCH * SIZE * SIZE * 3 / 1e6:.0f} MB) > 126 MB L2; "
                                 f"per-step activation traffic also exceeds L2",
                           "cuda_graph": True},
                "tflops": value * eng.flops_per_image / 1e12,
                "frac_of_bf16_peak": value * eng.flops_per_image / 1e12 / world / peaks["bf16_tflops"],
                "e2e": {"value": e2e_v, "unit": UNIT, "h2d_bytes_per_step": world * BATCH * SIZE * SIZE * 3,
                        "d2h_bytes_per_step": world * BATCH * SIZE * SIZE, "ms_per_step": ms_e2e / args.steps,
                        "api": "pinned host uint8 NHWC -> H2D -> Unet.predict_mask(x, 0.5, out=) -> D2H -> pinned host uint8 masks; "
                               "double-buffered over H2D / compute / D2H streams",
                        "stage_ms": {"h2d": stage_ms[0], "compute": stage_ms[1], "d2h": stage_ms[2],
                                     "how": "CUDA events around each stage on its own stream, mean per step, max over ranks"},
                        "steady_state": {"value": world * BATCH / (steady_ms * 1e-3) if steady_ms > 0 else None, "ms_per_step": steady_ms,
                                         "how": "mean interval between consecutive completed downloads; `value` above also carries the "
                                                "one-off pipeline fill (first upload) and drain (last download) of the K timed steps"},
                        "mask_checksum": e2e_checksum},
                "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "mask_checksum": checksum}
        if sustained:
            v_sus = world * BATCH * sustained["steps"] / (sustained["ms_total"] * 1e-3)
            tf_sus = v_sus * eng.flops_per_image / 1e12 / world
            sustained.update({"value": v_sus, "unit": UNIT, "tflops_per_gpu": tf_sus,
                              "frac_of_bf16_sustained_peak": tf_sus / peaks["bf16_tflops_sustained"] if peaks["bf16_tflops_sustained"] else None,
                              "frac_of_bf16_burst_peak": tf_sus / peaks["bf16_tflops"],
                              "peak_sustained_tflops": peaks["bf16_tflops_sustained"]})
            line["sustained"] = sustained
        if gpu_control:
            line["gpu_control"] = gpu_control
        if world == 1 and not args.no_cpu_baseline:
            per = 2 if SIZE <= 512 else 1
            v, ms, cores = cpu_reference_run(steps=5 if SIZE <= 512 else 3, warmup=1, images_per_step=per)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"timed steps x {per} images (of the {BATCH}-image batch), fp32, oracle port of "
                                              "the reference CPU path, all host threads"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier(device_ids=[local])
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-control", action="store_true")
    ap.add_argument("--no-sustained", action="store_true")
    ap.add_argument("--config", type=int, choices=sorted(CONFIGS), default=2,
                    help="BASELINE.json config (1-based): 2 = r34 512 B16 (default line), 3 = r34 1024 B64 split over the GPUs, "
                         "4 = r50 768 B32, 5 = r34 512 B16 training step")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.config == 5:
        return run_train(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
