"""Where the time of one training step (BASELINE config 5) goes: torch.profiler (CUPTI) kernel table of TrainStep.step,
grouped by kernel name, plus CUDA-event times of forward / loss / backward / optimizer.

    python tools/gpu_train_profile.py [--batch 16] [--size 512] [--steps 3] [--top 40]
"""
import argparse
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unet_watermark_b200.training import TrainStep  # noqa: E402
from unet_watermark_b200.unet_model import Unet  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--size", type=int, default=512)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--top", type=int, default=40)
a = ap.parse_args()
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = Unet("resnet34", encoder_weights=None).to(dev)
ts = TrainStep(m)
g = torch.Generator().manual_seed(1)
x = torch.randn(a.batch, 3, a.size, a.size, generator=g).to(dev)
t = (torch.rand(a.batch, a.size, a.size, generator=g) > 0.85).long().to(dev)
for _ in range(3):
    ts.step(x, t)
torch.cuda.synchronize()

# phases with CUDA events (the same code TrainStep.step runs)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
m.train()
ev[0].record()
ts.buckets.zero_grad()
out = m(x)
ev[1].record()
loss = ts.criterion(out, t.unsqueeze(1))
ev[2].record()
loss.backward()
ev[3].record()
ts.buckets.finish()
ts.optimizer.step()
ev[4].record()
torch.cuda.synchronize()
names = ["zero_grad + forward", "loss", "backward", "optimizer"]
for i, n in enumerate(names):
    print(f"{n:<22s} {ev[i].elapsed_time(ev[i + 1]):8.3f} ms")
print(f"{'step':<22s} {ev[0].elapsed_time(ev[4]):8.3f} ms", flush=True)

from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(a.steps):
        ts.step(x, t)
    torch.cuda.synchronize()
try:
    acc = collections.defaultdict(lambda: [0.0, 0])
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            k = e.name[:90]
            us = getattr(e, "device_time", None)
            us = getattr(e, "cuda_time", 0.0) if us is None else us
            acc[k][0] += us / a.steps
            acc[k][1] += 1
    tot = sum(v[0] for v in acc.values())
    print(f"\nCUDA kernel time per step: {tot / 1e3:.3f} ms in {sum(v[1] for v in acc.values()) // a.steps} launches")
    for k, (us, n) in sorted(acc.items(), key=lambda kv: -kv[1][0])[:a.top]:
        print(f"{us / 1e3:8.3f} ms {us / tot * 100:5.1f}%  x{n // a.steps:<4d} {k}")
except Exception as exc:  # noqa: BLE001 - profiler event fields differ between torch versions
    print("event aggregation failed:", exc)
    print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=a.top, max_name_column_width=90))
