"""Which part of the end-to-end training loop costs what: the same TrainStep timed (CUDA events, 10 steps each) with
device-resident inputs / a loss.item() per step / per-step uploads from pinned host memory / both (= bench.py's e2e)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unet_watermark_b200.training import TrainStep  # noqa: E402
from unet_watermark_b200.unet_model import Unet  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
m = Unet("resnet34", encoder_weights=None).to(dev)
ts = TrainStep(m)
g = torch.Generator().manual_seed(1)
hx = [torch.randn(16, 3, 512, 512, generator=g).pin_memory() for _ in range(4)]
ht = [(torch.rand(16, 512, 512, generator=g) > 0.85).long().pin_memory() for _ in range(4)]
dx = [h.to(dev) for h in hx]
dt = [h.to(dev) for h in ht]
for i in range(5):
    ts.step(dx[i % 4], dt[i % 4])
torch.cuda.synchronize()


def timed(name, fn, n=10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:<44s} {e0.elapsed_time(e1) / n:8.3f} ms/step", flush=True)


def up(i):
    return hx[i % 4].to(dev, non_blocking=True), ht[i % 4].to(dev, non_blocking=True)


for rep in range(2):
    timed("device inputs, no sync", lambda i: ts.step(dx[i % 4], dt[i % 4]))
    timed("device inputs, loss.item() each step", lambda i: ts.step(dx[i % 4], dt[i % 4]).item())
    timed("uploads, no sync", lambda i: ts.step(*up(i)))
    timed("uploads + loss.item() (bench e2e)", lambda i: ts.step(*up(i)).item())
    timed("uploads only", lambda i: up(i))
    timed("device inputs, torch.cuda.synchronize() each", lambda i: (ts.step(dx[i % 4], dt[i % 4]), torch.cuda.synchronize()))
