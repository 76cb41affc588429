"""Per-kernel device times of one eager forward (CUDA events around every launch), mean of --passes.

    python tools/gpu_layer_times.py [--encoder resnet34] [--size 512] [--batch 16]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unet_watermark_b200.unet_model import Unet  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--encoder", default="resnet34")
ap.add_argument("--size", type=int, default=512)
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--passes", type=int, default=5)
a = ap.parse_args()
dev = torch.device("cuda:0")
m = Unet(a.encoder, encoder_weights=None).to(dev).eval()
x = torch.randint(0, 256, (a.batch, a.size, a.size, 3), dtype=torch.uint8, device=dev)
for _ in range(3):
    m.predict_mask(x, 0.5)
eng = m.engine(a.batch, a.size, a.size)
acc = None
for _ in range(a.passes):
    rows = eng.profile(x)
    if acc is None:
        acc = [[n, 0.0, f, b] for n, _, f, b in rows]
    for r, (_, ms, _, _) in zip(acc, rows):
        r[1] += ms / a.passes
tot = sum(r[1] for r in acc)
print(f"{'kernel':<34s} {'us':>8s} {'share':>6s} {'TF/s':>8s} {'GB/s':>8s}")
for n, ms, f, b in acc:
    print(f"{n:<34s} {ms * 1e3:8.1f} {ms / tot * 100:5.1f}% {f / ms / 1e9 if ms else 0:8.1f} {b / ms / 1e6 if ms else 0:8.1f}")
print(f"{'total (eager, serialised)':<34s} {tot * 1e3:8.1f}")
