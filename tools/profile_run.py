"""Minimal workload for ncu: N eager forwards of Unet-resnet34 B=16 512x512 (config 2), uint8 in, mask out.

    python tools/profile_run.py [--passes 3] [--encoder resnet34] [--size 512] [--batch 16]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unet_watermark_b200.unet_model import Unet  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--passes", type=int, default=3)
ap.add_argument("--encoder", default="resnet34")
ap.add_argument("--size", type=int, default=512)
ap.add_argument("--batch", type=int, default=16)
a = ap.parse_args()
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = Unet(a.encoder, encoder_weights=None).to(dev).eval()
m.use_cuda_graph = False
x = torch.randint(0, 256, (a.batch, a.size, a.size, 3), dtype=torch.uint8, device=dev)
for _ in range(a.passes):
    mask = m.predict_mask(x, 0.5)
torch.cuda.synchronize()
print("ok", int(mask.sum()), m.engine(a.batch, a.size, a.size).kernels_per_forward, "kernels/forward")
