"""Timeline of CTA 0 of ONE kernel shape inside a real forward (bench-only).  The stem (4x4 taps):

    UWM_TRACE_KH=4 python tools/gpu_trace_model.py
"""
import os
os.environ.setdefault("UWM_TOOLS", "1")   # measurement build of the library (python -m unet_watermark_b200.build --tools)
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unet_watermark_b200 import _lib
from unet_watermark_b200.unet_model import Unet
dev = torch.device("cuda:0")
lib = _lib.load()
tr = torch.zeros(1024, dtype=torch.int64, device=dev)
_lib.check(lib.uwm_debug_set_trace(tr.data_ptr()))      # before the plan is instantiated: the pointer is baked in
m = Unet("resnet34", encoder_weights=None).to(dev).eval()
m.use_cuda_graph = False
x = torch.randint(0, 256, (16, 512, 512, 3), dtype=torch.uint8, device=dev)
for _ in range(3):
    m.predict_mask(x, 0.5)
torch.cuda.synchronize()
t = tr.cpu().tolist()
t0 = t[0]
rel = lambda v: (v - t0) if v else None  # noqa: E731
print(f"prologue done {rel(t[1])}")
print("loader  stage: slot-free / copies-issued")
for i in range(20):
    if t[16 + 2 * i]:
        print(f"   {i:3d}: {rel(t[16 + 2 * i]):8d} {rel(t[17 + 2 * i]):8d}")
print("mma     tile: acc-free / first-stage-landed / issued   |  epilogue: acc-full / stored")
for i in range(20):
    if t[160 + 3 * i]:
        print(f"   {i:3d}: {rel(t[160 + 3 * i]):8d} {rel(t[161 + 3 * i]):8d} {rel(t[162 + 3 * i]):8d}   | {rel(t[400 + 2 * i])} {rel(t[401 + 2 * i])}")
