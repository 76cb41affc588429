"""Timeline of CTA `--cta` ... (CTA 0) of one multi-layer chain launch inside a real forward (tools build).

    UWM_TRACE_CHAIN=0 python tools/gpu_trace_chain.py      # 0: the first chain of the plan (layer2), 1: layer3
"""
import os
os.environ.setdefault("UWM_TOOLS", "1")   # measurement build of the library (python -m unet_watermark_b200.build --tools)
os.environ.setdefault("UWM_TRACE_CHAIN", "0")
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unet_watermark_b200 import _lib
from unet_watermark_b200.unet_model import Unet
dev = torch.device("cuda:0")
lib = _lib.load()
tr = torch.zeros(640, dtype=torch.int64, device=dev)
_lib.check(lib.uwm_debug_set_trace(tr.data_ptr()))      # before the plan is instantiated: the pointer is baked in
m = Unet("resnet34", encoder_weights=None).to(dev).eval()
x = torch.randint(0, 256, (16, 512, 512, 3), dtype=torch.uint8, device=dev)
for _ in range(4):
    m.predict_mask(x, 0.5)
torch.cuda.synchronize()
t = tr.cpu().tolist()
t0 = t[0]
rel = lambda v: (v - t0) if v else None  # noqa: E731
print(f"chain {os.environ['UWM_TRACE_CHAIN']}: prologue done {rel(t[1])}; loader reached griddepcontrol.wait {rel(t[2])}, passed {rel(t[3])}")
print("item: dep wait begin/end | A stages slot-free/issued ... | mma acc-free / first-stage-landed / all issued | epilogue acc-full / stored")
for i in range(40):
    if not t[160 + 3 * i]:
        break
    st = []
    for s in range(2 * i, 2 * i + 2):
        st.append(f"{rel(t[16 + 2 * s])}/{rel(t[17 + 2 * s])}")
    print(f"{i:3d}: dep {rel(t[500 + 2 * i])}/{rel(t[501 + 2 * i])} | A {' '.join(st)} | mma {rel(t[160 + 3 * i])} {rel(t[161 + 3 * i])} {rel(t[162 + 3 * i])} | epi {rel(t[400 + 2 * i])} {rel(t[401 + 2 * i])}")
