"""Timeline of CTA `--cta` ... (CTA 0) of one multi-layer chain launch inside a real forward (tools build).

    UWM_TRACE_CHAIN=0 python tools/gpu_trace_chain.py      # 0: the first chain of the plan (layer2), 1: layer3
"""
import os
os.environ.setdefault("UWM_TOOLS", "1")   # measurement build of the library (python -m unet_watermark_b200.build --tools)
if "UWM_TRACE_LAUNCH" not in os.environ:          # UWM_TRACE_LAUNCH=k: the k-th launch of the plan (single or chain) instead
    os.environ.setdefault("UWM_TRACE_CHAIN", "0")
    os.environ.setdefault("UWM_CHAIN", "1")
os.environ["UWM_DBG"] = str(int(os.environ.get("UWM_DBG", "0")) | 8)      # every CTA stamps start / exit (globaltimer + clock64)
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unet_watermark_b200 import _lib
from unet_watermark_b200.unet_model import Unet
dev = torch.device("cuda:0")
lib = _lib.load()
tr = torch.zeros(2048, dtype=torch.int64, device=dev)
_lib.check(lib.uwm_debug_set_trace(tr.data_ptr()))      # before the plan is instantiated: the pointer is baked in
m = Unet("resnet34", encoder_weights=None).to(dev).eval()
x = torch.randint(0, 256, (16, 512, 512, 3), dtype=torch.uint8, device=dev)
for _ in range(4):
    m.predict_mask(x, 0.5)
torch.cuda.synchronize()
t = tr.cpu().tolist()
t0 = t[0]
rel = lambda v: (v - t0) if v else None  # noqa: E731
print(f"chain {os.environ.get('UWM_TRACE_CHAIN')} / launch {os.environ.get('UWM_TRACE_LAUNCH')}: prologue done {rel(t[1])}; loader reached griddepcontrol.wait {rel(t[2])}, passed {rel(t[3])}")
print("item: dep wait begin/end | A stages slot-free/issued ... | mma acc-free / first-stage-landed / all issued | epilogue acc-full / stored")
for i in range(40):
    if not t[160 + 3 * i]:
        break
    st = []
    for s in range(2 * i, 2 * i + 2):
        st.append(f"{rel(t[16 + 2 * s])}/{rel(t[17 + 2 * s])}")
    print(f"{i:3d}: dep {rel(t[500 + 2 * i])}/{rel(t[501 + 2 * i])} | A {' '.join(st)} | mma {rel(t[160 + 3 * i])} {rel(t[161 + 3 * i])} {rel(t[162 + 3 * i])} | epi {rel(t[400 + 2 * i])} {rel(t[401 + 2 * i])} done {rel(t[560 + i]) if 560 + i < 600 else None}")

# CTA lifetimes: start / exit in ns (globaltimer) and SM cycles (clock64) -> effective SM clock while the chain runs
import statistics
st, ex, clk = [], [], []
for b in range(148):
    s0, s1, c0, c1 = t[1024 + 4 * b], t[1025 + 4 * b], t[1026 + 4 * b], t[1027 + 4 * b]
    if s0 and s1:
        st.append(s0); ex.append(s1); clk.append((c1 - c0) / max(s1 - s0, 1))
if st:
    base = min(st)
    print(f"CTAs {len(st)}: start min/med/max {0}/{statistics.median(st) - base:.0f}/{max(st) - base} ns, "
          f"exit min/med/max {min(ex) - base}/{statistics.median(ex) - base:.0f}/{max(ex) - base} ns, "
          f"effective SM clock (cycles/ns) min/med/max {min(clk):.3f}/{statistics.median(clk):.3f}/{max(clk):.3f}")
