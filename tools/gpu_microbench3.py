"""Cycles per repetition of single synchronisation primitives (one warp, nothing else on the SM)."""
import os
os.environ.setdefault("UWM_TOOLS", "1")   # measurement build of the library (python -m unet_watermark_b200.build --tools)
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unet_watermark_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda:0")
names = ["empty loop", "tcgen05.fence::before_thread_sync", "tcgen05.fence::after_thread_sync", "fence.proxy.async",
         "mbarrier.arrive (lane 0)", "__syncwarp", "try_wait completed (lane 0)", "lane-0 try_wait + __syncwarp",
         "tcgen05.commit (lane 0)", "clock64 + st.global (lane 0)"]
out = torch.zeros(16, dtype=torch.int64, device=dev)
iters = 2000
for w, n in enumerate(names):
    for _ in range(2):
        _lib.check(lib.uwm_debug_prim_cost(w, iters, out.data_ptr(), None))
    torch.cuda.synchronize()
    print(f"{n:<40s} {out[0].item() / iters:8.1f} cycles", flush=True)
