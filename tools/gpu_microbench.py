"""tcgen05.mma rate on this GPU: cycles per group of 4 MMAs (M=128 x N x K=16 each) under different
synchronisation patterns.  mode 0 raw; 1 +commit per group; 3 +wait(complete)+commit; 4 full handshake."""
import os
os.environ.setdefault("UWM_TOOLS", "1")   # measurement build of the library (python -m unet_watermark_b200.build --tools)
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unet_watermark_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda:0")
blocks = 148
for mode in (0, 1, 3, 4):
    for n in (16, 64, 128):
        for stages in (2, 4, 8):
            if mode < 4 and stages != 4:
                continue
            cyc = torch.zeros(blocks, dtype=torch.int64, device=dev)
            iters = 2000
            for _ in range(2):
                _lib.check(lib.uwm_debug_mma_rate(n, iters, stages, mode, blocks, cyc.data_ptr(), None))
            torch.cuda.synchronize()
            c = cyc.float().mean().item() / iters
            print(f"mode={mode} N={n:3d} stages={stages}: {c:7.1f} cycles per 4-MMA group (ideal {4*max(41,n/2):.0f})", flush=True)
