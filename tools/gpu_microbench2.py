import os
os.environ.setdefault("UWM_TOOLS", "1")   # measurement build of the library (python -m unet_watermark_b200.build --tools)
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unet_watermark_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda:0")
names = {0: "try_wait, all lanes, depth1", 1: "test_wait spin, all lanes, depth1", 2: "try_wait, lane0, depth1",
         3: "test_wait spin, lane0, depth1", 4: "try_wait, all lanes, depth4", 5: "test_wait spin, all lanes, depth4",
         6: "try_wait, lane0, depth4", 7: "test_wait spin, lane0, depth4"}
for v in range(8):
    cyc = torch.zeros(148, dtype=torch.int64, device=dev)
    iters = 5000
    for _ in range(2):
        _lib.check(lib.uwm_debug_handshake(iters, v, 148, cyc.data_ptr(), None))
    torch.cuda.synchronize()
    print(f"variant {v} ({names[v]}): {cyc.float().mean().item()/iters:7.1f} cycles per iteration", flush=True)
