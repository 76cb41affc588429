"""Time single conv layers (all distinct r34@512^2 B=16 shapes) through the C ABI.  UWM_DBG=1|2 isolates TMA / MMA."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unet_watermark_b200 import ops, packing

SHAPES = [  # name, n, h, w, cin, cout, k, stride
    ("layer1 64->64 @128", 16, 128, 128, 64, 64, 3, 1),
    ("layer2 128->128 @64", 16, 64, 64, 128, 128, 3, 1),
    ("layer3 256->256 @32", 16, 32, 32, 256, 256, 3, 1),
    ("layer4 512->512 @16", 16, 16, 16, 512, 512, 3, 1),
    ("dec0 768->256 @32", 16, 32, 32, 768, 256, 3, 1),
    ("dec2 192->64 @128", 16, 128, 128, 192, 64, 3, 1),
    ("dec3 128->32 @256", 16, 256, 256, 128, 32, 3, 1),
    ("dec4 32->16 @512", 16, 512, 512, 32, 16, 3, 1),
    ("dec4 16->16 @512", 16, 512, 512, 16, 16, 3, 1),
]
dev = torch.device("cuda:0")
for name, n, h, w, cin, cout, k, st in SHAPES:
    x = torch.randn(n, h, w, cin, device=dev).to(torch.bfloat16)
    wp = packing.pack_taps(torch.randn(cout, cin, k, k, device=dev) / (cin * k * k) ** 0.5)
    b = torch.zeros(cout, device=dev)
    out = torch.empty(n, h // st, w // st, cout, dtype=torch.bfloat16, device=dev)
    for _ in range(3):
        ops.conv2d(x, wp, b, k, k, st, k // 2, relu=True, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(10):
        ops.conv2d(x, wp, b, k, k, st, k // 2, relu=True, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    fl = 2.0 * n * (h // st) * (w // st) * cin * cout * k * k
    by = 2.0 * (x.numel() + out.numel())
    print(f"{name:<24s} {ms*1e3:8.1f} us  {fl/ms/1e9:7.1f} TF/s  {by/ms/1e6:7.1f} GB/s(algorithmic)", flush=True)
