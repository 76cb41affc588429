"""Time single conv layers (all distinct r34@512^2 B=16 stride-1 shapes) through the C ABI.

    python tools/gpu_conv_bench.py [--dbg 0,1,2,4,7] [--filter dec4]

UWM_DBG bit mask (bench-only; results are garbage when set): 1 skip activation loads, 2 skip MMAs,
4 skip epilogue.  Comparing the columns shows which of loader / tensor pipe / epilogue bounds a layer.
"""
import argparse
import os
os.environ.setdefault("UWM_TOOLS", "1")   # measurement build of the library (python -m unet_watermark_b200.build --tools)
import sys

import torch

os.environ.setdefault("UWM_OP_PDL", "1")   # time single ops with programmatic dependent launch, as the plan runs them

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unet_watermark_b200 import ops, packing  # noqa: E402

SHAPES = [  # name, n, h, w, cin (x), cskip, cout, upsample
    ("layer1 64->64 @128", 16, 128, 128, 64, 0, 64, False),
    ("layer2 128->128 @64", 16, 64, 64, 128, 0, 128, False),
    ("layer3 256->256 @32", 16, 32, 32, 256, 0, 256, False),
    ("layer4 512->512 @16", 16, 16, 16, 512, 0, 512, False),
    ("dec0 up512+256->256 @32", 16, 16, 16, 512, 256, 256, True),
    ("dec1 up256+128->128 @64", 16, 32, 32, 256, 128, 128, True),
    ("dec2 up128+64->64 @128", 16, 64, 64, 128, 64, 64, True),
    ("dec3 up64+64->32 @256", 16, 128, 128, 64, 64, 32, True),
    ("dec3 spx up64+64->32 @256", 16, 128, 128, 64, 64, 32, "spx"),
    ("dec2 spx up128+64->64 @128", 16, 64, 64, 128, 64, 64, "spx"),
    ("dec0 xpart up512->256 @32", 16, 16, 16, 512, 0, 256, "par"),    # sub-pixel, one N tile per parity, + residual
    ("dec1 xpart up256->128 @64", 16, 32, 32, 256, 0, 128, "par"),
    ("dec3 32->32 @256", 16, 256, 256, 32, 0, 32, False),
    ("dec4 up32->16 @512", 16, 256, 256, 32, 0, 16, True),
    ("dec4 16->16 @512", 16, 512, 512, 16, 0, 16, False),
    ("dec4 s2d 16->16 @512", 16, 256, 256, 64, 0, 16, "s2d"),      # h, w, cx in space-to-depth terms
    ("head s2d 16->1 @512", 16, 256, 256, 64, 0, 1, "s2dhead"),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dbg", default="0")
    ap.add_argument("--filter", default="")
    ap.add_argument("--graph", action="store_true", help="time 20 launches captured in one CUDA graph (no host launch cost)")
    args = ap.parse_args()
    dbgs = [int(v) for v in args.dbg.split(",")]
    dev = torch.device("cuda:0")
    print(f"{'layer':<28s}" + "".join(f"  dbg={d:<2d} us" for d in dbgs) + "   TF/s(dbg0)  GB/s(dbg0, algorithmic)")
    for name, n, h, w, cx, cs, cout, up in SHAPES:
        if args.filter not in name:
            continue
        x = torch.randn(n, h, w, cx, device=dev).to(torch.bfloat16)
        ho, wo = (2 * h, 2 * w) if up else (h, w)
        skip = torch.randn(n, ho, wo, cs, device=dev).to(torch.bfloat16) if cs else None
        cin = cx + cs
        if up == "par":
            wt = torch.randn(cout, cin, 3, 3, device=dev) / (cin * 9) ** 0.5
            wp = packing.pack_up2x_shuffle(wt)
            b = torch.zeros(4 * cout, device=dev)
            out = torch.empty(n, ho, wo, cout, dtype=torch.bfloat16, device=dev)
            par_res = torch.randn(n, ho, wo, cout, device=dev).to(torch.bfloat16)
        elif up in ("s2d", "s2dhead"):
            ho, wo, cin = h, w, 16
            wt = torch.randn(cout, 16, 3, 3, device=dev) / 12
            wp = packing.pack_s2d_conv3x3(wt, 16 if cout == 1 else 0)
            b = torch.zeros(16 if cout == 1 else 4 * cout, device=dev)
            out = torch.empty(n, h, w, 64, dtype=torch.bfloat16, device=dev)
        else:
            wt = torch.randn(cout, cin, 3, 3, device=dev) / (cin * 9) ** 0.5
            wp = packing.pack_upcat_subpixel(wt, cx) if up == "spx" else packing.pack_taps(wt)
            b = torch.zeros(4 * cout if up == "spx" else cout, device=dev)
            out = torch.empty(n, ho, wo, cout, dtype=torch.bfloat16, device=dev)
        times = []

        def run():
            if up == "par":
                ops.conv2d_up2x_shuffle_res(x, wp, b, residual=par_res, relu=True, out=out)
            elif up == "s2d":
                ops.conv2d_s2d(x, wp, b, relu=True, out=out)
            elif up == "s2dhead":
                ops.head_s2d(x, wp, b, threshold=0.5, want_logits=False)
            elif up == "spx":
                ops.conv2d_upcat_subpixel(x, skip, wp, b, relu=True, out=out)
            else:
                ops.conv2d_upcat(x, skip, wp, b, relu=True, upsample=up, out=out)
        for d in dbgs:
            os.environ["UWM_DBG"] = str(d)
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            if args.graph:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    for _ in range(20):
                        run()
                g.replay()
                torch.cuda.synchronize()
                e0.record()
                g.replay()
                e1.record()
                torch.cuda.synchronize()
                times.append(e0.elapsed_time(e1) / 20)
            else:
                e0.record()
                for _ in range(10):
                    run()
                e1.record()
                torch.cuda.synchronize()
                times.append(e0.elapsed_time(e1) / 10)
        os.environ["UWM_DBG"] = "0"
        fl = 2.0 * n * ho * wo * cin * cout * 9 * (4 if up in ("s2d", "s2dhead") else 1)
        by = 2.0 * (x.numel() + out.numel() + (skip.numel() if skip is not None else 0))
        if up == "s2dhead":
            by = 2.0 * x.numel() + 4.0 * n * h * w
        ms = times[0]
        print(f"{name:<28s}" + "".join(f"  {t * 1e3:9.1f}" for t in times) +
              f"   {fl / ms / 1e9:8.1f}   {by / ms / 1e6:8.1f}", flush=True)


if __name__ == "__main__":
    main()
