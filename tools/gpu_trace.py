"""Timeline of CTA 0 of one halo conv launch (clock64 stamps written by the kernel, bench-only).

    python tools/gpu_trace.py --filter layer1
"""
import argparse
import os
os.environ.setdefault("UWM_TOOLS", "1")   # measurement build of the library (python -m unet_watermark_b200.build --tools)
import sys

import torch

os.environ.setdefault("UWM_OP_PDL", "1")   # time single ops with programmatic dependent launch, as the plan runs them

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.gpu_conv_bench import SHAPES  # noqa: E402
from unet_watermark_b200 import _lib, ops, packing  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--filter", default="layer1")
    ap.add_argument("--residual", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    lib = _lib.load()
    for name, n, h, w, cx, cs, cout, up in SHAPES:
        if args.filter not in name:
            continue
        if up == "par" or (up in ("s2d", "s2dhead") and "gap" in __file__):
            print(f"{name}: not driven by this tool (use tools/gpu_conv_bench.py)")
            continue
        x = torch.randn(n, h, w, cx, device=dev).to(torch.bfloat16)
        ho, wo = (2 * h, 2 * w) if up else (h, w)
        skip = torch.randn(n, ho, wo, cs, device=dev).to(torch.bfloat16) if cs else None
        cin = cx + cs
        wt = torch.randn(cout, 16 if str(up).startswith("s2d") else cin, 3, 3, device=dev) / (cin * 9) ** 0.5
        if str(up).startswith("s2d"):
            wp = packing.pack_s2d_conv3x3(wt, 16 if cout == 1 else 0)
            b = torch.zeros(16 if cout == 1 else 4 * cout, device=dev)
            ho, wo = h, w
        else:
            wp = packing.pack_upcat_subpixel(wt, cx) if up == "spx" else packing.pack_taps(wt)
            b = torch.zeros(4 * cout if up == "spx" else cout, device=dev)
        out = torch.empty(n, ho, wo, cout, dtype=torch.bfloat16, device=dev)
        res = torch.randn(n, ho, wo, cout, device=dev).to(torch.bfloat16) if args.residual else None

        s2d_out = torch.empty(n, h, w, 64, dtype=torch.bfloat16, device=dev) if up == "s2d" else None

        def run():
            if res is not None and not up and skip is None:
                ops.conv2d(x, wp, b, 3, 3, 1, 1, relu=True, residual=res, out=out)
            elif up == "s2d":
                ops.conv2d_s2d(x, wp, b, relu=True, out=s2d_out)
            elif up == "s2dhead":
                ops.head_s2d(x, wp, b, threshold=0.5, want_logits=False)
            elif up == "spx":
                ops.conv2d_upcat_subpixel(x, skip, wp, b, relu=True, out=out)
            else:
                ops.conv2d_upcat(x, skip, wp, b, relu=True, upsample=up, out=out)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        tr = torch.zeros(1024, dtype=torch.int64, device=dev)
        _lib.check(lib.uwm_debug_set_trace(tr.data_ptr()))
        run()
        torch.cuda.synchronize()
        _lib.check(lib.uwm_debug_set_trace(None))
        t = tr.cpu().tolist()
        t0 = t[0]
        rel = lambda v: (v - t0) if v else None  # noqa: E731
        print(f"== {name}  (cycles since kernel start of CTA 0)")
        print(f"prologue done {rel(t[1])}   loader: reached griddepcontrol.wait {rel(t[2])}, passed it {rel(t[3])}")
        print("loader  stage: slot-free / copies-issued")
        for i in range(72):
            if t[16 + 2 * i]:
                print(f"   {i:3d}: {rel(t[16 + 2 * i]):8d} {rel(t[17 + 2 * i]):8d}")
        print("mma     tile: acc-free / first-stage-landed / issued   |  epilogue: acc-full / stored")
        for i in range(80):
            if t[160 + 3 * i]:
                e0, e1 = rel(t[400 + 2 * i]), rel(t[401 + 2 * i])
                print(f"   {i:3d}: {rel(t[160 + 3 * i]):8d} {rel(t[161 + 3 * i]):8d} {rel(t[162 + 3 * i]):8d}   | {e0} {e1}")
                if i < 20 and any(t[600 + 16 * i + k] for k in range(16)):
                    print("        items (chunk in registers / staged): " + "  ".join(
                        f"{rel(t[600 + 16 * i + 2 * k])}/{rel(t[601 + 16 * i + 2 * k])}" for k in range(8) if t[600 + 16 * i + 2 * k]))


if __name__ == "__main__":
    main()
