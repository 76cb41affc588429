"""CTA lifetimes of consecutive launches of one conv layer (bench-only, UWM_DBG bit 8).

    python tools/gpu_trace_gap.py --filter layer2

Four launches are captured in one CUDA graph, each with its own trace buffer; every CTA stamps %globaltimer at
start and exit.  Printed per launch, relative to the first launch's first CTA start: first/median/last CTA start,
first/median/last CTA exit - i.e. how much of a launch's wall time lies outside the CTAs' own lifetimes.
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
    ap.add_argument("--filter", default="layer2")
    ap.add_argument("--launches", type=int, default=4)
    args = ap.parse_args()
    os.environ["UWM_DBG"] = str(int(os.environ.get("UWM_DBG", "0")) | 8)
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
        wt = torch.randn(cout, cin, 3, 3, device=dev) / (cin * 9) ** 0.5
        wp = packing.pack_upcat_subpixel(wt, cx) if up == "spx" else packing.pack_taps(wt)
        b = torch.zeros(4 * cout if up == "spx" else cout, device=dev)
        out = torch.empty(n, ho, wo, cout, dtype=torch.bfloat16, device=dev)

        def run():
            if up == "spx":
                ops.conv2d_upcat_subpixel(x, skip, wp, b, relu=True, out=out)
            else:
                ops.conv2d_upcat(x, skip, wp, b, relu=True, upsample=up, out=out)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        trs = [torch.zeros(2048, dtype=torch.int64, device=dev) for _ in range(args.launches)]
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for tr in trs:
                _lib.check(lib.uwm_debug_set_trace(tr.data_ptr()))
                run()
        _lib.check(lib.uwm_debug_set_trace(None))
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        print(f"== {name}: CTA start / exit in us since the first CTA start of launch 0")
        t0 = None
        prev_last_exit = None
        for i, tr in enumerate(trs):
            t = tr.cpu()[1024:].view(-1, 4)
            t = t[t[:, 0] > 0]
            st, en = t[:, 0].sort().values, t[:, 1].sort().values
            if t0 is None:
                t0 = int(st[0])
            f = lambda v: f"{(int(v) - t0) / 1e3:8.2f}"  # noqa: E731
            m = len(st) // 2
            gap = "" if prev_last_exit is None else f"  first start - prev last exit = {(int(st[0]) - prev_last_exit) / 1e3:6.2f} us"
            life = (t[:, 1] - t[:, 0]).float() / 1e3
            print(f"launch {i}: ctas {len(st):3d}  start {f(st[0])} {f(st[m])} {f(st[-1])}   exit {f(en[0])} {f(en[m])} {f(en[-1])}"
                  f"   lifetime us min/med/max {life.min():.2f}/{life.median():.2f}/{life.max():.2f}{gap}")
            mhz = ((t[:, 3] - t[:, 2]).float() / (t[:, 1] - t[:, 0]).float() * 1e3).median()
            print(f"          effective SM clock over the CTA lifetimes (clock64 / globaltimer): {mhz:.0f} MHz")
            prev_last_exit = int(en[-1])


if __name__ == "__main__":
    main()
