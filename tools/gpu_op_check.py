"""Run every per-operator parity case on the GPU and print one line per case (never aborts early).

    python tools/gpu_op_check.py [--filter substr]

Reference = plain PyTorch fp32 (cuDNN, TF32 off) on the same bf16-rounded inputs.
"""
import argparse
import os
import sys
import traceback

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.op_cases import CONV_CASES, S2D_CASES, S2P_CASES, SHUFFLE_CASES, SHUFFLE_RES_CASES, SPX_CASES, UPCAT_CASES  # noqa: E402
from unet_watermark_b200 import ops, packing  # noqa: E402


def conv_case(case, dev, seed=0):
    name, n, h, w, cin, cout, k, stride, pad, relu, use_res, in_extra, out_extra = case
    g = torch.Generator(device="cpu").manual_seed(seed)
    xbuf = torch.randn(n, h, w, cin + in_extra, generator=g).to(dev).to(torch.bfloat16)
    x = xbuf[..., in_extra:] if in_extra else xbuf
    wt = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).to(dev)
    bias = torch.randn(cout, generator=g).to(dev)
    ho = (h + 2 * pad - k) // stride + 1
    wo = (w + 2 * pad - k) // stride + 1
    res = torch.randn(n, ho, wo, cout, generator=g).to(dev).to(torch.bfloat16) if use_res else None
    wp = packing.pack_taps(wt)
    obuf = torch.full((n, ho, wo, cout + out_extra), 7.0, dtype=torch.bfloat16, device=dev)
    out = obuf[..., out_extra:] if out_extra else obuf
    ops.conv2d(x, wp, bias, k, k, stride, pad, relu=relu, residual=res, out=out)
    torch.cuda.synchronize()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wp.float().view(cout, k, k, cin).permute(0, 3, 1, 2), bias,
                   stride=stride, padding=pad)
    if res is not None:
        ref = ref + res.float().permute(0, 3, 1, 2)
    if relu:
        ref = ref.relu()
    ref = ref.permute(0, 2, 3, 1)
    err = (out.float() - ref).abs()
    tol = 1e-2 * ref.abs().clamp_min(1.0)
    bad = (err > tol).float().mean().item()
    untouched = True
    if out_extra:
        untouched = bool((obuf[..., :out_extra] == 7.0).all())
    return err.max().item(), ref.abs().max().item(), bad, untouched


def upcat_case(case, dev, seed=0):
    name, n, h, w, cx, cs, cout, up, relu, x_extra, s_extra = case
    g = torch.Generator(device="cpu").manual_seed(seed)
    xbuf = torch.randn(n, h, w, cx + x_extra, generator=g).to(dev).to(torch.bfloat16)
    x = xbuf[..., x_extra:] if x_extra else xbuf
    ho, wo = (2 * h, 2 * w) if up else (h, w)
    skip = None
    if cs:
        sbuf = torch.randn(n, ho, wo, cs + s_extra, generator=g).to(dev).to(torch.bfloat16)
        skip = sbuf[..., :cs] if s_extra else sbuf
    cin = cx + cs
    wt = (torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5).to(dev)
    bias = torch.randn(cout, generator=g).to(dev)
    wp = packing.pack_taps(wt)
    out = ops.conv2d_upcat(x, skip, wp, bias, relu=relu, upsample=up)
    torch.cuda.synchronize()
    xi = x.float().permute(0, 3, 1, 2)
    if up:
        xi = F.interpolate(xi, scale_factor=2, mode="nearest")
    if skip is not None:
        xi = torch.cat([xi, skip.float().permute(0, 3, 1, 2)], dim=1)
    ref = F.conv2d(xi, wp.float().view(cout, 3, 3, cin).permute(0, 3, 1, 2), bias, padding=1)
    if relu:
        ref = ref.relu()
    ref = ref.permute(0, 2, 3, 1)
    err = (out.float() - ref).abs()
    bad = (err > 1e-2 * ref.abs().clamp_min(1.0)).float().mean().item()
    return err.max().item(), ref.abs().max().item(), bad


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--filter", default="")
    args = ap.parse_args()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device("cuda:0")
    print(torch.cuda.get_device_name(0), flush=True)
    nfail = 0
    for case in CONV_CASES:
        if args.filter not in case[0]:
            continue
        try:
            e, m, bad, untouched = conv_case(case, dev)
            ok = bad == 0.0 and untouched
            print(f"{'OK  ' if ok else 'FAIL'} conv {case[0]:<24s} max_err={e:.4g} ref_max={m:.4g} bad_frac={bad:.4g} pad_untouched={untouched}", flush=True)
            nfail += (not ok)
        except Exception as ex:  # noqa: BLE001
            nfail += 1
            print(f"EXC  conv {case[0]}: {ex}", flush=True)
            traceback.print_exc()
            if "CUDA" in str(ex) or "fault" in str(ex):
                print("aborting after CUDA error")
                sys.exit(2)

    for case in UPCAT_CASES:
        if args.filter not in case[0]:
            continue
        try:
            e, m, bad = upcat_case(case, dev)
            ok = bad == 0.0
            print(f"{'OK  ' if ok else 'FAIL'} upcat {case[0]:<24s} max_err={e:.4g} ref_max={m:.4g} bad_frac={bad:.4g}", flush=True)
            nfail += (not ok)
        except Exception as ex:  # noqa: BLE001
            nfail += 1
            print(f"EXC  upcat {case[0]}: {ex}", flush=True)
            traceback.print_exc()
            if "CUDA" in str(ex) or "fault" in str(ex):
                print("aborting after CUDA error")
                sys.exit(2)

    for case in SHUFFLE_CASES:
        if args.filter not in case[0]:
            continue
        try:
            name, n, h, w, cin, cout, relu = case
            g = torch.Generator(device="cpu").manual_seed(0)
            x = torch.randn(n, h, w, cin, generator=g).to(dev).to(torch.bfloat16)
            wt = (torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5).to(dev)
            bias = torch.randn(cout, generator=g).to(dev)
            out = ops.conv2d_up2x_shuffle(x, packing.pack_up2x_shuffle(wt), bias.repeat(4).contiguous(), relu=relu)
            torch.cuda.synchronize()
            xi = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2, mode="nearest")
            ref = F.conv2d(xi, wt.to(torch.bfloat16).float(), bias, padding=1)
            ref = (ref.relu() if relu else ref).permute(0, 2, 3, 1)
            err = (out.float() - ref).abs()
            bad = (err > 1.5e-2 * ref.abs().clamp_min(1.0)).float().mean().item()
            print(f"{'OK  ' if bad == 0 else 'FAIL'} subpx {name:<24s} max_err={err.max().item():.4g} ref_max={ref.abs().max().item():.4g} bad_frac={bad:.4g}", flush=True)
            nfail += (bad != 0)
        except Exception as ex:  # noqa: BLE001
            nfail += 1
            print(f"EXC  subpx {case[0]}: {ex}", flush=True)
            traceback.print_exc()
            if "CUDA" in str(ex) or "fault" in str(ex):
                print("aborting after CUDA error")
                sys.exit(2)

    for case in SHUFFLE_RES_CASES:
        if args.filter not in case[0]:
            continue
        try:
            name, n, h, w, cin, cout, relu, with_res = case
            g = torch.Generator(device="cpu").manual_seed(0)
            x = torch.randn(n, h, w, cin, generator=g).to(dev).to(torch.bfloat16)
            wt = (torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5).to(dev)
            bias = torch.randn(cout, generator=g).to(dev)
            res = torch.randn(n, 2 * h, 2 * w, cout, generator=g).to(dev).to(torch.bfloat16) if with_res else None
            out = ops.conv2d_up2x_shuffle_res(x, packing.pack_up2x_shuffle(wt), bias.repeat(4).contiguous(), residual=res, relu=relu)
            torch.cuda.synchronize()
            xi = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2, mode="nearest")
            conv = F.conv2d(xi, wt.to(torch.bfloat16).float(), bias, padding=1)
            ref = conv + res.float().permute(0, 3, 1, 2) if res is not None else conv
            ref = (ref.relu() if relu else ref).permute(0, 2, 3, 1)
            err = (out.float() - ref).abs()
            scale = torch.maximum(ref.abs(), conv.abs().permute(0, 2, 3, 1)).clamp_min(1.0)
            bad = (err > 1.5e-2 * scale).float().mean().item()
            print(f"{'OK  ' if bad == 0 else 'FAIL'} subpar {name:<24s} max_err={err.max().item():.4g} ref_max={ref.abs().max().item():.4g} bad_frac={bad:.4g}", flush=True)
            nfail += (bad != 0)
        except Exception as ex:  # noqa: BLE001
            nfail += 1
            print(f"EXC  subpar {case[0]}: {ex}", flush=True)
            traceback.print_exc()
            if "CUDA" in str(ex) or "fault" in str(ex):
                print("aborting after CUDA error")
                sys.exit(2)

    for case in S2P_CASES:
        if args.filter not in case[0]:
            continue
        try:
            name, n, h, w, cin, cout, relu = case
            g = torch.Generator(device="cpu").manual_seed(0)
            x = torch.randn(n, h, w, cin, generator=g).to(dev).to(torch.bfloat16)
            wt = (torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5).to(dev)
            bias = torch.randn(cout, generator=g).to(dev)
            out = ops.conv2d_s2_planes(x, packing.pack_s2_planes(wt), bias, relu=relu)
            torch.cuda.synchronize()
            ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.to(torch.bfloat16).float(), bias, stride=2, padding=1)
            ref = (ref.relu() if relu else ref).permute(0, 2, 3, 1)
            err = (out.float() - ref).abs()
            bad = (err > 1e-2 * ref.abs().clamp_min(1.0)).float().mean().item()
            print(f"{'OK  ' if bad == 0 else 'FAIL'} s2pl  {name:<24s} max_err={err.max().item():.4g} ref_max={ref.abs().max().item():.4g} bad_frac={bad:.4g}", flush=True)
            nfail += (bad != 0)
        except Exception as ex:  # noqa: BLE001
            nfail += 1
            print(f"EXC  s2pl {case[0]}: {ex}", flush=True)
            traceback.print_exc()
            if "CUDA" in str(ex) or "fault" in str(ex):
                print("aborting after CUDA error")
                sys.exit(2)

    def s2d_nhwc(x):
        n, hh, ww, c = x.shape
        return x.reshape(n, hh // 2, 2, ww // 2, 2, c).permute(0, 1, 3, 2, 4, 5).reshape(n, hh // 2, ww // 2, 4 * c).contiguous()

    for case in S2D_CASES:
        if args.filter not in case[0]:
            continue
        try:
            name, n, h, w, relu = case
            g = torch.Generator(device="cpu").manual_seed(0)
            x = torch.randn(n, 2 * h, 2 * w, 16, generator=g).to(dev).to(torch.bfloat16)
            wt = (torch.randn(16, 16, 3, 3, generator=g) / 12).to(dev)
            bias = torch.randn(16, generator=g).to(dev)
            out = ops.conv2d_s2d(s2d_nhwc(x), packing.pack_s2d_conv3x3(wt), bias.repeat(4).contiguous(), relu=relu)
            torch.cuda.synchronize()
            ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.to(torch.bfloat16).float(), bias, padding=1)
            ref = s2d_nhwc((ref.relu() if relu else ref).permute(0, 2, 3, 1).contiguous())
            err = (out.float() - ref).abs()
            bad = (err > 1e-2 * ref.abs().clamp_min(1.0)).float().mean().item()
            print(f"{'OK  ' if bad == 0 else 'FAIL'} s2d   {name:<24s} max_err={err.max().item():.4g} ref_max={ref.abs().max().item():.4g} bad_frac={bad:.4g}", flush=True)
            nfail += (bad != 0)
            wh = (torch.randn(1, 16, 3, 3, generator=g) / 12).to(dev)
            bh = torch.randn(1, generator=g).to(dev)
            logits, mask = ops.head_s2d(s2d_nhwc(x), packing.pack_s2d_conv3x3(wh, 16), packing.pad_bias(bh, 16), threshold=0.5)
            torch.cuda.synchronize()
            refh = F.conv2d(x.float().permute(0, 3, 1, 2), wh.to(torch.bfloat16).float(), bh, padding=1)[:, 0]
            e = (logits - refh).abs().max().item()
            sure = refh.abs() > 2e-3
            okm = bool(torch.equal(mask[sure], ((refh > 0).to(torch.uint8) * 255)[sure]))
            print(f"{'OK  ' if (e <= 2e-3 and okm) else 'FAIL'} s2dhd {name:<24s} max_err={e:.4g} mask_ok={okm}", flush=True)
            nfail += (not (e <= 2e-3 and okm))
        except Exception as ex:  # noqa: BLE001
            nfail += 1
            print(f"EXC  s2d {case[0]}: {ex}", flush=True)
            traceback.print_exc()
            if "CUDA" in str(ex) or "fault" in str(ex):
                print("aborting after CUDA error")
                sys.exit(2)

    for case in SPX_CASES:
        if args.filter not in case[0]:
            continue
        try:
            name, n, h, w, cx, cs, cout, relu = case
            g = torch.Generator(device="cpu").manual_seed(0)
            x = torch.randn(n, h, w, cx, generator=g).to(dev).to(torch.bfloat16)
            skip = torch.randn(n, 2 * h, 2 * w, cs, generator=g).to(dev).to(torch.bfloat16)
            wt = (torch.randn(cout, cx + cs, 3, 3, generator=g) / ((cx + cs) * 9) ** 0.5).to(dev)
            bias = torch.randn(cout, generator=g).to(dev)
            out = ops.conv2d_upcat_subpixel(x, skip, packing.pack_upcat_subpixel(wt, cx), bias.repeat(4).contiguous(), relu=relu)
            torch.cuda.synchronize()
            xi = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2, mode="nearest")
            ref = F.conv2d(torch.cat([xi, skip.float().permute(0, 3, 1, 2)], 1), wt.to(torch.bfloat16).float(), bias, padding=1)
            ref = (ref.relu() if relu else ref).permute(0, 2, 3, 1)
            err = (out.float() - ref).abs()
            bad = (err > 1.5e-2 * ref.abs().clamp_min(1.0)).float().mean().item()
            print(f"{'OK  ' if bad == 0 else 'FAIL'} spx   {name:<24s} max_err={err.max().item():.4g} ref_max={ref.abs().max().item():.4g} bad_frac={bad:.4g}", flush=True)
            nfail += (bad != 0)
        except Exception as ex:  # noqa: BLE001
            nfail += 1
            print(f"EXC  spx {case[0]}: {ex}", flush=True)
            traceback.print_exc()
            if "CUDA" in str(ex) or "fault" in str(ex):
                print("aborting after CUDA error")
                sys.exit(2)

    # ---- glue ----
    try:
        g = torch.Generator().manual_seed(1)
        x = torch.randn(2, 32, 48, 64, generator=g).to(dev).to(torch.bfloat16)
        y = ops.maxpool3x3s2(x)
        ref = F.max_pool2d(x.float().permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1)
        ok = torch.equal(y.float(), ref)
        print(f"{'OK  ' if ok else 'FAIL'} maxpool exact={ok}", flush=True); nfail += (not ok)

        catbuf = torch.zeros(2, 64, 96, 64 + 32, dtype=torch.bfloat16, device=dev)
        ops.upsample2x(x, out=catbuf[..., :64])
        ref = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2, mode="nearest").permute(0, 2, 3, 1)
        ok = torch.equal(catbuf[..., :64].float(), ref) and bool((catbuf[..., 64:] == 0).all())
        print(f"{'OK  ' if ok else 'FAIL'} upsample exact={ok}", flush=True); nfail += (not ok)

        xin = torch.randn(2, 3, 64, 96, generator=g).to(dev)
        xs = ops.prep_input(xin)
        ref = torch.zeros(2, 32, 48, 16, device=dev)
        for ph in range(2):
            for pw in range(2):
                ref[..., (ph * 2 + pw) * 3:(ph * 2 + pw) * 3 + 3] = xin[:, :, ph::2, pw::2].permute(0, 2, 3, 1)
        ok = torch.equal(xs.float(), ref.to(torch.bfloat16).float())
        print(f"{'OK  ' if ok else 'FAIL'} prep_f32 exact={ok}", flush=True); nfail += (not ok)

        u8 = torch.randint(0, 256, (2, 64, 96, 3), generator=g, dtype=torch.uint8).to(dev)
        xs8 = ops.prep_input(u8)
        mean = torch.tensor([0.485, 0.456, 0.406], device=dev)
        std = torch.tensor([0.229, 0.224, 0.225], device=dev)
        xn = ((u8.float() / 255.0 - mean) / std).permute(0, 3, 1, 2).contiguous()
        ref8 = ops.prep_input(xn)
        d = (xs8.float() - ref8.float()).abs().max().item()
        ok = d <= 2e-2
        print(f"{'OK  ' if ok else 'FAIL'} prep_u8 max_diff_vs_f32_path={d:.4g}", flush=True); nfail += (not ok)

        # stem: 7x7/s2/p3 conv == 4x4/s1/p2 conv over the s2d tensor
        wt = (torch.randn(64, 3, 7, 7, generator=g) / 147 ** 0.5).to(dev)
        b = torch.randn(64, generator=g).to(dev)
        wp = packing.pack_stem_s2d(wt)
        y = ops.conv2d(xs, wp, b, 4, 4, 1, 2, relu=True)[:, :32, :48]
        # reference on the bf16-rounded input and bf16-rounded weights
        xr = xin.to(torch.bfloat16).float()
        ref = F.conv2d(xr, wt.to(torch.bfloat16).float(), b, stride=2, padding=3).relu().permute(0, 2, 3, 1)
        err = (y.float() - ref).abs()
        bad = (err > 1e-2 * ref.abs().clamp_min(1.0)).float().mean().item()
        print(f"{'OK  ' if bad == 0 else 'FAIL'} stem max_err={err.max().item():.4g} bad_frac={bad:.4g}", flush=True)
        nfail += (bad != 0)

        # head
        xh = torch.randn(2, 64, 96, 16, generator=g).to(dev).to(torch.bfloat16)
        wh = (torch.randn(1, 16, 3, 3, generator=g) / 12.0).to(dev)
        bh = torch.tensor([0.1], device=dev)
        wph = packing.pack_taps(wh)
        logits, mask = ops.head(xh, wph, packing.pad_bias(bh, 16), threshold=0.5)
        ref = F.conv2d(xh.float().permute(0, 3, 1, 2), wph.float()[:1].view(1, 3, 3, 16).permute(0, 3, 1, 2), bh,
                       padding=1)[:, 0]
        err = (logits - ref).abs().max().item()
        mref = (ref > 0).to(torch.uint8) * 255
        agree = (mask == mref).float().mean().item()
        near = ((mask != mref) & (ref.abs() > 1e-3)).sum().item()
        ok = err < 1e-3 and near == 0
        print(f"{'OK  ' if ok else 'FAIL'} head max_err={err:.4g} mask_agree={agree:.6f} wrong_outside_band={near}", flush=True)
        nfail += (not ok)
    except Exception as ex:  # noqa: BLE001
        nfail += 1
        print(f"EXC  glue: {ex}", flush=True)
        traceback.print_exc()
    print(f"SUMMARY failures={nfail}", flush=True)
    return 1 if nfail else 0


if __name__ == "__main__":
    sys.exit(main())
