"""End-to-end parity + first timing of the whole-model plan on the GPU.

    UWM_KEEP_ALL=1 python tools/gpu_model_check.py [--encoder resnet34] [--size 128] [--batch 2] [--bench]
"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import unet_oracle as O  # noqa: E402  (checker only)
from unet_watermark_b200.unet_model import Unet  # noqa: E402


def stats(name, a, b):
    d = (a - b).abs()
    print(f"  {name:<22s} max_abs={d.max().item():.4g}  mean_abs={d.mean().item():.4g}  ref_absmax={b.abs().max().item():.4g} "
          f"ref_std={b.std().item():.4g}  rel_max={d.max().item() / max(b.abs().max().item(), 1e-9):.4g}", flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--encoder", default="resnet34")
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--bench", action="store_true")
    ap.add_argument("--bench-size", type=int, default=512)
    ap.add_argument("--bench-batch", type=int, default=16)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    print(torch.cuda.get_device_name(0), flush=True)

    ref = O.build(args.encoder, seed=0, random_bn=True)
    model = Unet(args.encoder, encoder_weights=None)
    model.load_state_dict(ref.state_dict(), strict=True)
    model = model.to(dev).eval()

    x = O.image_like_input(args.batch, args.size, seed=3)
    with torch.no_grad():
        y32 = ref(x)
    yemu, feats = O.forward_bf16_emulated(ref, x, return_features=True)
    model.use_cuda_graph = False
    y = model(x.to(dev)).cpu()
    print(f"[{args.encoder} {args.size}x{args.size} B={args.batch}] logits vs oracle:", flush=True)
    stats("vs bf16-emulated", y, yemu)
    stats("vs fp32", y, y32)
    stats("emulated vs fp32", yemu, y32)
    if os.environ.get("UWM_KEEP_ALL") == "1":
        eng = model.engine(args.batch, args.size, args.size)
        for name, f in feats.items():
            t = eng.read_tensor(name, args.batch).float().cpu().permute(0, 3, 1, 2)
            stats(name, t, f)
    # graph path must give identical bits
    model.use_cuda_graph = True
    yg = model(x.to(dev)).cpu()
    yg2 = model(x.to(dev)).cpu()
    print("  graph == eager:", torch.equal(yg, y), " graph replay stable:", torch.equal(yg, yg2), flush=True)
    # masks
    m_sig = model.predict_mask(x.to(dev), 0.5, sigmoid=True).cpu()
    m_raw = model.predict_mask(x.to(dev), 0.5, sigmoid=False).cpu()
    print("  mask(sigmoid) == (logits>0):", torch.equal(m_sig, (y[:, 0] > 0).to(torch.uint8) * 255),
          " mask(raw) == (logits>0.5):", torch.equal(m_raw, (y[:, 0] > 0.5).to(torch.uint8) * 255),
          " agreement with fp32 oracle:", (m_sig == O.binarize(y32[:, 0])).float().mean().item(), flush=True)
    # u8 input path
    u8 = O.image_like_u8(args.batch, args.size, seed=3)
    mean = torch.tensor(O.IMAGENET_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(O.IMAGENET_STD).view(1, 3, 1, 1)
    xn = (u8.permute(0, 3, 1, 2).float() / 255.0 - mean) / std
    mu, lu = model.predict_mask(u8.to(dev), 0.5, return_logits=True)
    lf = model(xn.to(dev))
    stats("u8 path vs f32 path", lu.cpu(), lf.cpu())

    if args.bench:
        B, S = args.bench_batch, args.bench_size
        xb = O.image_like_u8(B, S, seed=1).to(dev)
        eng = model.engine(B, S, S)
        print(f"bench {args.encoder} B={B} {S}x{S}: workspace {eng.workspace_bytes / 2**20:.1f} MiB, "
              f"{eng.kernels_per_forward} kernels, {eng.flops_per_image / 1e9:.3f} GFLOP/img", flush=True)
        for _ in range(3):
            model.predict_mask(xb)
        torch.cuda.synchronize()
        for graph in (False, True):
            model.use_cuda_graph = graph
            for _ in range(3):
                model.predict_mask(xb)
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            n = 20
            e0.record()
            for _ in range(n):
                model.predict_mask(xb)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            tf = eng.flops_per_image * B / ms / 1e9
            print(f"  graph={graph}: {ms:.3f} ms/batch  {B / ms * 1e3:.0f} img/s  {tf:.1f} TFLOP/s", flush=True)
        prof = eng.profile(xb)
        tot = sum(p[1] for p in prof)
        print(f"  per-kernel (eager, events): total {tot:.3f} ms")
        for nm, ms, fl, by in prof:
            print(f"    {nm:<34s} {ms * 1e3:8.1f} us  {fl / ms / 1e9 if ms > 0 else 0:8.1f} TF/s  {by / ms / 1e6 if ms > 0 else 0:8.1f} GB/s",
                  flush=True)
        # torch/cuDNN control on the same box: oracle module in bf16 channels_last
        ctrl = ref.to(dev).to(torch.bfloat16).to(memory_format=torch.channels_last)
        xc = O.image_like_input(B, S, seed=1).to(dev).to(torch.bfloat16).to(memory_format=torch.channels_last)
        torch.backends.cudnn.benchmark = True
        with torch.no_grad():
            for _ in range(5):
                ctrl(xc)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            for _ in range(20):
                ctrl(xc)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"  GPU control (torch bf16 channels_last cuDNN): {ms:.3f} ms/batch  {B / ms * 1e3:.0f} img/s", flush=True)


if __name__ == "__main__":
    main()
