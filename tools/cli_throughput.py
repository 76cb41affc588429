"""Throughput of the CLI surface itself (`main.py predict`): files in, `<stem>_mask.png` files out.

    python tools/cli_throughput.py [--images 256] [--size-mix photo|net] [--batch-size 16] [--workers 8]
                                   [--mask-type auto|watermark] [--no-post-process] [--ref-images 4]

Builds a synthetic folder (JPEG + PNG files of mixed sizes, seeded), saves a random-init checkpoint in the reference
trainer's format, runs `unet_watermark_b200.cli.main(["predict", ...])` and reports images/s end to end (decode on CPU
threads -> GPU resize / network / upscale / post-processing -> PNG encode on CPU threads).  Next to it, the
reference's own loop (src/predict.py:588-664: one image per forward, cv2 on the host) restated with the oracle
network on the host cores, on a bounded sample.  Prints one JSON line.
"""
import argparse
import glob
import json
import os
import shutil
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def make_folder(path, n, mix, seed=0):
    import cv2
    rng = np.random.default_rng(seed)
    os.makedirs(path, exist_ok=True)
    photo = [(640, 480), (1280, 720), (1920, 1080), (800, 1200), (1024, 1024), (512, 512), (3000, 2000)]
    total_px = 0
    for i in range(n):
        w, h = (512, 512) if mix == "net" else photo[int(rng.integers(0, len(photo)))]
        low = rng.random((max(h // 32, 2), max(w // 32, 2), 3)).astype(np.float32)
        img = cv2.resize(low, (w, h), interpolation=cv2.INTER_CUBIC)
        img = np.clip(img + 0.03 * rng.standard_normal((h, w, 3)).astype(np.float32), 0, 1)
        cv2.rectangle(img, (w // 4, h // 3), (w // 2, h // 2), (0.9, 0.9, 0.9), -1)          # a bright "watermark"
        u8 = (img * 255).astype(np.uint8)
        ext = "png" if i % 4 == 0 else "jpg"
        cv2.imwrite(os.path.join(path, f"img_{i:05d}.{ext}"), u8)
        total_px += w * h
    return total_px


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=256)
    ap.add_argument("--size-mix", choices=["photo", "net"], default="photo")
    ap.add_argument("--batch-size", type=int, default=16)
    ap.add_argument("--workers", type=int, default=os.cpu_count() or 8)
    ap.add_argument("--mask-type", default="auto")
    ap.add_argument("--no-post-process", action="store_true")
    ap.add_argument("--ref-images", type=int, default=4)
    args = ap.parse_args()

    from oracle import unet_oracle as O
    from unet_watermark_b200 import cli
    work = tempfile.mkdtemp(prefix="uwm_cli_")
    try:
        inp, out = os.path.join(work, "in"), os.path.join(work, "out")
        t0 = time.time()
        px = make_folder(inp, args.images, args.size_mix)
        t_make = time.time() - t0
        ref = O.build("resnet34", seed=0, random_bn=True)
        ckpt = os.path.join(work, "model.pth")
        torch.save({"epoch": 1, "model_state_dict": ref.state_dict(), "val_loss": 0.0, "val_metrics": {}}, ckpt)
        argv = ["predict", "--input", inp, "--output", out, "--model", ckpt, "--model-name", "Unet",
                "--batch-size", str(args.batch_size), "--workers", str(args.workers), "--mask-type", args.mask_type]
        if args.no_post_process:
            argv.append("--no-post-process")
        # warm-up run on a few files (engine build, CUDA graph capture, cv2 thread pools), then the timed run
        warm_in = os.path.join(work, "warm")
        os.makedirs(warm_in)
        for p in sorted(glob.glob(os.path.join(inp, "*")))[:args.batch_size]:
            shutil.copy(p, warm_in)
        cli.main(["predict", "--input", warm_in, "--output", os.path.join(work, "warm_out"), *argv[5:]])
        torch.cuda.synchronize()
        t0 = time.time()
        rc = cli.main(argv)
        torch.cuda.synchronize()
        dt = time.time() - t0
        n_out = len(glob.glob(os.path.join(out, "*_mask.png")))
        line = {"metric": "images/sec through `main.py predict` (files in, mask PNGs out)", "value": n_out / dt,
                "unit": "images/s", "images": n_out, "seconds": dt, "megapixels_per_s": px / 1e6 / dt, "rc": rc,
                "config": {"size_mix": args.size_mix, "batch_size": args.batch_size, "decode_threads": args.workers,
                           "mask_type": args.mask_type, "post_process": not args.no_post_process,
                           "host_cores": os.cpu_count(), "folder_build_s": t_make,
                           "note": "includes model load + engine build + graph capture of the timed process run"}}
        # the reference loop on the host cores: one image per forward, cv2 pre/post (bounded sample)
        if args.ref_images > 0:
            import cv2
            from tests import cv2_reference as R
            torch.set_num_threads(os.cpu_count() or 1)
            files = sorted(glob.glob(os.path.join(inp, "*")))[:args.ref_images]
            t0 = time.time()
            for p in files:
                image = cv2.imread(p)
                rgb = cv2.cvtColor(image, cv2.COLOR_BGR2RGB)
                x = O.val_transform(rgb, 512).unsqueeze(0)
                with torch.no_grad():
                    m = ref(x)[0, 0].numpy()
                m = cv2.resize(m, (image.shape[1], image.shape[0]))
                b = (m > 0.5).astype(np.uint8) * 255
                if not args.no_post_process:
                    t = R.detect_watermark_type(rgb, b) if args.mask_type == "auto" else args.mask_type
                    b = R.optimize_mask(b, t)
                cv2.imwrite(os.path.join(work, "ref_mask.png"), b)
            dtr = time.time() - t0
            line["reference_loop"] = {"value": len(files) / dtr, "unit": "images/s", "images": len(files), "cores": os.cpu_count(),
                                      "kind": "port (oracle network + the reference's cv2 calls, src/predict.py:588-664)"}
        print(json.dumps(line), flush=True)
    finally:
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
