"""Sub-batch sweep: images/s of ``Unet.predict_mask`` on one GPU when a batch of B images runs as consecutive forwards of
``sub`` images on slices of the same buffers (Engine.forward ``sub_batch``; results are bit-identical, checked here).

    python tools/gpu_subbatch_sweep.py [--json gpurun_out/subbatch.json]

One process for every (encoder, size, batch) of BASELINE.json configs 2-4; inputs rotate through a pool larger than the
126 MB L2; every (buffer, sub-batch) pair's CUDA graph is captured before the timed region; CUDA events on the launching stream.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unet_watermark_b200.unet_model import Unet  # noqa: E402

CASES = [  # encoder, size, batch, sub-batches (0 = whole batch in one plan)
    ("resnet50", 768, 32, (0, 16, 8, 4, 2)),
    ("resnet34", 1024, 64, (0, 16, 8, 4)),
    ("resnet34", 512, 16, (0, 8, 4)),
]

ap = argparse.ArgumentParser()
ap.add_argument("--json", default="")
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--only", default="", help="encoder name filter")
a = ap.parse_args()
dev = torch.device("cuda:0")
out = []
for enc, size, batch, subs in CASES:
    if a.only and a.only != enc:
        continue
    torch.manual_seed(0)
    m = Unet(enc, encoder_weights=None).to(dev).eval()
    n_pool = max(2, min(12, -(-160_000_000 // (batch * size * size * 3))))
    g = torch.Generator().manual_seed(1)
    pool = [torch.randint(0, 256, (batch, size, size, 3), dtype=torch.uint8, generator=g).to(dev) for _ in range(n_pool)]
    masks = [torch.empty(batch, size, size, dtype=torch.uint8, device=dev) for _ in range(2)]
    eng = m.engine(batch, size, size)
    ref = None
    for sb in subs:
        m.sub_batch = sb
        for i in range(n_pool + 3):
            m.predict_mask(pool[i % n_pool], 0.5, out=masks[i % 2])
        torch.cuda.synchronize()
        cs = int(m.predict_mask(pool[0], 0.5).sum().item())
        ref = cs if ref is None else ref
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = None
        for _ in range(2):
            e0.record()
            for i in range(a.steps):
                m.predict_mask(pool[i % n_pool], 0.5, out=masks[i % 2])
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.steps
            best = ms if best is None else min(best, ms)
        tf = eng.flops_per_image * batch / best / 1e9
        row = {"encoder": enc, "size": size, "batch": batch, "sub_batch": sb, "ms_per_batch": best,
               "images_per_s": batch / best * 1e3, "tflops": tf, "mask_checksum_equal": cs == ref}
        out.append(row)
        print(f"{enc} {size}x{size} B={batch} sub={sb or batch:>3d}: {best:8.3f} ms  {row['images_per_s']:9.0f} img/s  "
              f"{tf:7.1f} TF/s  identical={cs == ref}", flush=True)
    del m, eng, pool, masks
    torch.cuda.empty_cache()
if a.json:
    with open(a.json, "w") as f:
        json.dump(out, f, indent=1)
