"""A/B of the multi-layer chain launches (UWM_CHAIN=0|1): bit-identical logits on several shapes, per-launch times.

    python tools/gpu_chain_ab.py            (spawns one subprocess per setting: the switch is read once per process)
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHAPES = [("resnet34", 96, 160, 3), ("resnet34", 128, 128, 2), ("resnet34", 64, 64, 5), ("resnet34", 512, 512, 16),
          ("resnet34", 1024, 1024, 4), ("resnet34", 256, 384, 7)]


def child(out_dir):
    import torch
    sys.path.insert(0, ROOT)
    from oracle import unet_oracle as O
    from unet_watermark_b200.unet_model import Unet
    dev = torch.device("cuda:0")
    ref = O.build("resnet34", seed=0, random_bn=True)
    m = Unet("resnet34", encoder_weights=None)
    m.load_state_dict(ref.state_dict())
    m = m.to(dev).eval()
    for enc, h, w, b in SHAPES:
        x = O.image_like_u8(b, (h, w), seed=h + w + b).to(dev)
        mask, logits = m.predict_mask(x, 0.5, return_logits=True)
        l2 = m.predict_mask(x, 0.5, return_logits=True)[1]
        assert torch.equal(logits, l2), "non-deterministic"
        torch.save(logits.cpu(), os.path.join(out_dir, f"logits_{h}x{w}x{b}.pt"))
    eng = m.engine(16, 512, 512)
    x = O.image_like_u8(16, 512, seed=1).to(dev)
    prof = None
    for _ in range(4):
        prof = eng.profile(x, 0.5)
    rows = [(n, ms * 1e3) for n, ms, fl, by in prof]
    json.dump(rows, open(os.path.join(out_dir, "profile.json"), "w"))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        return child(sys.argv[2])
    import torch
    outs = {}
    for setting in ("0", "1"):
        d = os.path.join(ROOT, "gpurun_out", f"chain_ab_{setting}")
        os.makedirs(d, exist_ok=True)
        env = dict(os.environ, UWM_CHAIN=setting)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", d], env=env, capture_output=True, text=True, timeout=600)
        if r.returncode != 0:
            print(f"UWM_CHAIN={setting} failed:\n{r.stdout[-2000:]}\n{r.stderr[-3000:]}")
            return 1
        outs[setting] = d
    ok = True
    for enc, h, w, b in SHAPES:
        a = torch.load(os.path.join(outs["0"], f"logits_{h}x{w}x{b}.pt"))
        c = torch.load(os.path.join(outs["1"], f"logits_{h}x{w}x{b}.pt"))
        same = torch.equal(a, c)
        ok &= same
        print(f"{h}x{w} B={b}: chain == single launches bit for bit: {same}" + ("" if same else f"  max|d| {(a - c).abs().max().item():.4f}"))
    for d in outs.values():                       # the logits are large: gpurun_out/ only travels back below 64 MiB
        for f in os.listdir(d):
            if f.endswith(".pt"):
                os.remove(os.path.join(d, f))
    p0 = json.load(open(os.path.join(outs["0"], "profile.json")))
    p1 = json.load(open(os.path.join(outs["1"], "profile.json")))
    print(f"eager per-launch sum: single {sum(t for _, t in p0):.1f} us ({len(p0)} launches), chains {sum(t for _, t in p1):.1f} us ({len(p1)} launches)")
    for n, t in p1:
        if n.startswith("chain["):
            first, last = n.split("] ")[1].split(" .. ")
            names = [x for x, _ in p0]
            i0 = names.index(first)
            cnt = int(n[6:n.index("]")])
            single = sum(t0 for _, t0 in p0[i0:i0 + cnt])
            print(f"  {n}: {t:.1f} us as one launch vs {single:.1f} us as {cnt} launches")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
