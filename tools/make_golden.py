"""Generate tests/golden/*.npz from the oracle (run in the dev container; commits small fixtures).

The reference cannot be imported here (segmentation_models_pytorch is not installed),
so these vectors pin the ORACLE RESTATEMENT against itself (regression) and give the GPU tests a
machine-independent target; they do not pin it against smp — "parity unpinned" stands.

    python tools/make_golden.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import unet_oracle as O  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)          # deterministic summation order
    for enc, size in (("resnet34", 64), ("resnet50", 64)):
        m = O.build(enc, seed=0, random_bn=True)
        x = O.image_like_input(2, size, seed=5)
        with torch.no_grad():
            y32 = m(x)
        yemu, feats = O.forward_bf16_emulated(m, x, return_features=True)
        np.savez_compressed(
            os.path.join(OUT, f"unet_{enc}_{size}.npz"),
            model_seed=0, input_seed=5, size=size, batch=2,
            logits_fp32=y32.numpy(), logits_bf16emu=yemu.numpy(),
            stem_bf16emu=feats["encoder.stem"][:, :8].numpy(),          # first 8 channels only (size)
            layer4_absmean=float(feats["encoder.layer4"].abs().mean()),
            n_params=sum(p.numel() for p in m.parameters()), n_entries=len(m.state_dict()))
        print(enc, y32.abs().max().item(), (y32 - yemu).abs().max().item())
    # preprocessing golden: a deterministic 37x53 RGB gradient image -> val_transform(32)
    yy, xx = np.mgrid[0:37, 0:53]
    img = np.stack([(xx * 5) % 256, (yy * 7) % 256, ((xx + yy) * 3) % 256], -1).astype(np.uint8)
    t = O.val_transform(img, 32)
    np.savez_compressed(os.path.join(OUT, "val_transform_37x53_to_32.npz"), image=img, out=t.numpy())
    # loss golden
    g = torch.Generator().manual_seed(11)
    logits = torch.randn(2, 1, 8, 8, generator=g) * 3
    target = (torch.rand(2, 1, 8, 8, generator=g) > 0.7).float()
    np.savez_compressed(os.path.join(OUT, "dice_bce.npz"), logits=logits.numpy(), target=target.numpy(),
                        dice=float(O.dice_loss_binary(logits, target)), combo=float(O.dice_bce_loss(logits, target)))
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
