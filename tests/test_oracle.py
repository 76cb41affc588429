"""The oracle restatement against the structural known answers of smp.Unet
and against the committed golden vectors (tests/golden, made by tools/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import unet_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("enc,params,entries", [("resnet34", 24_436_369, 278), ("resnet50", 32_521_105, 380)])
def test_structural_kat(enc, params, entries):
    m = O.Unet(enc)
    sd = m.state_dict()
    assert sum(p.numel() for p in m.parameters()) == params
    assert len(sd) == entries
    assert sd["segmentation_head.0.weight"].shape == (1, 16, 3, 3)
    assert sd["segmentation_head.0.bias"].shape == (1,)
    assert "decoder.blocks.4.conv2.1.num_batches_tracked" in sd
    assert not any(k.startswith(("decoder.center", "encoder.fc", "segmentation_head.1", "segmentation_head.2"))
                   for k in sd)
    if enc == "resnet34":
        assert sd["encoder.conv1.weight"].shape == (64, 3, 7, 7)
        assert sd["decoder.blocks.0.conv1.0.weight"].shape == (256, 768, 3, 3)
        assert sd["decoder.blocks.4.conv2.0.weight"].shape == (16, 16, 3, 3)
        assert sd["encoder.layer2.0.downsample.0.weight"].shape == (128, 64, 1, 1)
        assert sum(v.numel() for k, v in sd.items() if "running" in k or "tracked" in k) == 19_054
    else:
        assert sd["decoder.blocks.0.conv1.0.weight"].shape == (256, 3072, 3, 3)


def test_flops_kat_resnet34_512():
    m = O.build("resnet34")
    assert abs(O.conv_flops_per_image(m, 512, 512) / 1e9 - 62.512) < 1e-3


def test_shape_check_and_output_shape():
    m = O.build("resnet34")
    with pytest.raises(RuntimeError, match="divisible by 32"):
        m(torch.zeros(1, 3, 100, 64))
    with torch.no_grad():
        assert m(torch.zeros(2, 3, 64, 96)).shape == (2, 1, 64, 96)


def test_decoder_concat_order_upsampled_first():
    blk = O.DecoderBlock(4, 2, 3).eval()
    x, skip = torch.randn(1, 4, 2, 2), torch.randn(1, 2, 4, 4)
    with torch.no_grad():
        up = x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)        # out[i,j] = in[i//2, j//2]
        ref = blk.conv2(blk.conv1(torch.cat([up, skip], 1)))
        assert torch.allclose(blk(x, skip), ref)
        assert not torch.allclose(blk(x, skip), blk.conv2(blk.conv1(torch.cat([skip, up], 1))))


@pytest.mark.parametrize("enc", ["resnet34", "resnet50"])
def test_golden_vectors(enc):
    g = np.load(os.path.join(GOLD, f"unet_{enc}_64.npz"))
    m = O.build(enc, seed=int(g["model_seed"]), random_bn=True)
    assert sum(p.numel() for p in m.parameters()) == int(g["n_params"])
    x = O.image_like_input(int(g["batch"]), int(g["size"]), seed=int(g["input_seed"]))
    with torch.no_grad():
        y = m(x).numpy()
    scale = np.abs(g["logits_fp32"]).max()
    assert np.abs(y - g["logits_fp32"]).max() <= 1e-3 * scale
    yemu, feats = O.forward_bf16_emulated(m, x, return_features=True)
    # the emulation rounds to bf16 after every layer: a different CPU summation order can flip a few roundings
    d = np.abs(yemu.numpy() - g["logits_bf16emu"])
    assert d.max() <= 8e-2 * scale and d.mean() <= 1e-2 * scale
    assert np.abs(feats["encoder.stem"][:, :8].numpy() - g["stem_bf16emu"]).max() <= 2e-2


def test_bf16_emulation_tracks_fp32():
    m = O.build("resnet34", seed=1, random_bn=True)
    x = O.image_like_input(1, 64, seed=2)
    with torch.no_grad():
        y = m(x)
    ye = O.forward_bf16_emulated(m, x)
    assert (y - ye).abs().max() <= 0.08 * y.abs().max()
    assert (y - ye).abs().mean() <= 0.05 * y.std()


def test_val_transform_golden():
    g = np.load(os.path.join(GOLD, "val_transform_37x53_to_32.npz"))
    out = O.val_transform(g["image"], 32).numpy()
    assert out.shape == (3, 32, 32)
    assert np.abs(out - g["out"]).max() < 1e-5


def test_binarize_conventions():
    y = torch.tensor([[-1.0, 0.2, 0.6, 3.0]])
    assert O.binarize(y, 0.5, sigmoid=False).tolist() == [[0, 0, 255, 255]]      # raw > 0.5  (predict.py:624)
    assert O.binarize(y, 0.5, sigmoid=True).tolist() == [[0, 255, 255, 255]]     # sigmoid > 0.5 <=> logit > 0


def test_dice_bce_golden_and_closed_form():
    g = np.load(os.path.join(GOLD, "dice_bce.npz"))
    lg, tg = torch.from_numpy(g["logits"]), torch.from_numpy(g["target"])
    assert abs(float(O.dice_loss_binary(lg, tg)) - float(g["dice"])) < 1e-6
    assert abs(float(O.dice_bce_loss(lg, tg)) - float(g["combo"])) < 1e-6
    p = torch.sigmoid(lg)
    dice = (2 * (p * tg).sum() + 1e-5) / ((p + tg).sum() + 1e-5)      # batch-global sums
    assert abs(float(O.dice_loss_binary(lg, tg)) - float(1 - dice)) < 1e-6
    assert float(O.dice_loss_binary(lg, torch.zeros_like(tg))) == 0.0    # empty target masks the class out
