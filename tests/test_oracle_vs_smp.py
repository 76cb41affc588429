"""Pins for the oracle restatement.

1. ENCODER PIN (always runs): the oracle's encoder features are bit-for-bit the stage outputs of the stock
   ``torchvision.models.resnet34/50`` classes (the class smp's ResNetEncoder subclasses) carrying the same weights.
2. SMP PIN (runs only where ``segmentation_models_pytorch`` is importable): build ``smp.Unet`` with the reference's
   exact kwargs (reference src/models/unet_model.py:64-71, defaults src/configs/config.py:15-22), copy the state
   dict and require identical key lists and outputs to 1e-5.  smp is an un-vendored, un-pinned pip dependency of
   the reference (requirements.txt:17) and is absent from this image: here the test SKIPS, loudly, and the oracle's
   decoder/head stay "parity unpinned" (DESIGN.md §1c).
"""
import pytest
import torch
import torchvision

from oracle import unet_oracle as O


@pytest.mark.parametrize("enc", ["resnet34", "resnet50"])
def test_encoder_features_equal_torchvision_bit_for_bit(enc):
    ref = O.build(enc, seed=3, random_bn=True)
    tv = getattr(torchvision.models, enc)(weights=None)
    sd = {k[len("encoder."):]: v for k, v in ref.state_dict().items() if k.startswith("encoder.")}
    sd["fc.weight"], sd["fc.bias"] = tv.fc.weight.detach(), tv.fc.bias.detach()      # smp deletes the classifier
    tv.load_state_dict(sd, strict=True)
    tv.eval()
    x = O.image_like_input(2, 64, seed=11)
    with torch.no_grad():
        feats = ref.encoder(x)
        # torchvision ResNet._forward_impl, stage by stage
        y = tv.relu(tv.bn1(tv.conv1(x)))
        stages = [x, y]
        y = tv.layer1(tv.maxpool(y)); stages.append(y)
        for layer in (tv.layer2, tv.layer3, tv.layer4):
            y = layer(y); stages.append(y)
    assert len(feats) == 6
    for a, b in zip(feats, stages):
        assert a.shape == b.shape and torch.equal(a, b)
    # stage shapes of smp's encoder table (out_channels, strides 1..32)
    assert [f.shape[1] for f in feats] == list(O.ENCODERS[enc][2])
    assert [x.shape[-1] // f.shape[-1] for f in feats] == [1, 2, 4, 8, 16, 32]


def test_encoder_construction_rng_stream_matches_torchvision():
    """Seeded construction draws the same random numbers as torchvision's ResNet (smp builds it through
    ``super().__init__``), so 'identical random-init weights' means the same thing on both sides."""
    torch.manual_seed(5)
    enc = O.ResNetEncoder(O.ENCODERS["resnet34"][2], block=O.ENCODERS["resnet34"][0], layers=O.ENCODERS["resnet34"][1])
    torch.manual_seed(5)
    tv = torchvision.models.resnet34(weights=None)
    for k, v in enc.state_dict().items():
        assert torch.equal(v, tv.state_dict()[k]), k


@pytest.mark.parametrize("enc", ["resnet34", "resnet50"])
def test_oracle_equals_smp_unet(enc):
    smp = pytest.importorskip(
        "segmentation_models_pytorch",
        reason="segmentation_models_pytorch is not installed (un-vendored dependency of the reference, "
               "requirements.txt:17; no network): the oracle's decoder/head restatement stays UNPINNED")
    ref = O.build(enc, seed=0, random_bn=True)
    # reference src/models/unet_model.py:64-71 with the defaults of src/configs/config.py:15-22
    model = smp.Unet(encoder_name=enc, encoder_weights=None, in_channels=3, classes=1, activation=None,
                     encoder_depth=5, decoder_channels=[256, 128, 64, 32, 16])
    assert list(model.state_dict().keys()) == list(ref.state_dict().keys())
    model.load_state_dict(ref.state_dict(), strict=True)
    model.eval()
    x = O.image_like_input(2, 64, seed=4)
    with torch.no_grad():
        assert torch.allclose(model(x), ref(x), atol=1e-5, rtol=1e-5)
    # seeded initialisation draws the same numbers (decoder kaiming-uniform, head xavier-uniform)
    torch.manual_seed(9)
    a = smp.Unet(encoder_name=enc, encoder_weights=None)
    torch.manual_seed(9)
    b = O.Unet(enc)
    for (ka, va), (kb, vb) in zip(a.state_dict().items(), b.state_dict().items()):
        assert ka == kb and torch.equal(va, vb), ka
