"""Host-side logic of the product module (no GPU): factory surface, state-dict layout, weight packing."""
import copy
import pickle

import pytest
import torch
import torch.nn.functional as F

from oracle import unet_oracle as O
from unet_watermark_b200 import packing
from unet_watermark_b200.config import get_cfg_defaults
from unet_watermark_b200.unet_model import SMPModelFactory, Unet, WatermarkSegmentationModel, create_model_from_config


@pytest.mark.parametrize("enc", ["resnet34", "resnet50"])
def test_state_dict_layout_equals_oracle_and_loads_strict(enc):
    torch.manual_seed(0)
    ref = O.Unet(enc)
    torch.manual_seed(0)
    ours = Unet(enc, encoder_weights=None)
    sa, sb = ref.state_dict(), ours.state_dict()
    assert list(sa.keys()) == list(sb.keys())
    assert all(sa[k].shape == sb[k].shape and sa[k].dtype == sb[k].dtype for k in sa)
    assert all(torch.equal(sa[k], sb[k]) for k in sa), "seeded construction must reproduce smp's init stream"
    ours.load_state_dict(O.build(enc, seed=3, random_bn=True).state_dict(), strict=True)
    ref.load_state_dict(ours.state_dict(), strict=True)


def test_factory_surface_and_errors():
    with pytest.raises(ValueError, match="Unsupported model: Foo"):
        SMPModelFactory.create_model("Foo")
    with pytest.raises(NotImplementedError):
        SMPModelFactory.create_model("MAnet", encoder_weights=None)
    with pytest.raises(KeyError):
        SMPModelFactory.create_model("Unet", encoder_name="vgg16", encoder_weights=None)
    with pytest.raises(NotImplementedError):
        SMPModelFactory.create_model("Unet", encoder_weights=None, classes=3)
    m = SMPModelFactory.create_model("Unet", "resnet34", None, 3, 1, None, decoder_channels=[256, 128, 64, 32, 16],
                                     encoder_depth=5)
    assert isinstance(m, Unet)
    assert set(SMPModelFactory.SUPPORTED_MODELS) == {"Unet", "UnetPlusPlus", "MAnet", "Linknet", "FPN", "PSPNet",
                                                     "PAN", "DeepLabV3", "DeepLabV3Plus"}
    assert "resnet34" in SMPModelFactory.get_available_encoders()
    assert SMPModelFactory.get_encoder_info("resnet50")["out_channels"] == (3, 64, 256, 512, 1024, 2048)
    assert "error" in SMPModelFactory.get_encoder_info("nope")


def test_create_model_from_config_and_wrapper():
    cfg = get_cfg_defaults()
    cfg.MODEL.NAME = "Unet"
    cfg.MODEL.ENCODER_WEIGHTS = None
    m = create_model_from_config(cfg)
    assert isinstance(m, Unet) and m.decoder_channels == (256, 128, 64, 32, 16)
    w = WatermarkSegmentationModel(cfg)
    info = w.get_model_info()
    assert info["total_params"] == 24_436_369 and info["model_name"] == "Unet"
    assert all(k.startswith("model.") for k in w.state_dict())


def test_no_cpu_fallback():
    m = Unet("resnet34", encoder_weights=None).eval()
    with pytest.raises(RuntimeError, match="no CPU"):
        m(torch.zeros(1, 3, 64, 64))
    with pytest.raises(RuntimeError, match="no CPU"):
        m.predict_mask(torch.zeros(1, 64, 64, 3, dtype=torch.uint8))


def test_module_is_copyable_and_picklable():
    m = Unet("resnet34", encoder_weights=None).eval()
    m2 = copy.deepcopy(m)
    m3 = pickle.loads(pickle.dumps(m))
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), m2.state_dict().values()))
    assert list(m3.state_dict()) == list(m.state_dict())


def test_fold_bn_equals_conv_bn_eval():
    torch.manual_seed(0)
    conv = torch.nn.Conv2d(8, 12, 3, padding=1, bias=False)
    bn = torch.nn.BatchNorm2d(12).eval()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5); bn.bias.normal_(); bn.running_mean.normal_(); bn.running_var.uniform_(0.5, 2)
    x = torch.randn(2, 8, 9, 9)
    w, b = packing.fold_bn(conv.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps)
    with torch.no_grad():
        assert torch.allclose(F.conv2d(x, w, b, padding=1), bn(conv(x)), atol=1e-5)


def test_pack_taps_layout_is_tap_major_channel_minor():
    w = torch.randn(5, 16, 3, 3)
    p = packing.pack_taps(w)
    assert p.shape == (16, 9 * 16) and p.dtype == torch.bfloat16
    assert torch.equal(p[5:], torch.zeros_like(p[5:]))
    x = torch.randn(1, 16, 6, 6)
    cols = F.unfold(x, 3, padding=1).view(1, 16, 9, 36).permute(0, 2, 1, 3).reshape(1, 144, 36)   # (tap, c) order
    y = (p[:5].float() @ cols[0]).view(5, 6, 6)
    ref = F.conv2d(x, w.to(torch.bfloat16).float(), padding=1)[0]
    assert torch.allclose(y, ref, atol=1e-4)


def test_pack_stem_s2d_is_the_same_convolution():
    """conv7x7/s2/p3 == conv4x4/s1 (taps -2..1) over the 2x2 space-to-depth input with the repacked weights."""
    torch.manual_seed(1)
    w = torch.randn(64, 3, 7, 7)
    x = torch.randn(2, 3, 32, 48)
    ref = F.conv2d(x, w.to(torch.bfloat16).float(), stride=2, padding=3)
    xs = torch.zeros(2, 16, 16, 24)
    for ph in range(2):
        for pw in range(2):
            xs[:, (ph * 2 + pw) * 3:(ph * 2 + pw) * 3 + 3] = x[:, :, ph::2, pw::2]
    wp = packing.pack_stem_s2d(w).float().view(64, 4, 4, 16).permute(0, 3, 1, 2)       # [64,16,4,4]
    y = F.conv2d(F.pad(xs, (2, 1, 2, 1)), wp)                                           # taps -2..1
    assert y.shape == ref.shape
    assert torch.allclose(y, ref, atol=1e-3)


def test_unetplusplus_reference_default_architecture_keys_and_known_answers():
    """smp.UnetPlusPlus layout - 26 078 609 parameters for resnet34 (smp's published size), nested
    block names x_{depth}_{layer}, strict state-dict exchange with the oracle restatement, reference default config."""
    from oracle import unet_oracle as O
    from unet_watermark_b200.config import get_cfg_defaults
    from unet_watermark_b200.unet_model import create_model_from_config
    cfg = get_cfg_defaults()                       # MODEL.NAME: UnetPlusPlus (reference src/configs/config.py:15)
    cfg.MODEL.ENCODER_WEIGHTS = None
    m = create_model_from_config(cfg)
    assert type(m).__name__ == "UnetPlusPlus"
    assert sum(p.numel() for p in m.parameters()) == 26_078_609
    ref = O.build("resnet34", arch="UnetPlusPlus", seed=1, random_bn=True)
    assert list(m.state_dict().keys()) == list(ref.state_dict().keys()) and len(m.state_dict()) == 350
    m.load_state_dict(ref.state_dict(), strict=True)
    names = sorted(k for k in m.decoder.blocks.keys())
    assert names == sorted(["x_0_0", "x_0_1", "x_1_1", "x_0_2", "x_1_2", "x_2_2", "x_0_3", "x_1_3", "x_2_3", "x_3_3", "x_0_4"])
    assert m.decoder.blocks["x_0_3"].conv1[0].weight.shape == (32, 320, 3, 3)
    assert m.decoder.blocks["x_1_3"].conv1[0].weight.shape == (64, 256, 3, 3)
    with torch.no_grad():
        assert ref(torch.zeros(1, 3, 64, 96)).shape == (1, 1, 64, 96)
    with pytest.raises(RuntimeError, match="no CPU"):
        m.eval()(torch.zeros(1, 3, 64, 64))
    assert sum(p.numel() for p in O.build("resnet50", arch="UnetPlusPlus").parameters()) == 48_985_745


def test_sub_batch_policy(monkeypatch):
    """Unet._sub_batch_for: explicit attribute > UWM_SUBBATCH > per-encoder pixel budget; never >= the batch."""
    m = Unet("resnet34", encoder_weights=None)
    monkeypatch.delenv("UWM_SUBBATCH", raising=False)
    monkeypatch.setattr(Unet, "_SUB_BATCH_PIXELS", {})
    assert m._sub_batch_for(16, 512, 512) == 0
    monkeypatch.setenv("UWM_SUBBATCH", "8")
    assert m._sub_batch_for(16, 512, 512) == 8 and m._sub_batch_for(8, 512, 512) == 0 and m._sub_batch_for(4, 512, 512) == 0
    m.sub_batch = 4
    assert m._sub_batch_for(16, 512, 512) == 4
    m.sub_batch = 0
    assert m._sub_batch_for(16, 512, 512) == 0
    m.sub_batch = None
    monkeypatch.delenv("UWM_SUBBATCH")
    monkeypatch.setattr(Unet, "_SUB_BATCH_PIXELS", {"resnet34": 8 * 1024 * 1024})
    assert m._sub_batch_for(64, 1024, 1024) == 8 and m._sub_batch_for(16, 512, 512) == 0 and m._sub_batch_for(64, 512, 512) == 32
    assert m._sub_batch_for(4, 4096, 4096) == 1


@pytest.mark.parametrize("shape", [(32, 16, 3), (16, 48, 1), (64, 64, 3)])
def test_dgrad_weights_is_the_transposed_flipped_filter(shape):
    """training.dgrad_weights: conv(gy, flipped / transposed filter) == autograd's data gradient of a 'same' conv."""
    from unet_watermark_b200.training import dgrad_weights, native_dgrad_applies
    cout, cin, k = shape
    g = torch.Generator().manual_seed(cout + cin)
    x = torch.randn(2, cin, 9, 7, generator=g, requires_grad=True)
    w = torch.randn(cout, cin, k, k, generator=g)
    y = F.conv2d(x, w, padding=k // 2)
    gy = torch.randn(y.shape, generator=g)
    y.backward(gy)
    wp = dgrad_weights(w)                                          # UWM_PACK_TAPS: [Cin][k*k][Cout]
    assert wp.shape == (cin, k * k * cout) and wp.is_contiguous()
    gx = F.conv2d(gy, wp.view(cin, k, k, cout).permute(0, 3, 1, 2), padding=k // 2)
    assert torch.allclose(gx, x.grad, atol=1e-4, rtol=1e-4)
    assert native_dgrad_applies(w.shape, 1, k // 2) and not native_dgrad_applies(w.shape, 2, k // 2)
    assert not native_dgrad_applies((cout, 3, 7, 7), 1, 3) and not native_dgrad_applies((1, 16, 3, 3), 1, 1)
