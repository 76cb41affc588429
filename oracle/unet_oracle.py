"""ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.

CPU restatement (plain PyTorch fp32 on torchvision's ResNet) of the arithmetic the reference runs
for its UNet mask-inference path.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this package; ``unet_watermark_b200``
never does.

PARITY UNPINNED: the reference keeps all of this arithmetic in the third-party, un-vendored,
un-pinned dependency ``segmentation-models-pytorch>=0.3.0`` (reference requirements.txt:17), which
is not installed in this image and cannot be fetched (no network), and the reference ships no
golden vectors, known-answer tests or fixtures for the path.  This file
therefore restates smp's published ``Unet`` algorithm (smp 0.3.x–0.5.x, identical for this
configuration) and is anchored on the reference's own call sites:

  * constructor arguments ........ reference src/models/unet_model.py:29-73, :93-120
  * defaults ..................... reference src/configs/config.py:15-22
  * load / eval / forward ........ reference src/predict.py:68-99, :338-345, :610-617
  * threshold + uint8 ............ reference src/predict.py:624-625 (raw output > thr) and
                                   reference src/scripts/watermark_filter.py:135-150 (sigmoid first)
  * preprocessing ................ reference src/utils/dataset.py:389-395
  * checkpoint dict format ....... reference src/train.py:428-435

and on structural known answers of smp's Unet (tests/test_oracle.py): 24 436 369 parameters /
278 state-dict entries for resnet34, 32 521 105 / 380 for resnet50, the exact key names, and
62.512 GFLOP per 512x512 image.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torchvision.models.resnet import BasicBlock, Bottleneck, ResNet

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)

ENCODERS = {
    # name: (block, layers, out_channels)       smp encoders/resnet.py table
    "resnet18": (BasicBlock, [2, 2, 2, 2], (3, 64, 64, 128, 256, 512)),
    "resnet34": (BasicBlock, [3, 4, 6, 3], (3, 64, 64, 128, 256, 512)),
    "resnet50": (Bottleneck, [3, 4, 6, 3], (3, 64, 256, 512, 1024, 2048)),
    "resnet101": (Bottleneck, [3, 4, 23, 3], (3, 64, 256, 512, 1024, 2048)),
}


class ResNetEncoder(ResNet):
    """smp ResNetEncoder: torchvision ResNet minus fc/avgpool, returning the 6 stage outputs."""

    def __init__(self, out_channels: Sequence[int], depth: int = 5, **kwargs):
        super().__init__(**kwargs)
        self._depth = depth
        self._out_channels = tuple(out_channels)
        self._in_channels = 3
        del self.fc
        del self.avgpool

    @property
    def out_channels(self):
        return self._out_channels[: self._depth + 1]

    def forward(self, x: torch.Tensor) -> List[torch.Tensor]:
        feats = [x]                                        # stage 0: identity
        x = self.relu(self.bn1(self.conv1(x)))             # stage 1: /2
        feats.append(x)
        x = self.layer1(self.maxpool(x))                   # stage 2: /4
        feats.append(x)
        for layer in (self.layer2, self.layer3, self.layer4):
            x = layer(x)
            feats.append(x)
        return feats[: self._depth + 1]

    def load_state_dict(self, state_dict, **kwargs):
        state_dict = dict(state_dict)
        state_dict.pop("fc.bias", None)
        state_dict.pop("fc.weight", None)
        return super().load_state_dict(state_dict, **kwargs)


class Conv2dReLU(nn.Sequential):
    """smp base.modules.Conv2dReLU with use_batchnorm=True: conv(no bias) -> BN -> ReLU."""

    def __init__(self, cin: int, cout: int, kernel_size: int, padding: int = 0):
        super().__init__(nn.Conv2d(cin, cout, kernel_size, padding=padding, bias=False),
                         nn.BatchNorm2d(cout), nn.ReLU(inplace=True))


class DecoderBlock(nn.Module):
    def __init__(self, cin: int, cskip: int, cout: int):
        super().__init__()
        self.conv1 = Conv2dReLU(cin + cskip, cout, 3, padding=1)
        self.attention1 = nn.Identity()
        self.conv2 = Conv2dReLU(cout, cout, 3, padding=1)
        self.attention2 = nn.Identity()

    def forward(self, x: torch.Tensor, skip: Optional[torch.Tensor] = None) -> torch.Tensor:
        x = F.interpolate(x, scale_factor=2, mode="nearest")
        if skip is not None:
            x = torch.cat([x, skip], dim=1)                # upsampled first, skip second
            x = self.attention1(x)
        x = self.conv1(x)
        x = self.conv2(x)
        return self.attention2(x)


class UnetDecoder(nn.Module):
    def __init__(self, encoder_channels: Sequence[int], decoder_channels: Sequence[int], n_blocks: int = 5):
        super().__init__()
        if n_blocks != len(decoder_channels):
            raise ValueError(f"Model depth is {n_blocks}, but you provide `decoder_channels` for "
                             f"{len(decoder_channels)} blocks.")
        enc = list(encoder_channels[1:])[::-1]
        head = enc[0]
        ins = [head] + list(decoder_channels[:-1])
        skips = enc[1:] + [0]
        self.center = nn.Identity()
        self.blocks = nn.ModuleList(DecoderBlock(i, s, o) for i, s, o in zip(ins, skips, decoder_channels))

    def forward(self, *features: torch.Tensor) -> torch.Tensor:
        feats = list(features[1:])[::-1]
        x = self.center(feats[0])
        skips = feats[1:]
        for i, blk in enumerate(self.blocks):
            x = blk(x, skips[i] if i < len(skips) else None)
        return x


class Activation(nn.Module):
    def __init__(self, name):
        super().__init__()
        if name is None or name == "identity":
            self.activation = nn.Identity()
        elif name == "sigmoid":
            self.activation = nn.Sigmoid()
        else:
            raise ValueError(f"oracle restates activation None|'identity'|'sigmoid' only; got {name}")

    def forward(self, x):
        return self.activation(x)


class SegmentationHead(nn.Sequential):
    def __init__(self, cin: int, classes: int, activation=None, kernel_size: int = 3):
        super().__init__(nn.Conv2d(cin, classes, kernel_size, padding=kernel_size // 2), nn.Identity(),
                         Activation(activation))


def _init_decoder(module: nn.Module):
    for m in module.modules():
        if isinstance(m, nn.Conv2d):
            nn.init.kaiming_uniform_(m.weight, mode="fan_in", nonlinearity="relu")
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.BatchNorm2d):
            nn.init.constant_(m.weight, 1)
            nn.init.constant_(m.bias, 0)


def _init_head(module: nn.Module):
    for m in module.modules():
        if isinstance(m, (nn.Linear, nn.Conv2d)):
            nn.init.xavier_uniform_(m.weight)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)


class Unet(nn.Module):
    """Restatement of ``smp.Unet(encoder_name, encoder_depth=5, encoder_weights, decoder_use_batchnorm=True,
    decoder_channels, decoder_attention_type=None, in_channels=3, classes, activation)``."""

    def __init__(self, encoder_name: str = "resnet34", encoder_depth: int = 5, encoder_weights=None,
                 decoder_channels: Sequence[int] = (256, 128, 64, 32, 16), in_channels: int = 3, classes: int = 1,
                 activation=None):
        super().__init__()
        if encoder_name not in ENCODERS:
            raise KeyError(f"Wrong encoder name `{encoder_name}`, supported encoders: {list(ENCODERS)}")
        if in_channels != 3 or encoder_depth != 5:
            raise NotImplementedError("oracle restates in_channels=3, encoder_depth=5")
        block, layers, out_ch = ENCODERS[encoder_name]
        # encoder_weights (e.g. "imagenet") would download in smp; the oracle never downloads.
        self.encoder = ResNetEncoder(out_ch, depth=encoder_depth, block=block, layers=layers)
        self.decoder = UnetDecoder(self.encoder.out_channels, list(decoder_channels), n_blocks=encoder_depth)
        self.segmentation_head = SegmentationHead(decoder_channels[-1], classes, activation, kernel_size=3)
        self.classification_head = None
        self.name = f"u-{encoder_name}"
        _init_decoder(self.decoder)
        _init_head(self.segmentation_head)

    @staticmethod
    def check_input_shape(x: torch.Tensor):
        h, w = x.shape[-2:]
        if h % 32 != 0 or w % 32 != 0:
            nh = (h // 32 + 1) * 32 if h % 32 else h
            nw = (w // 32 + 1) * 32 if w % 32 else w
            raise RuntimeError(f"Wrong input shape height={h}, width={w}. Expected image height and width "
                               f"divisible by 32. Consider pad your images to shape ({nh}, {nw}).")

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        self.check_input_shape(x)
        feats = self.encoder(x)
        return self.segmentation_head(self.decoder(*feats))


class UnetPlusPlusDecoder(nn.Module):
    """smp ``UnetPlusPlusDecoder`` (decoders/unetplusplus/decoder.py), restated: nested dense skip pathways.
    Blocks ``x_{depth}_{layer}``; known answer 26 078 609 parameters for resnet34 (tests/test_oracle.py)."""

    def __init__(self, encoder_channels: Sequence[int], decoder_channels: Sequence[int], n_blocks: int = 5):
        super().__init__()
        if n_blocks != len(decoder_channels):
            raise ValueError(f"Model depth is {n_blocks}, but you provide `decoder_channels` for "
                             f"{len(decoder_channels)} blocks.")
        enc = list(encoder_channels[1:])[::-1]
        self.in_channels = [enc[0]] + list(decoder_channels[:-1])
        self.skip_channels = list(enc[1:]) + [0]
        self.out_channels = list(decoder_channels)
        self.center = nn.Identity()
        blocks = {}
        for layer_idx in range(len(self.in_channels) - 1):
            for depth_idx in range(layer_idx + 1):
                if depth_idx == 0:
                    in_ch = self.in_channels[layer_idx]
                    skip_ch = self.skip_channels[layer_idx] * (layer_idx + 1)
                    out_ch = self.out_channels[layer_idx]
                else:
                    out_ch = self.skip_channels[layer_idx]
                    skip_ch = self.skip_channels[layer_idx] * (layer_idx + 1 - depth_idx)
                    in_ch = self.skip_channels[layer_idx - 1]
                blocks[f"x_{depth_idx}_{layer_idx}"] = DecoderBlock(in_ch, skip_ch, out_ch)
        blocks[f"x_{0}_{len(self.in_channels) - 1}"] = DecoderBlock(self.in_channels[-1], 0, self.out_channels[-1])
        self.blocks = nn.ModuleDict(blocks)
        self.depth = len(self.in_channels) - 1

    def forward(self, *features: torch.Tensor) -> torch.Tensor:
        features = list(features[1:])[::-1]
        dense_x = {}
        for layer_idx in range(len(self.in_channels) - 1):
            for depth_idx in range(self.depth - layer_idx):
                if layer_idx == 0:
                    out = self.blocks[f"x_{depth_idx}_{depth_idx}"](features[depth_idx], features[depth_idx + 1])
                    dense_x[f"x_{depth_idx}_{depth_idx}"] = out
                else:
                    dense_l_i = depth_idx + layer_idx
                    cat = [dense_x[f"x_{idx}_{dense_l_i}"] for idx in range(depth_idx + 1, dense_l_i + 1)]
                    cat = torch.cat(cat + [features[dense_l_i + 1]], dim=1)
                    dense_x[f"x_{depth_idx}_{dense_l_i}"] = self.blocks[f"x_{depth_idx}_{dense_l_i}"](
                        dense_x[f"x_{depth_idx}_{dense_l_i - 1}"], cat)
        dense_x[f"x_{0}_{self.depth}"] = self.blocks[f"x_{0}_{self.depth}"](dense_x[f"x_{0}_{self.depth - 1}"])
        return dense_x[f"x_{0}_{self.depth}"]


class UnetPlusPlus(Unet):
    """Restatement of ``smp.UnetPlusPlus`` (the reference's default MODEL.NAME, src/configs/config.py:15): the Unet
    encoder and head with the nested decoder."""

    def __init__(self, encoder_name: str = "resnet34", decoder_channels: Sequence[int] = (256, 128, 64, 32, 16), **kw):
        super().__init__(encoder_name, decoder_channels=decoder_channels, **kw)
        self.decoder = UnetPlusPlusDecoder(self.encoder.out_channels, list(decoder_channels), n_blocks=5)
        self.name = f"unetplusplus-{encoder_name}"
        _init_decoder(self.decoder)


# ----------------------------------------------------------------------------------------------
# bf16-emulating forward: the same arithmetic with the SAME quantisation points as the CUDA path
# (BN folded in fp32, weights rounded once to bf16, every activation rounded once to bf16 after
# the fp32 epilogue, fp32 accumulation, fp32 logits).  Differences to the kernels are then only
# fp32 summation order (and the rare 1-ulp bf16 rounding flips it causes downstream).
# ----------------------------------------------------------------------------------------------
def _r(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).float()


def _fold(conv: nn.Conv2d, bn: nn.BatchNorm2d):
    scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    w = conv.weight.detach().float() * scale.view(-1, 1, 1, 1)
    b = bn.bias.detach().float() - bn.running_mean.detach().float() * scale
    return _r(w), b


def _cbr(x, conv, bn, relu=True, residual=None):
    w, b = _fold(conv, bn)
    y = F.conv2d(x, w, b, stride=conv.stride, padding=conv.padding)
    if residual is not None:
        y = y + residual
    if relu:
        y = y.relu()
    return _r(y)


@torch.no_grad()
def forward_bf16_emulated(model: Unet, x: torch.Tensor, return_features: bool = False):
    """fp32-accumulate / bf16-storage emulation of ``model.eval()(x)``; returns fp32 logits."""
    enc = model.encoder
    feats = {}
    x = _r(x.float())
    x = _cbr(x, enc.conv1, enc.bn1)
    f = [None, x]
    feats["encoder.stem"] = x
    x = F.max_pool2d(x, 3, 2, 1)
    feats["encoder.maxpool"] = x
    for li, layer in enumerate((enc.layer1, enc.layer2, enc.layer3, enc.layer4)):
        for blk in layer:
            idt = x
            if blk.downsample is not None:
                idt = _cbr(x, blk.downsample[0], blk.downsample[1], relu=False)
            if isinstance(blk, BasicBlock):
                t = _cbr(x, blk.conv1, blk.bn1)
                x = _cbr(t, blk.conv2, blk.bn2, residual=idt)
            else:
                t = _cbr(x, blk.conv1, blk.bn1)
                t = _cbr(t, blk.conv2, blk.bn2)
                x = _cbr(t, blk.conv3, blk.bn3, residual=idt)
        f.append(x)
        feats[f"encoder.layer{li + 1}"] = x
    skips = f[1:][::-1]
    x = skips[0]
    for i, blk in enumerate(model.decoder.blocks):
        x = F.interpolate(x, scale_factor=2, mode="nearest")
        if i + 1 < len(skips):
            x = torch.cat([x, skips[i + 1]], dim=1)
        x = _cbr(x, blk.conv1[0], blk.conv1[1])
        x = _cbr(x, blk.conv2[0], blk.conv2[1])
        feats[f"decoder.blocks.{i}"] = x
    hc = model.segmentation_head[0]
    logits = F.conv2d(x, _r(hc.weight.detach().float()), hc.bias.detach().float(), padding=hc.padding)
    if return_features:
        return logits, feats
    return logits


# ----------------------------------------------------------------------------------------------
# pre / post-processing restatements
# ----------------------------------------------------------------------------------------------
def val_transform(image_rgb_u8: np.ndarray, img_size: int = 512) -> torch.Tensor:
    """get_val_transform (reference src/utils/dataset.py:389-395): A.Resize(bilinear) ->
    A.Normalize(ImageNet, max_pixel_value=255) -> ToTensorV2.  uint8 HWC RGB -> fp32 [3,S,S]."""
    import cv2
    img = cv2.resize(image_rgb_u8, (img_size, img_size), interpolation=cv2.INTER_LINEAR)
    mean = np.array(IMAGENET_MEAN, dtype=np.float32) * 255.0
    denom = np.reciprocal(np.array(IMAGENET_STD, dtype=np.float32) * 255.0)
    img = (img.astype(np.float32) - mean) * denom
    return torch.from_numpy(np.ascontiguousarray(img.transpose(2, 0, 1)))


def binarize(output: torch.Tensor, threshold: float = 0.5, sigmoid: bool = True) -> torch.Tensor:
    """`(mask > thr) * 255` as uint8 (reference src/predict.py:624-625); with sigmoid=True the
    watermark_filter.py:136 convention (sigmoid first)."""
    y = torch.sigmoid(output) if sigmoid else output
    return (y > threshold).to(torch.uint8) * 255


def resize_and_binarize(mask_f32: np.ndarray, original_wh, threshold: float = 0.5) -> np.ndarray:
    """reference src/predict.py:620-625: bilinear cv2.resize of the float mask to the original size,
    then threshold -> uint8 {0,255}."""
    import cv2
    m = cv2.resize(mask_f32, tuple(original_wh))
    return (m > threshold).astype(np.uint8) * 255


# ----------------------------------------------------------------------------------------------
# losses of the optional training step (reference src/utils/losses.py:18-31; smp.losses restated)
# ----------------------------------------------------------------------------------------------
def dice_loss_binary(logits: torch.Tensor, target: torch.Tensor, smooth: float = 1e-5, eps: float = 1e-7):
    """smp DiceLoss(mode='binary', from_logits=True): batch-global sums over dims (0,2)."""
    bs = target.size(0)
    p = F.logsigmoid(logits).exp().view(bs, 1, -1)
    t = target.view(bs, 1, -1).type_as(p)
    inter = torch.sum(p * t, dim=(0, 2))
    card = torch.sum(p + t, dim=(0, 2))
    dice = (2.0 * inter + smooth) / (card + smooth).clamp_min(eps)
    loss = (1.0 - dice) * (t.sum((0, 2)) > 0).to(p.dtype)
    return loss.mean()


def dice_bce_loss(logits, target, dice_weight=0.5, bce_weight=0.5, smooth=1e-5):
    """config-5 composition: DICE_WEIGHT*Dice + BCE_WEIGHT*BCEWithLogits (reference src/configs/config.py:61-62)."""
    return dice_weight * dice_loss_binary(logits, target, smooth) + \
        bce_weight * F.binary_cross_entropy_with_logits(logits, target.type_as(logits))


# ----------------------------------------------------------------------------------------------
# fixtures
# ----------------------------------------------------------------------------------------------
def image_like_input(batch: int, size_hw, seed: int = 0, normalise: bool = True) -> torch.Tensor:
    """Seeded image-like fp32 [B,3,H,W]: bicubic-upsampled U[0,1] field on an S/32 grid + 0.03 N(0,1),
    clamped to [0,1], then ImageNet-normalised (what get_val_transform feeds the net)."""
    h, w = size_hw if isinstance(size_hw, (tuple, list)) else (size_hw, size_hw)
    g = torch.Generator().manual_seed(seed)
    low = torch.rand(batch, 3, max(h // 32, 2), max(w // 32, 2), generator=g)
    x = F.interpolate(low, size=(h, w), mode="bicubic", align_corners=False)
    x = (x + 0.03 * torch.randn(batch, 3, h, w, generator=g)).clamp_(0, 1)
    if normalise:
        mean = torch.tensor(IMAGENET_MEAN).view(1, 3, 1, 1)
        std = torch.tensor(IMAGENET_STD).view(1, 3, 1, 1)
        x = (x - mean) / std
    return x.contiguous()


def image_like_u8(batch: int, size_hw, seed: int = 0) -> torch.Tensor:
    """uint8 NHWC RGB version of image_like_input (pre-normalisation)."""
    x = image_like_input(batch, size_hw, seed, normalise=False)
    return (x * 255.0).round().clamp_(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous()


def randomize_bn(model: nn.Module, seed: int = 0) -> nn.Module:
    """Non-trivial BN affine + running stats so that a folding bug cannot hide behind identity BN."""
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, nn.BatchNorm2d):
            n = m.num_features
            with torch.no_grad():
                m.weight.copy_(0.8 + 0.4 * torch.rand(n, generator=g))
                m.bias.copy_(0.1 * torch.randn(n, generator=g))
                m.running_mean.copy_(0.1 * torch.randn(n, generator=g))
                m.running_var.copy_(0.7 + 0.6 * torch.rand(n, generator=g))
    return model


def build(encoder_name="resnet34", decoder_channels=(256, 128, 64, 32, 16), seed: int = 0, random_bn: bool = False,
          activation=None, arch: str = "Unet") -> Unet:
    torch.manual_seed(seed)
    cls = {"Unet": Unet, "UnetPlusPlus": UnetPlusPlus}[arch]
    m = cls(encoder_name, decoder_channels=decoder_channels, activation=activation)
    if random_bn:
        randomize_bn(m, seed + 1)
    return m.eval()


def conv_flops_per_image(model: Unet, h: int, w: int) -> float:
    """2*MACs of every Conv2d for one [1,3,h,w] input (hooks; equals the analytic count)."""
    total = 0.0
    hooks = []

    def hook(mod, inp, out):
        nonlocal total
        total += 2.0 * out.shape[1] * out.shape[2] * out.shape[3] * mod.in_channels * \
            mod.kernel_size[0] * mod.kernel_size[1] / mod.groups

    for m in model.modules():
        if isinstance(m, nn.Conv2d):
            hooks.append(m.register_forward_hook(hook))
    with torch.no_grad():
        model(torch.zeros(1, 3, h, w))
    for hk in hooks:
        hk.remove()
    return total
