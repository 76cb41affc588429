"""``UnetPlusPlus`` - the reference's DEFAULT architecture (reference src/configs/config.py:15, all three YAMLs;
factory entry src/models/unet_model.py:19) - on the B200 kernels.

Parameter layout = smp's ``UnetPlusPlus`` (ResNet encoder, ``decoder.blocks.x_{depth}_{layer}.conv{1,2}.{0,1}.*``,
``segmentation_head.0.*``), so reference checkpoints load with ``strict=True``; known answer: 26 078 609 parameters
for resnet34 (tests).  smp's nested decoder, restated (smp is an un-vendored dependency):

    features = encoder stages reversed: f0 = layer4 (/32), f1 = layer3, f2 = layer2, f3 = layer1, f4 = stem (/2)
    x_{d}_{d}   = Block(f_d, skip = f_{d+1})                                               d = 0..3
    x_{d}_{L}   = Block(x_{d}_{L-1}, skip = cat(x_{d+1}_{L}, ..., x_{L}_{L}, f_{L+1}))      L > d
    x_{0}_{4}   = Block(x_{0}_{3})                    (no skip);   Block = up2x(nearest) -> cat -> Conv2dReLU x 2

Unlike ``Unet`` (one static plan in C, CUDA-graph replay) this runs as a sequence of single-operator C-ABI calls from
Python.  No concat and no upsampled tensor is materialised: every node writes its output straight into its channel
slice of the per-level buffer ``C_L = [x_1_L | ... | x_L_L | f_{L+1}]`` (the encoder stages write the last slice), so
the skip operand of node (d, L) is the contiguous suffix of ``C_L`` starting at x_{d+1}_L, and
``uwm_conv2d_upcat_nhwc_bf16`` reads (up2x(x), that suffix) directly.  Eval / inference only.
"""
from __future__ import annotations

import logging
from typing import Callable, Dict, Optional, Sequence, Union

import torch
import torch.nn as nn

from . import ops, packing
from .unet_model import _ENCODERS, _Activation, _DecoderBlock, _ResNetEncoder

logger = logging.getLogger(__name__)


class _UnetPlusPlusDecoder(nn.Module):
    def __init__(self, encoder_channels, decoder_channels):
        super().__init__()
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
                blocks[f"x_{depth_idx}_{layer_idx}"] = _DecoderBlock(in_ch, skip_ch, out_ch)
        blocks[f"x_{0}_{len(self.in_channels) - 1}"] = _DecoderBlock(self.in_channels[-1], 0, self.out_channels[-1])
        self.blocks = nn.ModuleDict(blocks)
        self.depth = len(self.in_channels) - 1


class UnetPlusPlus(nn.Module):
    """B200-native ``smp.UnetPlusPlus`` (ResNet-34/50 encoder, depth 5, batch-norm decoder, 1-class head), inference."""

    def __init__(self, encoder_name: str = "resnet34", encoder_depth: int = 5, encoder_weights: Optional[str] = "imagenet",
                 decoder_use_batchnorm: bool = True, decoder_channels: Sequence[int] = (256, 128, 64, 32, 16),
                 decoder_attention_type: Optional[str] = None, in_channels: int = 3, classes: int = 1,
                 activation: Optional[Union[str, Callable]] = None, aux_params: Optional[dict] = None, **kwargs):
        super().__init__()
        if encoder_name not in _ENCODERS:
            raise KeyError(f"Wrong encoder name `{encoder_name}`, supported encoders: {list(_ENCODERS)}")
        bad = []
        if encoder_depth != 5: bad.append(f"encoder_depth={encoder_depth}")
        if decoder_use_batchnorm is not True: bad.append("decoder without batch-norm")
        if decoder_attention_type is not None: bad.append(f"decoder_attention_type={decoder_attention_type}")
        if in_channels != 3: bad.append(f"in_channels={in_channels}")
        if classes != 1: bad.append(f"classes={classes}")
        if activation not in (None, "identity", "sigmoid"): bad.append(f"activation={activation}")
        if aux_params is not None: bad.append("aux_params")
        if len(decoder_channels) != 5:
            raise ValueError(f"Model depth is {encoder_depth}, but you provide `decoder_channels` for "
                             f"{len(decoder_channels)} blocks.")
        if any(int(c) % 16 for c in decoder_channels): bad.append(f"decoder_channels={list(decoder_channels)} (multiples of 16)")
        if bad:
            raise NotImplementedError("unet_watermark_b200.UnetPlusPlus: unsupported: " + ", ".join(bad))
        if encoder_weights is not None:
            logger.info("encoder_weights=%r ignored: no download on the B200 path; load a checkpoint", encoder_weights)
        kind, layers, out_ch = _ENCODERS[encoder_name]
        self.encoder_name = encoder_name
        self.decoder_channels = tuple(int(c) for c in decoder_channels)
        self.activation_name = "sigmoid" if activation == "sigmoid" else None
        self.encoder = _ResNetEncoder(kind, layers, out_ch)
        self.decoder = _UnetPlusPlusDecoder(out_ch, self.decoder_channels)
        self.segmentation_head = nn.Sequential(nn.Conv2d(self.decoder_channels[-1], classes, 3, padding=1), nn.Identity(),
                                               _Activation(activation))
        self.classification_head = None
        self.name = f"unetplusplus-{encoder_name}"
        for m in self.decoder.modules():                       # smp initialize(): kaiming-uniform decoder, xavier head
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_uniform_(m.weight, mode="fan_in", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1); nn.init.constant_(m.bias, 0)
        nn.init.xavier_uniform_(self.segmentation_head[0].weight)
        nn.init.constant_(self.segmentation_head[0].bias, 0)
        self._gen = 0
        self._packed: Dict[str, tuple] = {}
        self._packed_key = None

    # ------------------------------------------------------------------ weights
    def _apply(self, fn, *a, **k):
        self._gen += 1
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, state_dict, *a, **k):
        state_dict = {k_: v for k_, v in state_dict.items() if not k_.startswith("encoder.fc.")}
        self._gen += 1
        return super().load_state_dict(state_dict, *a, **k)

    def refresh_weights(self):
        self._gen += 1

    def __getstate__(self):
        d = self.__dict__.copy()
        d["_packed"], d["_packed_key"] = {}, None
        return d

    def _weights(self, dev):
        key = (self._gen, str(dev), sum(t._version for t in list(self.parameters()) + list(self.buffers())))
        if self._packed_key == key:
            return self._packed
        P: Dict[str, tuple] = {}

        def put(name, conv, bn, stem=False):
            w = conv.weight.detach().float().cpu()
            if bn is not None:
                w, b = packing.fold_bn(w, bn.weight.cpu(), bn.bias.cpu(), bn.running_mean.cpu(), bn.running_var.cpu(), bn.eps)
            else:
                b = conv.bias.detach().float().cpu()
            cout_pad = (w.shape[0] + 15) // 16 * 16
            wp = packing.pack_stem_s2d(w, cout_pad) if stem else packing.pack_taps(w, cout_pad)
            # the 7x7 / stride-2 stem runs as a 4x4 / stride-1 conv (pad 2) over the 2x2 space-to-depth input
            geom = (4, 1, 2) if stem else (conv.kernel_size[0], conv.stride[0], conv.padding[0])
            P[name] = (wp.to(dev), packing.pad_bias(b, cout_pad).to(dev), *geom)
        enc = self.encoder
        put("stem", enc.conv1, enc.bn1, stem=True)
        for li, layer in enumerate((enc.layer1, enc.layer2, enc.layer3, enc.layer4)):
            for bi, blk in enumerate(layer):
                pre = f"l{li + 1}.{bi}"
                put(pre + ".c1", blk.conv1, blk.bn1); put(pre + ".c2", blk.conv2, blk.bn2)
                if hasattr(blk, "conv3"):
                    put(pre + ".c3", blk.conv3, blk.bn3)
                if blk.downsample is not None:
                    put(pre + ".ds", blk.downsample[0], blk.downsample[1])
        for name, blk in self.decoder.blocks.items():
            put(name + ".c1", blk.conv1[0], blk.conv1[1]); put(name + ".c2", blk.conv2[0], blk.conv2[1])
        put("head", self.segmentation_head[0], None)
        self._packed, self._packed_key = P, key
        return P

    # ------------------------------------------------------------------ forward
    @staticmethod
    def check_input_shape(h: int, w: int):
        if h % 32 != 0 or w % 32 != 0:
            nh = (h // 32 + 1) * 32 if h % 32 else h
            nw = (w // 32 + 1) * 32 if w % 32 else w
            raise RuntimeError(f"Wrong input shape height={h}, width={w}. Expected image height and width "
                               f"divisible by 32. Consider pad your images to shape ({nh}, {nw}).")

    def _conv(self, P, name, x, relu=True, residual=None, out=None):
        wp, b, k, s, p = P[name]
        return ops.conv2d(x, wp, b, k, k, s, p, relu=relu, residual=residual, out=out)

    def _run(self, x: torch.Tensor, want_logits: bool, threshold, sigmoid_threshold: bool, force_sigmoid: bool = False):
        if not x.is_cuda:
            raise RuntimeError("unet_watermark_b200.UnetPlusPlus runs only on CUDA (sm_100a) tensors; there is no CPU "
                               "fallback. Move the model and the input to a B200 (`.to('cuda')`).")
        if self.training:
            raise NotImplementedError("unet_watermark_b200.UnetPlusPlus implements eval-mode inference; call model.eval() "
                                      "(the training step of BASELINE config 5 is built for 'Unet')")
        if x.dtype == torch.uint8:
            if x.dim() != 4 or x.shape[-1] != 3:
                raise ValueError("uint8 input must be NHWC RGB [B,H,W,3]")
            n, h, w = x.shape[:3]
        else:
            if x.dim() != 4 or x.shape[1] != 3:
                raise ValueError(f"expected input [B,3,H,W], got {tuple(x.shape)}")
            x = x.float()
            n, h, w = x.shape[0], x.shape[2], x.shape[3]
        self.check_input_shape(h, w)
        dev = x.device
        with torch.cuda.device(dev):
            P = self._weights(dev)
            enc = self.encoder
            sk = self.decoder.skip_channels                              # r34: [256, 128, 64, 64, 0]
            bf = dict(dtype=torch.bfloat16, device=dev)
            # per-level buffers C_L = [x_1_L | ... | x_L_L | f_{L+1}] at the resolution of f_{L+1} (/16, /8, /4, /2)
            C = [torch.empty(n, h >> (4 - L), w >> (4 - L), (L + 1) * sk[L], **bf) for L in range(4)]
            tail = [C[L][..., L * sk[L]:] for L in range(4)]             # where encoder feature f_{L+1} lives
            # ---- encoder (features written into the level buffers) ----
            stem = self._conv(P, "stem", ops.prep_input(x))[:, :h // 2, :w // 2]   # 4x4 s2d stem yields one extra row / column
            tail[3].copy_(stem)
            y = ops.maxpool3x3s2(tail[3])
            stage_out = [tail[2], tail[1], tail[0], None]                # layer1 -> f3, layer2 -> f2, layer3 -> f1, layer4 -> f0
            for li, layer in enumerate((enc.layer1, enc.layer2, enc.layer3, enc.layer4)):
                for bi, blk in enumerate(layer):
                    pre = f"l{li + 1}.{bi}"
                    last = bi == len(layer) - 1
                    idt = self._conv(P, pre + ".ds", y, relu=False) if blk.downsample is not None else y
                    t = self._conv(P, pre + ".c1", y)
                    if hasattr(blk, "conv3"):
                        t = self._conv(P, pre + ".c2", t)
                        y = self._conv(P, pre + ".c3", t, residual=idt, out=stage_out[li] if last else None)
                    else:
                        y = self._conv(P, pre + ".c2", t, residual=idt, out=stage_out[li] if last else None)
            f = [y, tail[0], tail[1], tail[2], tail[3]]                  # f0 .. f4
            # ---- nested decoder ----
            X: Dict[tuple, torch.Tensor] = {}

            def block(d, L, xin, skip):
                name = f"x_{d}_{L}"
                wp, b, _, _, _ = P[name + ".c1"]
                t = ops.conv2d_upcat(xin, skip, wp, b, relu=True, upsample=True)
                out = C[L][..., (d - 1) * sk[L]:d * sk[L]] if (d >= 1 and L < 4) else None
                X[(d, L)] = self._conv(P, name + ".c2", t, out=out)

            depth = self.decoder.depth                                   # 4
            for layer_idx in range(depth):
                for d in range(depth - layer_idx):
                    L = d + layer_idx
                    if layer_idx == 0:
                        block(d, d, f[d], tail[d])
                    else:
                        block(d, L, X[(d, L - 1)], C[L][..., d * sk[L]:])
            block(0, depth, X[(0, depth - 1)], None)
            thr = None
            if threshold is not None:
                thr = float(threshold)
            logits, mask = ops.head(X[(0, depth)], P["head"][0], P["head"][1], threshold=thr,
                                    thr_on_logits=not sigmoid_threshold, want_logits=want_logits,
                                    apply_sigmoid=(self.activation_name == "sigmoid" or force_sigmoid))
        return (logits.unsqueeze(1) if logits is not None else None), mask

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        with torch.no_grad():
            return self._run(x, True, None, True)[0]

    @torch.no_grad()
    def predict(self, x):
        if self.training:
            self.eval()
        return self.forward(x)

    @torch.no_grad()
    def predict_mask(self, x, threshold: float = 0.5, sigmoid: bool = True, return_logits: bool = False, out=None):
        """uint8 [B,H,W] {0,255} mask (same conventions as ``Unet.predict_mask``)."""
        if self.activation_name == "sigmoid" and not sigmoid:
            sigmoid = True
        logits, mask = self._run(x, return_logits, threshold, sigmoid)
        if out is not None:
            out.copy_(mask); mask = out
        return (mask, logits) if return_logits else mask

    @torch.no_grad()
    def predict_proba(self, x):
        return self._run(x, True, None, True, force_sigmoid=True)[0]
