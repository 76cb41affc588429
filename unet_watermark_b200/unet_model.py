"""Drop-in mirror of the reference model seam (reference src/models/unet_model.py).

Same public names, signatures and errors as the reference:
``SMPModelFactory.create_model`` (:29-73), ``create_model_from_config`` (:93-120),
``WatermarkSegmentationModel`` (:123-150).  For ``('Unet', 'resnet34'|'resnet50')`` the factory
returns :class:`Unet`, an ``nn.Module`` whose parameters/buffers have exactly smp's names and
shapes so existing ``.pth`` files load with ``strict=True`` — but whose
``forward`` runs the hand-written sm_100a kernels of ``libuwm_b200.so`` instead of
torch/cuDNN modules.  The ``nn.Conv2d`` / ``nn.BatchNorm2d`` objects below are parameter
containers only: their ``forward`` is never called on the inference path.
"""
from __future__ import annotations

import logging
import os
from typing import Callable, Dict, List, Optional, Sequence, Union

import torch
import torch.nn as nn

from .engine import Engine

logger = logging.getLogger(__name__)

_ENCODERS = {
    # name: (block kind, blocks per stage, out_channels)   (smp encoder table for torchvision ResNets)
    "resnet34": ("basic", (3, 4, 6, 3), (3, 64, 64, 128, 256, 512)),
    "resnet50": ("bottleneck", (3, 4, 6, 3), (3, 64, 256, 512, 1024, 2048)),
}


def _conv(cin, cout, k, stride=1, padding=0, bias=False):
    return nn.Conv2d(cin, cout, k, stride=stride, padding=padding, bias=bias)


class _BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride, downsample):
        super().__init__()
        self.conv1 = _conv(inplanes, planes, 3, stride, 1)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = _conv(planes, planes, 3, 1, 1)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = downsample
        self.stride = stride


class _Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride, downsample):
        super().__init__()
        self.conv1 = _conv(inplanes, planes, 1)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = _conv(planes, planes, 3, stride, 1)     # torchvision v1.5: stride on the 3x3
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = _conv(planes, planes * 4, 1)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride


class _ResNetEncoder(nn.Module):
    """Parameter layout of smp's ResNetEncoder (torchvision ResNet without fc/avgpool)."""

    def __init__(self, kind: str, layers: Sequence[int], out_channels: Sequence[int]):
        super().__init__()
        block = _BasicBlock if kind == "basic" else _Bottleneck
        self.out_channels = tuple(out_channels)
        self.inplanes = 64
        self.conv1 = _conv(3, 64, 7, 2, 3)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, 2, 1)
        self.layer1 = self._make_layer(block, 64, layers[0], 1)
        self.layer2 = self._make_layer(block, 128, layers[1], 2)
        self.layer3 = self._make_layer(block, 256, layers[2], 2)
        self.layer4 = self._make_layer(block, 512, layers[3], 2)
        # torchvision also builds (and smp then deletes) the classifier; creating it keeps the RNG
        # stream of a seeded construction identical to smp's.
        _ = nn.Linear(512 * block.expansion, 1000)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def _make_layer(self, block, planes, blocks, stride):
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            downsample = nn.Sequential(_conv(self.inplanes, planes * block.expansion, 1, stride),
                                       nn.BatchNorm2d(planes * block.expansion))
        layers = [block(self.inplanes, planes, stride, downsample)]
        self.inplanes = planes * block.expansion
        for _ in range(1, blocks):
            layers.append(block(self.inplanes, planes, 1, None))
        return nn.Sequential(*layers)


class _Conv2dReLU(nn.Sequential):
    def __init__(self, cin, cout):
        super().__init__(_conv(cin, cout, 3, 1, 1), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))


class _DecoderBlock(nn.Module):
    def __init__(self, cin, cskip, cout):
        super().__init__()
        self.conv1 = _Conv2dReLU(cin + cskip, cout)
        self.attention1 = nn.Identity()
        self.conv2 = _Conv2dReLU(cout, cout)
        self.attention2 = nn.Identity()


class _UnetDecoder(nn.Module):
    def __init__(self, encoder_channels, decoder_channels):
        super().__init__()
        enc = list(encoder_channels[1:])[::-1]
        ins = [enc[0]] + list(decoder_channels[:-1])
        skips = enc[1:] + [0]
        self.center = nn.Identity()
        self.blocks = nn.ModuleList(_DecoderBlock(i, s, o) for i, s, o in zip(ins, skips, decoder_channels))


class _Activation(nn.Module):
    def __init__(self, name):
        super().__init__()
        self.activation = nn.Sigmoid() if name == "sigmoid" else nn.Identity()


class Unet(nn.Module):
    """B200-native ``smp.Unet`` (ResNet-34/50 encoder, depth 5, batch-norm decoder, 1-class head).

    ``forward(x)``: ``x`` float32 ``[B,3,H,W]`` on a CUDA device, H and W divisible by 32 →
    float32 ``[B,1,H,W]`` logits (probabilities if ``activation='sigmoid'``), like the reference
    model under ``model.eval()`` + ``torch.no_grad()`` (reference src/predict.py:338-345).
    """

    def __init__(self, encoder_name: str = "resnet34", encoder_depth: int = 5,
                 encoder_weights: Optional[str] = "imagenet", decoder_use_batchnorm: bool = True,
                 decoder_channels: Sequence[int] = (256, 128, 64, 32, 16),
                 decoder_attention_type: Optional[str] = None, in_channels: int = 3, classes: int = 1,
                 activation: Optional[Union[str, Callable]] = None, aux_params: Optional[dict] = None, **kwargs):
        super().__init__()
        if encoder_name not in _ENCODERS:
            raise KeyError(f"Wrong encoder name `{encoder_name}`, supported encoders: {list(_ENCODERS)}")
        unsupported = []
        if encoder_depth != 5:
            unsupported.append(f"encoder_depth={encoder_depth}")
        if decoder_use_batchnorm is not True or kwargs.get("decoder_use_norm", True) not in (True, "batchnorm"):
            unsupported.append("decoder without batch-norm")
        if decoder_attention_type is not None:
            unsupported.append(f"decoder_attention_type={decoder_attention_type}")
        if in_channels != 3:
            unsupported.append(f"in_channels={in_channels}")
        if classes != 1:
            unsupported.append(f"classes={classes}")
        if activation not in (None, "identity", "sigmoid"):
            unsupported.append(f"activation={activation}")
        if aux_params is not None:
            unsupported.append("aux_params")
        if len(decoder_channels) != 5:
            raise ValueError(f"Model depth is {encoder_depth}, but you provide `decoder_channels` for "
                             f"{len(decoder_channels)} blocks.")
        if any(int(c) % 16 for c in decoder_channels):
            unsupported.append(f"decoder_channels={list(decoder_channels)} (multiples of 16 required)")
        if unsupported:
            raise NotImplementedError("unet_watermark_b200.Unet implements the reference's mask path only; "
                                      "unsupported: " + ", ".join(unsupported))
        if encoder_weights is not None:
            # smp would download ImageNet weights here (reference src/configs/config.py:17); the predict path
            # overwrites them with the .pth right after (reference src/predict.py:75-81) and this
            # implementation never touches the network.
            logger.info("encoder_weights=%r ignored: no download on the B200 path; load a checkpoint", encoder_weights)
        kind, layers, out_ch = _ENCODERS[encoder_name]
        self.encoder_name = encoder_name
        self.decoder_channels = tuple(int(c) for c in decoder_channels)
        self.activation_name = "sigmoid" if activation == "sigmoid" else None
        self.encoder = _ResNetEncoder(kind, layers, out_ch)
        self.decoder = _UnetDecoder(out_ch, self.decoder_channels)
        self.segmentation_head = nn.Sequential(nn.Conv2d(self.decoder_channels[-1], classes, 3, padding=1),
                                               nn.Identity(), _Activation(activation))
        self.classification_head = None
        self.name = f"u-{encoder_name}"
        self._initialize()
        self._engines: Dict[tuple, Engine] = {}
        self._weights_gen = 0          # bumped whenever parameters may have been replaced or edited
        self._tensors_gen = -1
        self._tensors: list = []
        self.use_cuda_graph = True
        self.sub_batch: Optional[int] = None      # None: UWM_SUBBATCH or the measured rule of _sub_batch_for; 0: never

    # smp SegmentationModel.initialize(): decoder kaiming-uniform, head xavier-uniform
    def _initialize(self):
        for m in self.decoder.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_uniform_(m.weight, mode="fan_in", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        hc = self.segmentation_head[0]
        nn.init.xavier_uniform_(hc.weight)
        nn.init.constant_(hc.bias, 0)

    # ------------------------------------------------------------------ engine management
    @staticmethod
    def check_input_shape(h: int, w: int):
        if h % 32 != 0 or w % 32 != 0:
            nh = (h // 32 + 1) * 32 if h % 32 else h
            nw = (w // 32 + 1) * 32 if w % 32 else w
            raise RuntimeError(f"Wrong input shape height={h}, width={w}. Expected image height and width "
                               f"divisible by 32. Consider pad your images to shape ({nh}, {nw}).")

    def _token(self):
        """Fingerprint of "weights changed": a generation counter bumped by load_state_dict / .to() / refresh_weights
        plus the autograd version counter of EVERY parameter and buffer (in-place updates - optimizer steps,
        ``p.add_()``, ``p.copy_()`` - bump it).  Writes through ``p.data`` bypass the version counter: call
        ``refresh_weights()`` after those."""
        if self._tensors_gen != self._weights_gen:
            self._tensors = [t for t in list(self.parameters()) + list(self.buffers())]
            self._tensors_gen = self._weights_gen
        v = 0
        for t in self._tensors:
            v += t._version
        return (self._weights_gen, v)

    def refresh_weights(self):
        """Re-fold BatchNorm and re-upload packed weights (needed after edits through ``.data``)."""
        self._weights_gen += 1

    def _engine_for(self, b: int, h: int, w: int, device: torch.device) -> Engine:
        key = (h, w, device.index if device.index is not None else torch.cuda.current_device())
        eng = self._engines.get(key)
        if eng is None or eng.max_batch < b:
            if eng is not None:
                eng.close()
            eng = Engine(self.encoder_name, self.decoder_channels, h, w, max(b, 1), device)
            self._engines[key] = eng
        tok = self._token()
        if not eng.weights_loaded or getattr(eng, "_token", None) != tok:      # compared per engine
            eng.load_weights(self.state_dict())
            eng._token = tok
        return eng

    # hand-over tensors of a whole batch beyond this many input pixels no longer fit the 126 MB L2; (encoder ->
    # (pixel budget per sub-forward), from the sub-batch sweep of tools/gpu_subbatch_sweep.py (profiles/)
    _SUB_BATCH_PIXELS: Dict[str, int] = {}

    def _sub_batch_for(self, b: int, h: int, w: int) -> int:
        """Images per forward launch sequence (see Engine.forward): explicit attribute, else ``UWM_SUBBATCH``, else
        the measured per-encoder pixel budget.  0 / >= b: the whole batch in one plan."""
        sb = self.sub_batch
        if sb is None:
            env = os.environ.get("UWM_SUBBATCH", "")
            if env:
                sb = int(env)
            else:
                budget = self._SUB_BATCH_PIXELS.get(self.encoder_name, 0)
                sb = max(1, budget // (h * w)) if budget else 0
        return 0 if not sb or sb >= b else int(sb)

    def _apply(self, fn, *a, **k):
        self._weights_gen += 1
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, state_dict, *a, **k):
        state_dict = {k_: v for k_, v in state_dict.items()
                      if not k_.startswith("encoder.fc.")}       # smp ResNetEncoder drops fc.*
        self._weights_gen += 1
        return super().load_state_dict(state_dict, *a, **k)

    def __getstate__(self):
        # engines hold C handles / device workspaces: never pickled or deep-copied
        d = self.__dict__.copy()
        d["_engines"] = {}
        d["_tensors"] = []
        d["_tensors_gen"] = -1
        return d

    # ------------------------------------------------------------------ forward
    def _run(self, x: torch.Tensor, want_logits: bool, threshold, sigmoid_threshold: bool,
             force_sigmoid: bool = False, mask_out: Optional[torch.Tensor] = None):
        if not x.is_cuda:
            raise RuntimeError("unet_watermark_b200.Unet runs only on CUDA (sm_100a) tensors; there is no CPU "
                               "fallback. Move the model and the input to a B200 (`.to('cuda')`).")
        if self.training:
            raise RuntimeError("internal: the inference plan (BatchNorm folded) was reached in train mode")
        if x.dtype == torch.uint8:
            if x.dim() != 4 or x.shape[-1] != 3:
                raise ValueError("uint8 input must be NHWC RGB [B,H,W,3]")
            b, h, w = x.shape[0], x.shape[1], x.shape[2]
        else:
            if x.dim() != 4 or x.shape[1] != 3:
                raise ValueError(f"expected input [B,3,H,W], got {tuple(x.shape)}")
            x = x.float()
            b, h, w = x.shape[0], x.shape[2], x.shape[3]
        self.check_input_shape(h, w)
        eng = self._engine_for(b, h, w, x.device)
        with torch.cuda.device(x.device):
            return eng.forward(x, want_logits=want_logits, threshold=threshold,
                               sigmoid_threshold=sigmoid_threshold,
                               apply_sigmoid=(self.activation_name == "sigmoid" or force_sigmoid),
                               use_graph=self.use_cuda_graph, mask_out=mask_out,
                               sub_batch=self._sub_batch_for(b, h, w))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """eval mode: the fused inference plan (BatchNorm folded, no autograd), as the reference runs it under
        ``model.eval()`` + ``torch.no_grad()``.  train mode (reference src/train.py:70,89): BatchNorm with batch
        statistics and autograd enabled - the convolutions still run on the tcgen05 kernels (training.py)."""
        if self.training:
            from .training import forward_train
            return forward_train(self, x)
        with torch.no_grad():
            logits, _ = self._run(x, True, None, True)
        return logits

    @torch.no_grad()
    def predict(self, x: torch.Tensor) -> torch.Tensor:
        """smp SegmentationModel.predict: eval + no_grad forward."""
        if self.training:
            self.eval()
        return self.forward(x)

    @torch.no_grad()
    def predict_mask(self, x: torch.Tensor, threshold: float = 0.5, sigmoid: bool = True,
                     return_logits: bool = False, out: Optional[torch.Tensor] = None):
        """Fused forward + threshold: uint8 ``[B,H,W]`` mask with values {0,255}.

        ``sigmoid=True``: ``sigmoid(logit) > threshold`` (reference src/scripts/watermark_filter.py:136-150,
        evaluated as ``logit > log(t/(1-t))``); ``sigmoid=False``: raw output ``> threshold``
        (reference src/predict.py:624-625).  ``x`` may be fp32 NCHW (normalised) or uint8 NHWC RGB
        (normalisation fused on the GPU).  ``out``: optional caller-owned uint8 ``[B,H,W]`` CUDA buffer for the
        mask; with stable input/output buffers every call replays one cached CUDA graph."""
        if self.activation_name == "sigmoid" and not sigmoid:
            # the model output already is a probability: compare it against thr  <=>  logit > logit(thr)
            sigmoid = True
        if out is not None and (out.dtype != torch.uint8 or not out.is_cuda or not out.is_contiguous()
                                or tuple(out.shape) != (x.shape[0],) + tuple(x.shape[1:3] if x.dtype == torch.uint8 else x.shape[2:4])):
            raise ValueError("out must be a contiguous CUDA uint8 tensor of shape [B,H,W]")
        logits, mask = self._run(x, return_logits, threshold, sigmoid, mask_out=out)
        return (mask, logits) if return_logits else mask

    @torch.no_grad()
    def predict_proba(self, x: torch.Tensor) -> torch.Tensor:
        """fp32 ``[B,1,H,W]`` sigmoid(logits), the sigmoid fused into the head kernel
        (reference src/scripts/watermark_filter.py:136)."""
        probs, _ = self._run(x, True, None, True, force_sigmoid=True)
        return probs

    def engine(self, b: int, h: int, w: int, device=None) -> Engine:
        dev = torch.device(device) if device is not None else next(self.parameters()).device
        return self._engine_for(b, h, w, dev)


class _OutOfScope:
    """Architectures the reference factory lists but this hot-path build does not implement."""

    def __init__(self, name):
        self.name = name

    def __call__(self, *a, **k):
        raise NotImplementedError(f"{self.name}: 'Unet' and 'UnetPlusPlus' are implemented by the B200 mask-inference "
                                  "path (SURVEY.md §8, DESIGN.md 'out of scope')")


def _unetplusplus(*a, **k):
    from .unetpp import UnetPlusPlus          # imported lazily: unetpp builds on this module's encoder / blocks
    return UnetPlusPlus(*a, **k)


class SMPModelFactory:
    """Mirror of reference src/models/unet_model.py:11-91."""

    SUPPORTED_MODELS = {
        "Unet": Unet,
        "UnetPlusPlus": _unetplusplus,
        "MAnet": _OutOfScope("MAnet"),
        "Linknet": _OutOfScope("Linknet"),
        "FPN": _OutOfScope("FPN"),
        "PSPNet": _OutOfScope("PSPNet"),
        "PAN": _OutOfScope("PAN"),
        "DeepLabV3": _OutOfScope("DeepLabV3"),
        "DeepLabV3Plus": _OutOfScope("DeepLabV3Plus"),
    }

    @classmethod
    def create_model(cls, model_name: str, encoder_name: str = "resnet34",
                     encoder_weights: Optional[str] = "imagenet", in_channels: int = 3, classes: int = 1,
                     activation: Optional[Union[str, Callable]] = None, **kwargs) -> nn.Module:
        if model_name not in cls.SUPPORTED_MODELS:
            raise ValueError(f"Unsupported model: {model_name}. "
                             f"Supported models: {list(cls.SUPPORTED_MODELS.keys())}")
        model_class = cls.SUPPORTED_MODELS[model_name]
        return model_class(encoder_name=encoder_name, encoder_weights=encoder_weights, in_channels=in_channels,
                           classes=classes, activation=activation, **kwargs)

    @classmethod
    def get_available_encoders(cls) -> List[str]:
        return list(_ENCODERS.keys())

    @classmethod
    def get_encoder_info(cls, encoder_name: str) -> dict:
        try:
            kind, layers, out_ch = _ENCODERS[encoder_name]
            return {"name": encoder_name, "params": {"block": kind, "layers": list(layers)},
                    "out_channels": out_ch}
        except Exception as e:  # noqa: BLE001 - same contract as the reference (:89-90)
            return {"error": str(e)}


def create_model_from_config(cfg) -> nn.Module:
    """Mirror of reference src/models/unet_model.py:93-120."""
    model_params = {
        "model_name": cfg.MODEL.NAME,
        "encoder_name": cfg.MODEL.ENCODER_NAME,
        "encoder_weights": cfg.MODEL.ENCODER_WEIGHTS,
        "in_channels": cfg.MODEL.IN_CHANNELS,
        "classes": cfg.MODEL.CLASSES,
        "activation": cfg.MODEL.ACTIVATION,
    }
    if hasattr(cfg.MODEL, "ENCODER_DEPTH"):
        model_params["encoder_depth"] = cfg.MODEL.ENCODER_DEPTH
    if hasattr(cfg.MODEL, "DECODER_CHANNELS"):
        model_params["decoder_channels"] = cfg.MODEL.DECODER_CHANNELS
    return SMPModelFactory.create_model(**model_params)


class WatermarkSegmentationModel(nn.Module):
    """Mirror of reference src/models/unet_model.py:123-150."""

    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        self.model = create_model_from_config(cfg)

    def forward(self, x):
        return self.model(x)

    def get_model_info(self) -> dict:
        total_params = sum(p.numel() for p in self.parameters())
        trainable_params = sum(p.numel() for p in self.parameters() if p.requires_grad)
        return {
            "model_name": self.cfg.MODEL.NAME,
            "encoder_name": self.cfg.MODEL.ENCODER_NAME,
            "total_params": total_params,
            "trainable_params": trainable_params,
            "input_channels": self.cfg.MODEL.IN_CHANNELS,
            "output_classes": self.cfg.MODEL.CLASSES,
        }
